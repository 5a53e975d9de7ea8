#!/bin/bash
# where does the weight-stationary recurrence spend its time?  config print + ncu --set full of the text-encoder launch; dep-sched A/B
export STAIR_LSTM_WS=1 STAIR_DEBUG=1
python profiles/micro_lstm_one.py text > gpurun_out/lstm_one_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lstm_ws -s 4 -c 1 -o gpurun_out/lstm_ws_text_r2 python profiles/micro_lstm_one.py text > gpurun_out/ncu_lstm_ws.log 2>&1; echo "ncu rc=$?"; cat gpurun_out/lstm_one_plain.log
unset STAIR_LSTM_WS
python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_r2e.log 2>&1; echo "pytest (dep sched) rc=$?"; tail -4 gpurun_out/gpu_tests_r2e.log
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
B, T, V = 4096, 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
lib = L.lib()
def run(ph, n=20):
    for _ in range(4):
        model.forward_batch(batch, phases=ph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        st = model.forward_batch(batch, phases=ph)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, st
model.forward_batch(batch)
ref = None
for dep in (0, 1, 0, 1):
    lib.stair_set_dep_sched(dep)
    for lanes in (4, 6, 8):
        lib.stair_set_lanes(lanes)
        m, _ = run(L.FWD_MODULES)
        f, st = run(L.FWD_ALL)
        lg = st.logits.clone()
        if ref is None: ref = lg
        print('dep_sched %d lanes %d: modules %.3f ms, whole forward %.3f ms, logits equal %s' % (dep, lanes, m, f, bool(torch.equal(lg, ref))), flush=True)
PY
