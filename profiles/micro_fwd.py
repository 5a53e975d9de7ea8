"""One resident 4096-question batch, a few inference forwards (for ncu kernel filters); LSTM_CG=8 selects the 16-epilogue-warp
recurrence variant and checks it against the default."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
B, T, V = 4096, 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
L.lib().stair_set_lanes(int(os.environ.get('LANES', 6)))
L.lib().stair_set_gemm_epi2(int(os.environ.get('GEMM_EPI2', 1)))
L.lib().stair_set_gemm_wide_min(int(os.environ.get('GEMM_WIDE_MIN', 1)))
ref = None
for cg in [2] + ([int(os.environ['LSTM_CG'])] if 'LSTM_CG' in os.environ else []):
    L.lib().stair_lstm_colgroups(cg)
    for _ in range(4):
        st = model.forward_batch(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        st = model.forward_batch(batch)
    e1.record(); torch.cuda.synchronize()
    lg = st.logits.float().clone()
    print('colgroups %d: %.3f ms per forward' % (cg, e0.elapsed_time(e1) / 20), flush=True)
    if lg is not None:
        if ref is None:
            ref = lg
        else:
            print('max |dlogit| vs default: %.3g' % float((lg - ref).abs().max()))
L.lib().stair_lstm_colgroups(2)
