"""Time the fused BiLSTM recurrence alone (csrc/lstm_fused.cu) at the bench shape, comparing epilogue warp counts."""
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
lib = L.lib()
def run(ph, n=10):
    for _ in range(3):
        model.forward_batch(batch, phases=ph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        model.forward_batch(batch, phases=ph)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for rows, cg in [(128, 2), (128, 4), (64, 4)]:
    lib.stair_lstm_rows(rows)
    lib.stair_lstm_colgroups(cg)
    both = run(L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    text = run(L.FWD_ENCODE_TEXT)
    video = run(L.FWD_ENCODE_VIDEO)
    st = model.forward_batch(batch, phases=L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    torch.cuda.synchronize()
    out = (st.vid[:B * T * 512].float().clone(), st.qfeat[:B * 512].float().clone())
    if ref is None:
        ref = out
    err = max(float((a - b).abs().max()) for a, b in zip(out, ref))
    print('rows/CTA %d colgroups %d: video+text %.3f ms   text only %.3f   video only %.3f   (GEMMs included; max diff vs cg=2 %.3g)' % (rows, cg, both, text, video, err), flush=True)
lib.stair_lstm_colgroups(2)
lib.stair_lstm_rows(64)
