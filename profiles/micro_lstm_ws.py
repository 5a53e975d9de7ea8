"""A/B of the bf16 recurrence kernels at the bench shape (4096 questions, H = 512): the streaming kernel (csrc/lstm_fused.cu) vs the
weight-stationary cluster kernel (csrc/lstm_ws.cu).
Times are whole encoder phases (input-projection GEMM + recurrence), CUDA events; the differences between rows are the recurrence itself.
Also checks that all kernels produce the same encoder outputs."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T, V = 8, 4096
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


outs = {}
for name, ws, rows, cg in (('streaming (lstm_fused)', 0, 64, 4), ('streaming 64 rows, 16 gate-serial warps', 0, 64, 8),
                           ('streaming 128 rows, 8 warps (round 1)', 0, 128, 2),
                           ('weight-stationary (lstm_ws)', 1, 64, 4)):
    lib.stair_set_lstm_ws(ws)
    lib.stair_lstm_rows(rows)
    lib.stair_lstm_colgroups(cg)
    both = run(L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    text = run(L.FWD_ENCODE_TEXT)
    video = run(L.FWD_ENCODE_VIDEO)
    full = run(L.FWD_ALL)
    st = model.forward_batch(batch, phases=L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    torch.cuda.synchronize()
    outs[name] = (st.vid[:B * T * 512].float().clone(), st.tokfeat[:batch.n_tok * 512].float().clone(), st.qfeat[:B * 512].float().clone())
    print('%-42s video+text %.3f ms   text only %.3f   video only %.3f   whole forward %.3f   (GEMMs included; gemm error flag %d)'
          % (name, both, text, video, full, lib.stair_gemm_error_flag()), flush=True)
lib.stair_set_lstm_ws(0)
lib.stair_lstm_rows(64)
lib.stair_lstm_colgroups(4)
ref = outs['streaming (lstm_fused)']
for name, o in outs.items():
    print('%-42s max |x - streaming|: video_feat %.3g, token_feature %.3g, question_feature %.3g'
          % (name, *(float((a - b).abs().max()) for a, b in zip(o, ref))))
