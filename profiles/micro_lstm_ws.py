"""A/B of the two bf16 recurrence kernels at the bench shape (4096 questions, H = 512): the streaming kernel (csrc/lstm_fused.cu) vs the
weight-stationary cluster kernel (csrc/lstm_ws.cu).  Times are whole encoder phases (input-projection GEMM + recurrence), CUDA events;
the difference of the two columns is the recurrence itself.  Also checks that both kernels produce the same encoder outputs."""
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
for ws in (0, 1):
    lib.stair_set_lstm_ws(ws)
    both = run(L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    text = run(L.FWD_ENCODE_TEXT)
    video = run(L.FWD_ENCODE_VIDEO)
    full = run(L.FWD_ALL)
    st = model.forward_batch(batch, phases=L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    torch.cuda.synchronize()
    outs[ws] = (st.vid[:B * T * 512].float().clone(), st.tokfeat[:batch.n_tok * 512].float().clone(), st.qfeat[:B * 512].float().clone())
    print('%-28s video+text %.3f ms   text only %.3f   video only %.3f   whole forward %.3f   (GEMMs included; gemm error flag %d)'
          % ('weight-stationary (lstm_ws)' if ws else 'streaming (lstm_fused)', both, text, video, full, lib.stair_gemm_error_flag()), flush=True)
lib.stair_set_lstm_ws(1)
for name, a, b in zip(('video_feat', 'token_feature', 'question_feature'), outs[0], outs[1]):
    print('%-18s max |ws - fused| = %.3g   (max |x| %.3g)' % (name, float((a - b).abs().max()), float(a.abs().max())))
