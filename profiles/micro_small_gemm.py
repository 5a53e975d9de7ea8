"""A/B of the two-CTAs-per-SM GEMM form for the module-sized GEMMs (stair_set_gemm_small) at the bench shape: whole forward and the
module phase alone, CUDA events over 20 forwards, two passes; checks the logits are bit-identical."""
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


def run(ph, n=20):
    for _ in range(3):
        model.forward_batch(batch, phases=ph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        st = model.forward_batch(batch, phases=ph)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, st


ref = None
for rep in range(2):
    for small in (0, 1):
        lib.stair_set_gemm_small(small)
        full, st = run(L.FWD_ALL)
        lg = st.logits.float().clone()
        mods, _ = run(L.FWD_MODULES)
        dec, _ = run(L.FWD_DECODE)
        if ref is None:
            ref = lg
        print('gemm_small %d: whole forward %.3f ms, modules alone %.3f ms, decoder alone %.3f ms, max |dlogit| vs first %.3g, gemm error flag %d'
              % (small, full, mods, dec, float((lg - ref).abs().max()), lib.stair_gemm_error_flag()), flush=True)
lib.stair_set_gemm_small(1)
