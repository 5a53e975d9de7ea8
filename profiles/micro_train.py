"""One resident 4096-question window, N training steps (forward with history + losses + backward + Adam); for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate
from stair_b200.train import NMNTrainStep, FusedAdam

B, T, V = int(os.environ.get('B', 4096)), 8, 4096
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
cfg = syn.model_config(T=T, V=V, dropout=float(os.environ.get('DROPOUT', 0.25)))
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().train()
qs = syn.make_questions(B, T, V, seed=1234, with_gold=True)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
from stair_b200 import _lib as L
L.lib().stair_set_bwd_lanes(int(os.environ.get('BWD_LANES', 4)))
L.lib().stair_set_bptt_impl(int(os.environ.get('BPTT_IMPL', 0)))
L.lib().stair_set_gemm_wide_min(int(os.environ.get('GEMM_WIDE_MIN', 1)))
L.lib().stair_set_gemm_pair_mn(int(os.environ.get('PAIR_MN', 1)))       # CTA-pair kernel for the MN-major weight-gradient GEMMs
L.lib().stair_set_gemm_pair(int(os.environ.get('PAIR', 1)))
step, opt = NMNTrainStep(model), FusedAdam(model)
plan = step.plan(batch)
for i in range(steps):
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    out = step.run(plan)
    e1.record()
    opt.step(); opt.zero_grad()
    e2.record()
    torch.cuda.synchronize()
    print('step %d: fwd+bwd %.2f ms, adam %.2f ms, loss %.4f, launches %d' % (i, e0.elapsed_time(e1), e1.elapsed_time(e2), float(out['loss']), step.last_launches), flush=True)
