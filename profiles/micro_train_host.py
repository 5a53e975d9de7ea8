"""Host-side profile (cProfile) of the 32-question training window (the reference's accumulation window): where do the ~3 ms per step go
on the host?  argv: [B]"""
import sys, os, cProfile, pstats, io, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
from stair_b200.train import NMNTrainStep, FusedAdam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
T, V = 8, 4096
cfg = dict(syn.model_config(T=T, V=V), dropout=0.25)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().train()
tmpl = list(syn.TEMPLATES) + ['and_equals_xor', 'compare_xor_equals']
qs = syn.make_questions(max(B, 32), T, V, seed=4321, with_gold=True, templates=tmpl)
step = NMNTrainStep(model)
opt = FusedAdam(model, lr=2e-4)
plan = step.plan(qs[:B])


def one():
    step.run(plan); opt.step(); opt.zero_grad()


for _ in range(5):
    one()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(50):
    one()
t_host = (time.perf_counter() - t0) / 50
torch.cuda.synchronize()
t_all = (time.perf_counter() - t0) / 50
print('B=%d: host enqueue %.3f ms per step, with the final synchronize %.3f ms per step' % (B, 1e3 * t_host, 1e3 * t_all))
pr = cProfile.Profile()
pr.enable()
for _ in range(50):
    one()
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats('cumulative').print_stats(28)
print(s.getvalue()[:6000])
