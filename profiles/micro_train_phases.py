"""Where does the training step go?  CUDA events around the C entry points of one 4096-question window (bench's training workload):
forward with history, backward of losses + decoder + modules, backward of the encoders (BPTT + weight gradients), Adam.
argv: [B]"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
from stair_b200.train import NMNTrainStep, FusedAdam

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T, V = 8, 4096
cfg = dict(syn.model_config(T=T, V=V), dropout=0.25)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().train()
tmpl = list(syn.TEMPLATES) + ['and_equals_xor', 'compare_xor_equals']
qs = syn.make_questions(B, T, V, seed=4321, with_gold=True, templates=tmpl)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
step = NMNTrainStep(model, overlap_allreduce=True)
step.split_backward = True
opt = FusedAdam(model, lr=2e-4)
plan = step.plan(batch)

real = L.lib()
if 'BWD_LANES' in os.environ:
    real.stair_set_bwd_lanes(int(os.environ['BWD_LANES']))
if 'DEP' in os.environ:
    real.stair_set_dep_sched(int(os.environ['DEP']))
marks = []


class Proxy:
    def __getattr__(self, name):
        fn = getattr(real, name)
        if name in ('stair_nmn_forward_train', 'stair_nmn_backward_phases', 'stair_adam_multi'):
            def wrapped(*a):
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(); rc = fn(*a); e1.record()
                marks.append((name + (str(a[4].value) if name == 'stair_nmn_backward_phases' else ''), e0, e1, int(real.stair_last_launch_count())))
                return rc
            return wrapped
        return fn


def one():
    step.run(plan); opt.step(); opt.zero_grad()


for _ in range(3):
    one()
torch.cuda.synchronize()
L._lib = Proxy()
tot0, tot1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
N = 5
tot0.record()
for _ in range(N):
    one()
tot1.record(); torch.cuda.synchronize()
L._lib = real
agg = {}
for name, e0, e1, n in marks:
    a = agg.setdefault(name, [0.0, 0, 0]); a[0] += e0.elapsed_time(e1); a[1] += 1; a[2] = n
print('training step, B=%d, bwd lanes %s, dependency scheduling %s: %.3f ms per step (events around the whole step)' % (B, os.environ.get('BWD_LANES', 'default'), os.environ.get('DEP', 'default'), tot0.elapsed_time(tot1) / N))
for name, (ms, k, n) in agg.items():
    print('  %-32s %.3f ms per step  (%d launches)' % (name, ms / N, n))
