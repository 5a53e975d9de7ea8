"""A few I3D-configuration forwards (T = 64, V = 1024, layouts of >= 12 modules) for ncu launch lists."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate
B, T, V = 4096, 64, 1024
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=777, templates=list(syn.LONG_TEMPLATES))
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
for _ in range(3):
    model.forward_batch(batch)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    model.forward_batch(batch)
e1.record(); torch.cuda.synchronize()
print('I3D forward %.3f ms' % (e0.elapsed_time(e1) / 3))
