import sys, os
sys.path.insert(0, '/root/repo')
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate
B, T, V = 4096, 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
for _ in range(4):
    model.forward_batch(batch)
torch.cuda.synchronize()
