"""Where the end-to-end step goes: raw H2D time of the pinned batch, host time of forward_batch, pipelined step for several chunk counts."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate_chunks

B, T, V = 4096, 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
for nch in (1, 2, 4, 8, 16):
    chunks = collate_chunks(qs, nch, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
    nbytes = sum(c.h2d_bytes() for c in chunks)
    for _ in range(3):
        model.forward_pipelined(chunks)[0].cpu()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        for c in chunks:
            c.to('cuda')
        torch.cuda.synchronize()
    h2d = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(10):
        model.forward_pipelined(chunks)[0].cpu()
    torch.cuda.synchronize()
    e2e = (time.perf_counter() - t0) / 10
    t0 = time.perf_counter()
    for _ in range(10):
        for c in chunks:
            model.forward_batch(c)
    host = (time.perf_counter() - t0) / 10
    torch.cuda.synchronize()
    print('chunks %2d: H2D alone %.2f ms (%.1f GB/s)   pipelined e2e %.2f ms   host time of the forward_batch calls %.2f ms' % (nch, h2d * 1e3, nbytes / h2d / 1e9, e2e * 1e3, host * 1e3), flush=True)
