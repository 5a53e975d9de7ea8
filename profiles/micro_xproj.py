"""The dominant kernel alone: video input projection [B*T, V] x [V, 4*H] (+bias), bf16 in/out, for `ncu --set full` captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L
M, N, K = 32768, 2048, 4096
A = torch.randn(M, K, device='cuda').bfloat16()
W = (torch.randn(N, K, device='cuda') * 0.02).bfloat16()
b = torch.zeros(N, device='cuda')
out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.gemm(A, W, bias=b, out=out); e1.record(); torch.cuda.synchronize()
    print('launch %d: %.1f us  %.1f TFLOP/s' % (i, e0.elapsed_time(e1) * 1e3, 2.0 * M * N * K / (e0.elapsed_time(e1) * 1e-3) / 1e12), flush=True)
