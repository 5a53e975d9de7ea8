"""GPU-side duration of small GEMM launches without host launch overhead: 100 launches captured in a CUDA graph."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L

dev = 'cuda'
x = torch.zeros(1024, device=dev)
for (M, N, K, odt) in [(128, 128, 64, torch.bfloat16), (128, 128, 512, torch.bfloat16), (4096, 512, 512, torch.bfloat16),
                       (4096, 1024, 256, torch.float32), (32768, 512, 512, torch.bfloat16), (65492, 2048, 304, torch.bfloat16)]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=odt)
    res = {}
    for mode in ('gemm', 'gemm+tiny', 'tiny'):
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            for _ in range(3):
                L.gemm(A, W, out=out); x.add_(1.0)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(100):
                if mode != 'tiny':
                    L.gemm(A, W, out=out)
                if mode != 'gemm':
                    x.add_(1.0)
        for _ in range(3):
            g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        res[mode] = e0.elapsed_time(e1) / 500 * 1e3
    print('M=%6d N=%5d K=%4d: gemm %.2f us | gemm+tiny kernel %.2f us | tiny kernel alone %.2f us' % (M, N, K, res['gemm'], res['gemm+tiny'], res['tiny']))
