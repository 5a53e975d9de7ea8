"""Where does a module-sized GEMM launch spend its time?  globaltimer stamps of block 0 (debug hook stair_gemm_debug_timeline) and the
in-stream cost per launch of a chain of 20 identical launches, for the two-CTAs-per-SM form (gemm_small 1) and the one-CTA-per-SM forms."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L

dev = 'cuda'
buf = torch.zeros(8, dtype=torch.int64, device=dev)
lib = L.lib()
for (M, N, K) in [(128, 128, 64), (410, 512, 1536), (3280, 512, 512), (6544, 512, 512), (16376, 512, 512), (29480, 512, 512)]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for small in (0, 1):
        lib.stair_set_gemm_small(small)
        for _ in range(3):
            L.gemm(A, W, bias=bias, out=out, act=L.ACT_RELU)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            L.gemm(A, W, bias=bias, out=out, act=L.ACT_RELU)
        e1.record(); torch.cuda.synchronize()
        chain = e0.elapsed_time(e1) * 1e3 / 20
        lib.stair_gemm_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
        L.gemm(A, W, bias=bias, out=out, act=L.ACT_RELU)
        torch.cuda.synchronize()
        lib.stair_gemm_debug_timeline(ctypes.c_void_p(0))
        t = buf.cpu().tolist()
        print('M=%5d N=%d K=%4d small=%d  %5.1f us per launch in a chain (%.0f TFLOP/s) | block 0: start->setup %.2f | ->first smem full %.2f | ->accum ready %.2f | ->epilogue done %.2f | ->all warps joined %.2f us'
              % (M, N, K, small, chain, 2.0 * M * N * K / chain / 1e6, (t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3, (t[4] - t[3]) / 1e3, (t[5] - t[4]) / 1e3), flush=True)
lib.stair_set_gemm_small(1)
