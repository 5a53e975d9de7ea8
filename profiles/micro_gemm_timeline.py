"""Where does a tiny GEMM launch spend its time?  globaltimer stamps of block 0 (debug hook stair_gemm_debug_timeline)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L

dev = 'cuda'
buf = torch.zeros(8, dtype=torch.int64, device=dev)
for (M, N, K) in [(128, 128, 64), (128, 128, 512), (4096, 512, 512), (4096, 1024, 256)]:
    A = torch.randn(M, K, device=dev).bfloat16(); W = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3):
        L.gemm(A, W, out=out)
    torch.cuda.synchronize()
    L.lib().stair_gemm_debug_timeline(ctypes.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); L.gemm(A, W, out=out); e1.record()
    torch.cuda.synchronize()
    L.lib().stair_gemm_debug_timeline(ctypes.c_void_p(0))
    t = buf.cpu().tolist()
    print('M=%d N=%d K=%d event %.1f us | start->setup %.2f | ->first smem full %.2f | ->accum ready %.2f | ->epilogue done %.2f | ->all warps joined %.2f us'
          % (M, N, K, e0.elapsed_time(e1) * 1e3, (t[1] - t[0]) / 1e3, (t[2] - t[1]) / 1e3, (t[3] - t[2]) / 1e3, (t[4] - t[3]) / 1e3, (t[5] - t[4]) / 1e3))
