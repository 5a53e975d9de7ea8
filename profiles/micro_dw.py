"""The weight-gradient contraction in isolation: dW[512,512] += dZ^T . X over 19648 rows (one Localize Linear of the B=4096 window),
MN-major operands in place (stair_gemm_bf16_tn) vs the round-1 path (two bf16 transposes + K-major GEMM)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L

lib = L.lib()
M, N, K = 512, 512, 19648
dz = torch.randn(K, M, device='cuda').bfloat16()
x = torch.randn(K, N, device='cuda').bfloat16()
C = torch.zeros(M, N, device='cuda')
ll = ctypes.c_longlong


def tn():
    L.check(lib.stair_gemm_bf16_tn(L.ptr(dz), ll(M), L.i32(K), L.ptr(x), ll(N), L.i32(K), L.i32(1), L.ptr(C), ll(N), L.i32(M), L.i32(N), L.i32(K),
                                   L.i32(1), L.stream_ptr()), 'tn')


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for wide in (0, 1):
    lib.stair_set_gemm_dw_wide(wide)
    print('128 x %d tiles: %.1f us' % (256 if wide else 128, timed(tn)), flush=True)
t = timed(tn)
ref = dz.float().t() @ x.float()
C.zero_(); tn(); torch.cuda.synchronize()
err = float((C - ref).abs().max()) / float(ref.abs().max())
print('dW = dZ^T.X  [%d,%d] over %d rows, MN-major in place: %.1f us  (%.0f TFLOP/s, %.0f GB/s of operand reads)  rel err %.2e'
      % (M, N, K, t, 2.0 * M * N * K / t / 1e6, (dz.numel() + x.numel()) * 2 / t / 1e3, err), flush=True)
