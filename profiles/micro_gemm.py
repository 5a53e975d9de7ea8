"""Micro-benchmark of stair_gemm_bf16 launch floor and small-K throughput (CUDA events, back-to-back launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    dev = 'cuda'
    for (M, N, K, odt) in [(128, 128, 64, torch.bfloat16), (128, 128, 512, torch.bfloat16), (4096, 1024, 256, torch.float32),
                           (4096, 512, 512, torch.bfloat16), (32768, 512, 512, torch.bfloat16), (65536, 512, 512, torch.bfloat16),
                           (65492, 2048, 300, torch.bfloat16), (32768, 2048, 4096, torch.bfloat16)]:
        Kp = (K + 7) // 8 * 8
        A = torch.randn(M, Kp, device=dev).bfloat16()
        W = torch.randn(N, Kp, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev, dtype=odt)
        us = timeit(lambda: L.gemm(A, W, out=out, K=K))
        x = torch.empty(1024, device=dev)
        us_alt = timeit(lambda: (L.gemm(A, W, out=out, K=K), x.add_(1.0)))
        tf = 2.0 * M * N * K / us / 1e6
        print('M=%6d N=%5d K=%5d %s: %8.1f us  %7.1f TFLOP/s   (alternating with a tiny torch kernel: %8.1f us)' % (M, N, K, str(odt)[6:], us, tf, us_alt))


if __name__ == '__main__':
    main()
