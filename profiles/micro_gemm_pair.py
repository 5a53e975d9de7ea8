"""A/B of the CTA-pair (cta_group::2) GEMM against the single-CTA kernel on the path's big contractions (CUDA events, back-to-back
launches on an otherwise idle GPU; the video projection is the roofline kernel of bench.py)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import _lib as L
from micro_gemm import timeit


def main():
    dev = 'cuda'
    lib = L.lib()
    shapes = [(32768, 2048, 4096, 'video projection (RX, 4096 questions)'), (65492, 2048, 300, 'text projection'),
              (262144, 2048, 1024, 'video projection (I3D)'), (32768, 512, 512, 'module Linear, 4096 x 8 frame rows'),
              (19648, 512, 512, 'module Linear, 2456 x 8'), (4096, 1024, 1024, 'decoder.0'), (8192, 8192, 8192, 'square 8192')]
    for (M, N, K, what) in shapes:
        Kp = (K + 7) // 8 * 8
        A = torch.randn(M, Kp, device=dev).bfloat16()
        W = (torch.randn(N, Kp, device=dev) * K ** -0.5).bfloat16()
        bias = torch.randn(N, device=dev)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        res = {}
        for mode in (0, 2):
            lib.stair_set_gemm_pair(mode)
            res[mode] = timeit(lambda: L.gemm(A, W, bias=bias, out=out, K=K), n=30)
            res[(mode, 'out')] = out.clone()
        lib.stair_set_gemm_pair(1)
        same = torch.equal(res[(0, 'out')], res[(2, 'out')])
        tf = lambda us: 2.0 * M * N * K / us / 1e6      # noqa: E731
        print('%-42s M=%6d N=%5d K=%5d  single-CTA %8.1f us %7.1f TF/s | CTA pair %8.1f us %7.1f TF/s  (x%.3f, bit-identical %s)'
              % (what, M, N, K, res[0], tf(res[0]), res[2], tf(res[2]), res[0] / res[2], same))


if __name__ == '__main__':
    main()
