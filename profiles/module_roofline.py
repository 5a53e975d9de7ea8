"""Achieved HBM bandwidth of the memory-bound module kernels through their C-ABI single-operator entry points.

For every kernel: algorithmic bytes (each input read once + each exposed output written once) / CUDA-event time.
``measure(n)`` returns one row per kernel for ``n`` instances; bench.py calls it at the instance count of the bench step's own groups
(launch-latency bound: 17-34 MB per launch) and at a streaming size (32768 instances: inputs larger than the 126 MB L2).
Peak = MEASURED_PEAKS.json hbm_gbs (else the profiling recipe's fallback 6650 GB/s).

    python profiles/module_roofline.py            # prints the table at n = 2048 and n = 32768 (+ the raw-feature ingest)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import torch  # noqa: E402

from stair_b200 import _lib as L  # noqa: E402

L2_BYTES = 126e6


def hbm_peak():
    pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    return (json.load(open(pk))['hbm_gbs'], 'measured') if os.path.exists(pk) else (6650.0, 'fallback')


_flush = None


def timed(fn, reps=20, use_flush=True):
    """Mean CUDA-event time of ``fn`` (seconds).  Small working sets: a 256 MB write between repetitions evicts the inputs from L2.
    Working sets of >= 2x L2 are streamed without it ("inputs larger than L2"): the flush leaves L2 full of dirty lines whose write-back
    (up to 126 MB) would be charged to the timed kernel — a 40 % penalty for a read-only 270 MB pass."""
    global _flush
    if use_flush and _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if use_flush:
            _flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e-3


def measure(n, T=8, H=512, K=1, reps=20, kernels=None, dtype=torch.bfloat16):
    """-> [{'kernel', 'n', 'bytes', 'ms', 'gbs', 'frac'}] for the module row kernels at ``n`` instances (bf16 activations)."""
    lib = L.lib()
    dev = 'cuda'
    peak, _ = hbm_peak()
    st = L.stream_ptr()
    dc = L.i32(L.dtype_code(dtype))
    esz = 2 if dtype == torch.bfloat16 else 4
    rows = []

    def row(name, nbytes, fn, n_=None):
        if kernels is not None and name not in kernels:
            return
        t = timed(fn, reps=reps, use_flush=nbytes < 2 * L2_BYTES)
        rows.append({'kernel': name, 'n': n_ or n, 'bytes': int(nbytes), 'ms': t * 1e3, 'gbs': nbytes / t / 1e9, 'frac': nbytes / t / 1e9 / peak})

    f = torch.randn(n * T, H, device=dev).to(dtype)
    kw = torch.randn(n * K, H, device=dev).to(dtype)
    att = torch.empty(n * K * T, device=dev)
    row('cos_att', f.numel() * esz + kw.numel() * esz + att.numel() * 4,
        lambda: L.check(lib.stair_cos_attention(dc, L.ptr(f), L.ptr(kw), L.i32(K), L.i32(T), L.i32(H), L.ptr(att), L.i32(n), st), 'cos'))
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    out = torch.empty_like(f)
    row('layernorm', 2 * f.numel() * esz,
        lambda: L.check(lib.stair_layernorm(dc, L.ptr(f), L.ptr(g), L.ptr(b), L.ptr(out), L.i64(n * T), L.i32(H), st), 'ln'))
    agg = torch.empty(n, H, device=dev, dtype=dtype)
    row('sum_frames', f.numel() * esz + agg.numel() * esz,
        lambda: L.check(lib.stair_sum_frames(dc, L.ptr(f), L.ptr(agg), L.i32(n), L.i32(T), L.i32(H), st), 'sum'))
    del out
    vid = torch.randn(2 * n, T, H, device=dev).to(dtype)
    idx = torch.arange(n, device=dev, dtype=torch.int32)
    a1 = torch.rand(n, T, device=dev)
    row('attn_video', 2 * n * T * H * esz + n * T * 4,
        lambda: L.check(lib.stair_attn_video(dc, L.ptr(vid), L.ptr(idx), L.ptr(a1), L.ptr(idx), L.i32(n), L.i32(n), L.i32(T), L.i32(H), st), 'av'))
    vec = torch.randn(n, H, device=dev).to(dtype)
    row('exists_frame', n * T * H * esz + n * H * esz + n * T * 4,
        lambda: L.check(lib.stair_exists_frame(dc, L.ptr(vid), L.ptr(idx), L.ptr(vec), L.ptr(idx), L.ptr(att), L.i32(0), L.i32(n), L.i32(T), L.i32(H), st), 'ef'))
    del vid
    w, bb = torch.randn(H, device=dev), torch.zeros(1, device=dev)
    row('hasitem_tail', n * T * H * esz + n * T * 4,
        lambda: L.check(lib.stair_hasitem_tail(dc, L.ptr(f), L.ptr(w), L.ptr(bb), L.ptr(att), L.i32(0), L.i32(n), L.i32(T), L.i32(H), st), 'hi'))
    a2 = torch.rand(n, T, device=dev); o2 = torch.empty(n, T, device=dev); beta = torch.rand(T, device=dev)
    row('relate', 2 * n * T * 4,
        lambda: L.check(lib.stair_relate(L.ptr(a2), L.ptr(beta), L.i32(1), L.ptr(o2), L.i32(n), L.i32(T), st), 'rel'))
    row('relate_scan', 2 * n * T * 4,
        lambda: L.check(lib.stair_relate_scan(L.ptr(a2), L.i32(1), L.ptr(o2), L.i32(n), L.i32(T), st), 'scan'))
    # l2normalize reads [n, H] rows once and writes fp32: 3 KB per row, so the streaming point needs 8x the rows of the [T, H] kernels to be
    # larger than L2 (at n = 32768 the pass is 100 MB: it fits, gets flushed, and the number measures the flush, not the kernel)
    n2 = n * T
    vec2 = torch.randn(n2, H, device=dev).to(dtype)
    nrm = torch.empty(n2, H, device=dev)
    row('l2normalize', n2 * H * esz + n2 * H * 4,
        lambda: L.check(lib.stair_l2normalize(dc, L.ptr(vec2), L.ptr(nrm), L.i32(n2), L.i32(H), st), 'l2'), n_=n2)
    del vec2, nrm
    lg = torch.randn(n, 172, device=dev); am = torch.empty(n, device=dev, dtype=torch.int32)
    row('argmax', n * 172 * 4 + n * 4,
        lambda: L.check(lib.stair_argmax(L.ptr(lg), L.ptr(am), L.i32(n), L.i32(172), st), 'am'))
    return rows


def main():
    lib = L.lib()
    lib.stair_set_row_stream(int(os.environ.get('ROW_STREAM', 0)))     # HasItem tail: 0 = register-staged kernel (product), 1 = TMA-staged streaming kernel
    lib.stair_set_cos_impl(int(os.environ.get('COS_IMPL', 0)))       # 0 = instance-major cosine maps (product), 1 = row-major
    peak, src = hbm_peak()
    for n in (2048, 32768):
        for r in measure(n):
            print('%-34s n=%7d  %8.2f MB  %8.1f us  %7.1f GB/s  %5.1f%% of %.0f (%s)' % (r['kernel'], r['n'], r['bytes'] / 1e6, r['ms'] * 1e3, r['gbs'],
                                                                                          100 * r['frac'], peak, src), flush=True)
    # ---- raw-feature ingest (csrc/ingest.cu): appearance [B,8,16,2048] mean-pooled + motion [B,8,2048] -> [B,8,4096] bf16 ------------------
    from stair_b200 import ingest
    dev = 'cuda'
    for B, dt in ((256, torch.float32), (1024, torch.float32), (1024, torch.bfloat16)):
        app = torch.rand(B, 8, 16, 2048, device=dev, dtype=torch.float32).to(dt)
        mot = torch.rand(B, 8, 2048, device=dev, dtype=torch.float32).to(dt)
        out = torch.empty(B, 8, 4096, device=dev, dtype=torch.bfloat16)
        nb = app.numel() * app.element_size() + mot.numel() * mot.element_size() + out.numel() * 2
        t = timed(lambda: ingest.pool_concat(app, mot, out=out), use_flush=nb < 2 * L2_BYTES)
        print('%-34s n=%7d  %8.2f MB  %8.1f us  %7.1f GB/s  %5.1f%% of %.0f' % ('ingest pool+concat (%s in)' % ('fp32' if dt == torch.float32 else 'bf16'),
                                                                                B, nb / 1e6, t * 1e6, nb / t / 1e9, 100 * nb / t / 1e9 / peak, peak), flush=True)
        del app, mot, out


if __name__ == '__main__':
    main()
