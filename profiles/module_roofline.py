"""Achieved HBM bandwidth of the memory-bound module kernels through their C-ABI single-operator entry points.

For every kernel: algorithmic bytes (each input read once + each exposed output written once) / CUDA-event time,
at the instance count of one bench group (B=4096 mix: ~410-2900 instances) and at a streaming size (32768 instances, inputs
larger than L2).  Peak = MEASURED_PEAKS.json hbm_gbs (else the profiling recipe's fallback 6650 GB/s).
"""
import ctypes, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from stair_b200 import _lib as L

T, H, K = 8, 512, 1
pk = os.path.join(ROOT, 'MEASURED_PEAKS.json')
PEAK = json.load(open(pk))['hbm_gbs'] if os.path.exists(pk) else 6650.0
lib = L.lib()
lib.stair_set_row_stream(int(os.environ.get('ROW_STREAM', 0)))     # HasItem tail: 0 = register-staged kernel (product), 1 = TMA-staged streaming kernel
lib.stair_set_cos_impl(int(os.environ.get('COS_IMPL', 0)))       # 0 = instance-major cosine maps (product), 1 = row-major
dev = 'cuda'
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, reps=20, use_flush=True):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        # small working sets: flush so that inputs do not stay in the 126 MB L2 between repetitions.  Working sets of >= 2x L2 are
        # streamed without a flush ("inputs larger than L2"): the write-flush leaves L2 full of dirty lines whose write-back (up to
        # 126 MB) is charged to the timed kernel — a 40 % penalty for a read-only 270 MB pass.
        if use_flush:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e-3


def row(name, n, nbytes, fn):
    t = timed(fn, use_flush=nbytes < 2 * 126e6)
    print('%-34s n=%6d  %8.2f MB  %8.1f us  %7.1f GB/s  %5.1f%% of %.0f' % (name, n, nbytes / 1e6, t * 1e6, nbytes / t / 1e9, 100 * nbytes / t / 1e9 / PEAK, PEAK), flush=True)


for n in (2048, 32768):
    st = L.stream_ptr()
    f = torch.randn(n * T, H, device=dev).to(torch.bfloat16)
    kw = torch.randn(n * K, H, device=dev).to(torch.bfloat16)
    att = torch.empty(n * K * T, device=dev)
    row('cos_att (Localize map)', n, f.numel() * 2 + kw.numel() * 2 + att.numel() * 4,
        lambda: L.check(lib.stair_cos_attention(L.i32(0), L.ptr(f), L.ptr(kw), L.i32(K), L.i32(T), L.i32(H), L.ptr(att), L.i32(n), st), 'cos'))
    g, b = torch.ones(H, device=dev), torch.zeros(H, device=dev)
    out = torch.empty_like(f)
    row('layernorm (Temporal)', n, 2 * f.numel() * 2,
        lambda: L.check(lib.stair_layernorm(L.i32(0), L.ptr(f), L.ptr(g), L.ptr(b), L.ptr(out), L.i64(n * T), L.i32(H), st), 'ln'))
    agg = torch.empty(n, H, device=dev, dtype=torch.bfloat16)
    row('sum_frames (Filter)', n, f.numel() * 2 + agg.numel() * 2,
        lambda: L.check(lib.stair_sum_frames(L.i32(0), L.ptr(f), L.ptr(agg), L.i32(n), L.i32(T), L.i32(H), st), 'sum'))
    vid = torch.randn(2 * n, T, H, device=dev).to(torch.bfloat16)
    idx = torch.arange(n, device=dev, dtype=torch.int32)
    a1 = torch.rand(n, T, device=dev)
    row('attn_video', n, 2 * n * T * H * 2 + n * T * 4,
        lambda: L.check(lib.stair_attn_video(L.i32(0), L.ptr(vid), L.ptr(idx), L.ptr(a1), L.ptr(idx), L.i32(n), L.i32(n), L.i32(T), L.i32(H), st), 'av'))
    vec = torch.randn(n, H, device=dev).to(torch.bfloat16)
    row('exists_frame', n, n * T * H * 2 + n * H * 2 + n * T * 4,
        lambda: L.check(lib.stair_exists_frame(L.i32(0), L.ptr(vid), L.ptr(idx), L.ptr(vec), L.ptr(idx), L.ptr(att), L.i32(0), L.i32(n), L.i32(T), L.i32(H), st), 'ef'))
    w, bb = torch.randn(H, device=dev), torch.zeros(1, device=dev)
    row('hasitem_tail', n, n * T * H * 2 + n * T * 4,
        lambda: L.check(lib.stair_hasitem_tail(L.i32(0), L.ptr(f), L.ptr(w), L.ptr(bb), L.ptr(att), L.i32(0), L.i32(n), L.i32(T), L.i32(H), st), 'hi'))
    a2 = torch.rand(n, T, device=dev); o2 = torch.empty(n, T, device=dev); beta = torch.rand(T, device=dev)
    row('relate (softmax_T)', n, 2 * n * T * 4,
        lambda: L.check(lib.stair_relate(L.ptr(a2), L.ptr(beta), L.i32(1), L.ptr(o2), L.i32(n), L.i32(T), st), 'rel'))
    row('relate_scan before', n, 2 * n * T * 4,
        lambda: L.check(lib.stair_relate_scan(L.ptr(a2), L.i32(1), L.ptr(o2), L.i32(n), L.i32(T), st), 'scan'))
    nrm = torch.empty(n, H, device=dev)
    row('l2normalize (heads)', n, n * H * 2 + n * H * 4,
        lambda: L.check(lib.stair_l2normalize(L.i32(0), L.ptr(vec), L.ptr(nrm), L.i32(n), L.i32(H), st), 'l2'))
    lg = torch.randn(n, 172, device=dev); am = torch.empty(n, device=dev, dtype=torch.int32)
    row('argmax (answers)', n, n * 172 * 4 + n * 4,
        lambda: L.check(lib.stair_argmax(L.ptr(lg), L.ptr(am), L.i32(n), L.i32(172), st), 'am'))

# ---- raw-feature ingest (csrc/ingest.cu): appearance [B,8,16,2048] mean-pooled + motion [B,8,2048] -> [B,8,4096] bf16 ------------------
from stair_b200 import ingest
for B, dt in ((256, torch.float32), (1024, torch.float32), (1024, torch.bfloat16)):
    app = torch.rand(B, 8, 16, 2048, device=dev, dtype=torch.float32).to(dt)
    mot = torch.rand(B, 8, 2048, device=dev, dtype=torch.float32).to(dt)
    out = torch.empty(B, 8, 4096, device=dev, dtype=torch.bfloat16)
    nb = app.numel() * app.element_size() + mot.numel() * mot.element_size() + out.numel() * 2
    row('ingest pool+concat (%s in)' % ('fp32' if dt == torch.float32 else 'bf16'), B, nb, lambda: ingest.pool_concat(app, mot, out=out))
    del app, mot, out
