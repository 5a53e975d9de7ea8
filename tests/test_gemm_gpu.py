"""GPU: tcgen05/TMA GEMM (stair_gemm_bf16) against a torch fp32 reference of the same contraction."""
import pytest
import torch

from stair_b200 import _lib as L

pytestmark = pytest.mark.gpu


def _ref(A, W, bias, row_scale, act):
    y = A.float() @ W.float().t()
    if row_scale is not None:
        y = y * row_scale[:, None]
    if bias is not None:
        y = y + bias
    return torch.relu(y) if act else y


CASES = [
    # M, N, K, out dtype, bias, relu, row_scale
    (128, 128, 64, torch.float32, False, False, False),
    (128, 128, 512, torch.float32, True, False, False),
    (256, 512, 512, torch.bfloat16, True, True, False),
    (1000, 512, 512, torch.bfloat16, True, True, True),
    (4096, 2048, 4096, torch.bfloat16, True, False, False),
    (300, 172, 1024, torch.float32, True, False, False),     # decoder head: ragged N
    (333, 64, 192, torch.float32, True, True, False),        # small-config shapes (H=64)
    (77, 2048, 300, torch.bfloat16, True, False, False),     # text projection: K=300 (row pitch 304)
    (5000, 1024, 256, torch.float32, False, False, False),   # LSTM recurrent projection
    (20000, 512, 1536, torch.bfloat16, True, True, False),
]


@pytest.mark.parametrize('impl', [0, 1], ids=['tc', 'simt'])
@pytest.mark.parametrize('case', CASES)
def test_gemm_matches_fp32_reference(case, impl):
    M, N, K, odt, use_bias, relu, use_rs = case
    if impl == 1 and M * N * K > 3e9:
        pytest.skip('SIMT debug kernel: small cases only')
    g = torch.Generator(device='cuda').manual_seed(M * 7 + N * 3 + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, device='cuda', dtype=torch.bfloat16)
    W = torch.zeros(N, Kp, device='cuda', dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W[:, :K] = (torch.randn(N, K, device='cuda', generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g) if use_bias else None
    rs = torch.rand(M, device='cuda', generator=g) if use_rs else None
    L.lib().stair_set_gemm_impl(impl)
    try:
        out = L.gemm(A, W, bias=bias, out_dtype=odt, act=L.ACT_RELU if relu else L.ACT_NONE, row_scale=rs, K=K)
        torch.cuda.synchronize()
    finally:
        L.lib().stair_set_gemm_impl(0)
    ref = _ref(A[:, :K], W[:, :K], bias, rs, relu)
    tol = 2e-2 if odt == torch.bfloat16 else 2e-4
    err = (out.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    assert err <= tol * max(scale, 1.0), 'max err %g (scale %g)' % (err, scale)


def test_gemm_accumulate_and_strided_output():
    M, N, K = 512, 256, 128
    A = torch.randn(M, K, device='cuda').bfloat16()
    W = torch.randn(N, K, device='cuda').bfloat16()
    big = torch.ones(M, 2 * N, device='cuda')
    L.gemm(A, W, out=big[:, N:], accumulate=True, ldc=2 * N)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + 1
    assert torch.equal(big[:, :N], torch.ones(M, N, device='cuda'))
    assert (big[:, N:] - ref).abs().max().item() < 1e-3


def test_gemm_split3_is_fp32_grade():
    """bf16x3 planes: six plane products accumulate to an fp32-accurate contraction."""
    M, N, K = 700, 512, 512
    g = torch.Generator(device='cuda').manual_seed(5)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) * K ** -0.5

    def split(x):
        p0 = x.bfloat16(); r = x - p0.float()
        p1 = r.bfloat16(); r = r - p1.float()
        return torch.cat([p0, p1, r.bfloat16()], dim=0).contiguous()
    for impl in (1, 0):
        L.lib().stair_set_gemm_impl(impl)
        try:
            out = L.gemm(split(A), split(W), out_dtype=torch.float32, M=M, N=N, K=K, nplanes=3, a_plane_rows=M, w_plane_rows=N)
            torch.cuda.synchronize()
        finally:
            L.lib().stair_set_gemm_impl(0)
        ref = (A.double() @ W.double().t()).float()
        err = (out - ref).abs().max().item()
        assert err < 2e-5 * ref.abs().max().item() + 1e-6, 'impl %d err %g' % (impl, err)


def test_gemm_small_k():
    """K < one 64-wide TMA box (LSTM recurrent projection of the H=64 fixtures: K = 32)."""
    M, N, K = 300, 128, 32
    g = torch.Generator(device='cuda').manual_seed(3)
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W = torch.randn(N, K, device='cuda', generator=g).bfloat16()
    out = L.gemm(A, W, out_dtype=torch.float32)
    torch.cuda.synchronize()
    assert (out - A.float() @ W.float().t()).abs().max().item() < 1e-3


@pytest.mark.parametrize('T', [8, 16, 64, 128])
@pytest.mark.parametrize('H', [64, 512])
def test_gemm_slot_gather(T, H):
    """A rows gathered from a [slots, T, H] arena by TMA 3-D boxes == gather-then-GEMM."""
    slots, n, N = 97, 211, 192
    g = torch.Generator(device='cuda').manual_seed(T + H)
    arena = torch.randn(slots, T, H, device='cuda', generator=g).bfloat16()
    idx = torch.randint(0, slots, (n,), device='cuda', generator=g, dtype=torch.int32)
    W = (torch.randn(N, H, device='cuda', generator=g) * H ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    rs = torch.rand(n * T, device='cuda', generator=g)
    out = L.gemm_gather(arena, idx, T, W, bias=bias, out_dtype=torch.float32, act=L.ACT_RELU, row_scale=rs)
    torch.cuda.synchronize()
    ref = torch.relu((arena[idx.long()].reshape(n * T, H).float() @ W.float().t()) * rs[:, None] + bias)
    assert (out - ref).abs().max().item() < 2e-3


TN_CASES = [
    # M (out features), N (in features), K (rows), planes, accumulate
    (128, 128, 64, 1, False),
    (512, 512, 19648, 1, True),        # a module Linear's dW over 2456 x 8 frame rows (split-K)
    (512, 1536, 4096, 1, True),
    (172, 1024, 4096, 1, False),       # decoder head: M not a multiple of 64
    (1024, 304, 777, 1, True),         # text W_ih: N = padded 300, K not a multiple of 64
    (64, 192, 333, 3, True),           # strict fp32: three bf16 planes, plane stride > K
    (2048, 4096, 8192, 1, False),
]


@pytest.mark.parametrize('impl', [0, 1, 2], ids=['tc', 'simt', 'tc-pair'])
@pytest.mark.parametrize('case', TN_CASES + [(256, 256, 64, 1, True), (700, 768, 5000, 1, True), (130, 512, 300, 3, False)])
def test_gemm_tn_matches_fp32_reference(case, impl):
    """C = A^T . W with both operands read in place as MN-major tiles (the weight-gradient contraction, no transposed copies).
    'tc' = the product dispatch (CTA-pair kernel with split-K where eligible, single-CTA kernel otherwise), 'tc-pair' = the CTA-pair
    kernel forced wherever it is legal (N % 256 == 0; incl. ragged M whose peer CTA rows are partly or wholly past M), 'simt' = debug."""
    import ctypes
    M, N, K, planes, acc = case
    if impl == 1 and M * N * K > 3e9:
        pytest.skip('SIMT debug kernel: small cases only')
    pair_mode = 2 if impl == 2 else 1
    impl = 0 if impl == 2 else impl
    g = torch.Generator(device='cuda').manual_seed(M + 3 * N + 7 * K)
    lda, ldw = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    prow = K + 5                                        # plane stride in rows (> K: rows past K belong to nobody)
    a32 = torch.randn(K, M, device='cuda', generator=g) * K ** -0.5
    w32 = torch.randn(K, N, device='cuda', generator=g)

    def pack(x, ld):
        buf = torch.full((planes, prow, ld), 7.0, device='cuda', dtype=torch.bfloat16)      # poison outside the valid region
        r = x.clone()
        for p in range(planes):
            h = r.bfloat16()
            buf[p, :K, :x.shape[1]] = h
            buf[p, :K, x.shape[1]:] = 0
            r = r - h.float()
        return buf

    A, W = pack(a32, lda), pack(w32, ldw)
    if planes == 1:
        ref = A[0, :K, :M].float().t() @ W[0, :K, :N].float()
    else:
        ref = a32.double().t() @ w32.double()
    C0 = torch.randn(M, N, device='cuda', generator=g)
    C = C0.clone()
    lib = L.lib()
    lib.stair_set_gemm_impl(impl)
    lib.stair_set_gemm_pair(pair_mode)
    try:
        rc = lib.stair_gemm_bf16_tn(L.ptr(A), ctypes.c_longlong(lda), L.i32(prow), L.ptr(W), ctypes.c_longlong(ldw), L.i32(prow), L.i32(planes),
                                    L.ptr(C), ctypes.c_longlong(N), L.i32(M), L.i32(N), L.i32(K), L.i32(1 if acc else 0), L.stream_ptr())
        torch.cuda.synchronize()
    finally:
        lib.stair_set_gemm_impl(0)
        lib.stair_set_gemm_pair(1)
    assert rc == 0 and lib.stair_gemm_error_flag() == 0
    want = (ref.float() + C0) if acc else ref.float()
    err = (C - want).abs().max().item()
    assert err <= 2e-4 * max(want.abs().max().item(), 1.0), 'max err %g' % err


@pytest.mark.parametrize('shape', [(1000, 512, 512), (333, 256, 300), (4096, 2048, 320), (130, 384, 64), (2048, 512, 1024)])
@pytest.mark.parametrize('odt', [torch.bfloat16, torch.float32])
def test_two_epilogue_warp_sets_are_bit_identical(shape, odt):
    """stair_set_gemm_epi2(1) (product for K <= 512): a second set of four epilogue warps takes the upper half of every tile's columns.
    Same accumulators, same epilogue arithmetic -> bit-identical to the single-set kernel, including ragged M / N and TMA-clipped boxes;
    K = 1024 stays on the single-set kernel either way."""
    M, N, K = shape
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, device='cuda', dtype=torch.bfloat16)
    W = torch.zeros(N, Kp, device='cuda', dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W[:, :K] = (torch.randn(N, K, device='cuda', generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    rs = torch.rand(M, device='cuda', generator=g)
    outs = []
    try:
        for on in (0, 1):
            L.lib().stair_set_gemm_epi2(on)
            outs.append(L.gemm(A, W, bias=bias, out_dtype=odt, act=L.ACT_RELU, row_scale=rs, K=K).clone())
            torch.cuda.synchronize()
    finally:
        L.lib().stair_set_gemm_epi2(1)
    assert torch.equal(outs[0], outs[1])
    ref = _ref(A[:, :K], W[:, :K], bias, rs, True)
    assert (outs[1].float() - ref).abs().max().item() <= (2e-2 if odt == torch.bfloat16 else 2e-4) * max(ref.abs().max().item(), 1.0)


PAIR_SHAPES = [(256, 256, 64), (1000, 512, 512), (333, 256, 300), (4096, 2048, 320), (130, 768, 64), (2048, 512, 1024), (4096, 2048, 4096),
               (257, 256, 128)]


@pytest.mark.parametrize('shape', PAIR_SHAPES)
@pytest.mark.parametrize('odt', [torch.bfloat16, torch.float32])
def test_cta_pair_kernel_is_bit_identical_to_the_single_cta_kernel(shape, odt):
    """stair_set_gemm_pair: tcgen05 cta_group::2 — a 2-CTA cluster per 256 x 256 tile, M = 256 MMAs issued by the leader, the B tile split
    across the pair.  Every accumulator element is the same k-ordered fp32 sum as in the single-CTA kernel and the epilogue code is
    shared, so the outputs must be bit-identical (ragged M incl. a peer CTA whose rows are all past M, short and long K, bias / ReLU /
    row scale, both output types); and both match the fp32 reference."""
    M, N, K = shape
    g = torch.Generator(device='cuda').manual_seed(M + N + K)
    Kp = (K + 7) // 8 * 8
    A = torch.zeros(M, Kp, device='cuda', dtype=torch.bfloat16)
    W = torch.zeros(N, Kp, device='cuda', dtype=torch.bfloat16)
    A[:, :K] = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W[:, :K] = (torch.randn(N, K, device='cuda', generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    rs = torch.rand(M, device='cuda', generator=g)
    lib = L.lib()
    outs = []
    try:
        for mode in (0, 2):
            lib.stair_set_gemm_pair(mode)
            for _ in range(2):                                   # twice: barrier phases / TMEM alloc-dealloc across launches
                out = L.gemm(A, W, bias=bias, out_dtype=odt, act=L.ACT_RELU, row_scale=rs, K=K)
            torch.cuda.synchronize()
            outs.append(out.clone())
    finally:
        lib.stair_set_gemm_pair(1)
    assert lib.stair_gemm_error_flag() == 0
    assert torch.equal(outs[0], outs[1])
    ref = _ref(A[:, :K], W[:, :K], bias, rs, True)
    assert (outs[1].float() - ref).abs().max().item() <= (2e-2 if odt == torch.bfloat16 else 2e-4) * max(ref.abs().max().item(), 1.0)


def test_cta_pair_kernel_split3_planes():
    """The strict fp32 mode (three bf16 planes, six plane products into one accumulator) through the CTA-pair kernel."""
    M, N, K = 700, 512, 512
    g = torch.Generator(device='cuda').manual_seed(5)
    A = torch.randn(M, K, device='cuda', generator=g)
    W = torch.randn(N, K, device='cuda', generator=g) * K ** -0.5

    def split(x):
        p0 = x.bfloat16(); r = x - p0.float()
        p1 = r.bfloat16(); r = r - p1.float()
        return torch.cat([p0, p1, r.bfloat16()], dim=0).contiguous()
    lib = L.lib()
    outs = []
    try:
        for mode in (0, 2):
            lib.stair_set_gemm_pair(mode)
            outs.append(L.gemm(split(A), split(W), out_dtype=torch.float32, M=M, N=N, K=K, nplanes=3, a_plane_rows=M, w_plane_rows=N).clone())
            torch.cuda.synchronize()
    finally:
        lib.stair_set_gemm_pair(1)
    assert torch.equal(outs[0], outs[1])
    ref = (A.double() @ W.double().t()).float()
    assert (outs[1] - ref).abs().max().item() < 2e-5 * ref.abs().max().item() + 1e-6


@pytest.mark.parametrize('T,n,N', [(8, 211, 512), (8, 16, 256), (8, 17, 256), (8, 4100, 512), (64, 37, 512), (16, 1000, 256), (128, 9, 256)])
def test_cta_pair_kernel_gathers_frame_slots(T, n, N):
    """Gathered A operand (rows of frame-arena slots, one 3-D TMA box per slot issued by each CTA of the pair for its own 128 rows) through
    the CTA-pair kernel: bit-identical to the single-CTA gather kernel and equal to gather-then-GEMM, incl. a last M tile whose second
    CTA holds no slot at all (n = 16, T = 8: 128 rows) or a single one (n = 17)."""
    slots, H = 97, 512
    g = torch.Generator(device='cuda').manual_seed(T * 1000 + n)
    arena = torch.randn(slots, T, H, device='cuda', generator=g).bfloat16()
    idx = torch.randint(0, slots, (n,), device='cuda', generator=g, dtype=torch.int32)
    W = (torch.randn(N, H, device='cuda', generator=g) * H ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    rs = torch.rand(n * T, device='cuda', generator=g)
    lib = L.lib()
    outs = []
    try:
        for mode in (0, 2):
            lib.stair_set_gemm_pair(mode)
            for _ in range(2):
                out = L.gemm_gather(arena, idx, T, W, bias=bias, out_dtype=torch.bfloat16, act=L.ACT_RELU, row_scale=rs)
            torch.cuda.synchronize()
            outs.append(out.clone())
    finally:
        lib.stair_set_gemm_pair(1)
    assert lib.stair_gemm_error_flag() == 0
    assert torch.equal(outs[0], outs[1])
    ref = torch.relu((arena[idx.long()].reshape(n * T, H).float() @ W.float().t()) * rs[:, None] + bias)
    assert (outs[1].float() - ref).abs().max().item() <= 2e-2 * max(ref.abs().max().item(), 1.0)


@pytest.mark.parametrize('T,n,N,K', [(8, 409, 512, 512), (8, 16, 256, 512), (8, 17, 512, 64), (16, 100, 512, 512), (32, 33, 256, 512),
                                      (64, 37, 512, 512), (64, 2, 512, 512), (128, 9, 256, 512), (8, 3000, 512, 512), (4, 77, 192, 320)])
@pytest.mark.parametrize('pair', [0, 2])
def test_frame_sum_epilogue(T, n, N, K, pair):
    """GemmArgs::sum_out — Filter's sum over the T frames of an instance (modules.py:374-376) inside the GEMM epilogue: equals the two-kernel
    path (GEMM with bf16 output, then a frame-ordered fp32 sum rounded to bf16) bit for bit when an instance lies inside one lane quarter
    (T <= 32) and to one bf16 ulp when its 32-row partial sums are combined (T = 64, 128); single-CTA and CTA-pair kernels."""
    g = torch.Generator(device='cuda').manual_seed(T * 7 + n)
    M = n * T
    A = torch.randn(M, K, device='cuda', generator=g).bfloat16()
    W = (torch.randn(N, K, device='cuda', generator=g) * K ** -0.5).bfloat16()
    bias = torch.randn(N, device='cuda', generator=g)
    lib = L.lib()
    out = torch.full((n, N), float('nan'), device='cuda', dtype=torch.bfloat16)
    try:
        lib.stair_set_gemm_pair(pair)
        for _ in range(2):
            L.check(lib.stair_gemm_bf16_framesum(L.ptr(A), L.i64(K), L.ptr(W), L.i64(K), L.ptr(bias), L.ptr(out), L.i64(N), L.i32(M), L.i32(N), L.i32(K),
                                                 L.i32(L.ACT_RELU), L.i32(T), L.stream_ptr()), 'stair_gemm_bf16_framesum')
        full = L.gemm(A, W, bias=bias, out_dtype=torch.bfloat16, act=L.ACT_RELU)
        torch.cuda.synchronize()
    finally:
        lib.stair_set_gemm_pair(1)
    assert lib.stair_gemm_error_flag() == 0
    x = full.float().view(n, T, N)
    ref = torch.zeros(n, N, device='cuda')
    for t in range(T):                                               # frame order, fp32
        ref = ref + x[:, t]
    ref = ref.bfloat16()
    if T <= 32:
        assert torch.equal(out, ref)
    else:
        assert (out.float() - ref.float()).abs().max().item() <= 2 ** -7 * ref.float().abs().max().item()
