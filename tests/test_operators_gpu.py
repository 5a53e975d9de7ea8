"""GPU parity of (1) the per-operator ``forward(*params)`` surface (reference: video_nmn/modules.py:7-465, one class per module,
``NAME_TO_MODULE`` :446-465) and (2) the exported single-operator C entry points (include/stair_b200.h "single operators"),
each against the CPU oracle's per-operator methods / ATen restatements on the same seeded inputs.

(1) goes through ``stair_op_forward`` — the same group code and kernels the batched interpreter runs; (2) calls the C ABI directly
with ctypes.  Bars: fp32 strict rtol 2e-4 / atol 2e-5; bf16 storage ``|err| <= 1e-2 * max|ref| + 1e-3``; argmax of attention maps
bit-exact in fp32 (and in bf16 where the reference's top-2 margin exceeds the tolerance); the temporal mask scans
(``stair_relate_scan`` vs the reference's own ``TemporalModule.relate_`` fixture) <= 1e-6 with bit-exact argmax.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import nmn_oracle as orc
from stair_b200 import VideoNMN, synthetic as syn, _lib as L, NAME_TO_MODULE
from tests import golden_util as gu

pytestmark = pytest.mark.gpu

STRICT = dict(rtol=2e-4, atol=2e-5)
BF16_REL, BF16_ABS = 1e-2, 1e-3


def _close(got, want, precision, what):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    assert got.shape == want.shape, '%s: shape %s vs %s' % (what, tuple(got.shape), tuple(want.shape))
    if precision == 'fp32':
        torch.testing.assert_close(got, want, msg=lambda m: '%s: %s' % (what, m), **STRICT)
    else:
        scale = max(float(want.abs().max()), 1e-3)
        err = float((got - want).abs().max())
        assert err <= BF16_REL * scale + BF16_ABS, '%s: max err %g vs scale %g' % (what, err, scale)


def _argmax_equal(got, want, precision, what):
    got, want = got.detach().float().cpu(), want.detach().float().cpu()
    ok = torch.ones(want.shape[:-1], dtype=torch.bool)
    if precision == 'bf16':
        top = want.topk(2, dim=-1).values
        ok = (top[..., 0] - top[..., 1]) > 2 * (BF16_REL * float(want.abs().max()) + BF16_ABS)
    assert torch.equal(got.argmax(-1)[ok], want.argmax(-1)[ok]), '%s: attention argmax' % what
    return int(ok.sum())


def _pair(T, precision, hidden=128, seed=0):
    cfg = syn.model_config(T=T, V=64, hidden=hidden, object_types=16)
    torch.manual_seed(seed)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision=precision)
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    return cfg, model.cuda().eval(), orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES)


def _rnd(gen, shape, precision, scale=1.0, positive=False):
    """fp32 operand; in bf16 mode pre-rounded to bf16 so the oracle sees exactly what the arenas hold."""
    x = torch.randn(shape, generator=gen) * scale
    if positive:
        x = x.abs()
    return x.bfloat16().float() if precision == 'bf16' else x


@pytest.mark.parametrize('T', [8, 64])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_every_operator_forward_matches_the_oracle(T, precision):
    cfg, model, oracle = _pair(T, precision)
    H, n = cfg['hidden_size'], 37
    gen = torch.Generator().manual_seed(100 + T)
    sub = model.submodules
    feat = _rnd(gen, (n, T, H), precision)
    feat2 = _rnd(gen, (n, T, H), precision)
    v1, v2, v3 = (_rnd(gen, (n, H), precision) for _ in range(3))
    a1 = torch.rand((n, T), generator=gen) * 0.98
    a2 = torch.rand((n, T), generator=gen) * 0.98
    aK = torch.rand((n, 2, T), generator=gen) * 0.98
    dev = lambda *xs: [x.cuda() if isinstance(x, torch.Tensor) else x for x in xs]       # noqa: E731

    def run(name, *params, att=False):
        got = sub[name].forward_batched(*dev(*params))
        want = torch.stack([getattr(oracle, name)(*[p if isinstance(p, str) else p[i] for p in params]) for i in range(n)])
        _close(got, want, precision, '%s%s' % (name, [p for p in params if isinstance(p, str)]))
        if att:
            _argmax_equal(got, want, precision, name)
        # one instance with the reference's own (unbatched) shapes
        one = sub[name](*dev(*[p if isinstance(p, str) else p[3] for p in params]))
        assert one.shape == want[3].shape, name
        _close(one, want[3], precision, name + ' (single instance)')
        return got, want

    with torch.no_grad():
        run('And', v1, v2); run('And', a1, a2); run('And', aK, aK.flip(0))
        run('XorFrame', a1, a2); run('XorFrame', v1, v2)
        run('AttnVideo', feat, a1)
        got, want = run('Choose', v1, v2, v3)
        run('Compare', v1, v2); run('Equals', v1, v2); run('Xor', v1, v2); run('ToAction', v1, v2); run('Exists', v1, v2)
        got, _ = run('Array2', v1, v2)
        assert got.shape == (n, 2, H)
        run('ExistsFrame', v1, feat, att=True)
        for kw in (v1, 'actions', 'objects', 'relations'):
            run('Filter', feat, kw)
        for kw in (v1, 'relations', 'actions'):
            run('FilterFrame', feat, kw)
        with pytest.raises(KeyError):                                    # no 'objects' branch in the reference either (modules.py:398-414)
            sub['FilterFrame'].forward_batched(feat.cuda(), 'objects')
        run('HasItem', feat, att=True)
        got, _ = run('Localize', feat, v1, att=True)
        assert got.shape == (n, 1, T)
        kw2 = torch.stack([v1, v2], 1)
        got, _ = run('Localize', feat, kw2, att=True)
        assert got.shape == (n, 2, T)
        run('Relate', 'forward', a1, att=True); run('Relate', 'backward', a1, att=True)
        for mode in ('max', 'min'):
            run('Superlative', mode, v1, feat); run('Superlative', mode, kw2, feat); run('Superlative', mode, feat2, feat)
        for mode in ('while', 'before', 'after', 'between'):
            att_in = aK if mode == 'between' else aK[:, :1]
            run('Temporal', mode, feat, att_in)
            # the stateful head (modules.py:287-288): pretrain_head() returns the related attention of the last forward
            want_rel = oracle.temporal_relate(mode, att_in[3].mean(0))
            _close(sub['Temporal'].pretrain_head(), want_rel, precision, 'Temporal.pretrain_head after a single-instance forward')


def test_name_to_module_mirrors_the_reference_registry():
    """video_nmn/modules.py:446-465: a dict name -> class, in the reference's order; module_net.py:27-35 constructor protocol."""
    assert list(NAME_TO_MODULE) == ['And', 'AttnVideo', 'Choose', 'Compare', 'Equals', 'Exists', 'ExistsFrame', 'Filter', 'FilterFrame',
                                    'HasItem', 'Localize', 'Relate', 'Superlative', 'Temporal', 'ToAction', 'Xor', 'XorFrame', 'Array2']
    cfg, model, _ = _pair(8, 'fp32')
    for name, cls in NAME_TO_MODULE.items():
        assert type(model.submodules[name]) is cls
    assert model.submodules['Superlative'].localize_module is model.submodules['Localize']
    # an operator that is not attached to a model cannot run (there is no CPU / torch path)
    with pytest.raises(L.StairError):
        NAME_TO_MODULE['And'](cfg)(torch.zeros(4, device='cuda'), torch.zeros(4, device='cuda'))


# ---------------------------------------------------------------------------------------------------------------------
# exported single-operator C entry points
# ---------------------------------------------------------------------------------------------------------------------
def _st():
    return L.stream_ptr()


MODES = {'while': 0, 'before': 1, 'after': 2, 'between': 3}


def _relate_scan(att, mode):
    n, K, T = att.shape
    out = torch.empty((n, T), dtype=torch.float32, device='cuda')
    att_dev = att.cuda().contiguous()
    L.check(L.lib().stair_relate_scan(L.ptr(att_dev), L.i32(MODES[mode]), L.ptr(out), L.i32(n), L.i32(T), _st()), 'stair_relate_scan')
    return out.cpu()


def test_relate_scan_matches_the_reference_fixture():
    """stair_relate_scan vs TemporalModule.relate_ of the unmodified reference (tests/golden/relate_scan.npz)."""
    fx = np.load(os.path.join(gu.GOLDEN_DIR, 'relate_scan.npz'))
    for T in (8, 64):
        for mode in MODES:
            att = torch.from_numpy(fx['T%d/%s/att' % (T, mode)])
            want = torch.from_numpy(fx['T%d/%s/out' % (T, mode)])
            got = _relate_scan(att, mode)
            assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max())), (T, mode)
            assert torch.equal(got.argmax(-1), want.argmax(-1)), (T, mode)
            if mode == 'while':
                assert torch.equal(got, want)


@pytest.mark.parametrize('T', [8, 13, 64, 150])
def test_relate_scan_random_instances(T):
    """>= 1k random instances per mode (incl. T that is not a multiple of the warp size and T > 32: carried scans) vs the oracle."""
    gen = torch.Generator().manual_seed(T)
    n = 1536
    for mode in MODES:
        K = 2 if mode == 'between' else 1
        att = torch.rand((n, K, T), generator=gen) * 1.4 - 0.4
        want = torch.stack([orc.OracleNMN.relate_scan(a, mode) for a in att])
        got = _relate_scan(att, mode)
        assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max())), mode
        # argmax: the scans are monotone (ties at the plateau edges are decided by 1-ulp summation-order effects), so compare
        # wherever the reference's top-2 margin is above the value tolerance
        top = want.topk(2, dim=-1).values
        ok = (top[:, 0] - top[:, 1]) > 4e-6 * max(1.0, float(want.abs().max()))
        assert torch.equal(got.argmax(-1)[ok], want.argmax(-1)[ok]), mode
        assert int(ok.sum()) > n // 4 or mode in ('before', 'after')
    # the drop-in's TemporalModule.relate_ (device) == the reference formula on one instance
    _, model, _ = _pair(8, 'fp32')
    a = torch.rand((2, T), generator=gen)
    got = model.submodules['Temporal'].relate_(a.cuda(), 'between').cpu()
    want = orc.OracleNMN.relate_scan(a, 'between')
    assert float((got - want).abs().max()) <= 1e-6 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('T', [8, 64])
def test_single_operator_entry_points(dtype, T):
    lib = L.lib()
    H, n = 512, 1100
    gen = torch.Generator().manual_seed(7 + T)
    precision = 'fp32' if dtype == torch.float32 else 'bf16'
    dc = L.i32(L.dtype_code(dtype))

    def rnd(*shape):
        x = torch.randn(shape, generator=gen)
        return x.to(dtype).float()

    keep = []

    def dev(x, dt=None):
        """Upload and HOLD a reference: ``L.ptr(x.cuda())`` on a temporary would let the caching allocator hand the block to the next
        temporary of the same call before the kernel has run."""
        y = (x.to(dt) if dt is not None else x).cuda()
        keep.append(y)
        return L.ptr(y)

    # stair_cos_attention (Localize tail, modules.py:205-216): f [n*T, H], k [n*K, H] -> att [n][K][T]
    for K in (1, 2):
        f, k = rnd(n, T, H), rnd(n, K, H)
        att = torch.empty((n, K, T), dtype=torch.float32, device='cuda')
        L.check(lib.stair_cos_attention(dc, dev(f, dtype), dev(k, dtype), L.i32(K), L.i32(T), L.i32(H), L.ptr(att),
                                        L.i32(n), _st()), 'stair_cos_attention')
        want = (orc._cos(f.unsqueeze(1), k.unsqueeze(2)) + 1) * 0.49
        _close(att, want, precision, 'cos_attention K=%d' % K)
        assert _argmax_equal(att, want, precision, 'cos_attention') > 0
    # stair_relate (modules.py:417-435)
    a, beta = torch.rand((n, T), generator=gen), torch.rand(T, generator=gen)
    for sign in (1, -1):
        out = torch.empty((n, T), dtype=torch.float32, device='cuda')
        L.check(lib.stair_relate(dev(a), dev(beta), L.i32(sign), L.ptr(out), L.i32(n), L.i32(T), _st()), 'stair_relate')
        want = torch.softmax(a + sign * beta, dim=-1)
        torch.testing.assert_close(out.cpu(), want, rtol=1e-5, atol=1e-7)
        assert torch.equal(out.cpu().argmax(-1), want.argmax(-1))
    # stair_layernorm (modules.py:283,327)
    x, gam, bet = rnd(n * T, H), torch.randn(H, generator=gen), torch.randn(H, generator=gen)
    out = torch.empty((n * T, H), dtype=dtype, device='cuda')
    L.check(lib.stair_layernorm(dc, dev(x, dtype), dev(gam), dev(bet), L.ptr(out), L.i64(n * T), L.i32(H), _st()),
            'stair_layernorm')
    _close(out, F.layer_norm(x, (H,), gam, bet, 1e-5), precision, 'layernorm')
    # stair_sum_frames (Filter aggregation, modules.py:374)
    x = rnd(n, T, H)
    out = torch.empty((n, H), dtype=dtype, device='cuda')
    L.check(lib.stair_sum_frames(dc, dev(x, dtype), L.ptr(out), L.i32(n), L.i32(T), L.i32(H), _st()), 'stair_sum_frames')
    _close(out, x.sum(1), precision, 'sum_frames')
    # stair_attn_video (modules.py:330-340) and stair_exists_frame (modules.py:162-178) on an arena with permuted indices
    vid = rnd(2 * n, T, H)
    vid_dev = vid.to(dtype).cuda()
    att = torch.rand((n, T), generator=gen)
    fi = torch.randperm(n, generator=gen).int()
    ai = torch.randperm(n, generator=gen).int()
    L.check(lib.stair_attn_video(dc, L.ptr(vid_dev), dev(fi), dev(att), dev(ai), L.i32(n), L.i32(n), L.i32(T), L.i32(H), _st()),
            'stair_attn_video')
    want = att[ai.long()].unsqueeze(-1) * vid[fi.long()]
    _close(vid_dev[n:], want, precision, 'attn_video')
    assert torch.equal(vid_dev[:n].float().cpu(), vid[:n])                                # inputs untouched
    kw = rnd(n, H)
    ki = torch.randperm(n, generator=gen).int()
    out = torch.zeros((n + 5, T), dtype=torch.float32, device='cuda')
    L.check(lib.stair_exists_frame(dc, L.ptr(vid_dev), dev(fi), dev(kw, dtype), dev(ki), L.ptr(out), L.i32(5), L.i32(n),
                                   L.i32(T), L.i32(H), _st()), 'stair_exists_frame')
    want = (orc._cos(vid[fi.long()], kw[ki.long()].unsqueeze(1)) + 1) * 0.49
    _close(out[5:], want, precision, 'exists_frame')
    _argmax_equal(out[5:], want, precision, 'exists_frame')
    assert float(out[:5].abs().max()) == 0.0
    # stair_hasitem_tail (modules.py:128-129)
    x, w, b = rnd(n, T, H), torch.randn(H, generator=gen) * 0.05, torch.randn(1, generator=gen)
    out = torch.empty((n, T), dtype=torch.float32, device='cuda')
    L.check(lib.stair_hasitem_tail(dc, dev(x, dtype), dev(w), dev(b), L.ptr(out), L.i32(0), L.i32(n), L.i32(T), L.i32(H), _st()),
            'stair_hasitem_tail')
    want = torch.sigmoid(x @ w + b)
    _close(out, want, precision, 'hasitem_tail')
    # stair_argmax (train_module.py:252; first maximal index) incl. exact ties
    z = torch.randn((n, 172), generator=gen)
    z[::7, 5] = z[::7].max(1).values
    z[::7, 100] = z[::7, 5]
    out = torch.empty(n, dtype=torch.int32, device='cuda')
    L.check(lib.stair_argmax(dev(z), L.ptr(out), L.i32(n), L.i32(172), _st()), 'stair_argmax')
    assert torch.equal(out.cpu().long(), z.argmax(1))
    # stair_l2normalize (module_net.py:211-216)
    x = rnd(n, H)
    x[3] = 0
    out = torch.empty((n, H), dtype=torch.float32, device='cuda')
    L.check(lib.stair_l2normalize(dc, dev(x, dtype), L.ptr(out), L.i32(n), L.i32(H), _st()), 'stair_l2normalize')
    torch.testing.assert_close(out.cpu(), F.normalize(x, dim=1, eps=1e-12), rtol=1e-5, atol=1e-7)
