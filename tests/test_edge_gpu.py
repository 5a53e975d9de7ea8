"""GPU edge cases and size-independent properties of the batched forward (through the C ABI).

  * ragged / odd batches against the CPU oracle (fp32 strict): 1 question, batches that do not fill a 64-row recurrence block or a
    128-row GEMM tile, one-word and 40-word questions, a batch of a single layout;
  * at BASELINE.json's full size (4096 questions, H = 512, bf16): the results of a question do not depend on where it sits in the
    batch (permutation invariance, bit-exact) nor on what it is batched with (split invariance, bit-exact);
  * a batch larger than the executor's chunk caps (> 65 536 frame rows per group) equals its halves.
"""
import numpy as np
import pytest
import torch

from oracle import nmn_oracle as orc
from stair_b200 import VideoNMN, collate, synthetic as syn

pytestmark = pytest.mark.gpu


def _with_length(d, L, rng):
    d = dict(d)
    d['question'] = torch.from_numpy((rng.standard_normal((L, d['question'].shape[1])) * 0.4).astype(np.float32))
    spans = {}
    for i in d['prog_str_to_question_tokens']:
        w = int(rng.integers(1, min(3, L) + 1))
        s = int(rng.integers(0, L - w + 1))
        spans[i] = (s, s + w)
    d['prog_str_to_question_tokens'] = spans
    return d


@pytest.mark.parametrize('B', [1, 3, 65, 129])
def test_odd_batches_and_extreme_question_lengths_against_oracle(B):
    T, V, hid = 8, 192, 128
    cfg = syn.model_config(T=T, V=V, hidden=hid, object_types=16)
    torch.manual_seed(B)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32')
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    rng = np.random.default_rng(100 + B)
    qs = syn.make_questions(B, T, V, seed=B, templates=list(syn.ALL_TEMPLATES), object_types=16)
    qs = [_with_length(d, (1, 40, 2, 25)[i % 4], rng) if i % 2 == 0 else d for i, d in enumerate(qs)]     # 1-word and 40-word questions
    oracle = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES)
    with torch.no_grad():
        want = torch.stack([oracle(d, return_res_by_step=False, test_mode=True)['logits'] for d in qs])
    model = model.cuda().eval()
    out = model(qs, return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    model.check_status(out['state'])
    got = out['logits'].cpu()
    assert got.shape == want.shape
    assert float((got - want).abs().max()) <= 2e-4 * max(1.0, float(want.abs().max()))
    assert torch.equal(out['answers'].cpu().long(), want.argmax(1))


def test_single_layout_batch_and_single_dict():
    T, V, hid = 8, 128, 64
    cfg = syn.model_config(T=T, V=V, hidden=hid, object_types=16)
    torch.manual_seed(0)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32').cuda().eval()
    qs = syn.make_questions(70, T, V, seed=3, templates=['xor_between'], object_types=16)
    batched = model(qs, return_res_by_step=False, test_mode=True)['logits']
    single = torch.stack([model(d, return_res_by_step=False, test_mode=True)['logits'] for d in qs[:5]])
    torch.testing.assert_close(batched[:5], single, rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        collate([])


def test_full_size_permutation_and_split_invariance():
    """BASELINE configs[1] size.  A question's logits are bit-identical wherever it sits in the batch and whatever it is batched with."""
    B, T, V = 4096, 8, 4096
    cfg = syn.model_config(T=T, V=V)
    torch.manual_seed(0)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    qs = syn.make_questions(B, T, V, seed=1234)
    base = model(qs, return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    model.check_status(base['state'])
    logits, answers = base['logits'].clone(), base['answers'].clone()
    assert bool(torch.isfinite(logits).all())
    assert torch.equal(answers.long(), logits.argmax(1))
    perm = torch.from_numpy(np.random.default_rng(5).permutation(B))
    out_p = model([qs[i] for i in perm.tolist()], return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    assert torch.equal(out_p['logits'], logits[perm.cuda()])
    assert torch.equal(out_p['answers'], answers[perm.cuda()])
    parts = [model(qs[a:b], return_res_by_step=False, test_mode=True)['logits'].clone() for a, b in ((0, 1000), (1000, 1001), (1001, B))]
    assert torch.equal(torch.cat(parts), logits)


def test_batch_beyond_the_chunk_caps_equals_its_halves():
    """9000 questions of one 12-module layout: every group exceeds the 65 536-frame-row chunk cap of the executor."""
    B, T, V, hid = 9000, 8, 128, 128
    cfg = syn.model_config(T=T, V=V, hidden=hid, object_types=16)
    torch.manual_seed(1)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    qs = syn.make_questions(B, T, V, seed=8, templates=['xor_between', 'compare'], object_types=16)
    whole = model(qs, return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    model.check_status(whole['state'])
    logits = whole['logits'].clone()
    halves = [model(qs[a:b], return_res_by_step=False, test_mode=True)['logits'].clone() for a, b in ((0, 4500), (4500, B))]
    assert torch.equal(torch.cat(halves), logits)


def test_returned_intermediates_survive_later_forwards():
    """res_by_step / result_of_each_step tensors must be independent of the model's arena cache (the reference returns fresh
    tensors; evaluate.py:65-117 keeps Filter outputs across calls).  A second, different forward must not change what the first returned,
    and a stale ForwardState is refused instead of silently read."""
    from stair_b200 import _lib as L
    from stair_b200.nmn import OutputViews
    T, V, hid = 8, 128, 64
    cfg = syn.model_config(T=T, V=V, hidden=hid, object_types=16)
    torch.manual_seed(4)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32').cuda().eval()
    qa = syn.make_questions(34, T, V, seed=1, templates=list(syn.ALL_TEMPLATES), object_types=16)
    qb = syn.make_questions(34, T, V, seed=2, templates=list(syn.ALL_TEMPLATES), object_types=16)
    first = model(qa, return_res_by_step=True, return_result_of_each_step=True, test_mode=True)
    keep = [[(r.clone() if isinstance(r, torch.Tensor) else r) for _, r in steps] for steps in first['result_of_each_step']]
    keep_res = [{k: v[1].clone() for k, v in d.items()} for d in first['res_by_step']]
    model(qb, return_res_by_step=True, return_result_of_each_step=True, test_mode=True)           # overwrites the arenas
    torch.cuda.synchronize()
    for steps, kept in zip(first['result_of_each_step'], keep):
        for (_, r), k in zip(steps, kept):
            assert (r == k) if isinstance(r, str) else torch.equal(r, k)
    for d, kept in zip(first['res_by_step'], keep_res):
        for k, v in d.items():
            assert torch.equal(v[1], kept[k])
    with pytest.raises(L.StairError):
        OutputViews(model, first['state'], frozenset())


def test_explicit_init_and_shutdown_of_the_library_runtime_objects():
    """include/stair_b200.h "Conventions": the lane streams / events / pinned error word are the only objects the library owns;
    stair_init() creates them, stair_shutdown() destroys them, and a later forward re-creates them lazily with identical results."""
    from stair_b200 import _lib as L
    lib = L.lib()
    T, V, hid = 8, 128, 64
    cfg = syn.model_config(T=T, V=V, hidden=hid, object_types=16)
    torch.manual_seed(2)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    qs = syn.make_questions(200, T, V, seed=11, templates=list(syn.ALL_TEMPLATES), object_types=16)
    assert lib.stair_init() == 0 and lib.stair_init() == 0                 # idempotent
    a = model(qs, return_res_by_step=False, test_mode=True)['logits'].clone()
    assert lib.stair_shutdown() == 0
    b = model(qs, return_res_by_step=False, test_mode=True)['logits'].clone()
    torch.cuda.synchronize()
    assert torch.equal(a, b) and lib.stair_gemm_error_flag() == 0


def test_module_scheduling_modes_agree_and_the_timeline_hook_reports_every_group():
    """The module phase can run wave by wave or by data dependency (StairBatch.group_deps), on 1..8 lanes: same kernels, same
    operands, so the forward is bit-identical in every mode.  The backward adds into shared gradient slots with atomics when groups
    run concurrently, so its gradients agree to fp32 summation-order noise.  stair_debug_timeline brackets every group with events."""
    import ctypes
    from stair_b200 import _lib as L
    from stair_b200.train import NMNTrainStep
    lib = L.lib()
    T, V, hid = 8, 128, 64
    cfg = dict(syn.model_config(T=T, V=V, hidden=hid, object_types=16), dropout=0.0)
    torch.manual_seed(6)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    qs = syn.make_questions(300, T, V, seed=21, templates=list(syn.ALL_TEMPLATES), object_types=16, with_gold=True)
    batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
    try:
        outs, grads = {}, {}
        for dep, lanes in ((1, 8), (0, 8), (1, 3), (0, 1)):
            lib.stair_set_dep_sched(dep); lib.stair_set_lanes(lanes); lib.stair_set_bwd_lanes(lanes)
            outs[(dep, lanes)] = model.forward_batch(batch).logits.clone()
            step = NMNTrainStep(model.train(), distributed=False)
            step.run(step.plan(batch), assign_grads=False, dropout_seed=5)
            grads[(dep, lanes)] = step.last['flat'].clone()
            model.eval()
        ref = outs[(1, 8)]
        assert all(torch.equal(o, ref) for o in outs.values())
        g0 = grads[(0, 1)]                                                   # one lane: no concurrent adds
        for k, g in grads.items():
            assert float((g - g0).abs().max()) <= 1e-4 * float(g0.abs().max()), k
        # timeline hook (dependency scheduling only)
        lib.stair_set_dep_sched(1); lib.stair_set_lanes(8)
        lib.stair_debug_timeline(1)
        st = model.forward_batch(batch)
        cap = 96
        t0 = np.zeros(cap, np.float32); t1 = np.zeros(cap, np.float32)
        lane, op, cnt, var = (np.zeros(cap, np.int32) for _ in range(4))
        n = lib.stair_debug_timeline_read(*(a.ctypes.data_as(ctypes.c_void_p) for a in (t0, t1, lane, op, cnt, var)), cap)
        assert n == batch.n_groups
        assert (t1[:n] >= t0[:n]).all() and (t0[:n] >= 0).all() and (lane[:n] >= 0).all() and (lane[:n] < 8).all()
        from stair_b200 import layout as LY
        groups, _, _ = LY.build_groups(batch, frozenset())
        assert [int(c) for c in cnt[:n]] == [groups[g].count for g in range(n)] and [int(o) for o in op[:n]] == [groups[g].op for g in range(n)]
        model.check_status(st)
    finally:
        lib.stair_debug_timeline(0); lib.stair_set_dep_sched(1); lib.stair_set_lanes(8); lib.stair_set_bwd_lanes(8)


@pytest.mark.parametrize('T,V', [(8, 256), (64, 128)])
def test_optional_kernel_forms_do_not_change_the_forward(T, V):
    """Comparison switches of the C ABI (include/stair_b200.h): the frame sum of Filter inside the GEMM epilogue (stair_set_fuse_sum), the
    gathered-A form of the CTA-pair GEMM (stair_set_gemm_pair_gather), the two-CTAs-per-SM GEMM (stair_set_gemm_small) and the
    weight-stationary recurrence (stair_set_lstm_ws), the text recurrence in batch order instead of length-sorted (stair_set_text_sort) compute the same values as the product configuration: logits bit-identical."""
    from stair_b200 import _lib as L
    lib = L.lib()
    cfg = syn.model_config(T=T, V=V, hidden=512, object_types=16)
    torch.manual_seed(9)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    qs = syn.make_questions(1500, T, V, seed=5, templates=list(syn.ALL_TEMPLATES), object_types=16)
    batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
    ref = model.forward_batch(batch).logits.clone()
    switches = (('stair_set_fuse_sum', 1, 0), ('stair_set_gemm_pair_gather', 0, 1), ('stair_set_gemm_small', 1, 0), ('stair_set_lstm_ws', 1, 0),
                ('stair_set_gemm_pair', 2, 1), ('stair_set_gemm_pair', 0, 1), ('stair_set_text_sort', 0, 1))
    for name, on, default in switches:
        try:
            getattr(lib, name)(on)
            out = model.forward_batch(batch).logits.clone()
            torch.cuda.synchronize()
        finally:
            getattr(lib, name)(default)
        assert lib.stair_gemm_error_flag() == 0
        assert torch.equal(out, ref), (name, on, float((out - ref).abs().max()))


_LAUNCHES = {}


def outs_launches(model, on, device_sort, outs):
    """The device sort adds exactly two launches to the forward; the other two modes launch the same kernels."""
    _LAUNCHES.setdefault(id(model), model.last_launches)
    return _LAUNCHES[id(model)] + (2 if (on and device_sort) else 0)


@pytest.mark.parametrize('B', [1, 63, 333, 1100])
def test_length_sorted_text_recurrence_is_bit_identical_to_batch_order(B):
    """The inference text recurrence runs over length-sorted questions (stair_set_text_sort, default on; csrc/lstm_fused.cu
    text_sort_kernel; schedule from collate or from the library's device sort): token_feature, question_feature and the logits equal the batch-order run bit for bit — lengths 1 .. 40 incl. many
    equal ones, batches that do not fill a 64-question block, and the text-only phase (encode_question)."""
    from stair_b200 import _lib as L
    lib = L.lib()
    T, V = 8, 128
    cfg = syn.model_config(T=T, V=V, hidden=256, object_types=16)
    torch.manual_seed(B)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    rng = np.random.default_rng(B)
    qs = syn.make_questions(B, T, V, seed=B, templates=list(syn.ALL_TEMPLATES), object_types=16)
    qs = [_with_length(d, int(rng.integers(1, 41)), rng) if i % 3 else d for i, d in enumerate(qs)]
    batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
    outs = []
    for on, device_sort in ((1, False), (0, False), (1, True)):      # collate's schedule | batch order | the library's device counting sort
        try:
            lib.stair_set_text_sort(on)
            model.device_text_sort = device_sort
            st = model.forward_batch(batch)
            torch.cuda.synchronize()
            model.check_status(st)
            outs.append((st.logits.clone(), st.tokfeat[:batch.n_tok * 256].clone(), st.qfeat[:B * 256].clone()))
            assert model.last_launches == outs_launches(model, on, device_sort, outs)
        finally:
            lib.stair_set_text_sort(1)
            model.device_text_sort = False
    for a, b in ((outs[0], outs[1]), (outs[0], outs[2])):
        for x, y, name in zip(a, b, ('logits', 'token_feature', 'question_feature')):
            assert torch.equal(x, y), (name, float((x.float() - y.float()).abs().max()))
    assert float(outs[0][2].float().abs().max()) > 0


@pytest.mark.parametrize('T', [16, 32, 128])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_other_frame_counts_against_oracle(T, precision):
    """Frame counts other than the two benchmark configurations: T = 16 / 32 (Linear(T, T) relate, modules.py:271-277) and T = 128
    (Conv1d relate with k = 32, 32, 65, modules.py:255-266; one instance spans a whole 128-row GEMM tile) — templates and random layouts
    against the CPU oracle."""
    V, hid = 96, 128
    cfg = syn.model_config(T=T, V=V, hidden=hid, object_types=16)
    torch.manual_seed(T)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision=precision)
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    qs = syn.make_questions(34, T, V, seed=T, templates=list(syn.ALL_TEMPLATES), object_types=16) + syn.make_random_questions(40, T, V, seed=T + 1)
    oracle = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES)
    with torch.no_grad():
        want = torch.stack([oracle(d, return_res_by_step=False, test_mode=True)['logits'] for d in qs])
    model = model.cuda().eval()
    out = model(qs, return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    model.check_status(out['state'])
    got = out['logits'].cpu()
    scale = max(1.0, float(want.abs().max()))
    if precision == 'fp32':
        assert float((got - want).abs().max()) <= 2e-4 * scale
        assert torch.equal(out['answers'].cpu().long(), want.argmax(1))
    else:
        assert float((got - want).abs().max()) <= 3e-2 * float(want.abs().max()) + 2e-3


@pytest.mark.parametrize('dims', [dict(V=100, text_size=52, answer_vocab=37, hidden=384), dict(V=200, text_size=768, answer_vocab=1000, hidden=256),
                                  dict(V=4096, text_size=300, answer_vocab=172, hidden=320)])
@pytest.mark.parametrize('precision', ['fp32', 'bf16'])
def test_other_model_dimensions_against_oracle(dims, precision):
    """Model dimensions other than the benchmark's: feature / embedding sizes that are not multiples of 8 (staged operands instead of direct
    TMA), a BERT-sized text embedding, small and large answer vocabularies, hidden sizes 256 / 320 / 384 (h = 160 is not a fused-recurrence
    size: per-step LSTM path; 384 is not a multiple of 256: no CTA-pair GEMM) — templates and random layouts against the CPU oracle."""
    T = 8
    cfg = syn.model_config(T=T, V=dims['V'], hidden=dims['hidden'], text_size=dims['text_size'], answer_vocab=dims['answer_vocab'], object_types=16)
    torch.manual_seed(dims['V'])
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision=precision)
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    kw = dict(text_size=dims['text_size'], answer_vocab=dims['answer_vocab'])
    qs = syn.make_questions(34, T, dims['V'], seed=1, templates=list(syn.ALL_TEMPLATES), object_types=16, **kw) + \
        syn.make_random_questions(30, T, dims['V'], seed=2, **kw)
    oracle = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES)
    with torch.no_grad():
        want = torch.stack([oracle(d, return_res_by_step=False, test_mode=True)['logits'] for d in qs])
    model = model.cuda().eval()
    out = model(qs, return_res_by_step=False, test_mode=True)
    torch.cuda.synchronize()
    model.check_status(out['state'])
    got = out['logits'].cpu()
    assert got.shape == want.shape
    if precision == 'fp32':
        assert float((got - want).abs().max()) <= 2e-4 * max(1.0, float(want.abs().max()))
        assert torch.equal(out['answers'].cpu().long(), want.argmax(1))
    else:
        assert float((got - want).abs().max()) <= 3e-2 * float(want.abs().max()) + 2e-3
