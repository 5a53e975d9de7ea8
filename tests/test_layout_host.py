"""CPU: host layout compiler (stair_b200/layout.py) against the reference's layout helpers recorded in
tests/golden/layouts.json (parse_program / get_childrens_and_parents / stat_module_levels / program_is_valid outputs of the
unmodified reference) and against the oracle restatement."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import nmn_oracle as orc
from stair_b200 import layout as LY, synthetic as syn
from tests import golden_util as gu

LAY = json.load(open(os.path.join(gu.GOLDEN_DIR, 'layouts.json')))


def test_arity_table_matches_reference():
    assert dict(LY.NARY) == LAY['nary']


@pytest.mark.parametrize('name', sorted(LAY['templates']))
def test_children_levels_validity(name):
    t = LAY['templates'][name]
    ch, pa = LY.children_and_parents(t['tokens'])
    assert ch == t['children'] and pa == t['parents']
    assert LY.module_levels(t['tokens']) == t['levels']
    assert LY.program_is_valid(t['tokens']) == t['valid']


def test_invalid_programs():
    for prog, res in zip(LAY['invalid'], LAY['invalid_results']):
        assert LY.program_is_valid(prog) == res


@pytest.mark.parametrize('name', sorted(syn.ALL_TEMPLATES))
def test_compiled_layout_is_consistent_with_reference_tree(name):
    tokens = syn.ALL_TEMPLATES[name][1]
    lay = LY.compile_layout(tokens)
    ch, _ = orc.children_and_parents(tokens)
    lv = orc.module_levels(tokens)
    for nd in range(lay.n):
        i = lay.token_of_node[nd]
        assert lay.level[nd] == lv[i]
        if tokens[i] in LY.OP_OF:
            assert lay.op[nd] == LY.OP_OF[tokens[i]]
            assert lay.param_tokens[i] == ch[i]                       # same children, same (pop) order
            tensor_children = [c for c in ch[i] if tokens[c] == 'video' or lay.node_of_token[c] >= 0]
            args = [a for a in lay.args[:, nd] if a != -1]
            want = [-2 if tokens[c] == 'video' else lay.node_of_token[c] for c in tensor_children]
            assert list(args) == want
    assert lay.token_of_node[lay.root] == 0


def test_interpreter_errors_mirror_reference():
    with pytest.raises(IndexError):                                   # stack.pop() on an empty stack, module_net.py:102
        LY.Layout(['Exists', 'table'])
    with pytest.raises(AssertionError):                               # assert len(stack) == 1, module_net.py:135
        LY.Layout(['table', 'Filter', 'video', 'objects'])
    with pytest.raises(KeyError):                                     # FilterFrame has no 'objects' branch, modules.py:384
        LY.Layout(['Filter', 'FilterFrame', 'video', 'objects', 'objects'])


def test_attention_rank_follows_the_reference():
    """Localize returns a 2-D [K, T] map, ExistsFrame / HasItem / Relate(1-D) a 1-D [T] one; the reference's AttnVideo / Relate / Temporal
    behave differently on the two (modules.py:318,340,427-435) — the layout compiler tracks the rank."""
    # Relate on a Localize map -> variant 2 (softmax without beta), result keeps the [1, T] shape and feeds Temporal
    lay = LY.Layout(['Filter', 'Temporal', 'while', 'video', 'Relate', 'forward', 'Localize', 'video', 'x', 'objects'])
    rel = lay.node_of_token[4]
    assert lay.variant[rel] == 2 and lay.out_rank2[rel] and lay.out_rank2[lay.node_of_token[6]]
    lay = LY.Layout(['Filter', 'AttnVideo', 'video', 'Relate', 'backward', 'ExistsFrame', 'x', 'video', 'objects'])
    rel = lay.node_of_token[3]
    assert lay.variant[rel] == 1 and not lay.out_rank2[rel]
    with pytest.raises(TypeError):      # attn.unsqueeze(1) * feat does not broadcast for a [1, T] map
        LY.Layout(['Filter', 'AttnVideo', 'video', 'Localize', 'video', 'x', 'objects'])
    with pytest.raises(TypeError):      # mean over dim 0 of a 1-D map is a scalar
        LY.Layout(['Filter', 'Temporal', 'before', 'video', 'ExistsFrame', 'x', 'video', 'objects'])
    with pytest.raises(TypeError):      # [2, T] + beta[:2] does not broadcast
        LY.Layout(['Filter', 'Temporal', 'while', 'video', 'Relate', 'forward', 'Localize', 'video', 'Array2', 'a', 'b', 'objects'])


def test_collate_tables_and_grouping():
    qs = syn.make_questions(64, 8, 32, seed=3, templates=list(syn.ALL_TEMPLATES))
    b = LY.collate(qs)
    assert b.B == 64 and b.n_nodes == sum(LY.compile_layout(q['nmn_program_list']).n for q in qs)
    assert b.video.shape == (64, 8, 32) and b.question.shape[0] == b.n_tok
    q_off = b.host_tab('q_off').numpy()
    assert q_off[0] == 0 and q_off[-1] == b.n_tok and np.all(np.diff(q_off) == [q['question'].shape[0] for q in qs])
    gid = b.host_tab('node_gid').numpy()
    assert np.array_equal(np.bincount(gid, minlength=b.n_groups), b.group_counts)
    node_q = b.host_tab('node_q').numpy()
    node_arg = b.host_tab('node_arg').numpy().reshape(3, -1)
    for k in range(3):                                                # argument edges stay inside the question
        m = node_arg[k] >= 0
        assert np.array_equal(node_q[node_arg[k][m]], node_q[m])
    # children are always scheduled in an earlier group than their parent (level-synchronous execution is valid)
    for k in range(3):
        m = node_arg[k] >= 0
        assert np.all(gid[node_arg[k][m]] < gid[m])
    groups, tab, sizes = LY.build_groups(b, frozenset(syn.PRETRAIN_MODULES))
    assert sum(g.count for g in groups) == b.n_nodes
    assert [g.node_off for g in groups] == list(np.concatenate([[0], np.cumsum(b.group_counts)[:-1]]))
    assert [g.level for g in groups] == sorted(g.level for g in groups)
    # spans: every content word got its (start, end); None span = whole question
    qs[0]['prog_str_to_question_tokens'] = {k: (None, None) for k in qs[0]['prog_str_to_question_tokens']}
    b2 = LY.collate(qs[:1])
    span = b2.host_tab('node_span').numpy().reshape(2, -1)
    lay = b2.layouts[0]
    assert all(span[0, nd] == -1 for nd in lay.word_nodes)
    del qs[1]['prog_str_to_question_tokens'][next(iter(qs[1]['prog_str_to_question_tokens']))]
    with pytest.raises(KeyError):                                     # module_net.py:128 KeyError on a missing span
        LY.collate(qs[1:2])


def test_state_dict_matches_reference_keys_and_shapes():
    from stair_b200 import VideoNMN
    for fixture in ('rx_small', 'i3d_small'):
        cfg, weights, _, meta, _ = gu.load(fixture)
        m = VideoNMN(cfg, pretrain_modules=set(meta['pretrain_modules']))
        sd = m.state_dict()
        assert set(sd) == set(weights)
        for k, v in weights.items():
            assert tuple(sd[k].shape) == tuple(v.shape), k
        m.load_state_dict(weights)
        # Superlative shares the Localize module (module_net.py:31-32)
        assert m.submodules['Superlative'].localize_module is m.submodules['Localize']


def test_product_path_fails_loudly_without_cuda():
    from stair_b200 import VideoNMN, _lib
    if torch.cuda.is_available():
        pytest.skip('CUDA present')
    cfg, weights, questions, meta, _ = gu.load('rx_small')
    m = VideoNMN(cfg)
    with pytest.raises(_lib.StairError):
        m(questions[0][0])


def test_wave_schedule_respects_dependencies_and_merges_groups():
    """Batch-level wave schedule (layout.schedule_waves): children strictly before parents, never more groups than the
    ASAP levels of utils/program_parser.py:307-321, and a cyclic (op, variant) dependency graph falls back to ASAP."""
    import numpy as np
    from stair_b200 import layout as LY, synthetic as syn
    lays = [LY.compile_layout(t[1]) for t in syn.ALL_TEMPLATES.values()]
    waves = LY.schedule_waves(lays)
    asap, merged = set(), set()
    for lay, w in zip(lays, waves):
        for nd in range(lay.n):
            for a in lay.args[:, nd]:
                if a >= 0:
                    assert w[a] < w[nd]
            assert w[nd] >= lay.level[nd]
            asap.add((int(lay.level[nd]), int(lay.op[nd]), int(lay.variant[nd])))
            merged.add((int(w[nd]), int(lay.op[nd]), int(lay.variant[nd])))
    assert len(merged) < len(asap)
    # all instances of one (op, variant) share a wave when the kind graph is acyclic
    per_kind = {}
    for lay, w in zip(lays, waves):
        for nd in range(lay.n):
            per_kind.setdefault((int(lay.op[nd]), int(lay.variant[nd])), set()).add(int(w[nd]))
    assert sum(len(v) > 1 for v in per_kind.values()) <= 2
    # Exists feeding And feeding ... and And feeding Exists in another layout: cyclic kinds -> ASAP fallback, still valid
    a = LY.compile_layout(['And', 'Exists', 'x', 'Filter', 'video', 'objects', 'Exists', 'y', 'Filter', 'video', 'objects'])
    b = LY.compile_layout(['Exists', 'x', 'And', 'Filter', 'video', 'objects', 'Filter', 'video', 'actions'])
    for lay, w in zip((a, b), LY.schedule_waves([a, b])):
        assert np.array_equal(w, lay.level)
    # collate uses the merged schedule by default and the ASAP one on request; both give the same node -> question tables
    qs = syn.make_questions(40, 8, 32, seed=3, templates=list(syn.ALL_TEMPLATES))
    m, p = LY.collate(qs), LY.collate(qs, merge_waves=False)
    assert m.n_groups < p.n_groups and m.n_nodes == p.n_nodes
    assert np.array_equal(m.host_tab('node_q').numpy(), p.host_tab('node_q').numpy())


def test_native_collate_staging_is_bit_identical_to_torch():
    """stair_host_collate_rows (csrc/host_collate.cu, host threads, no GPU): the dtype-converting row copies of a whole batch == one torch
    copy per question (video_nmn/dataset.py:463-476 collate_fn / to_device staging), for every non-NaN fp32 bit pattern."""
    g = torch.Generator().manual_seed(0)
    bits = torch.randint(-2 ** 31, 2 ** 31 - 1, (64, 1000), generator=g, dtype=torch.int64).to(torch.int32).view(torch.float32)
    bits[0, :6] = torch.tensor([float('inf'), -float('inf'), -0.0, 1e-40, 3.4e38, -3.4e38])
    d = torch.empty(64, 1000, dtype=torch.bfloat16)
    LY._stage_rows(d, [bits[i:i + 1].contiguous() for i in range(64)], list(range(64)))
    want = bits.to(torch.bfloat16)
    ok = ~torch.isnan(bits)
    assert torch.equal(d.view(torch.int16)[ok], want.view(torch.int16)[ok])
    assert bool(torch.isnan(d.float()[~ok]).all())
    back = torch.empty(64, 1000, dtype=torch.float32)
    LY._stage_rows(back, [d[i:i + 1].contiguous() for i in range(64)], list(range(64)))
    assert torch.equal(back.view(torch.int32)[ok], d.float().view(torch.int32)[ok])
    # ragged rows at arbitrary offsets, and through collate itself
    qs = syn.make_questions(40, 8, 32, seed=3, templates=list(syn.ALL_TEMPLATES))
    b = LY.collate(qs, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
    assert torch.equal(b.video, torch.stack([q['video_features'] for q in qs]).to(torch.bfloat16))
    assert torch.equal(b.question, torch.cat([q['question'] for q in qs]).to(torch.bfloat16))
    b32 = LY.collate(qs)
    assert torch.equal(b32.video, torch.stack([q['video_features'] for q in qs])) and b32.video.dtype == torch.float32


def test_length_sorted_text_schedule():
    """layout.length_sorted_schedule (StairBatch.q_order / q_soff / tok_src): a permutation in descending length, offsets in that order and
    the batch-order row of every sorted token row — and collate stores it in the packed integer table."""
    from stair_b200.layout import length_sorted_schedule
    rng = np.random.default_rng(3)
    for B in (1, 2, 65, 700):
        lens = rng.integers(1, 41, size=B)
        q_off = np.concatenate([[0], np.cumsum(lens)])
        order, soff, tok_src = length_sorted_schedule(q_off)
        assert sorted(order.tolist()) == list(range(B))
        assert (np.diff(lens[order]) <= 0).all()
        assert soff[0] == 0 and soff[-1] == q_off[-1] and (np.diff(soff) == lens[order]).all()
        for p in range(B):
            q = order[p]
            assert (tok_src[soff[p]:soff[p + 1]] == np.arange(q_off[q], q_off[q + 1])).all()
    qs = syn.make_questions(23, 8, 32, seed=4)
    b = LY.collate(qs)
    order, soff, tok_src = length_sorted_schedule(b.host_tab('q_off').numpy())
    assert (b.host_tab('q_order').numpy() == order).all() and (b.host_tab('q_soff').numpy() == soff).all() and (b.host_tab('tok_src').numpy() == tok_src).all()


def test_collate_word_spans_fast_and_slow_paths_agree():
    """collate writes the word spans of all questions of one layout in one assignment; (None, None), one-sided None and negative spans
    (python slice semantics of ``token_feature[s:t]``, module_net.py:128-129) take the per-element path — both == a per-question
    restatement, for batches that mix them."""
    rng = np.random.default_rng(5)
    qs = syn.make_questions(120, 8, 16, seed=6, templates=list(syn.ALL_TEMPLATES))
    for i, d in enumerate(qs):
        sp = dict(d['prog_str_to_question_tokens'])
        Lq = int(d['question'].shape[0])
        for k in list(sp):
            r = rng.integers(0, 8)
            if i % 3 == 0 and r == 0:
                sp[k] = (None, None)
            elif i % 3 == 0 and r == 1:
                sp[k] = (None, int(rng.integers(1, Lq + 1)))
            elif i % 3 == 0 and r == 2:
                sp[k] = (int(rng.integers(-Lq, 0)), None)
            elif i % 3 == 0 and r == 3:
                sp[k] = (-2, -1)
        d['prog_str_to_question_tokens'] = sp
    b = LY.collate(qs)
    span = b.host_tab('node_span').numpy().reshape(2, -1)
    for qi, (lay, d) in enumerate(zip(b.layouts, qs)):
        Lq = int(d['question'].shape[0])
        for nd in lay.word_nodes:
            s_, t_ = d['prog_str_to_question_tokens'][lay.token_of_node[nd]]
            if s_ is None and t_ is None:
                want = (-1, -1)
            else:
                want = slice(s_, t_).indices(Lq)[:2] if (s_ is None or t_ is None or s_ < 0 or t_ < 0) else (s_, t_)
            got = (int(span[0, b.node_start[qi] + nd]), int(span[1, b.node_start[qi] + nd]))
            assert got == tuple(want), (qi, nd, got, want)
    others = np.ones(b.n_nodes, bool)
    for qi, lay in enumerate(b.layouts):
        others[b.node_start[qi] + np.asarray(lay.word_nodes, np.int64)] = False
    assert (span[:, others] == -1).all()
