"""CPU: host side of the training step — loss-row collation (train_module.py:350-373 inclusion rules), span_to_attention,
touched-parameter bookkeeping — against the golden fixtures written by the unmodified reference."""
import numpy as np
import pytest
import torch

from stair_b200 import VideoNMN, collate
from stair_b200 import _lib as L
from stair_b200.params import grad_targets
from stair_b200.train import collate_losses, span_to_attention, touched_slots
from tests import golden_util as gu


@pytest.fixture(scope='module', params=['rx_small', 'i3d_small'])
def fx(request):
    return gu.load(request.param)


def test_span_to_attention_known_answers(fx):
    for case in fx[3]['span_to_attention']:
        got = span_to_attention(tuple(case['gold']), case['T'])
        assert max(abs(float(a) - b) for a, b in zip(got, case['out'])) < 1e-6


def test_loss_rows_follow_the_reference_inclusion_rules(fx):
    cfg, weights, questions, meta, grads = fx
    batch = collate([d for d, _, _ in questions])
    rows = collate_losses(batch, set(meta['pretrain_modules']), cfg['max_video_length'])
    logs = meta['window']['logs']
    assert rows.counts['Localize'] == len(logs['Localize'])
    assert rows.counts['Temporal'] == len(logs['Temporal'])
    assert rows.counts['ExistsFrame'] == len(logs['ExistsFrame'])
    assert rows.counts['Exists/Xor'] == len(logs['Exists']) + len(logs['Xor'])
    assert rows.counts['Equals'] == len(logs['Equals'])
    assert rows.counts['contrastive'] == len(logs['Filter']) + len(logs['Superlative']) + len(logs['ToAction'])
    assert len(logs['FilterFrame']) == 0                         # excluded from training by default (args.py:62)
    ga = len(questions)
    assert all(abs(w - 1.0 / ga) < 1e-9 for w in rows.bin_w + rows.con_w)
    # the root module is never supervised (module_net.py:110, i != 0)
    roots = {int(batch.node_start[q]) + lay.root for q, lay in enumerate(batch.layouts)}
    assert not roots & set(rows.att_node + rows.bin_node + rows.con_node)


def test_touched_slots_equal_the_parameters_the_reference_gives_gradients(fx):
    cfg, weights, questions, meta, grads = fx
    model = VideoNMN(cfg, pretrain_modules=set(meta['pretrain_modules']))
    batch = collate([d for d, _, _ in questions])
    rows = collate_losses(batch, model.pretrain_modules, cfg['max_video_length'])
    touched = touched_slots(batch, rows, cfg['have_pretrain_head'])
    tg = grad_targets(model.submodules, cfg)
    name_of = {id(p): k for k, p in model.named_parameters()}
    with_grad = {name_of[id(p)] for wid in touched for p, _ in tg[wid][1]}
    ref_with_grad = {k for k, g in grads.items() if float(g.abs().max()) > 0}
    assert ref_with_grad <= with_grad
    ref_none = {k for k in meta['params_without_grad'] if not k.startswith('submodules.Superlative.localize_module.')}
    assert not (with_grad & ref_none)


def test_grad_slots_cover_every_trainable_tensor_once():
    from stair_b200 import synthetic as syn
    cfg = syn.model_config(T=8, V=128, hidden=64, object_types=16)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES)
    tg = grad_targets(model.submodules, cfg)
    covered = {}
    for wid, (numel, targets) in tg.items():
        for p, o in targets:
            assert o + p.numel() <= numel
            covered[id(p)] = covered.get(id(p), 0) + 1
    names = {id(p): k for k, p in model.named_parameters()}
    missing = sorted(k for i, k in names.items() if i not in covered)
    # Filter.attention is dead in the reference forward (Softmax over a size-1 dim): zero gradient, no slot;
    # FilterFrame.pretrain_head is only trained by the FilterFrame criterion, which is excluded by default (args.py:62)
    assert all(k.startswith(('submodules.Filter.attention.', 'submodules.FilterFrame.pretrain_head.')) for k in missing), missing
    assert all(v == 1 for v in covered.values())
    assert L.W['COUNT'] > max(tg)


def test_fast_parameter_walk_equals_module_parameters():
    """params.all_parameters is the per-forward replacement of nn.Module.parameters(): same tensors, same order, shared modules
    (Superlative.localize_module is Localize, module_net.py:33) listed once."""
    from stair_b200 import synthetic as syn
    from stair_b200.params import all_parameters
    cfg = syn.model_config(T=8, V=128, hidden=64, object_types=16)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES)
    for root in (model, model.submodules):
        assert [id(p) for p in all_parameters(root)] == [id(p) for p in root.parameters()]


def test_flat_gradient_pieces_tile_the_buffer():
    """NMNTrainStep hands out the gradients as views of ONE flat buffer cut by a single split_with_sizes call: the pieces must tile
    the buffer and every parameter's piece must have its size; only the twin LSTM biases share a piece."""
    from stair_b200 import synthetic as syn
    from stair_b200.train import NMNTrainStep
    cfg = syn.model_config(T=8, V=128, hidden=64, object_types=16)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES)
    step = NMNTrainStep(model, distributed=False)
    cuts, entries = step._grad_views()
    _, offsets, total = step._layout()
    assert sum(cuts) == total and all(c > 0 for c in cuts)
    starts = np.concatenate([[0], np.cumsum(cuts)[:-1]])
    names = {id(p): k for k, p in model.named_parameters()}
    shared = 0
    for wid, piece, prm, dup in entries:
        assert cuts[piece] == prm.numel()
        assert offsets[wid] <= starts[piece] < offsets[wid] + max(1, prm.numel()) + (1 << 30)
        shared += int(dup)
        if dup:
            assert 'bias_hh' in names[id(prm)]
    assert shared == 4                                      # b_hh of 2 encoders x 2 directions


def test_create_attention_from_frame_interval_known_answers():
    """module_net.py:190-208 (the worked example in its comments: (0.2, 5.8) -> gold[0] = 0.8, gold[1:5] = 1, gold[5] = 0.8)."""
    from stair_b200.train import create_attention_from_frame_interval
    g = create_attention_from_frame_interval([(0.2, 5.8), (2.0, 3.5), (-1.0, 99.0)], 3, 8)
    np.testing.assert_allclose(g[0], [0.8, 1, 1, 1, 1, 0.8, 0, 0], atol=1e-6)
    np.testing.assert_allclose(g[1], [0, 0, 1, 0.5, 0, 0, 0, 0], atol=1e-6)          # ceil(2.0) = 2 = floor: gold[1] += 0, gold[2:3] = 1, gold[3] += 0.5
    np.testing.assert_allclose(g[2], [0.999, 1, 1, 1, 1, 1, 1, 0.999], atol=1e-5)    # clamped to (0.001, T - 0.001)
