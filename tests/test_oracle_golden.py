"""Pin the CPU oracle (oracle/nmn_oracle.py) to the golden fixtures produced by the unmodified reference."""
import json
import os

import pytest
import torch

from oracle import nmn_oracle as orc
from tests import golden_util as gu

FIXTURES = ['rx_small', 'i3d_small']
TOL = dict(rtol=2e-5, atol=2e-6)


@pytest.fixture(scope='module', params=FIXTURES)
def fx(request):
    return gu.load(request.param)


def test_layout_helpers_match_reference():
    lay = json.load(open(os.path.join(gu.GOLDEN_DIR, 'layouts.json')))
    assert {k: v for k, v in orc.NARY.items()} == lay['nary']
    for name, t in lay['templates'].items():
        ch, pa = orc.children_and_parents(t['tokens'])
        assert ch == t['children'], name
        assert pa == t['parents'], name
        assert orc.module_levels(t['tokens']) == t['levels'], name
        assert orc.program_is_valid(t['tokens']) == t['valid'], name
    for prog, res in zip(lay['invalid'], lay['invalid_results']):
        assert orc.program_is_valid(prog) == res


def test_forward_every_intermediate(fx):
    cfg, weights, questions, meta, _ = fx
    model = orc.OracleNMN(cfg, weights, meta['pretrain_modules'])
    for data, ref, q in questions:
        with torch.no_grad():
            out = model(data, return_res_by_step=True, return_result_of_each_step=True)
        torch.testing.assert_close(out['logits'], ref['logits'], **TOL)
        assert int(out['logits'].argmax()) == int(ref['logits'].argmax())
        assert len(out['result_of_each_step']) == len(ref['steps'])
        for j, ((_, got), want) in enumerate(zip(out['result_of_each_step'], ref['steps'])):
            if isinstance(want, str):
                assert got == want
            else:
                torch.testing.assert_close(got, want, msg=lambda m: '%s step %d: %s' % (q['template'], j, m), **TOL)
                if want.dim() >= 1 and want.size(-1) == cfg['max_video_length'] and want.numel() <= 2 * want.size(-1):
                    assert torch.equal(got.argmax(-1), want.argmax(-1))       # attention argmax index
        assert set(out['res_by_step']) == set(ref['res_by_step'])
        for k, (m, t) in ref['res_by_step'].items():
            assert out['res_by_step'][k][0] == m
            torch.testing.assert_close(out['res_by_step'][k][1], t, **TOL)
        for k, reps in ref['gold_reps'].items():
            for (n, t), (n2, t2) in zip(reps, out['sg_res_by_step'][k]):
                assert n == n2
                torch.testing.assert_close(t2, t, **TOL)


def test_span_to_attention(fx):
    for case in fx[3]['span_to_attention']:
        got = orc.span_to_attention(tuple(case['gold']), case['T'])
        torch.testing.assert_close(got, torch.tensor(case['out']), rtol=1e-6, atol=1e-7)


def test_window_loss_and_gradients(fx):
    cfg, weights, questions, meta, grads = fx
    w = {k: v.clone().requires_grad_(True) for k, v in weights.items()}
    for k in list(w):        # Superlative.localize_module.* aliases Localize.* (module_net.py:31-32)
        if k.startswith('submodules.Superlative.localize_module.'):
            w[k] = w[k.replace('Superlative.localize_module', 'Localize')]
    model = orc.OracleNMN(cfg, w, meta['pretrain_modules'])
    crit = orc.OracleCriterion({'obj_%d' % i: i for i in range(cfg['object_types'])})
    batch = [d for d, _, _ in questions]
    total, logs, outs = orc.window_loss(model, crit, batch)
    assert abs(float(total) - meta['window']['loss']) < 2e-5 * abs(meta['window']['loss'])
    for name, vals in meta['window']['logs'].items():
        assert len(vals) == len(logs[name]), name
        for a, b in zip(sorted(vals), sorted(logs[name])):
            assert abs(a - b) <= 2e-5 * max(1.0, abs(a)), name
    # FilterFrame criterion (excluded from training by default, args.py:62) checked separately
    for it, step, val in meta['window']['filterframe_losses']:
        res = outs[it]['res_by_step'][step][1]
        got = float(crit('FilterFrame', res, outs[it]['sg_res_by_step'][step]))
        assert abs(got - val) <= 2e-5 * max(1.0, abs(val))
    total.backward()
    for k, g in grads.items():
        got = w[k].grad
        assert got is not None, k
        scale = max(float(g.abs().max()), 1e-6)
        assert float((got - g).abs().max()) <= 1e-4 * scale + 1e-7, k
    for k in meta['params_without_grad']:
        if k in w and not k.startswith('submodules.Superlative.localize_module.'):
            assert w[k].grad is None or float(w[k].grad.abs().max()) == 0.0, k


def test_relate_scan_pinned_to_the_reference():
    """oracle.relate_scan == the reference's own TemporalModule.relate_ (modules.py:290-308) on the committed fixture
    (tests/golden/relate_scan.npz, written by make_golden.py from the unmodified reference object): all four modes, T = 8 and 64."""
    import numpy as np
    fx = np.load(os.path.join(gu.GOLDEN_DIR, 'relate_scan.npz'))
    n = 0
    for T in (8, 64):
        for mode in ('while', 'before', 'after', 'between'):
            att = torch.from_numpy(fx['T%d/%s/att' % (T, mode)])
            want = torch.from_numpy(fx['T%d/%s/out' % (T, mode)])
            got = torch.stack([orc.OracleNMN.relate_scan(a, mode) for a in att])
            assert torch.equal(got, want), (T, mode)          # same ATen ops in the same order: bit-identical
            n += len(att)
    assert n == 8 * 24


def test_golden_window_supervises_equals_and_xor(fx):
    """criterion_equals / criterion_exists-on-Xor (train_module.py:92-107) only run for NON-root Equals / Xor nodes; the fixtures
    must contain such rows so that the loss and gradient tests above (and the GPU ones) cover them."""
    logs = fx[3]['window']['logs']
    assert len(logs['Equals']) >= 2 and len(logs['Xor']) >= 2 and len(logs['Exists']) >= 5
    assert 'submodules.Equals.pretrain_head.weight' in fx[4] and 'submodules.Xor.pretrain_head.weight' in fx[4]
    assert float(fx[4]['submodules.Equals.pretrain_head.weight'].abs().max()) > 0
    assert float(fx[4]['submodules.Xor.pretrain_head.weight'].abs().max()) > 0


def test_relate_scan_matches_reference_formula():
    # modules.py:290-308 is dead code in the reference forward; restated and checked on a hand example.
    a = torch.tensor([0.1, -0.2, 0.5, 0.0, 0.3])
    torch.testing.assert_close(orc.OracleNMN.relate_scan(a, 'before'), torch.tensor([0.1, 0.1, 0.6, 0.6, 0.9]))
    torch.testing.assert_close(orc.OracleNMN.relate_scan(a, 'after'), torch.tensor([0.9, 0.8, 0.8, 0.3, 0.3]))
    ab = torch.stack([a, a.flip(0)])
    got = orc.OracleNMN.relate_scan(ab, 'between')
    b = torch.relu(a.flip(0))
    want = torch.max(torch.min(torch.cumsum(torch.relu(a), 0), torch.cumsum(torch.relu(a).flip(0), 0).flip(0)),
                     torch.min(torch.cumsum(b, 0), torch.cumsum(b.flip(0), 0).flip(0)))
    torch.testing.assert_close(got, want)


def test_aten_lstm_variant_equals_loop(fx):
    """The CPU-baseline variant (encoders through torch.nn.LSTM, the reference's own call) == the explicit restatement."""
    cfg, weights, questions, meta, _ = fx
    a = orc.OracleNMN(cfg, weights, meta['pretrain_modules'])
    b = orc.OracleNMN(cfg, weights, meta['pretrain_modules'], aten_lstm=True)
    for data, ref, _ in questions[:4]:
        with torch.no_grad():
            la = a(data, return_res_by_step=False, test_mode=True)['logits']
            lb = b(data, return_res_by_step=False, test_mode=True)['logits']
        torch.testing.assert_close(la, lb, **TOL)
        torch.testing.assert_close(lb, ref['logits'], **TOL)


@pytest.mark.parametrize('name', ['rx_small', 'i3d_small'])
def test_oracle_matches_the_reference_on_random_layouts(name):
    """150 random well-typed layouts per configuration (all 18 operators, up to 17 module calls, compositions none of the probed AGQA
    templates contain): the oracle's logits equal the unmodified reference's (tests/golden/make_random_golden.py)."""
    cfg, weights, qs, want, meta = gu.load_random(name)
    oracle = orc.OracleNMN(cfg, weights, meta['pretrain_modules'])
    with torch.no_grad():
        got = torch.stack([oracle(d, return_res_by_step=False, test_mode=True)['logits'] for d in qs])
    torch.testing.assert_close(got, want, **TOL)
    assert torch.equal(got.argmax(1), want.argmax(1))


@pytest.mark.parametrize('name', ['rx_small', 'i3d_small'])
def test_oracle_window_on_random_layouts(name):
    """Row L on random layouts: a 40-question window with supervision on every supervisable non-root module (all nine criteria fire) —
    the oracle's window loss, per-criterion logs and autograd gradients equal the reference's (make_random_golden.py)."""
    cfg, weights, qs, tm, grads, meta = gu.load_random_train(name)
    w = {k: v.clone().requires_grad_(True) for k, v in weights.items()}
    for k in list(w):
        if k.startswith('submodules.Superlative.localize_module.'):
            w[k] = w[k.replace('Superlative.localize_module', 'Localize')]
    model = orc.OracleNMN(cfg, w, meta['pretrain_modules'])
    crit = orc.OracleCriterion({'obj_%d' % i: i for i in range(cfg['object_types'])})
    total, logs, _ = orc.window_loss(model, crit, qs)
    assert abs(float(total) - tm['loss']) < 2e-5 * abs(tm['loss'])
    for mname, vals in tm['logs'].items():
        assert len(vals) == len(logs[mname]), mname
        for a, b in zip(sorted(vals), sorted(logs[mname])):
            assert abs(a - b) <= 2e-5 * max(1.0, abs(a)), mname
    total.backward()
    for k, g in grads.items():
        got = w[k].grad
        assert got is not None, k
        assert float((got - g).abs().max()) <= 1e-4 * max(float(g.abs().max()), 1e-6) + 1e-7, k
    for k in tm['params_without_grad']:
        if k in w and not k.startswith('submodules.Superlative.localize_module.'):
            assert w[k].grad is None or float(w[k].grad.abs().max()) == 0.0, k
