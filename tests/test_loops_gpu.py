"""GPU: the batched evaluate / train drivers and the checkpoint round trip.

evaluate: accuracy == the oracle's per-question argmax accuracy with the reference's <UNK> rule (evaluate.py:41-48).
train: a few windows of Adam steps reduce the window loss; save -> load -> same logits; resumed optimizer state continues
identically to an uninterrupted run."""
import os

import pytest
import torch

from oracle import nmn_oracle as orc
from stair_b200 import VideoNMN, synthetic as syn
from stair_b200 import loops

pytestmark = pytest.mark.gpu

CFG = dict(T=8, V=128, hidden=64, object_types=16)


def _model(seed=0, precision='fp32'):
    cfg = syn.model_config(**CFG)
    torch.manual_seed(seed)
    return cfg, VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision=precision)


def test_evaluate_matches_reference_accuracy_rule():
    cfg, model = _model()
    weights = {k: v.detach().clone() for k, v in model.state_dict().items()}
    model = model.cuda().eval()
    oracle = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES)
    qs = syn.make_questions(40, 8, 128, seed=21, templates=list(syn.ALL_TEMPLATES), answer_vocab=172)
    with torch.no_grad():
        preds = [int(oracle(d, return_res_by_step=False, test_mode=True)['logits'].argmax()) for d in qs]
    for i, d in enumerate(qs):                                   # make about half of the golds "correct", some of them <UNK>
        if i % 2 == 0:
            d['answer'] = torch.tensor(preds[i])
    unk = preds[0]
    want = sum(int(p == int(d['answer']) and int(d['answer']) != unk) for p, d in zip(preds, qs)) / len(qs)
    acc, out = loops.evaluate([qs[:25], qs[25:]], model, unk_token_id=unk)
    assert abs(acc - want) < 1e-12 and len(out['preds']) == 40 and out['qa_ids'][0] == qs[0]['qa_id']
    acc2, _ = loops.evaluate([qs], model, unk_token_id=unk, pipelined_chunks=3)
    assert abs(acc2 - want) < 1e-12


def test_train_reduces_loss_and_checkpoint_round_trip(tmp_path):
    cfg, model = _model(1)
    model = model.cuda()
    qs = syn.make_questions(32, 8, 128, seed=5, templates=list(syn.ALL_TEMPLATES), with_gold=True, object_types=16)
    windows = [qs] * 12
    state = loops.train(windows, model, lr=2e-3, scheduler_kwargs=dict(total_iters=100), report_interval=4)
    totals = [sum(l.values()) for l in state['losses']]
    assert len(totals) == 12 and totals[-1] < 0.9 * totals[0], totals
    assert state['global_steps'] == 12 and abs(state['scheduler'].get_last_lr()[0] - 2e-3 * (1.0 - 0.9 * 12 / 100)) < 1e-9
    ckpt = str(tmp_path / 'ckpt')
    loops.save_checkpoint(ckpt, model, state)
    assert sorted(os.listdir(ckpt)) == ['config.json', 'pytorch_model.bin', 'trainer_state.pt']
    sd = torch.load(os.path.join(ckpt, 'pytorch_model.bin'))
    assert len(sd) == 119                                        # the reference's state_dict keys
    model2, state2 = loops.load_checkpoint(ckpt, VideoNMN, precision='fp32', pretrain_modules=syn.PRETRAIN_MODULES, with_trainer_state=True, lr=2e-3)
    model.eval(); model2.eval()
    a = model(qs, return_res_by_step=False, test_mode=True)['logits']
    b = model2(qs, return_res_by_step=False, test_mode=True)['logits']
    assert torch.equal(a, b)
    # resume == uninterrupted
    loops.train([qs] * 2, model, state=state)
    loops.train([qs] * 2, model2, state=state2)
    model.eval(); model2.eval()
    a = model(qs, return_res_by_step=False, test_mode=True)['logits']
    b = model2(qs, return_res_by_step=False, test_mode=True)['logits']
    torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-5)
