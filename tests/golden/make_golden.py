"""Generate the committed golden fixtures by running the UNMODIFIED reference (``/root/reference``).

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

For each config it (1) builds the reference ``VideoNMN`` (video_nmn/module_net.py) under
``torch.manual_seed(0)``, (2) runs it on seeded synthetic questions from ``stair_b200.synthetic``
(one per layout template), (3) runs the reference ``CriterionByModule`` + the gradient-accumulation-window
logic of ``train_module.py:341-412`` and ``backward()``, and stores inputs, weights, every intermediate,
``res_by_step``, logits, losses and gradients as ``<name>.npz`` + ``<name>.json``.
It also records the reference's layout helpers (``parse_program``, ``get_childrens_and_parents``,
``stat_module_levels``, ``program_is_valid``) on every template.

The h5py / nltk stand-ins are the SURVEY.md Appendix A recipe: the hot path never calls them, they are only
needed so that ``video_nmn.dataset`` (imported for one constant, module_net.py:8) can be imported.
"""
import io
import json
import os
import sys
import tempfile
import types
import contextlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

sys.modules['h5py'] = types.ModuleType('h5py')
_nltk, _corpus, _tok = (types.ModuleType(n) for n in ('nltk', 'nltk.corpus', 'nltk.tokenize'))
_corpus.stopwords = type('SW', (), {'words': staticmethod(lambda lang: [])})()
_tok.word_tokenize = lambda s: s.split()
_nltk.corpus, _nltk.tokenize = _corpus, _tok
sys.modules.update({'nltk': _nltk, 'nltk.corpus': _corpus, 'nltk.tokenize': _tok})

import torch  # noqa: E402
from video_nmn.module_net import VideoNMN  # noqa: E402  (reference)
from utils.program_parser import (parse_program, get_childrens_and_parents, stat_module_levels,  # noqa: E402
                                  program_is_valid, nary_mappings)
import train_module as ref_train  # noqa: E402  (reference)

from stair_b200 import synthetic as syn  # noqa: E402

CONFIGS = {
    # linear-mode Temporal (T <= 32), RX-like
    'rx_small': dict(T=8, V=128, hidden=64, text_size=300, answer_vocab=172, object_types=16, seed=4321),
    # conv-mode Temporal (T > 32): k = 16,16,33 (modules.py:255-266), I3D-like
    'i3d_small': dict(T=64, V=128, hidden=64, text_size=300, answer_vocab=172, object_types=16, seed=8765),
}


def tensor_items(prefix, value, store):
    """Flatten result_of_each_step / res_by_step values into the npz store; return a JSON descriptor."""
    if isinstance(value, torch.Tensor):
        store[prefix] = value.detach().cpu().numpy()
        return {'t': prefix}
    if isinstance(value, str):
        return {'s': value}
    if value is None:
        return None
    raise TypeError(type(value))


def gold_to_json(gold):
    out = {}
    for k, v in gold.items():
        if isinstance(v, bool):
            out[str(k)] = {'bool': v}
        elif isinstance(v, dict):
            out[str(k)] = {'dict': {n: list(iv) for n, iv in v.items()}}
        elif isinstance(v, list):
            out[str(k)] = {'classes': [n for n, _ in v]}
        elif isinstance(v, tuple) and isinstance(v[0], tuple):
            out[str(k)] = {'intervals': [list(iv) for iv in v]}
        elif isinstance(v, tuple):
            out[str(k)] = {'interval': list(v)}
        else:
            raise TypeError(type(v))
    return out


def reference_window(model, crit, args, batch):
    """One gradient-accumulation window with the reference criterion and the loop semantics of train_module.py:341-412 (losses scaled by
    weight / gradient_accumulation, window-level contrastive negatives) followed by ``backward()``.  Returns (loss, logs, FilterFrame losses)."""
    model.zero_grad()
    batch_loss, logs = 0., {m: [] for m in crit.criterions}
    class_reps, neg_reps = {}, {}
    ff_losses = []
    for it, data in enumerate(batch):
        out = model(data, return_res_by_step=True)
        gold_by_step = out['sg_res_by_step']
        example_loss = 0.
        for step, (module, res) in out['res_by_step'].items():       # train_module.py:350-373
            if module == 'FilterFrame' and step in gold_by_step:
                ff_losses.append([it, int(step), float(crit(module, res, gold_by_step[step]))])
            if step not in gold_by_step or module in args.modules_no_intermediate_train or module not in crit.criterions:
                continue
            sg = gold_by_step[step]
            if sg is None:
                continue
            if module in ['Filter', 'Superlative', 'ToAction']:
                for cname, crep in sg:
                    class_reps.setdefault(cname, []).append((it, module, res))
                    neg_reps[cname] = crep
            else:
                loss = crit(module, res, sg)
                logs[module].append(float(loss))
                example_loss = example_loss + loss * args.module_loss_weight / args.gradient_accumulation
        loss = crit('decoder', out['logits'], data['answer'])           # :376-380
        logs['decoder'].append(float(loss))
        batch_loss = batch_loss + example_loss + loss * args.decoder_loss_weight / args.gradient_accumulation
    for cname, vals in class_reps.items():                              # :388-406
        for _, module, res in vals:
            pos = neg_reps[cname]
            neg = [v for k, v in neg_reps.items() if k != cname]
            gold = torch.cat([pos.unsqueeze(0), torch.stack(neg)]) if neg else pos.unsqueeze(0)
            loss = crit(module, res, gold)
            logs[module].append(float(loss))
            batch_loss = batch_loss + loss * args.module_loss_weight / args.gradient_accumulation
    batch_loss.backward()
    return batch_loss, logs, ff_losses


def make_criterion(object_types, ga):
    with tempfile.NamedTemporaryFile('w', suffix='.json', delete=False) as f:
        json.dump({'obj_%d' % i: i for i in range(object_types)}, f)
    args = types.SimpleNamespace(word2id_filename=f.name, module_loss_weight=1.0, decoder_loss_weight=1.0,
                                 gradient_accumulation=ga, modules_no_intermediate_train=['FilterFrame'])
    with contextlib.redirect_stdout(io.StringIO()):
        crit = ref_train.CriterionByModule(args)
    os.unlink(f.name)
    return crit, args


def run_config(name, c):
    cfg = syn.model_config(T=c['T'], V=c['V'], hidden=c['hidden'], text_size=c['text_size'],
                           answer_vocab=c['answer_vocab'], object_types=c['object_types'])
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        model = VideoNMN(cfg, pretrain_modules=set(syn.PRETRAIN_MODULES))
    model.eval()
    store, meta = {}, {'config': cfg, 'questions': [], 'pretrain_modules': sorted(syn.PRETRAIN_MODULES)}
    for k, v in model.state_dict().items():
        store['w/' + k] = v.detach().numpy()

    rng = np.random.default_rng(c['seed'])
    names = list(syn.ALL_TEMPLATES.keys())
    batch = [syn.make_question(rng, t, c['T'], c['V'], c['text_size'], c['answer_vocab'], with_gold=True,
                               object_types=c['object_types'], qa_id='g-%d' % i) for i, t in enumerate(names)]

    # ---- forward (audit mode: every intermediate) -------------------------------------------------
    for qi, data in enumerate(batch):
        with torch.no_grad():
            out = model(data, return_res_by_step=True, return_result_of_each_step=True, test_mode=False)
            out_raw = model(data, return_res_by_step=False, test_mode=True)
        assert torch.equal(out['logits'], out_raw['logits'])
        tokens = data['nmn_program_list']
        ch, pa = get_childrens_and_parents(tokens)
        q = {
            'template': data['template'], 'tokens': tokens, 'idx_list': data['nmn_program_idx'],
            'spans': {str(k): list(v) for k, v in data['prog_str_to_question_tokens'].items()},
            'answer': int(data['answer']), 'children': ch, 'parents': pa, 'levels': stat_module_levels(tokens),
            'valid': program_is_valid(tokens), 'gold': gold_to_json(data['sg_res_by_step']),
        }
        store['q%d/question' % qi] = data['question'].numpy()
        store['q%d/video' % qi] = data['video_features'].numpy()
        store['q%d/logits' % qi] = out['logits'].numpy()
        q['steps'] = [tensor_items('q%d/step%d' % (qi, j), res, store)
                      for j, (_, res) in enumerate(out['result_of_each_step'])]
        q['res_by_step'] = {str(k): {'module': m, **tensor_items('q%d/res%d' % (qi, k), r, store)}
                            for k, (m, r) in out['res_by_step'].items()}
        q['gold_reps'] = {}
        for k, v in out['sg_res_by_step'].items():
            if isinstance(v, list) and v and isinstance(v[0][1], torch.Tensor):
                for j, (cn, rep) in enumerate(v):
                    store['q%d/goldrep%d_%d' % (qi, k, j)] = rep.numpy()
                q['gold_reps'][str(k)] = [cn for cn, _ in v]
        meta['questions'].append(q)

    # ---- one gradient-accumulation window, reference criterion + loop semantics ---------------------
    crit, args = make_criterion(c['object_types'], len(batch))
    batch_loss, logs, ff_losses = reference_window(model, crit, args, batch)
    meta['window'] = {'loss': float(batch_loss), 'logs': logs, 'filterframe_losses': ff_losses}
    seen = set()
    for k, p in model.named_parameters():
        if id(p) in seen:
            continue
        seen.add(id(p))
        if p.grad is not None:
            store['g/' + k] = p.grad.numpy()
    meta['params_without_grad'] = [k for k, p in model.named_parameters() if p.grad is None]

    # span_to_attention known answers (train_module.py:67-81)
    meta['span_to_attention'] = [
        {'gold': [s, e], 'T': T, 'out': crit.span_to_attention((s, e), T).tolist()}
        for (s, e, T) in [(3.4, 6.1, 8), (0.0, 8.0, 8), (2.2, 2.7, 8), (7.9, 8.0, 8), (0.2, 5.8, 8), (10.5, 40.25, 64), (5.0, 5.0, 8)]]

    np.savez_compressed(os.path.join(HERE, name + '.npz'), **store)
    with open(os.path.join(HERE, name + '.json'), 'w') as fh:
        json.dump(meta, fh)
    print(name, 'questions', len(batch), 'window loss', float(batch_loss),
          'npz MB', os.path.getsize(os.path.join(HERE, name + '.npz')) / 1e6)


def layout_fixture():
    out = {'nary': dict(nary_mappings), 'templates': {}}
    for tname, (prog, tokens, idx) in syn.ALL_TEMPLATES.items():
        entry = {}
        if prog is not None:
            pl, more = parse_program(prog)
            entry.update({'program': prog, 'tokens': pl, 'idx_list': more['idx_list']})
            assert pl == tokens and more['idx_list'] == idx, tname
        else:
            entry.update({'program': None, 'tokens': tokens, 'idx_list': idx})
        ch, pa = get_childrens_and_parents(entry['tokens'])
        entry.update({'children': ch, 'parents': pa, 'levels': stat_module_levels(entry['tokens']),
                      'valid': program_is_valid(entry['tokens'])})
        out['templates'][tname] = entry
    # some invalid programs for program_is_valid
    out['invalid'] = [['Exists', 'table'], ['Filter', 'video', 'objects', 'video'], ['And'], []]
    out['invalid_results'] = [program_is_valid(p) for p in out['invalid']]
    with open(os.path.join(HERE, 'layouts.json'), 'w') as fh:
        json.dump(out, fh)
    print('layouts', len(out['templates']))


def relate_scan_fixture():
    """TemporalModule.relate_ (video_nmn/modules.py:290-308) of the unmodified reference on seeded random maps: the cumsum
    before / after / between masks the north star names (dead code in the reference forward, callable on the module object).
    Inputs contain negative values so the ReLU in front of the scans is exercised; 'while' returns its input untouched."""
    store = {}
    rng = np.random.default_rng(2468)
    for T in (8, 64):
        cfg = syn.model_config(T=T, V=32, hidden=16)
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):
            temporal = VideoNMN(cfg).submodules['Temporal']
        for mode in ('while', 'before', 'after', 'between'):
            K = 2 if mode == 'between' else 1
            att = torch.from_numpy(rng.uniform(-0.4, 0.98, size=(24, K, T)).astype(np.float32))
            with torch.no_grad():
                out = torch.stack([temporal.relate_(a, mode) for a in att])
            assert out.shape == (24, T), (mode, out.shape)
            store['T%d/%s/att' % (T, mode)] = att.numpy()
            store['T%d/%s/out' % (T, mode)] = out.numpy()
    np.savez_compressed(os.path.join(HERE, 'relate_scan.npz'), **store)
    print('relate_scan', len(store) // 2, 'cases')


if __name__ == '__main__':
    layout_fixture()
    relate_scan_fixture()
    for n, c in CONFIGS.items():
        run_config(n, c)
