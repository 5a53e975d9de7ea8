"""Load the committed golden fixtures (tests/golden/*.npz + *.json, written by make_golden.py)."""
import json
import os

import numpy as np
import torch

from stair_b200 import synthetic as syn

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _gold_from_json(g, text_size):
    out = {}
    for k, v in g.items():
        k = int(k)
        if 'bool' in v:
            out[k] = v['bool']
        elif 'dict' in v:
            out[k] = {n: tuple(iv) for n, iv in v['dict'].items()}
        elif 'classes' in v:
            out[k] = [(n, syn.class_embedding(syn.CLASS_POOL.index(n), text_size)) for n in v['classes']]
        elif 'intervals' in v:
            out[k] = tuple(tuple(iv) for iv in v['intervals'])
        elif 'interval' in v:
            out[k] = tuple(v['interval'])
    return out


def load(name):
    meta = json.load(open(os.path.join(GOLDEN_DIR, name + '.json')))
    npz = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    weights = {k[2:]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith('w/')}
    grads = {k[2:]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith('g/')}
    cfg = meta['config']
    questions = []
    for qi, q in enumerate(meta['questions']):
        data = {
            'question': torch.from_numpy(npz['q%d/question' % qi]),
            'video_features': torch.from_numpy(npz['q%d/video' % qi]),
            'prog_str_to_question_tokens': {int(k): tuple(v) for k, v in q['spans'].items()},
            'nmn_program_list': q['tokens'], 'nmn_program_idx': q['idx_list'],
            'answer': torch.tensor(q['answer']), 'qa_id': 'g-%d' % qi, 'template': q['template'],
            'sg_res_by_step': _gold_from_json(q['gold'], cfg['text_size']),
        }
        ref = {'logits': torch.from_numpy(npz['q%d/logits' % qi]), 'steps': [], 'res_by_step': {}, 'gold_reps': {}}
        for st in q['steps']:
            ref['steps'].append(st['s'] if 's' in st else torch.from_numpy(npz[st['t']]))
        for k, v in q['res_by_step'].items():
            ref['res_by_step'][int(k)] = (v['module'], torch.from_numpy(npz[v['t']]))
        for k, names in q['gold_reps'].items():
            ref['gold_reps'][int(k)] = [(n, torch.from_numpy(npz['q%d/goldrep%d_%d' % (qi, int(k), j)]))
                                        for j, n in enumerate(names)]
        questions.append((data, ref, q))
    return cfg, weights, questions, meta, grads


def load_random(name):
    """(cfg, weights, questions, reference logits [n, A]) of the random-layout fixture (tests/golden/make_random_golden.py): the seeded
    questions are regenerated and their token lists are checked against the ones the reference outputs were recorded for."""
    cfg, weights, _, meta, _ = load(name)
    rmeta = json.load(open(os.path.join(GOLDEN_DIR, 'random_layouts.json')))[name]
    logits = torch.from_numpy(np.load(os.path.join(GOLDEN_DIR, 'random_layouts.npz'))[name + '/logits'])
    qs = syn.make_random_questions(rmeta['n'], cfg['max_video_length'], cfg['video_size'], seed=rmeta['seed'], text_size=cfg['text_size'],
                                   answer_vocab=cfg['answer_vocab_length'])
    assert [d['nmn_program_list'] for d in qs] == rmeta['tokens'], \
        'synthetic.random_layout changed: regenerate tests/golden/random_layouts.* with make_random_golden.py'
    return cfg, weights, qs, logits, meta


def load_random_train(name):
    """(cfg, weights, window questions with gold, {'loss', 'logs', 'params_without_grad'}, reference gradients, meta) of the random-layout
    TRAINING window of tests/golden/make_random_golden.py."""
    cfg, weights, _, meta, _ = load(name)
    tm = json.load(open(os.path.join(GOLDEN_DIR, 'random_layouts.json')))[name]['train']
    npz = np.load(os.path.join(GOLDEN_DIR, 'random_layouts.npz'))
    pre = name + '/g/'
    grads = {k[len(pre):]: torch.from_numpy(npz[k]) for k in npz.files if k.startswith(pre)}
    qs = syn.make_random_questions(tm['n'], cfg['max_video_length'], cfg['video_size'], seed=tm['seed'], text_size=cfg['text_size'],
                                   answer_vocab=cfg['answer_vocab_length'], with_gold=True, object_types=cfg['object_types'])
    assert [d['nmn_program_list'] for d in qs] == tm['tokens'], \
        'synthetic.random_layout changed: regenerate tests/golden/random_layouts.* with make_random_golden.py'
    return cfg, weights, qs, tm, grads, meta
