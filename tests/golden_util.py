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
