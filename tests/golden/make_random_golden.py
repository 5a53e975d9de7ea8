"""Golden outputs of the UNMODIFIED reference on RANDOM well-typed layouts (``stair_b200.synthetic.random_layout``): beyond the probed AGQA
templates, any composition of the 18 operators the reference interpreter accepts.  Build container only:

    python tests/golden/make_random_golden.py

For the two small configurations of make_golden.py (their weights are reused from rx_small.npz / i3d_small.npz) it regenerates the seeded
questions, runs the reference ``VideoNMN.forward`` (video_nmn/module_net.py:65-145) on each and stores the logits together with the token
lists (so that a change of the generator is detected, not silently compared against stale outputs) in ``random_layouts.npz`` / ``.json``.
"""
import json
import os
import sys
import types

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

from stair_b200 import synthetic as syn  # noqa: E402
from tests import golden_util as gu  # noqa: E402
from tests.golden import make_golden as mg  # noqa: E402  (reference_window / make_criterion: the reference criterion + window loop)

N_LAYOUTS = 150
SEEDS = {'rx_small': 20250, 'i3d_small': 20251}
N_TRAIN = 40                                  # questions of the random-layout training window (with intermediate supervision)
TRAIN_SEEDS = {'rx_small': 30250, 'i3d_small': 30251}


def main():
    store, meta = {}, {}
    for name, seed in SEEDS.items():
        cfg, weights, _, _, _ = gu.load(name)
        model = VideoNMN(cfg, pretrain_modules=set(syn.PRETRAIN_MODULES))
        model.load_state_dict(weights)
        model.eval()
        qs = syn.make_random_questions(N_LAYOUTS, cfg['max_video_length'], cfg['video_size'], seed=seed, text_size=cfg['text_size'],
                                       answer_vocab=cfg['answer_vocab_length'])
        logits = []
        with torch.no_grad():
            for d in qs:
                out = model(d, return_res_by_step=False, test_mode=True)
                logits.append(out['logits'].reshape(-1).numpy())
        store[name + '/logits'] = np.stack(logits).astype(np.float32)
        meta[name] = {'seed': seed, 'n': N_LAYOUTS, 'tokens': [d['nmn_program_list'] for d in qs],
                      'modules': [sum(1 for t in d['nmn_program_list'] if t in syn.MODULE_ARITY) for d in qs]}
        # ---- one accumulation window of random layouts with gold for every supervisable non-root module: reference losses + gradients ----
        tq = syn.make_random_questions(N_TRAIN, cfg['max_video_length'], cfg['video_size'], seed=TRAIN_SEEDS[name], text_size=cfg['text_size'],
                                       answer_vocab=cfg['answer_vocab_length'], with_gold=True, object_types=cfg['object_types'])
        crit, args = mg.make_criterion(cfg['object_types'], len(tq))
        model.train()                                                   # dropout is 0 in these configs; train() as in train_module.py:341
        loss, logs, _ = mg.reference_window(model, crit, args, tq)
        model.eval()
        seen = set()
        for k, p in model.named_parameters():
            if id(p) in seen:
                continue
            seen.add(id(p))
            if p.grad is not None:
                store[name + '/g/' + k] = p.grad.numpy().copy()
        meta[name]['train'] = {'seed': TRAIN_SEEDS[name], 'n': N_TRAIN, 'loss': float(loss), 'logs': logs,
                               'tokens': [d['nmn_program_list'] for d in tq],
                               'params_without_grad': [k for k, p in model.named_parameters() if p.grad is None]}
        print(name, 'training window of', N_TRAIN, 'random layouts: loss %.6f' % float(loss), {m: len(v) for m, v in logs.items() if v})
        print(name, 'layouts', N_LAYOUTS, 'modules per layout: max %d mean %.1f' % (max(meta[name]['modules']), np.mean(meta[name]['modules'])),
              'distinct operators', len({t for d in qs for t in d['nmn_program_list'] if t in syn.MODULE_ARITY}))
    np.savez_compressed(os.path.join(HERE, 'random_layouts.npz'), **store)
    json.dump(meta, open(os.path.join(HERE, 'random_layouts.json'), 'w'))


if __name__ == '__main__':
    main()
