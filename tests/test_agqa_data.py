"""On-disk formats: the reference's example pickle (utils/agqa_lite.py:122-143), answer vocab, GloVe table and
npy feature directory (video_nmn/dataset.py:31-258) round-tripped through synthetic files into reference-schema data dicts and
``collate``.  Where /root/reference is mounted the items are also compared with the reference's own ``AGQADataset`` on the same files."""
import json
import os
import pickle
import sys
import types

import numpy as np
import pytest
import torch

from stair_b200 import agqa_data as AD, collate, synthetic as syn


def _write_files(tmp, n=24, T_raw=40, D=32, max_len=16, text=12):
    rng = np.random.default_rng(0)
    vocab_words = ['the', 'a', 'dish', 'food', 'holding', 'table', 'person', 'before', 'after', 'cup', 'door', 'opening', 'what', 'is', 'did']
    glove = {w: rng.standard_normal(text) for w in vocab_words}
    pickle.dump(glove, open(tmp / 'glove.pkl', 'wb'))
    os.makedirs(tmp / 'rgb')
    vids = ['V%03d' % i for i in range(6)]
    secs = {}
    for i, v in enumerate(vids):
        np.save(tmp / 'rgb' / (v + '.npy'), rng.standard_normal((T_raw + 2 * i, D)).astype(np.float32))
        secs[v] = 10.0 + i
    json.dump(secs, open(tmp / 'video_secs.json', 'w'))
    names = list(syn.TEMPLATES)
    examples = []
    for i in range(n):
        prog, tokens, idx = syn.TEMPLATES[names[i % len(names)]]
        words = ['what', 'is', 'the', 'person', 'holding', 'before', 'opening', 'the', 'door', 'zzz%d' % i]      # one out-of-vocabulary word
        spans = {}
        for p, tok in enumerate(tokens):
            if tok not in syn.MODULE_ARITY and tok not in syn.WORDS_TO_KEEP:
                s = int(rng.integers(0, len(words) - 2))
                spans[p] = (s, s + 2)
        if i == 4:
            spans[next(iter(spans))] = (None, None)                                                     # dropped from train / valid
        gold = {}
        for p, (tok, ix) in enumerate(zip(tokens, idx)):
            if ix is None or p == 0:
                continue
            if tok == 'Localize':
                gold[ix] = ((3.0, 12.0),) if tokens[p + 2] != 'Array2' else ((1.0, 5.0), (6.0, 20.0))
            elif tok == 'Temporal':
                gold[ix] = (2.0, 9.5)
            elif tok == 'Filter':
                gold[ix] = ['dish', 'cup'] if i % 2 else 'food'
            elif tok == 'Exists':
                gold[ix] = bool(i % 2)
            elif tok == 'FilterFrame':
                gold[ix] = {'dish': (0.0, 4.0)}
        examples.append({'question': ' '.join(words), 'answer': ['yes', 'no', 'dish', 'cup', 'never-seen'][i % 5], 'video_id': vids[i % len(vids)],
                         'program': prog, 'qa_id': 'qa-%d' % i, 'novel_comp': i % 2, 'more_steps': 0, 'nmn_program': list(tokens),
                         'nmn_program_idx': list(idx), 'sg_program': ['x'], 'sg_program_idx': [0], 'sg_res_by_step': gold if i != 7 else None,
                         'nmn_program_span_by_word': spans, 'nmn_program_span_by_char': {}})
    pickle.dump(examples, open(tmp / 'train.pkl', 'wb'))
    return dict(data_filename=str(tmp / 'train.pkl'), vocab_filename=str(tmp / 'vocab.json'), glove_filename=str(tmp / 'glove.pkl'),
                rgb_path=str(tmp / 'rgb'), video_secs_path=str(tmp / 'video_secs.json'), max_video_length=max_len), examples


def test_pickle_npy_glove_vocab_round_trip(tmp_path):
    kw, examples = _write_files(tmp_path)
    ds = AD.AGQADataset('train', seed=0, **kw)
    assert len(ds) == len(examples) - 1                                   # the (None, None) span example is dropped (dataset.py:53-54)
    voc = json.load(open(kw['vocab_filename']))
    assert [voc['id2word'][str(i)] for i in range(4)] == ['yes', 'no', 'before', 'after'] and voc['id2word'][str(len(voc['word2id']) - 1)] == '<UNK>'
    item = ds[0]
    assert set(item) == {'question', 'answer', 'video_features', 'prog_str_to_question_tokens', 'nmn_program_list', 'nmn_program_idx',
                         'sg_program_list', 'sg_res_by_step', 'qa_id', 'question_raw'}
    raw = np.load(os.path.join(kw['rgb_path'], examples[0]['video_id'] + '.npy'))
    want = torch.tensor(raw[::2][:kw['max_video_length']])
    assert torch.equal(item['video_features'], want)                      # every 2nd row, then truncated (dataset.py:138-141)
    assert item['question'].shape == (10, 12) and item['question'].dtype == torch.float32
    # gold intervals are rescaled from 3 fps source frames to the T feature frames (dataset.py:201-211)
    T = want.shape[0]
    src = json.load(open(kw['video_secs_path']))[examples[0]['video_id']] * 3
    for k, v in examples[0]['sg_res_by_step'].items():
        got = item['sg_res_by_step'][k]
        if isinstance(v, tuple) and isinstance(v[0], float):
            assert got == (v[0] / src * T, v[1] / src * T)
        elif isinstance(v, str):
            assert got[0][0] == v and got[0][1].shape[1] == 12
    test = AD.AGQADataset('test', seed=0, **kw)                           # a second open reads the vocab file back
    assert len(test) == len(examples) and 'sg_res_by_step' not in test[0]
    assert int(test[4]['answer']) == test.answer_vocab['word2id']['<UNK>'] or examples[4]['answer'] in test.answer_vocab['word2id']
    assert len(AD.AGQADataset('train', novel_comp=1, seed=0, **kw)) < len(ds)
    # the dicts collate (layout compile + staging) like synthetic ones; T is uniform here because every file has >= 2 * max_len rows
    batch = collate([ds[i] for i in range(len(ds))], video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
    assert batch.B == len(ds) and batch.T == kw['max_video_length'] and batch.answer.shape == (len(ds),)
    assert AD.frame_interval_change_fps((3.0, 6.0), 30, 10) == (1.0, 2.0)


def test_text_glove_and_missing_h5py(tmp_path):
    with open(tmp_path / 'glove.txt', 'w') as f:
        f.write('2 3\nthe 0.1 0.2 0.3\ncup -1 0 1\n')
    g = AD.load_glove(str(tmp_path / 'glove.txt'))
    assert set(g) == {'the', 'cup'} and np.allclose(g['cup'], [-1, 0, 1])
    try:
        import h5py  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError):
            AD.load_h5_features(str(tmp_path / 'app.h5'), None, {}, [], 8)


@pytest.mark.skipif(not os.path.isdir('/root/reference'), reason='reference not mounted')
def test_items_equal_the_reference_dataset_on_the_same_files(tmp_path):
    """The unmodified reference ``AGQADataset`` (video_nmn/dataset.py:31-258) reads the same synthetic files: every item field is equal
    (random vectors of out-of-vocabulary words excepted — the reference draws them unseeded)."""
    kw, examples = _write_files(tmp_path)
    sys.modules.setdefault('h5py', types.ModuleType('h5py'))
    nltk, corpus, tok = (types.ModuleType(n) for n in ('nltk', 'nltk.corpus', 'nltk.tokenize'))
    corpus.stopwords = type('SW', (), {'words': staticmethod(lambda lang: [])})()
    tok.word_tokenize = lambda s: s.split()
    nltk.corpus, nltk.tokenize = corpus, tok
    for k, v in (('nltk', nltk), ('nltk.corpus', corpus), ('nltk.tokenize', tok)):
        sys.modules.setdefault(k, v)
    sys.path.insert(0, '/root/reference')
    try:
        from video_nmn.dataset import AGQADataset as RefDataset
    finally:
        sys.path.remove('/root/reference')
    args = types.SimpleNamespace(debug=False, rgb_path=kw['rgb_path'], flow_path=None, video_secs_path=kw['video_secs_path'], str2num_path=None,
                                 train_filename=kw['data_filename'], valid_filename=kw['data_filename'], test_filename=kw['data_filename'],
                                 novel_comp=None, more_steps=None, vocab_filename=kw['vocab_filename'], max_video_length=kw['max_video_length'],
                                 glove_filename=kw['glove_filename'], shuffle_video=False)
    import contextlib
    import io
    mine = AD.AGQADataset('train', seed=0, **kw)
    with contextlib.redirect_stdout(io.StringIO()):
        ref = RefDataset(args, 'train')
    assert len(ref) == len(mine)
    for i in range(len(ref)):
        a, b = ref[i], mine[i]
        assert set(a) == set(b)
        assert torch.equal(a['video_features'], b['video_features']) and int(a['answer']) == int(b['answer'])
        assert a['nmn_program_list'] == b['nmn_program_list'] and a['prog_str_to_question_tokens'] == b['prog_str_to_question_tokens']
        assert torch.equal(a['question'][:-1], b['question'][:-1])        # the last word is out of vocabulary (random in both)
        assert set(a['sg_res_by_step']) == set(b['sg_res_by_step'])
        for k, va in a['sg_res_by_step'].items():
            vb = b['sg_res_by_step'][k]
            if isinstance(va, list) and va and isinstance(va[0][1], torch.Tensor):
                assert [n for n, _ in va] == [n for n, _ in vb] and all(torch.equal(x[1], y[1]) for x, y in zip(va, vb))
            else:
                assert va == vb
