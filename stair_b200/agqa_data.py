"""On-disk formats either side of the hot path: the reference's example pickle, answer vocabulary, GloVe table
and video-feature stores, read into the reference-schema ``data`` dicts that ``stair_b200.collate`` / ``VideoNMN.forward`` consume.

Reference: ``video_nmn/dataset.py:31-258`` (``AGQADataset``) and ``utils/agqa_lite.py:122-143`` (``convert_``: the pickle's example
schema).  Same file formats, same field names, same filtering and conversion rules — re-implemented as a plain keyword-argument class
(the reference threads an argparse namespace through) so that it can be tested on synthetic files; the arithmetic-free parts
(filtering, vocab, fps rescaling of gold intervals, phrase embedding) are host code like in the reference.

  example pickle   list of dicts with ``question, answer, video_id, qa_id, nmn_program, nmn_program_idx, nmn_program_span_by_word,
                   sg_program, sg_res_by_step`` (+ ``novel_comp``, ``more_steps``)                       agqa_lite.py:122-143
  answer vocab     json ``{'word2id': {...}, 'id2word': {...}}``, ids 0-3 = yes / no / before / after, last = ``<UNK>``  dataset.py:70-98
  GloVe            a pickle ``{word: np.ndarray}`` or the text format ``word v0 v1 ...`` with a ``count dim`` header      dataset.py:235-246
  video features   a directory of ``<video_id>.npy`` [n, D] (I3D): every 2nd row, then ``[:max_video_length]``           dataset.py:134-143
                   or h5 files ``ids`` + ``resnet_features`` [n, clips, frames, D] (mean over frames) and ``resnext_features``
                   [n, clips, D] (concatenated)                                                          dataset.py:145-172
                   (h5 needs ``h5py``, which this image does not ship: the reader imports it lazily and says so)
"""
from __future__ import annotations

import json
import os
import pickle
import random

import numpy as np
import torch

WORDS_TO_KEEP = ['forward', 'backward', 'while', 'between', 'before', 'after', 'max', 'min', 'start', 'end', 'video']     # dataset.py:23


def frame_interval_change_fps(interval, src_length, tgt_length):
    """dataset.py:261-264 — rescale a (start, end) interval from source frames to the T feature frames."""
    return (interval[0] / src_length * tgt_length, interval[1] / src_length * tgt_length)


def load_npy_features(directory, used_video_ids, max_video_length):
    """dataset.py:134-143 — ``<video_id>.npy`` -> rows 0, 2, 4, ... -> ``[:max_video_length]`` -> squeezed fp32/fp64 tensor as stored."""
    used = set(used_video_ids)
    feats = {}
    for fname in os.listdir(directory):
        video_id = fname.split('.')[0]
        if video_id in used:
            arr = np.load(os.path.join(directory, fname))
            arr = arr[np.arange(start=0, stop=arr.shape[0], step=2), :]
            if arr.shape[0] > max_video_length:
                arr = arr[:max_video_length]
            feats[video_id] = torch.tensor(arr).squeeze()
    return feats


def load_npy_features_device(directory, used_video_ids, max_video_length, device='cuda', out_dtype=torch.bfloat16):
    """The same reader with the reduction on the GPU (``stair_ingest_subsample``, csrc/ingest.cu): raw arrays of equal shape are uploaded
    as one batch and subsampled / truncated / converted there.  Returns ``{video_id: [T, D] tensor on device}``."""
    from . import ingest
    used = set(used_video_ids)
    by_shape = {}
    for fname in sorted(os.listdir(directory)):
        video_id = fname.split('.')[0]
        if video_id in used:
            arr = np.load(os.path.join(directory, fname))
            by_shape.setdefault(arr.shape, []).append((video_id, arr))
    feats = {}
    for shape, items in by_shape.items():
        raw = torch.from_numpy(np.stack([a for _, a in items]).astype(np.float32)).to(device)
        out = ingest.subsample(raw.reshape(len(items), shape[0], -1), max_video_length, step=2, out_dtype=out_dtype)
        for i, (video_id, _) in enumerate(items):
            feats[video_id] = out[i]
    return feats


def load_h5_features(appearance_path, motion_path, str2num, used_video_ids, max_video_length):
    """dataset.py:145-172 — TGIF-QA style h5 stores: appearance ``resnet_features[id]`` [clips, frames, D] -> ``[:max_video_length]`` ->
    mean over frames; motion ``resnext_features[id]`` [clips, D] -> ``[:max_video_length]`` -> concatenated."""
    try:
        import h5py
    except ImportError as e:                                              # not installed in this image; never silently skipped
        raise ImportError('reading %s needs h5py (video_nmn/dataset.py:8,146), which is not installed' % appearance_path) from e
    used = set(used_video_ids)
    feats = {}
    f_app = h5py.File(appearance_path, 'r')
    id2id = {id_: i for i, id_ in enumerate(f_app['ids'][()])}
    for video_id, id_ in str2num.items():
        if video_id in used:
            v = f_app['resnet_features'][id2id[id_]]
            if v.shape[0] > max_video_length:
                v = v[:max_video_length]
            feats[video_id] = torch.tensor(v).mean(dim=1)
    if motion_path is not None and os.path.isfile(motion_path):
        f_mot = h5py.File(motion_path, 'r')
        id2id = {id_: i for i, id_ in enumerate(f_mot['ids'][()])}
        for video_id, id_ in str2num.items():
            if video_id in used:
                v = f_mot['resnext_features'][id2id[id_]]
                if v.shape[0] > max_video_length:
                    v = v[:max_video_length]
                feats[video_id] = torch.cat([feats[video_id], torch.tensor(v)], dim=-1)
    return feats


def load_glove(glove_filename):
    """dataset.py:235-246 — ``{word: vector}`` from a pickle or from the text format (first line ``count dim``)."""
    if glove_filename.endswith('.pkl'):
        return pickle.load(open(glove_filename, 'rb'))
    table = {}
    for i, line in enumerate(open(glove_filename)):
        if i == 0:
            continue                                                       # header: count dim
        parts = line.rstrip('\n').split(' ')
        table[parts[0]] = np.array(list(map(float, parts[1:])))
    return table


def build_answer_vocab(examples):
    """dataset.py:72-85 — yes / no / before / after first, then answers by descending frequency, ``<UNK>`` last."""
    from collections import Counter
    counter = Counter(d['answer'] for d in examples)
    vocab = ['yes', 'no', 'before', 'after']
    given = set(vocab)
    for ans, _ in sorted(counter.items(), key=lambda x: -x[1]):
        if ans not in given:
            vocab.append(ans)
    vocab.append('<UNK>')
    return {'word2id': {w: i for i, w in enumerate(vocab)}, 'id2word': {i: w for i, w in enumerate(vocab)}}


class AGQADataset(torch.utils.data.Dataset):
    """``video_nmn.dataset.AGQADataset`` over the same files; ``__getitem__`` returns the same ``data`` dict (dataset.py:189-233).

    >>> ds = AGQADataset('train', data_filename='train.pkl', vocab_filename='vocab.json', glove_filename='glove.pkl',
    ...                  rgb_path='i3d_rgb/', video_secs_path='video_secs.json', max_video_length=64)
    >>> model(stair_b200.collate([ds[i] for i in range(4096)], pin_memory=True))
    """

    def __init__(self, split, data_filename, vocab_filename, glove_filename, rgb_path, video_secs_path=None, flow_path=None,
                 str2num_path=None, max_video_length=150, novel_comp=None, more_steps=None, debug=False, shuffle_video=False,
                 tokenizer=None, seed=None):
        assert max_video_length >= 2, 'why do you set max video length so small?'                       # dataset.py:101
        self.split, self.max_video_length, self.debug = split, max_video_length, debug
        self.tokenizer = tokenizer or (lambda s: s.split())             # the reference uses nltk.word_tokenize (absent here): inject it if needed
        self.rng = np.random.default_rng(seed)
        self.video_secs = json.load(open(video_secs_path)) if video_secs_path else {}
        data_ = pickle.load(open(data_filename, 'rb'))
        if split in ('train', 'valid'):                                   # dataset.py:48-56: drop examples with unaligned program words
            self.data = []
            for d in data_:
                if d['sg_res_by_step'] is None:
                    d['sg_res_by_step'] = {}
                if (None, None) in d['nmn_program_span_by_word'].values():
                    continue
                self.data.append(d)
        else:
            self.data = data_
        if novel_comp is not None:                                        # dataset.py:60-65 generalisation splits
            self.data = [d for d in self.data if d['novel_comp'] == novel_comp]
        if more_steps is not None:
            self.data = [d for d in self.data if d['more_steps'] == more_steps]
        if debug:
            self.data = random.sample(self.data, min(256, len(self.data)))
        if not os.path.exists(vocab_filename):                            # dataset.py:70-98
            self.answer_vocab = build_answer_vocab(self.data)
            json.dump(self.answer_vocab, open(vocab_filename, 'w'))
        else:
            self.answer_vocab = json.load(open(vocab_filename))
            self.answer_vocab['id2word'] = {int(k): v for k, v in self.answer_vocab['id2word'].items()}
            assert len(self.answer_vocab['id2word']) == len(self.answer_vocab['word2id'])
            assert [self.answer_vocab['id2word'][i] for i in range(4)] == ['yes', 'no', 'before', 'after']
        self.word_embeddings = load_glove(glove_filename)
        self.word_embedding_size = np.asarray(next(iter(self.word_embeddings.values()))).size
        used = list(set(d['video_id'] for d in self.data))
        self.shuffle_video = shuffle_video
        if shuffle_video:                                                 # dataset.py:108-114 ablation
            ids = list(range(len(used)))
            random.shuffle(ids)
            self.shuffle_video_mapping = {used[i]: used[ids[i]] for i in range(len(used))}
        if os.path.isdir(rgb_path):
            self.video_feats = load_npy_features(rgb_path, used, max_video_length)
        elif os.path.isfile(rgb_path):
            str2num = json.load(open(str2num_path))
            self.video_feats = load_h5_features(rgb_path, flow_path, str2num, used, max_video_length)
        else:
            raise ValueError('appearance path not given!')               # dataset.py:158

    def __len__(self):
        return len(self.data)

    def answer_vocab_length(self):
        return len(self.answer_vocab['word2id'])

    def embed_sent(self, sent):
        """dataset.py:248-255 — lower-cased words -> GloVe rows; unknown words get a fresh uniform random vector, like the reference."""
        words = self.tokenizer(sent.lower()) if isinstance(sent, str) else [s.lower() for s in sent]
        emb = [self.word_embeddings[w] if w in self.word_embeddings else self.rng.random(self.word_embedding_size) for w in words]
        return torch.tensor(np.asarray(emb), dtype=torch.float32)

    def __getitem__(self, idx):
        d = self.data[idx]
        video_id = self.shuffle_video_mapping[d['video_id']] if self.shuffle_video else d['video_id']
        video_features = self.video_feats[video_id]
        T = video_features.size(0)
        w2i = self.answer_vocab['word2id']
        ret = {'question': self.embed_sent(d['question']), 'answer': torch.tensor(w2i.get(d['answer'], w2i.get('<UNK>'))),
               'video_features': video_features, 'prog_str_to_question_tokens': d['nmn_program_span_by_word'],
               'nmn_program_list': d['nmn_program'], 'nmn_program_idx': d['nmn_program_idx'], 'qa_id': d['qa_id'], 'question_raw': d['question']}
        if self.split == 'test':                                          # dataset.py:189-197
            return ret
        src_length = self.video_secs[video_id] * 3                        # dataset.py:201: gold intervals are in 3 fps source frames
        gold = {}
        for key, value in d['sg_res_by_step'].items():                    # dataset.py:203-211
            if isinstance(value, (tuple, list)) and len(value) >= 1:
                if isinstance(value[0], float):
                    value = frame_interval_change_fps(value, src_length, T)
                if isinstance(value[0], tuple) and isinstance(value[0][0], float):
                    value = tuple(frame_interval_change_fps(v, src_length, T) for v in value)
            if isinstance(value, dict) and isinstance(list(value.values())[0], tuple) and isinstance(list(value.values())[0][0], float):
                value = {k: frame_interval_change_fps(v, src_length, T) for k, v in value.items()}
            if isinstance(value, str):                                    # dataset.py:213-221: class names -> (name, phrase embedding)
                value = [(value, self.embed_sent(value))]
            elif isinstance(value, list) and len(value) and isinstance(value[0], str):
                value = [(v, self.embed_sent(v)) for v in value]
            gold[key] = value
        ret.update({'sg_program_list': d.get('sg_program'), 'sg_res_by_step': gold})
        return ret


def collate_fn(examples):
    """Batched replacement of dataset.py:463-464 (``examples[0]``): one pinned ``NMNBatch`` per DataLoader batch."""
    from .layout import collate
    return collate(list(examples), pin_memory=torch.cuda.is_available())
