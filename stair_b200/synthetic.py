"""Synthetic AGQA2-shaped questions for the NMN hot path (tests, bench, golden fixtures).

The reference trains/evaluates on AGQA2 pickles + TGIF-QA style h5 features that are not on
the box, so every test and benchmark in this repo runs on seeded synthetic data of the same
shape.  The *layouts* are real: ``TEMPLATES`` holds the NMN prefix-token lists
(and original-program index lists) that the reference ``utils/program_parser.py:28-170``
``parse_program`` emits for ten AGQA program strings; ``tests/test_layout_host.py`` re-derives
them from the reference whenever ``/root/reference`` is mounted.

Data dict schema follows ``video_nmn/dataset.py:189-233`` (producer) /
``video_nmn/module_net.py:69-71`` (consumer).
"""
from __future__ import annotations

import numpy as np
import torch

# name -> (AGQA program string, nmn token list, idx_list) ; token lists probed from the reference.
TEMPLATES = {
    'exists_while': (
        "Exists(table, Iterate(Localize(while, [sitting on a bed]), Filter(frame, [objects])))",
        ['Exists', 'table', 'Filter', 'Temporal', 'while', 'video', 'Localize', 'video', 'sitting_on_a_bed', 'objects'],
        [0, 1, 2, 3, None, None, 4, None, 6, 10]),
    'choose': (
        "Choose(dish, food, Iterate(Localize(before, [opening a door]), Filter(frame, [relations, holding, objects])))",
        ['Choose', 'dish', 'food', 'Filter', 'Temporal', 'before', 'video', 'Localize', 'video', 'opening_a_door', 'holding'],
        [0, 1, 2, 3, 4, None, None, 5, None, 7, 12]),
    'and': (
        "AND(Exists(dish, Iterate(video, Filter(frame, [objects]))), Exists(food, Iterate(video, Filter(frame, [objects]))))",
        ['And', 'Exists', 'dish', 'Filter', 'video', 'objects', 'Exists', 'food', 'Filter', 'video', 'objects'],
        [0, 1, 2, 3, 4, 8, 9, 10, 11, 12, 16]),
    'equals': (
        "Equals(Query(class, OnlyItem(Iterate(video, Filter(frame, [relations, holding, objects])))), Query(class, OnlyItem(Iterate(video, Filter(frame, [relations, touching, objects])))))",
        ['Equals', 'Filter', 'video', 'holding', 'Filter', 'video', 'touching'],
        [0, 4, 5, 10, 15, 16, 21]),
    'toaction': (
        "ToAction(holding, Query(class, OnlyItem(Iterate(video, Filter(frame, [relations, holding, objects])))))",
        ['ToAction', 'holding', 'Filter', 'video', 'holding'],
        [0, 1, 5, 6, 11]),
    'superlative': (
        "Superlative(max, Filter(video, [actions]), Subtract(Query(end, action), Query(start, action)))",
        ['Superlative', 'max', 'FilterFrame', 'video', 'actions', 'video'],
        [0, 1, 2, 3, 5, None]),
    'compare': (
        "Compare([before, after], Exists(holding a dish, Iterate(Localize(temporal tag, [watching television]), Filter(frame, [actions]))))",
        ['Compare', 'Exists', 'holding_a_dish', 'Filter', 'Temporal', 'before', 'video', 'Localize', 'video', 'watching_television', 'actions',
         'Exists', 'holding_a_dish', 'Filter', 'Temporal', 'after', 'video', 'Localize', 'video', 'watching_television', 'actions'],
        [0, 4, 5, 6, 7, None, None, 8, None, 10, 14, 4, 5, 6, 7, None, None, 8, None, 10, 14]),
    'iterate_until': (
        "Query(class, OnlyItem(IterateUntil(forward, Localize(after, [eating some food]), HasItem(Filter(frame, [relations, taking, objects])), Filter(frame, [relations, taking, objects]))))",
        ['Filter', 'AttnVideo', 'Temporal', 'after', 'video', 'Localize', 'video', 'eating_some_food', 'Relate', 'forward', 'HasItem',
         'FilterFrame', 'video', 'taking', 'taking'],
        [3, None, 5, None, None, 6, None, 8, None, 4, 9, 10, 11, 14, 20]),
    'iterate_until_exists': (
        "Query(class, OnlyItem(IterateUntil(backward, video, Exists(dish, Filter(frame, [relations, holding, objects])), Filter(frame, [relations, holding, objects]))))",
        ['Filter', 'AttnVideo', 'video', 'Relate', 'backward', 'ExistsFrame', 'dish', 'FilterFrame', 'video', 'holding', 'holding'],
        [3, None, 5, None, 4, 6, 7, 8, 9, 12, 18]),
    'xor_between': (
        "XOR(Exists(food, Iterate(Localize(between, [A, B]), Filter(frame, [relation, holding, objects]))), Exists(Query(class, OnlyItem(Iterate(video, Filter(frame, [relations, opening, objects])))), Iterate(Localize(between, [A, B]), Filter(frame, [relation, holding, objects]))))",
        ['Xor', 'Exists', 'food', 'Filter', 'Temporal', 'between', 'video', 'Localize', 'video', 'Array2', 'A', 'B', 'holding',
         'Exists', 'Filter', 'video', 'opening', 'Filter', 'Temporal', 'between', 'video', 'Localize', 'video', 'Array2', 'A', 'B', 'holding'],
        [0, 1, 2, 3, 4, None, None, 5, None, 6, 7, 8, 13, 15, 19, 20, 25, 27, 28, None, None, 29, None, 30, 31, 32, 37]),
}

# Extra hand-written layouts (valid for the reference interpreter, module_net.py:94-133) that reach the
# operators / dynamic-typing cases the ten AGQA templates above do not: XorFrame, And on attention maps,
# Superlative over an Array2 of keywords (min mode), FilterFrame with string keywords, Filter 'relations', supervised (non-root)
# Equals / Xor.
EXTRA_TEMPLATES = {
    'xorframe': (
        None,
        ['Filter', 'AttnVideo', 'video', 'Relate', 'forward', 'XorFrame', 'ExistsFrame', 'dish', 'FilterFrame', 'video', 'holding',
         'ExistsFrame', 'cup', 'FilterFrame', 'video', 'relations', 'touching'],
        [0, None, 1, None, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14]),
    'and_frames': (
        None,
        ['Filter', 'AttnVideo', 'video', 'Relate', 'backward', 'And', 'ExistsFrame', 'dish', 'FilterFrame', 'video', 'actions',
         'HasItem', 'FilterFrame', 'video', 'holding', 'relations'],
        [0, None, 1, None, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13]),
    'superlative_min': (
        None,
        ['Superlative', 'min', 'Array2', 'running', 'sitting_down', 'video'],
        [0, 1, 2, 3, 4, None]),
    'toaction_superlative': (
        None,
        ['Equals', 'ToAction', 'holding', 'Filter', 'video', 'cup', 'Superlative', 'max', 'Array2', 'running', 'sitting_down',
         'Temporal', 'after', 'video', 'Localize', 'video', 'opening_a_door'],
        [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, None, None, 12, None, 13]),
    # non-root Equals / Xor: the only way criterion_equals (MSE on the Linear(H,1) head) and criterion_exists on the Xor head
    # (train_module.py:92-107) ever run — the root is trained by the decoder only (module_net.py:107-113), and in the ten AGQA
    # templates Equals / Xor are always the root
    'and_equals_xor': (
        None,
        ['And', 'Equals', 'Filter', 'video', 'holding', 'Filter', 'video', 'touching',
         'Xor', 'Exists', 'dish', 'Filter', 'video', 'objects', 'Exists', 'food', 'Filter', 'video', 'objects'],
        list(range(19))),
    # Relate on a 2-D [1, T] Localize map: the reference adds beta[:1] (a constant) and softmaxes over T (modules.py:427-435); the
    # result keeps the [1, T] shape, which Temporal accepts
    'relate_localize': (
        None,
        ['Filter', 'Temporal', 'while', 'video', 'Relate', 'forward', 'Localize', 'video', 'opening_a_door', 'objects'],
        [0, 1, None, None, 2, None, 3, None, 4, 5]),
    'compare_xor_equals': (
        None,
        ['Compare', 'Xor', 'Filter', 'video', 'actions', 'Filter', 'video', 'cup',
         'Equals', 'ToAction', 'holding', 'Filter', 'video', 'relations', 'Filter', 'video', 'dish'],
        list(range(17))),
}

ALL_TEMPLATES = {**TEMPLATES, **EXTRA_TEMPLATES}

# Long layouts for the I3D stress configuration (BASELINE.json configs[4]: "long program layouts (>= 12 modules)").  Not part of the golden
# fixtures: `xor_between` (12 module calls) is the longest AGQA template above; these two chain the same AGQA sub-programs further.
STRESS_TEMPLATES = {
    'and_between_until': (          # 15 module calls
        None,
        ['And', 'Exists', 'food', 'Filter', 'Temporal', 'between', 'video', 'Localize', 'video', 'Array2', 'A', 'B', 'holding',
         'Exists', 'Filter', 'AttnVideo', 'video', 'Relate', 'forward', 'HasItem', 'FilterFrame', 'video', 'taking', 'opening',
         'Filter', 'Temporal', 'after', 'video', 'Localize', 'video', 'C', 'holding'],
        list(range(32))),
    'compare_between': (            # 12 module calls
        None,
        ['Compare', 'Exists', 'dish', 'Filter', 'Temporal', 'between', 'video', 'Localize', 'video', 'Array2', 'A', 'B', 'actions',
         'Exists', 'Filter', 'video', 'cup', 'Filter', 'Temporal', 'before', 'video', 'Localize', 'video', 'Array2', 'C', 'D', 'relations'],
        list(range(27))),
}
LONG_TEMPLATES = ['xor_between', 'and_between_until', 'compare_between']       # every layout has >= 12 module calls

# arity table (utils/program_parser.py:16-23) restricted to the tokens the interpreter dispatches on.
MODULE_ARITY = {
    'And': 2, 'AttnVideo': 2, 'Choose': 3, 'Compare': 2, 'Equals': 2, 'Exists': 2, 'ExistsFrame': 2, 'Filter': 2,
    'FilterFrame': 2, 'HasItem': 1, 'Localize': 2, 'Relate': 2, 'Superlative': 3, 'Temporal': 3, 'ToAction': 2,
    'Xor': 2, 'XorFrame': 2, 'Array2': 2,
}
# video_nmn/dataset.py:23 WORDS_TO_KEEP | module_net.py:23 type keywords
WORDS_TO_KEEP = {'forward', 'backward', 'while', 'between', 'before', 'after', 'max', 'min', 'start', 'end', 'video',
                 'actions', 'objects', 'relations'}

CLASS_POOL = ['class_%02d' % i for i in range(64)]


def _content_positions(tokens):
    return [i for i, t in enumerate(tokens) if t not in MODULE_ARITY and t not in WORDS_TO_KEEP]


def make_question(rng: np.random.Generator, template: str, T: int, V: int, text_size: int = 300,
                  answer_vocab: int = 172, with_gold: bool = False, object_types: int = 256,
                  qa_id: str | None = None):
    """One reference-schema ``data`` dict (CPU fp32 tensors)."""
    _, tokens, idx_list = (ALL_TEMPLATES.get(template) or STRESS_TEMPLATES[template])
    L = int(rng.integers(8, 25))
    question = torch.from_numpy((rng.standard_normal((L, text_size)) * 0.4).astype(np.float32))
    video = torch.from_numpy(np.abs(rng.standard_normal((T, V))).astype(np.float32))
    spans = {}
    for i in _content_positions(tokens):
        w = int(rng.integers(1, 4))
        s = int(rng.integers(0, L - w + 1))
        spans[i] = (s, s + w)
    data = {
        'question': question, 'video_features': video, 'prog_str_to_question_tokens': spans,
        'nmn_program_list': list(tokens), 'nmn_program_idx': list(idx_list),
        'answer': torch.tensor(int(rng.integers(0, answer_vocab - 1))),
        'qa_id': qa_id or 'syn-%s' % template, 'question_raw': ' '.join(tokens), 'template': template,
    }
    if with_gold:
        data['sg_res_by_step'] = make_gold(rng, tokens, idx_list, T, text_size, object_types)
    return data


def _subtree_end(tokens, i):
    """Index one past the prefix-order subtree rooted at token i (arity table of utils/program_parser.py:16-23)."""
    need = 1
    while need:
        need += MODULE_ARITY.get(tokens[i], 0) - 1
        i += 1
    return i


def make_gold(rng, tokens, idx_list, T, text_size, object_types):
    """Random intermediate supervision with the value types ``CriterionByModule`` expects
    (train_module.py:83-194): bool / (s,e) / [(s,e)..] / {name:(s,e)} / [(class, glove[n_w,text])...]."""
    def interval():
        s = float(rng.uniform(0, T - 1)); e = float(rng.uniform(s + 0.05, T)); return (s, e)
    gold = {}
    for i, (tok, idx) in enumerate(zip(tokens, idx_list)):
        if idx is None or i == 0 or tok not in MODULE_ARITY:
            continue
        if tok in ('Exists', 'Xor', 'Equals'):
            gold[idx] = bool(rng.integers(0, 2))
        elif tok == 'Localize':
            # K = 2 iff the keyword argument (the token right after the feat argument's subtree) is an Array2
            k = 2 if tokens[_subtree_end(tokens, i + 1)] == 'Array2' else 1
            gold[idx] = tuple(interval() for _ in range(k))
        elif tok in ('Temporal', 'ExistsFrame'):
            gold[idx] = interval()
        elif tok == 'FilterFrame':
            gold[idx] = {'obj_%d' % int(rng.integers(0, object_types)): interval()
                         for _ in range(int(rng.integers(1, 3)))}
        elif tok in ('Filter', 'ToAction', 'Superlative'):
            names = rng.choice(len(CLASS_POOL), size=int(rng.integers(1, 3)), replace=False)
            gold[idx] = [(CLASS_POOL[int(n)], class_embedding(int(n), text_size)) for n in names]
    return gold


def class_embedding(class_id: int, text_size: int):
    """Deterministic stand-in GloVe phrase [n_w, text_size] for a class name (same name -> same tensor)."""
    r = np.random.default_rng(10_000 + class_id)
    n_w = 1 + class_id % 3
    return torch.from_numpy((r.standard_normal((n_w, text_size)) * 0.4).astype(np.float32))


def make_questions(n: int, T: int, V: int, seed: int = 1234, templates=None, text_size: int = 300,
                   with_gold: bool = False, answer_vocab: int = 172, object_types: int = 256):
    """``n`` questions cycling through ``templates`` (default: the ten AGQA templates)."""
    rng = np.random.default_rng(seed)
    names = list(templates or TEMPLATES.keys())
    return [make_question(rng, names[i % len(names)], T, V, text_size, answer_vocab, with_gold, object_types,
                          qa_id='syn-%d' % i) for i in range(n)]


# ---- random well-typed layouts (fuzzing beyond the probed templates) ---------------------------------------------------------
# Value types of the reference interpreter's stack (video_nmn/modules.py signatures):
#   VID [T,H] frame features, VEC [H], VEC2 [2,H] (Array2), ATT1 [T] (ExistsFrame / HasItem / Relate), ATTK [K,T] (Localize, K = 1 | 2)
def random_layout(rng: np.random.Generator, max_modules: int = 12):
    """A random NMN prefix program the reference interpreter accepts: (tokens, idx_list).  Top-down over the operators' argument
    types with a budget of module calls; leaves are 'video' and content words ('w<i>', which get question spans)."""
    budget = [int(rng.integers(2, max_modules + 1))]
    n_words = [0]

    def word():
        n_words[0] += 1
        return ['w%d' % n_words[0]]

    def pick(options):
        return options[int(rng.integers(0, len(options)))]

    def spend():
        if budget[0] <= 0:
            return False
        budget[0] -= 1
        return True

    def vec(depth=0):
        if depth > 5 or not spend():
            return word()
        kind = pick(['Filter', 'Filter', 'Exists', 'ToAction', 'Compare', 'Equals', 'Xor', 'And', 'Choose', 'Superlative', 'word'])
        if kind == 'word':
            budget[0] += 1
            return word()
        if kind == 'Filter':
            kw = pick([None, 'actions', 'objects', 'relations'])
            return ['Filter'] + vid(depth + 1) + ([kw] if kw else vec_leafy(depth + 1))
        if kind == 'Exists':
            return ['Exists'] + vec_leafy(depth + 1) + vec(depth + 1)
        if kind == 'ToAction':
            return ['ToAction'] + vec(depth + 1) + vec_leafy(depth + 1)
        if kind in ('Compare', 'Equals', 'Xor', 'And'):
            return [kind] + vec(depth + 1) + vec(depth + 1)
        if kind == 'Choose':
            return ['Choose'] + vec_leafy(depth + 1) + vec_leafy(depth + 1) + vec(depth + 1)
        actions = pick(['vec', 'vec2', 'vid'])
        a = vec_leafy(depth + 1) if actions == 'vec' else (vec2(depth + 1) if actions == 'vec2' else vid(depth + 1))
        return ['Superlative', pick(['max', 'min'])] + a + vid(depth + 1)

    def vec_leafy(depth):                                   # keyword-like operand: usually a phrase, sometimes a computed vector
        return word() if rng.random() < 0.8 else vec(depth)

    def vec2(depth):
        if not spend():
            budget[0] += 0
        return ['Array2'] + vec_leafy(depth + 1) + vec_leafy(depth + 1)

    def vid(depth=0):
        if depth > 5 or rng.random() < 0.35 or not spend():
            return ['video']
        kind = pick(['Temporal', 'Temporal', 'FilterFrame', 'AttnVideo'])
        if kind == 'Temporal':
            return ['Temporal', pick(['while', 'before', 'after', 'between'])] + vid(depth + 1) + attk(depth + 1)
        if kind == 'FilterFrame':
            kw = pick([None, 'relations', 'actions'])
            return ['FilterFrame'] + vid(depth + 1) + ([kw] if kw else vec_leafy(depth + 1))
        return ['AttnVideo'] + vid(depth + 1) + att1(depth + 1)

    def attk(depth):
        spend()
        return ['Localize'] + vid(depth + 1) + (vec_leafy(depth + 1) if rng.random() < 0.7 else vec2(depth + 1))

    def att1(depth):
        spend()
        kind = pick(['HasItem', 'ExistsFrame', 'ExistsFrame', 'Relate', 'And', 'XorFrame'] if depth < 5 else ['HasItem', 'ExistsFrame'])
        if kind == 'HasItem':
            return ['HasItem'] + vid(depth + 1)
        if kind == 'ExistsFrame':
            return ['ExistsFrame'] + vec_leafy(depth + 1) + vid(depth + 1)
        if kind == 'Relate':
            return ['Relate', pick(['forward', 'backward'])] + att1(depth + 1)
        return [kind] + att1(depth + 1) + att1(depth + 1)

    budget[0] -= 1
    root = pick(['Filter', 'Exists', 'Compare', 'Equals', 'Xor', 'And', 'Choose', 'ToAction', 'Superlative'])
    if root == 'Filter':
        kw = pick([None, 'actions', 'objects', 'relations'])
        tokens = ['Filter'] + vid(1) + ([kw] if kw else word())
    elif root == 'Exists':
        tokens = ['Exists'] + word() + vec(1)
    elif root == 'ToAction':
        tokens = ['ToAction'] + vec(1) + word()
    elif root == 'Choose':
        tokens = ['Choose'] + word() + word() + vec(1)
    elif root == 'Superlative':
        tokens = ['Superlative', pick(['max', 'min'])] + (word() if rng.random() < 0.5 else vid(1)) + vid(1)
    else:
        tokens = [root] + vec(1) + vec(1)
    idx_list, k = [], 0
    for t in tokens:
        if t in WORDS_TO_KEEP and t not in ('actions', 'objects', 'relations'):
            idx_list.append(None)
        else:
            idx_list.append(k)
            k += 1
    return tokens, idx_list


def make_random_questions(n: int, T: int, V: int, seed: int = 99, max_modules: int = 12, text_size: int = 300, answer_vocab: int = 172,
                          distinct: int | None = None, with_gold: bool = False, object_types: int = 256):
    """``n`` questions over ``distinct`` (default n) random well-typed layouts; reference data-dict schema (``with_gold``: random
    intermediate supervision for every supervisable non-root module, as ``make_gold``)."""
    rng = np.random.default_rng(seed)
    layouts = [random_layout(rng, max_modules) for _ in range(distinct or n)]
    out = []
    for i in range(n):
        tokens, idx_list = layouts[i % len(layouts)]
        L = int(rng.integers(8, 25))
        spans = {}
        for j in _content_positions(tokens):
            w = int(rng.integers(1, 4))
            s0 = int(rng.integers(0, L - w + 1))
            spans[j] = (s0, s0 + w)
        out.append({'question': torch.from_numpy((rng.standard_normal((L, text_size)) * 0.4).astype(np.float32)),
                    'video_features': torch.from_numpy(np.abs(rng.standard_normal((T, V))).astype(np.float32)),
                    'prog_str_to_question_tokens': spans, 'nmn_program_list': list(tokens), 'nmn_program_idx': list(idx_list),
                    'answer': torch.tensor(int(rng.integers(0, answer_vocab - 1))), 'qa_id': 'rnd-%d' % i,
                    'question_raw': ' '.join(tokens), 'template': 'random'})
        if with_gold:
            out[-1]['sg_res_by_step'] = make_gold(rng, tokens, idx_list, T, text_size, object_types)
    return out


def model_config(T: int = 8, V: int = 4096, hidden: int = 512, text_size: int = 300, dropout: float = 0.0,
                 answer_vocab: int = 172, object_types: int = 256):
    """Mirrors the dict built at train_module.py:304-310."""
    return {'hidden_size': hidden, 'video_size': V, 'text_size': text_size, 'dropout': dropout,
            'answer_vocab_length': answer_vocab, 'max_video_length': T, 'init_method': 'default', 'layer_norm': 1,
            'have_pretrain_head': True, 'object_types': object_types}


PRETRAIN_MODULES = {'Exists', 'Xor', 'Equals', 'Filter', 'ToAction', 'FilterFrame', 'ExistsFrame', 'Superlative',
                    'Localize', 'Temporal', 'decoder'}   # CriterionByModule.criterions keys, train_module.py:36-48
