"""Host-side layout compiler: NMN prefix-token programs -> typed node tables for the device executor.

The reference interprets ``nmn_program_list`` with a dynamically-typed Python stack, one question at a time
(video_nmn/module_net.py:94-133), with arities from ``nary_mappings`` (utils/program_parser.py:16-23).  Here every
distinct token list is compiled ONCE into a typed DAG (``Layout``): node = module call or content word, keyword
strings (``while``, ``forward``, ``objects`` ...) folded into the consuming module's *variant*, children/levels as in
utils/program_parser.py:182-200 / :307-321.  ``collate`` concatenates the cached layouts of a batch into SoA int32
tables; the stable grouping by (level, op, variant), output-slot assignment and argument resolution then happen on
the device (csrc/layout_group.cu).  The host only histograms the group keys (``np.bincount``) so that launch
dimensions are known without a device->host sync.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L

# utils/program_parser.py:16-23 — arity of every module token
NARY = {}
for _n, _names in ((1, 'Array1 HasItem OnlyItem Query'),
                   (2, 'Array2 AND XOR And Xor Compare Equals Exists Filter Iterate Localize ToAction Relate AttnVideo '
                       'FilterFrame ExistsFrame XorFrame'),
                   (3, 'Array3 Superlative Choose Temporal'), (4, 'IterateUntil')):
    for _t in _names.split():
        NARY[_t] = _n

# video_nmn/modules.py:446-465 — the modules the interpreter dispatches on, in NAME_TO_MODULE order
MODULE_NAMES = ['And', 'AttnVideo', 'Choose', 'Compare', 'Equals', 'Exists', 'ExistsFrame', 'Filter', 'FilterFrame', 'HasItem',
                'Localize', 'Relate', 'Superlative', 'Temporal', 'ToAction', 'Xor', 'XorFrame', 'Array2']
OP_OF = {name: L.OP[name.upper()] for name in MODULE_NAMES}
OP_WORD = L.OP['WORD']
OP_NAME = {v: k for k, v in OP_OF.items()}
OP_NAME[OP_WORD] = '<word>'

# video_nmn/dataset.py:23 | video_nmn/module_net.py:23-25
WORDS_TO_KEEP = frozenset(['forward', 'backward', 'while', 'between', 'before', 'after', 'max', 'min', 'start', 'end', 'video',
                           'actions', 'objects', 'relations'])

TEMPORAL_MODES = {'while': 0, 'before': 1, 'after': 2, 'between': 3}
FILTER_KINDS = {'actions': 1, 'objects': 2, 'relations': 3}                 # 0 = tensor keyword ('representation')
FILTERFRAME_KINDS = {'relations': 1, 'actions': 2}                          # no 'objects' branch in the reference (KeyError)

# value types on the interpreter stack
STR, VID, ATT, VEC, VEC2 = 'str', 'vid', 'att', 'vec', 'vec2'
ARENA_OF = {VID: 'vid', VEC: 'vec', VEC2: 'vec', ATT: 'att'}


def children_and_parents(tokens):
    """utils/program_parser.py:182-200 semantics (children in pop order, parent index per token)."""
    children, parents, stack = [[] for _ in tokens], [0] * len(tokens), []
    for i in range(len(tokens) - 1, -1, -1):
        for _ in range(NARY.get(tokens[i], 0)):
            children[i].append(stack.pop())
        stack.append(i)
    for i, ch in enumerate(children):
        for c in ch:
            parents[c] = i
    return children, parents


def module_levels(tokens):
    """utils/program_parser.py:307-321 semantics: leaves 0, module = 1 + max(children)."""
    lv, stack = [0] * len(tokens), []
    for i in range(len(tokens) - 1, -1, -1):
        n = NARY.get(tokens[i], 0)
        if n:
            args = [stack.pop() for _ in range(n)]
            lv[i] = 1 + max(args)
        stack.append(lv[i])
    return lv


def program_is_valid(tokens):
    """utils/program_parser.py:324-333."""
    depth = 0
    for tok in reversed(tokens):
        depth += 1 - NARY.get(tok, 0)
        if depth < 1:
            return False
    return depth == 1


class _Val:
    """A value on the simulated interpreter stack.  ``rank2`` distinguishes the two attention-map shapes of the reference:
    Localize returns a 2-D ``[K, T]`` map (modules.py:205-216), ExistsFrame / HasItem a 1-D ``[T]`` one (modules.py:138,170-177) —
    AttnVideo, Relate and Temporal behave differently (or fail) depending on which one they are given."""
    __slots__ = ('type', 'node', 'K', 'text', 'level', 'rank2')

    def __init__(self, type_, node=-1, K=1, text=None, level=0, rank2=False):
        self.type, self.node, self.K, self.text, self.level, self.rank2 = type_, node, K, text, level, bool(rank2) or K > 1


def _type_error(tok, what):
    return TypeError('%s: %s (the reference interpreter fails on this layout too)' % (tok, what))


class Layout:
    """One compiled token list.  Node i = i-th token that is a module call or a content word (token order)."""

    def __init__(self, tokens, submodules=MODULE_NAMES, words_to_keep=WORDS_TO_KEEP):
        self.tokens = tuple(tokens)
        n_tok = len(tokens)
        node_of_token, n = [-1] * n_tok, 0
        for i, tok in enumerate(tokens):
            if tok in submodules or tok not in words_to_keep:
                node_of_token[i] = n
                n += 1
        self.n = n
        self.node_of_token = node_of_token
        self.token_of_node = [i for i, x in enumerate(node_of_token) if x >= 0]
        op = np.zeros(n, np.int32); variant = np.zeros(n, np.int32); level = np.zeros(n, np.int32)
        args = np.full((3, n), -1, np.int32)
        out_type = [None] * n; out_K = [1] * n; out_rank2 = [False] * n
        self.param_tokens = [[] for _ in range(n_tok)]          # token positions of each call's params (pop order)
        stack = []
        for i in range(n_tok - 1, -1, -1):
            tok = tokens[i]
            if tok in submodules:                                # module_net.py:100-106
                ps = [stack.pop() for _ in range(NARY[tok])]     # IndexError on an empty stack, like the reference
                self.param_tokens[i] = [p.text[1] if p.type == STR else self.token_of_node[p.node] for p in ps]
                vals = [(_Val(VID, -2, level=0) if (p.type == STR and p.text[0] == 'video') else p) for p in ps]
                nd = node_of_token[i]
                v, a, t, K = self._resolve(tok, vals)
                op[nd], variant[nd] = OP_OF[tok], v
                for k, x in enumerate(a):
                    args[k, nd] = x
                lvl = 1 + max(p.level for p in ps)
                level[nd] = lvl
                out_type[nd], out_K[nd] = t, K
                out_rank2[nd] = self._att_rank2(tok, vals, t, K)
                stack.append(_Val(t, nd, K, level=lvl, rank2=out_rank2[nd]))
            elif tok in words_to_keep:                           # module_net.py:121-124
                stack.append(_Val(STR, text=(tok, i)))
            else:                                                # module_net.py:126-131: phrase embedding
                nd = node_of_token[i]
                op[nd] = OP_WORD
                out_type[nd] = VEC
                stack.append(_Val(VEC, nd))
        assert len(stack) == 1                                   # module_net.py:135
        root = stack[0]
        if root.type != VEC:
            raise _type_error(tokens[0], 'the program result must be a [hidden] vector for the decoder, got %s' % root.type)
        self.root = root.node
        self.op, self.variant, self.level, self.args = op, variant, level, args
        self.out_type, self.out_K, self.out_rank2 = out_type, out_K, out_rank2
        self.key = (level.astype(np.int64) * 32 + op) * 8 + variant          # ascending key = schedule order
        self.word_nodes = [nd for nd in range(n) if op[nd] == OP_WORD]
        self.is_module = op != OP_WORD

    @staticmethod
    def _att_rank2(tok, p, out_type, K):
        """Is an attention-map result 2-D ([K, T]) in the reference?  Localize always; elementwise / Relate results keep the
        rank of their (broadcast) operands."""
        if out_type != ATT:
            return False
        return K > 1 or tok == 'Localize' or (tok in ('And', 'XorFrame', 'Relate') and any(x.type == ATT and x.rank2 for x in p))

    @staticmethod
    def _resolve(tok, p):
        """-> (variant, arg nodes, out type, out K) with the reference's argument order (pop order)."""
        ty = [x.type for x in p]

        def need(cond, what):
            if not cond:
                raise _type_error(tok, what)
        if tok in ('And', 'XorFrame'):                                           # modules.py:7-12, 75-80
            need(ty[0] == ty[1] and ty[0] in (VEC, ATT) and p[0].K == p[1].K, 'operands must both be vectors or both attention maps')
            return (0 if ty[0] == VEC else p[0].K), [p[0].node, p[1].node], ty[0], p[0].K
        if tok == 'AttnVideo':                                                   # (feat, attn) modules.py:330-340
            need(ty[0] == VID and ty[1] == ATT and p[1].K == 1, 'expects (frame features, [T] attention)')
            # attn.unsqueeze(1) * feat (modules.py:340) does not broadcast for a 2-D [1, T] map (Localize output)
            need(not p[1].rank2, 'the attention must be a 1-D [T] map (ExistsFrame / HasItem / Relate), not a [K, T] Localize map')
            return 0, [p[0].node, p[1].node], VID, 1
        if tok == 'Choose':
            need(ty == [VEC, VEC, VEC], 'expects three vectors')
            return 0, [x.node for x in p], VEC, 1
        if tok in ('Compare', 'Equals', 'Xor', 'ToAction', 'Exists'):
            need(ty == [VEC, VEC], 'expects two vectors')
            return 0, [p[0].node, p[1].node], VEC, 1
        if tok == 'ExistsFrame':                                                 # (keyword, feat) modules.py:169
            need(ty == [VEC, VID], 'expects (keyword vector, frame features)')
            return 0, [p[0].node, p[1].node], ATT, 1
        if tok == 'Filter':                                                      # (feat, keyword) modules.py:361
            need(ty[0] == VID, 'expects frame features first')
            if ty[1] == STR:
                return FILTER_KINDS[p[1].text[0]], [p[0].node], VEC, 1           # KeyError on other strings, like param[keyword]
            need(ty[1] == VEC, 'keyword must be a vector or a type word')
            return 0, [p[0].node, p[1].node], VEC, 1
        if tok == 'FilterFrame':                                                 # (feat, keyword) modules.py:398
            need(ty[0] == VID, 'expects frame features first')
            if ty[1] == STR:
                return FILTERFRAME_KINDS[p[1].text[0]], [p[0].node], VID, 1      # 'objects' -> KeyError as in the reference
            need(ty[1] == VEC, 'keyword must be a vector or a type word')
            return 0, [p[0].node, p[1].node], VID, 1
        if tok == 'HasItem':
            need(ty[0] == VID, 'only frame features are supported (a [hidden] input yields an unusable 0-d tensor in the reference)')
            return 0, [p[0].node], ATT, 1
        if tok == 'Localize':                                                    # (feat, keyword) modules.py:194
            need(ty[0] == VID and ty[1] in (VEC, VEC2), 'expects (frame features, keyword vector(s))')
            K = 1 if ty[1] == VEC else 2
            return K - 1, [p[0].node, p[1].node], ATT, K
        if tok == 'Relate':                                                      # (mode, attn) modules.py:423
            need(ty[0] == STR and ty[1] == ATT and p[1].K == 1, 'expects (direction word, [T] attention)')
            # on a 2-D [1, T] map the reference adds beta[:attn.size(0)] = beta[:1], a constant, and the implicit-dim softmax runs over
            # T (modules.py:427-435): the result is softmax_T(attn) whatever the direction -> variants 2 / 3 (no beta)
            return (0 if p[0].text[0] == 'forward' else 1) + (2 if p[1].rank2 else 0), [p[1].node], ATT, 1
        if tok == 'Superlative':                                                 # (mode, actions, feat) modules.py:233
            need(ty[0] == STR and ty[1] in (VEC, VEC2, VID) and ty[2] == VID, 'expects (max|min, actions, frame features)')
            kind = {VEC: 0, VEC2: 1, VID: 2}[ty[1]]
            return (1 if p[0].text[0] == 'min' else 0) + 2 * kind, [p[1].node, p[2].node], VEC, 1
        if tok == 'Temporal':                                                    # (mode, feat, attention) modules.py:310
            need(ty[0] == STR and ty[1] == VID and ty[2] == ATT, 'expects (mode word, frame features, attention map)')
            # torch.mean(attention_scores, dim=0) (modules.py:318) turns a 1-D [T] map into a scalar: Linear(T, T) then fails; the
            # 'while' / conv modes would gate every frame with that one scalar — not a temporal map any more, not supported here
            need(p[2].rank2, 'the attention must be a 2-D [K, T] map (a Localize output); a 1-D [T] map is reduced to a scalar by '
                             'the reference (mean over dim 0)')
            return TEMPORAL_MODES[p[0].text[0]] * 2 + (p[2].K - 1), [p[1].node, p[2].node], VID, 1
        if tok == 'Array2':
            need(ty == [VEC, VEC], 'expects two vectors')
            return 0, [p[0].node, p[1].node], VEC2, 1
        raise KeyError(tok)


def schedule_waves(layouts, max_iter=64):
    """Batch-level schedule: wave index per node of every distinct layout.

    The reference semantics only need "children before parents" (module_net.py:94-133).  ASAP levels
    (``Layout.level``, = utils/program_parser.py:307-321) give one (level, op, variant) group per distinct depth at which a
    module type occurs in the batch; delaying nodes so that all instances of one (op, variant) share a wave whenever the
    dependencies allow merges those groups (fewer, larger launches).  Fixpoint of
        wave(node) = max(1 + max wave(children), max over the batch of wave(nodes with the same (op, variant)))
    which converges when the (op, variant) dependency graph of the batch is acyclic; otherwise ASAP levels are kept."""
    waves = [lay.level.astype(np.int64).copy() for lay in layouts]
    kinds = [lay.op.astype(np.int64) * 8 + lay.variant for lay in layouts]
    orders = [sorted(range(lay.n), key=lambda nd, lay=lay: -lay.token_of_node[nd]) for lay in layouts]     # children first
    for _ in range(max_iter):
        tgt = {}
        for w, k in zip(waves, kinds):
            for kk, ww in zip(k.tolist(), w.tolist()):
                if ww > tgt.get(kk, -1):
                    tgt[kk] = ww
        changed = False
        for lay, w, k, order in zip(layouts, waves, kinds, orders):
            for nd in order:
                lv = tgt[int(k[nd])]
                for a in lay.args[:, nd]:
                    if a >= 0 and w[a] + 1 > lv:
                        lv = w[a] + 1
                if lv != w[nd]:
                    w[nd] = lv
                    changed = True
        if not changed:
            return waves
    return [lay.level.astype(np.int64).copy() for lay in layouts]


_LAYOUT_CACHE = {}


def compile_layout(tokens) -> Layout:
    key = tuple(tokens)
    lay = _LAYOUT_CACHE.get(key)
    if lay is None:
        lay = _LAYOUT_CACHE[key] = Layout(key)
    return lay


def _out_units(op, variant, T):
    """(arena, units per instance) of a group's outputs."""
    name = OP_NAME[op]
    if op == OP_WORD:
        return 'vec', 1
    if name in ('And', 'XorFrame'):
        return ('vec', 1) if variant == 0 else ('att', variant)
    if name in ('AttnVideo', 'FilterFrame', 'Temporal'):
        return 'vid', 1
    if name in ('ExistsFrame', 'HasItem', 'Relate'):
        return 'att', 1
    if name == 'Localize':
        return 'att', variant + 1
    if name == 'Array2':
        return 'vec', 2
    return 'vec', 1


HEAD_KIND = {'Equals': 'small', 'Xor': 'small', 'Exists': 'small', 'Filter': 'vec', 'Superlative': 'vec', 'ToAction': 'vec',
             'FilterFrame': 'ff'}


class NMNBatch:
    """A collated batch (host tensors, optionally pinned) + the device copies made by ``to``.

    Host tensors: ``video`` [B,T,V], ``question`` [n_tok,text], ``itab_host`` (one packed int32 buffer holding q_off,
    node_gid, node_q, node_arg, node_span, root_node).  Python-side metadata (layouts, node offsets, spans, program
    index lists) is kept for rebuilding ``res_by_step`` / ``result_of_each_step`` in the reference's format.
    """

    def __init__(self):
        self.device = None

    # ---- offsets inside the packed int table
    def _slices(self):
        B, n = self.B, self.n_nodes
        o, out = 0, {}
        for name, size in (('q_off', B + 1), ('node_gid', n), ('node_q', n), ('node_arg', 3 * n), ('node_span', 2 * n),
                           ('root_node', B), ('q_order', B), ('q_soff', B + 1), ('tok_src', self.n_tok)):
            out[name] = (o, size)
            o += (size + 3) // 4 * 4
        out['_total'] = (0, o)
        return out

    def to(self, device, non_blocking=True):
        """Upload (H2D) the batch: 3 copies (video, question, int tables)."""
        self.device = torch.device(device)
        if self.device.type == 'cuda' and self.device.index is None:
            self.device = torch.device('cuda', torch.cuda.current_device())
        if self.video is None:                                   # raw features: pool + concat on the device (csrc/ingest.cu)
            from . import ingest
            app = self.appearance.to(self.device, non_blocking=non_blocking)
            mot = self.motion.to(self.device, non_blocking=non_blocking) if self.motion is not None else None
            self.video_dev = ingest.pool_concat(app, mot, out_dtype=self.video_dtype)
        else:
            self.video_dev = self.video.to(self.device, non_blocking=non_blocking)
        self.question_dev = self.question.to(self.device, non_blocking=non_blocking)
        self.itab_dev = self.itab_host.to(self.device, non_blocking=non_blocking)
        return self

    def h2d_bytes(self):
        feats = [self.video] if self.video is not None else [self.appearance] + ([self.motion] if self.motion is not None else [])
        return (sum(t.numel() * t.element_size() for t in feats) + self.question.numel() * self.question.element_size()
                + self.itab_host.numel() * 4)

    def tab_ptr(self, name):
        o, _ = self._slices()[name]
        return self.itab_dev.data_ptr() + 4 * o

    def host_tab(self, name):
        o, size = self._slices()[name]
        return self.itab_host[o:o + size]


def _stage_rows(dst, srcs, row_off):
    """dst[row_off[i] : row_off[i] + srcs[i].shape[0]] <- srcs[i] for a whole batch (dtype-converting row copies of the dataset's fp32
    tensors into the — usually pinned, usually bf16 — staging buffer).  One native call (``stair_host_collate_rows``, csrc/host_collate.cu:
    a small host thread pool, no GPU work) instead of one torch copy per question: 4096 RX questions 0.35-0.45 s -> ~0.03 s.
    Sources that are not contiguous CPU fp32 / bf16 tensors take the per-tensor torch copy."""
    n = len(srcs)
    if n == 0:
        return
    cols = int(dst.shape[-1]) if dst.dim() > 1 else 1
    width = dst[0].numel() if dst.dim() > 1 else 1                   # elements per dst row (trailing dims flattened)
    ok = dst.is_contiguous() and dst.device.type == 'cpu' and dst.dtype in (torch.float32, torch.bfloat16)
    if ok:
        sd = srcs[0].dtype
        cpu = torch.device('cpu')
        ok = sd in (torch.float32, torch.bfloat16) and all([t.dtype == sd and t.device == cpu and t.is_contiguous() for t in srcs])
    if not ok or n < 16:
        for t, r in zip(srcs, row_off):
            dst[r:r + t.shape[0]].copy_(t)
        return
    import ctypes
    ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in srcs])
    rows = (ctypes.c_longlong * n)(*[int(t.shape[0]) for t in srcs])
    offs = (ctypes.c_longlong * n)(*[int(r) for r in row_off])
    del cols
    L.check(L.lib().stair_host_collate_rows(ptrs, rows, offs, L.i32(n), ctypes.c_longlong(width), L.i32(L.dtype_code(sd)), L.vp(dst.data_ptr()),
                                            L.i32(L.dtype_code(dst.dtype)), L.i32(0)), 'stair_host_collate_rows')


def collate(examples, pin_memory=False, video_dtype=None, question_dtype=None, merge_waves=True) -> NMNBatch:
    """Collate reference-schema ``data`` dicts (video_nmn/dataset.py:189-233) into one ``NMNBatch``.

    Replaces the reference's ``collate_fn = examples[0]`` (video_nmn/dataset.py:463-464).  All questions of a batch
    must have the same number of frames T.
    """
    if isinstance(examples, dict):
        examples = [examples]
    B = len(examples)
    if B == 0:
        raise ValueError('empty batch')
    b = NMNBatch()
    b.B = B
    b.examples = examples
    layouts = [compile_layout(e['nmn_program_list']) for e in examples]
    b.layouts = layouts
    b.appearance = b.motion = None
    if 'video_features' not in examples[0] and 'appearance_features' in examples[0]:
        # raw TGIF-QA style features (video_nmn/dataset.py:145-172): appearance [T, F, Da] (+ motion [T, Dm]); the frame mean and the
        # concat run on the device when the batch is uploaded (stair_b200/ingest.py)
        a0 = examples[0]['appearance_features']
        T, F, Da = (int(x) for x in a0.shape)
        has_m = examples[0].get('motion_features') is not None
        Dm = int(examples[0]['motion_features'].shape[1]) if has_m else 0
        app = torch.empty((B, T, F, Da), dtype=a0.dtype, pin_memory=pin_memory)
        mot = torch.empty((B, T, Dm), dtype=a0.dtype, pin_memory=pin_memory) if has_m else None
        for i, e in enumerate(examples):
            if tuple(e['appearance_features'].shape) != (T, F, Da):
                raise ValueError('all questions of a batch must have the same [T, F, D] appearance features')
            app[i].copy_(e['appearance_features'])
            if has_m:
                mot[i].copy_(e['motion_features'])
        b.appearance, b.motion, b.video = app, mot, None
        b.T, b.V = T, Da + Dm
        b.video_dtype = video_dtype or torch.bfloat16
    else:
        v0 = examples[0]['video_features']
        T, V = int(v0.shape[0]), int(v0.shape[1])
        for e in examples:
            if tuple(e['video_features'].shape) != (T, V):
                raise ValueError('all questions of a batch must have the same [T, V] video features; bucket by length '
                                 '(got %s and %s)' % ((T, V), tuple(e['video_features'].shape)))
        b.T, b.V = T, V
        vdt = video_dtype or v0.dtype
        video = torch.empty((B, T, V), dtype=vdt, pin_memory=pin_memory)
        _stage_rows(video.view(B * T, V), [e['video_features'] for e in examples], [i * T for i in range(B)])
        b.video = video
        b.video_dtype = vdt
    lens = np.array([int(e['question'].shape[0]) for e in examples], np.int64)
    q_off = np.zeros(B + 1, np.int64)
    np.cumsum(lens, out=q_off[1:])
    b.n_tok, b.L_max = int(q_off[-1]), int(lens.max())
    text = int(examples[0]['question'].shape[1])
    qdt = question_dtype or examples[0]['question'].dtype
    question = torch.empty((b.n_tok, text), dtype=qdt, pin_memory=pin_memory)
    _stage_rows(question, [e['question'] for e in examples], q_off[:-1].tolist())
    b.question = question
    b.text_size = text

    sizes = np.array([lay.n for lay in layouts], np.int64)
    node_start = np.zeros(B + 1, np.int64)
    np.cumsum(sizes, out=node_start[1:])
    n = int(node_start[-1])
    b.n_nodes, b.node_start = n, node_start
    node_key = np.empty(n, np.int64); node_q = np.empty(n, np.int32)
    node_arg = np.full((3, n), -1, np.int32); node_span = np.full((2, n), -1, np.int32)
    root = np.empty(B, np.int32)
    # vectorised per distinct layout
    by_layout = {}
    for qi, lay in enumerate(layouts):
        by_layout.setdefault(id(lay), (lay, []))[1].append(qi)
    distinct = [lay for lay, _ in by_layout.values()]
    waves = schedule_waves(distinct) if merge_waves else [lay.level.astype(np.int64) for lay in distinct]
    b.wave_of_layout = {id(lay): w for lay, w in zip(distinct, waves)}
    for (lay, qs), wave in zip(by_layout.values(), waves):
        qs = np.asarray(qs, np.int64)
        starts = node_start[qs]
        pos = (starts[:, None] + np.arange(lay.n)[None, :]).reshape(-1)
        node_key[pos] = np.tile((wave * 32 + lay.op) * 8 + lay.variant, len(qs))          # ascending key = schedule order
        node_q[pos] = np.repeat(qs, lay.n).astype(np.int32)
        for k in range(3):
            a = lay.args[k]
            glob = np.where(a[None, :] >= 0, a[None, :] + starts[:, None], a[None, :])
            node_arg[k, pos] = glob.reshape(-1).astype(np.int32)
        root[qs] = (starts + lay.root).astype(np.int32)
    # word spans (per question data; module_net.py:128-129).  A missing entry raises KeyError like the reference.  Questions of one layout
    # have the same word nodes: their (start, end) pairs are fetched with one itemgetter call per question and written with one numpy
    # assignment per layout; None / negative entries (python slice semantics of token_feature[s:t]) take the per-element path.
    from operator import itemgetter
    from itertools import chain
    for lay, qidx in by_layout.values():
        if not lay.word_nodes:
            continue
        wn = list(lay.word_nodes)
        keys_w = [lay.token_of_node[nd] for nd in wn]
        get = itemgetter(*keys_w) if len(keys_w) > 1 else (lambda d, k=keys_w[0]: (d[k],))
        rows = [get(examples[qi]['prog_str_to_question_tokens']) for qi in qidx]
        qarr = np.asarray(qidx, np.int64)
        pos = node_start[qarr][:, None] + np.asarray(wn, np.int64)[None, :]
        try:                                                      # [questions, word nodes, 2]; a None inside raises TypeError
            flat = chain.from_iterable(chain.from_iterable(rows))
            arr = np.fromiter(flat, dtype=np.int64, count=len(qidx) * len(wn) * 2).reshape(len(qidx), len(wn), 2)
            if next(flat, None) is not None:
                raise ValueError
        except (TypeError, ValueError):
            arr = None
        if arr is not None and not (arr < 0).any():
            node_span[0, pos] = arr[:, :, 0]
            node_span[1, pos] = arr[:, :, 1]
            continue
        for i, qi in enumerate(qidx):
            for j in range(len(wn)):
                s_, t_ = rows[i][j]
                if s_ is None and t_ is None:
                    s_, t_ = -1, -1
                elif s_ is None or t_ is None or s_ < 0 or t_ < 0:
                    s_, t_, _ = slice(s_, t_).indices(int(lens[qi]))
                node_span[0, pos[i, j]], node_span[1, pos[i, j]] = s_, t_
    keys = np.unique(node_key)
    gid = np.searchsorted(keys, node_key).astype(np.int32)
    counts = np.bincount(gid, minlength=len(keys))
    b.group_keys, b.group_counts = keys, counts
    b.n_groups = len(keys)
    b.node_gid_host = gid
    b.group_deps = _group_deps(gid, node_arg, len(keys))

    sl = b._slices()
    itab = torch.empty(sl['_total'][1], dtype=torch.int32, pin_memory=pin_memory)
    it = itab.numpy()
    q_order, q_soff, tok_src = length_sorted_schedule(q_off)
    for name, arr in (('q_off', q_off), ('node_gid', gid), ('node_q', node_q), ('node_arg', node_arg.reshape(-1)),
                      ('node_span', node_span.reshape(-1)), ('root_node', root), ('q_order', q_order), ('q_soff', q_soff),
                      ('tok_src', tok_src)):
        o, size = sl[name]
        it[o:o + size] = arr
    b.itab_host = itab
    b.answer = None
    ans = [e.get('answer') for e in examples]
    if not any([a is None for a in ans]):                        # (identity test: `None in ans` would call Tensor.__eq__ per element)
        b.answer = torch.tensor([a.item() if isinstance(a, torch.Tensor) else int(a) for a in ans], dtype=torch.int64)
    return b


def length_sorted_schedule(q_off):
    """Schedule of the inference text recurrence (StairBatch.q_order / q_soff / tok_src): the questions in descending length (stable), the
    token offsets in that order and the batch-order token row of every sorted row.  A 64-question block of the recurrence then stops at
    its own longest question instead of the batch's (csrc/lstm_fused.cu); the outputs keep the batch's order."""
    q_off = np.asarray(q_off, np.int64)
    lens = np.diff(q_off)
    order = np.argsort(-lens, kind='stable')
    soff = np.zeros(len(q_off), np.int64)
    np.cumsum(lens[order], out=soff[1:])
    n_tok = int(q_off[-1])
    # row r of the sorted layout belongs to sorted position p = searchsorted(soff, r): its source row is q_off[order[p]] + (r - soff[p])
    shift = np.repeat(q_off[:-1][order] - soff[:-1], lens[order])
    tok_src = np.arange(n_tok, dtype=np.int64) + shift
    return order.astype(np.int32), soff.astype(np.int32), tok_src.astype(np.int32)


def _group_deps(gid, node_arg, n_groups):
    """[n_groups, MAX_GROUP_DEPS] int32: the groups whose outputs each group reads (-1 = unused; first entry -2 = more producers than fit,
    i.e. "wait for every earlier group").  Lets the executor start a group as soon as ITS producers are done (StairBatch.group_deps)."""
    deps = np.full((n_groups, L.MAX_GROUP_DEPS), -1, np.int32)
    pairs = []
    for k in range(3):
        m = node_arg[k] >= 0
        if m.any():
            pairs.append(gid[m].astype(np.int64) * n_groups + gid[node_arg[k][m]])
    if pairs:
        uniq = np.unique(np.concatenate(pairs))
        cons, prod = uniq // n_groups, uniq % n_groups
        fill = np.zeros(n_groups, np.int64)
        for c_, p_ in zip(cons.tolist(), prod.tolist()):
            if fill[c_] < 0:
                continue
            if fill[c_] == L.MAX_GROUP_DEPS:
                deps[c_, :] = -1
                deps[c_, 0] = -2
                fill[c_] = -1
                continue
            deps[c_, fill[c_]] = p_
            fill[c_] += 1
    return deps


def collate_chunks(examples, n_chunks, **kw):
    """Collate ``examples`` as ``n_chunks`` contiguous sub-batches (sizes differ by at most one) for the pipelined forward
    (``VideoNMN.forward_pipelined``): the host->device copy of chunk k+1 overlaps the execution of chunk k."""
    n = len(examples)
    n_chunks = max(1, min(int(n_chunks), n))
    base, rem = divmod(n, n_chunks)
    out, lo = [], 0
    for c in range(n_chunks):
        hi = lo + base + (1 if c < rem else 0)
        out.append(collate(examples[lo:hi], **kw))
        lo = hi
    return out


def build_groups(batch: NMNBatch, head_modules=frozenset()):
    """Group table (host ``StairGroup`` array + the device [4][n_groups] int table) and arena sizes for this batch."""
    ng = batch.n_groups
    groups = (L.StairGroup * ng)()
    tab = np.zeros((4, ng), np.int32)
    next_free = {'vid': batch.B, 'vec': 0, 'att': 0, 'small': 0, 'hvec': 0, 'ff': 0}
    off = 0
    for g in range(ng):
        key = int(batch.group_keys[g])
        variant, op, level = key % 8, (key // 8) % 32, key // 256
        cnt = int(batch.group_counts[g])
        arena, mult = _out_units(op, variant, batch.T)
        name = OP_NAME[op]
        head = 1 if name in head_modules and name in HEAD_KIND else 0
        aux = -1
        if name == 'Temporal':                                   # stash of related_attn (modules.py:288,321-325)
            aux = next_free['att']; next_free['att'] += cnt
        elif head:
            hk = {'small': 'small', 'vec': 'hvec', 'ff': 'ff'}[HEAD_KIND[name]]
            aux = next_free[hk]; next_free[hk] += cnt
        G = groups[g]
        G.op, G.variant, G.level, G.count, G.node_off = op, variant, level, cnt, off
        G.out_base, G.out_mult, G.aux_base, G.head = next_free[arena], mult, aux, head
        tab[:, g] = (off, G.out_base, mult, aux)
        next_free[arena] += cnt * mult
        off += cnt
    return groups, tab, next_free


def host_grouping(batch: NMNBatch):
    """Reference grouping computed on the host with numpy (stable argsort) — used by the tests to check the device
    counting sort bit-exactly."""
    perm = np.argsort(batch.node_gid_host, kind='stable').astype(np.int32)
    return perm
