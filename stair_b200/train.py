"""Training step of the batched NMN: intermediate-supervision losses + backward on the sm_100a library.

Reference: ``train_module.py:33-194`` (``CriterionByModule``), ``:341-412`` (loop: per-question module losses, decoder CE,
window-level contrastive losses, one ``backward()`` per gradient-accumulation window, Adam).  The reference runs one
question per iteration and accumulates 32 of them; here one *batch* is one accumulation window: every loss is scaled by
``module_loss_weight`` (or ``decoder_loss_weight``) ``/ gradient_accumulation`` exactly like ``train_module.py:372,379,403``
with ``gradient_accumulation`` defaulting to the (global) number of questions in the window.

Host side (this file) only *collates*: it applies the reference's inclusion rules (``module_net.py:107-113`` for
``res_by_step``; ``train_module.py:350-366`` for which modules are supervised and how) and writes flat loss-row tables;
the losses, their gradients and the whole backward pass run in ``libstair_b200.so`` (csrc/executor_bwd.cu,
csrc/train_kernels.cu).  Gradients land in ``parameter.grad`` with the reference's shapes, parameters of modules that
no question of the window used keep ``grad = None`` (so ``torch.optim.Adam`` skips them as in the reference — SURVEY.md
hard part 4).  There is no CPU path.
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib as L
from . import layout as LY
from .params import grad_targets

LOSS_SLOTS = ('Localize', 'Temporal', 'ExistsFrame', 'Exists/Xor', 'Equals', 'contrastive', 'decoder', 'FilterFrame')
CONTRASTIVE = ('Filter', 'Superlative', 'ToAction')
SUPERVISED = ('Exists', 'Xor', 'Equals', 'Filter', 'ToAction', 'FilterFrame', 'ExistsFrame', 'Superlative', 'Localize',
              'Temporal')                                   # CriterionByModule.criterions keys, train_module.py:36-48


def span_to_attention(gold, T):
    """train_module.py:67-81 — soft [T] mask from a float interval (host, numpy)."""
    g = np.zeros(T, np.float32)
    start, end = min(T - 0.002, max(0.001, gold[0])), min(T - 0.001, gold[1])
    si, ei = math.ceil(start), math.floor(end)
    if si < ei:
        g[si:ei] += 1
    if si <= ei:
        g[si - 1] += si - start
        g[ei] += end - ei
    else:
        g[ei] += end - start
    return g


def create_attention_from_frame_interval(frame_interval, K, T):
    """module_net.py:190-208 — soft gold attention [K, T] from K float intervals (host, numpy).  Unused by the reference's own training
    loop (which calls ``span_to_attention``); kept because it is part of the module_net.py surface.  Same arithmetic, including the
    wrap-around of ``gold[start_int - 1]`` when ``start_int == 0`` cannot happen (start is clamped to >= 0.001) and the IndexError the
    reference raises when ``floor(end) == T`` cannot either (end is clamped to <= T - 0.001)."""
    g = np.zeros((K, T), np.float32)
    for i in range(K):
        start, end = max(0.001, frame_interval[i][0]), min(T - 0.001, frame_interval[i][1])
        si, ei = math.ceil(start), math.floor(end)
        if si < ei:
            g[i, si:ei] += 1
        g[i, si - 1] += si - start
        g[i, ei] += end - ei
    return g


class LossRows:
    """Flat loss-row tables of one window (host numpy), see ``StairTrain`` in include/stair_b200.h."""

    def __init__(self):
        self.att_node, self.att_kind, self.att_slot, self.att_gold, self.att_w = [], [], [], [], []
        self.bin_node, self.bin_which, self.bin_label, self.bin_w = [], [], [], []
        self.con_node, self.con_cls, self.con_w = [], [], []
        self.ff_node, self.ff_gold, self.ff_w = [], [], []       # criterion_filterframe rows (only when FilterFrame is not excluded)
        self.class_emb = {}                  # class name -> word-embedding phrase [n_w, text] (last writer wins, :364)
        self.counts = {k: 0 for k in LOSS_SLOTS}


def filterframe_gold(gold, T, O, word2index):
    """train_module.py:146-154 — {entity: interval} -> [T, O] soft masks, every frame normalised by its sum (0/0 -> 0)."""
    g = np.zeros((T, O), np.float32)
    for key, val in gold.items():
        g[:, word2index[key]] = span_to_attention(val, T)
    ssum = g.sum(axis=1, keepdims=True)
    return np.where(ssum > 0, g / np.where(ssum > 0, ssum, 1), 0).astype(np.float32)


def collate_losses(batch: LY.NMNBatch, pretrain_modules, T, module_loss_weight=1.0, gradient_accumulation=None,
                   modules_no_intermediate_train=('FilterFrame',), word2index=None, object_types=0) -> LossRows:
    """train_module.py:350-373 over every question of the window -> loss rows."""
    ga = float(gradient_accumulation or batch.B)
    rows = LossRows()
    if module_loss_weight == 0:
        return rows
    mlw = module_loss_weight / ga
    for q, (lay, e) in enumerate(zip(batch.layouts, batch.examples)):
        gold_by_step = e.get('sg_res_by_step') or {}
        idx = e['nmn_program_idx']
        base = int(batch.node_start[q])
        res_by_step = {}                                                  # module_net.py:107-113: a dict keyed by the original
        for i in range(len(lay.tokens) - 1, 0, -1):                       # program index (i != 0) — duplicated subtrees (Compare
            tok = lay.tokens[i]                                           # rewrite) share an index and the last writer wins
            if tok in LY.OP_OF and idx[i] is not None and tok in pretrain_modules:
                res_by_step[idx[i]] = i
        for step, i in res_by_step.items():                               # train_module.py:350-373
            tok = lay.tokens[i]
            if step not in gold_by_step or tok in modules_no_intermediate_train or tok not in SUPERVISED:
                continue
            gold = gold_by_step[step]
            if gold is None:
                continue
            node = base + lay.node_of_token[i]
            if tok in CONTRASTIVE:                                          # :360-366
                for cname, emb in gold:
                    rows.con_node.append(node); rows.con_cls.append(cname); rows.con_w.append(mlw)
                    rows.class_emb[cname] = emb
                    rows.counts['contrastive'] += 1
            elif tok == 'Localize':                                         # criterion :173-182, mean over [K,T]
                K = lay.out_K[lay.node_of_token[i]]
                for k in range(K):
                    rows.att_node.append(node); rows.att_kind.append(k); rows.att_slot.append(0)
                    rows.att_gold.append(span_to_attention(gold[k], T)); rows.att_w.append(mlw / (K * T))
                rows.counts['Localize'] += 1
            elif tok in ('Temporal', 'ExistsFrame'):                        # :157-164,184-191
                rows.att_node.append(node); rows.att_kind.append(2 if tok == 'Temporal' else 0)
                rows.att_slot.append(1 if tok == 'Temporal' else 2)
                rows.att_gold.append(span_to_attention(gold, T)); rows.att_w.append(mlw / T)
                rows.counts[tok] += 1
            elif tok in ('Exists', 'Xor', 'Equals'):                        # :92-107
                rows.bin_node.append(node); rows.bin_which.append({'Equals': 0, 'Xor': 1, 'Exists': 2}[tok])
                rows.bin_label.append(int(gold)); rows.bin_w.append(mlw)
                rows.counts['Equals' if tok == 'Equals' else 'Exists/Xor'] += 1
            elif tok == 'FilterFrame':                                      # criterion :141-155 (excluded by default, args.py:62)
                if word2index is None or object_types <= 0:
                    raise ValueError('FilterFrame supervision needs word2id (CriterionByModule(word2id), train_module.py:33-55) and '
                                     'config["object_types"]')
                rows.ff_node.append(node); rows.ff_gold.append(filterframe_gold(gold, T, object_types, word2index))
                rows.ff_w.append(mlw / (T * object_types))
                rows.counts['FilterFrame'] += 1
            else:
                raise NotImplementedError('no criterion for module %s (train_module.py:36-48)' % tok)
    return rows


def _subtree_info(lay):
    """Per node of a compiled layout: ({(op, variant)} of the module calls in its subtree, reads the text encoder?, reads the video encoder?).
    Cached on the layout.  Children follow their parent in token order, so a reverse sweep sees them first."""
    info = getattr(lay, '_subtree_info', None)
    if info is None:
        kinds, text, video = [None] * lay.n, [False] * lay.n, [False] * lay.n
        for nd in range(lay.n - 1, -1, -1):
            if lay.op[nd] == LY.OP_WORD:
                kinds[nd], text[nd] = frozenset(), True
                continue
            k = {(int(lay.op[nd]), int(lay.variant[nd]))}
            for a in lay.args[:, nd]:
                if a == -2:
                    video[nd] = True
                elif a >= 0:
                    k |= kinds[a]
                    text[nd] |= text[a]
                    video[nd] |= video[a]
            kinds[nd] = frozenset(k)
        info = lay._subtree_info = (kinds, text, video)
    return info


def touched_slots(batch: LY.NMNBatch, rows: LossRows, have_heads: bool, decoder_active: bool = True):
    """Weight-table slots that receive a gradient in this window (host logic).  The reference leaves ``grad = None`` on every parameter
    no loss of the window reaches, and Adam skips those (no step count, no momentum drift — SURVEY hard part 4): a parameter is touched
    iff it belongs to a module call inside the subtree of a supervised node, or — when the decoder loss is active
    (``decoder_loss_weight != 0``, train_module.py:376) — to any module of any question, the decoder, and the encoders feeding them."""
    Wt = L.W
    s = set()

    def lin(prefix):
        s.update((Wt[prefix + '_W'], Wt[prefix + '_B']))

    live, text, video = set(), False, False
    if decoder_active:
        lin('DEC0'); lin('DEC1')
        text = True                                                    # question_feature feeds the decoder (module_net.py:136)
        seen = set()
        for lay in batch.layouts:
            if id(lay) in seen:
                continue
            seen.add(id(lay))
            kinds, tx, vd = _subtree_info(lay)
            live |= kinds[lay.root]
            video |= vd[lay.root]
    else:
        nodes = np.unique(np.asarray(list(rows.att_node) + list(rows.bin_node) + list(rows.con_node) + list(rows.ff_node), np.int64))
        qs = np.searchsorted(batch.node_start, nodes, side='right') - 1
        seen = set()
        for node, q in zip(nodes.tolist(), qs.tolist()):
            lay = batch.layouts[q]
            nd = node - int(batch.node_start[q])
            if (id(lay), nd) in seen:
                continue
            seen.add((id(lay), nd))
            kinds, tx, vd = _subtree_info(lay)
            live |= kinds[nd]
            text |= tx[nd]
            video |= vd[nd]
    if video:
        s.update(Wt['VENC' + x] for x in ('_WIH', '_B', '_WHH_F', '_WHH_R'))
    if text:
        s.update(Wt['TENC' + x] for x in ('_WIH', '_B', '_WHH_F', '_WHH_R'))
    for op, variant in sorted(live):
        name = LY.OP_NAME[op]
        if name in ('Localize', 'Superlative'):
            lin('LOC_V0'); lin('LOC_V1'); lin('LOC_K')
            if name == 'Superlative':
                lin('SUP_D')
        elif name == 'Temporal':
            lin('TEMP_D'); s.update((Wt['TEMP_LN_G'], Wt['TEMP_LN_B']))
            mode = variant >> 1
            if mode > 0:
                s.update(range(Wt['TEMP_REL_BEFORE'] + 6 * (mode - 1), Wt['TEMP_REL_BEFORE'] + 6 * mode))
        elif name == 'Filter':
            s.update(range(Wt['FILT_REPR'] + 4 * variant, Wt['FILT_REPR'] + 4 * variant + 4)); lin('FILT_D')
        elif name == 'FilterFrame':
            s.update(range(Wt['FF_REPR'] + 4 * variant, Wt['FF_REPR'] + 4 * variant + 4)); lin('FF_D')
            if variant == 0:
                lin('FF_ATT')
        elif name == 'HasItem':
            lin('HAS0'); lin('HAS1')
        elif name == 'Relate':
            s.add(Wt['REL_BETA'])
        elif name in ('Compare', 'Equals', 'Xor'):
            lin(name.upper())
        elif name == 'Exists':
            lin('EXISTS0'); lin('EXISTS1')
        elif name == 'ToAction':
            lin('TOACT0'); lin('TOACT1')
    if have_heads:
        for which in set(rows.bin_which):
            lin(('EQUALS_HEAD', 'XOR_HEAD', 'EXISTS_HEAD')[which])
        if rows.ff_node:
            lin('FF_HEAD')
    return s


class NMNTrainStep:
    """``model(batch, return_res_by_step=True)`` + ``CriterionByModule`` + ``batch_loss.backward()`` of the reference
    (train_module.py:345-408) for one window of questions, on the GPU.

    >>> step = NMNTrainStep(model)                       # model: stair_b200.VideoNMN on a CUDA device
    >>> opt = torch.optim.Adam(model.parameters(), lr=2e-4)
    >>> out = step(list_of_data_dicts)                   # fills parameter.grad, returns losses / logits / answers
    >>> opt.step(); opt.zero_grad()

    A loss weight of 0 means that loss is NOT computed for the window — the reference's iteration gates (train_module.py:349,376:
    ``train_module_before_iters`` / ``train_decoder_after_iters``) and its ``return_res_by_step=args.module_loss_weight != 0`` — so the
    parameters only that loss reaches keep ``grad = None`` and Adam skips them.

    Data-parallel (one process per GPU): pass ``process_group`` (or leave the default group initialised); every rank
    passes its own shard, ``gradient_accumulation`` defaults to the global window size, gradients are summed with NCCL
    all-reduces over the flat fp32 gradient buffer (decoder / module slots while the encoders still back-propagate, encoder slots
    after; ``overlap_allreduce=False`` = one all-reduce after the whole backward) and the *touched* flags are OR-reduced.
    """

    def __init__(self, model, module_loss_weight=1.0, decoder_loss_weight=1.0, gradient_accumulation=None,
                 modules_no_intermediate_train=('FilterFrame',), distributed=None, process_group=None, global_negatives=True,
                 dropout_seed=None, save_activations_budget=24 << 30, word2id=None, overlap_allreduce=True):
        self.model = model
        self.module_loss_weight, self.decoder_loss_weight = module_loss_weight, decoder_loss_weight
        self.gradient_accumulation = gradient_accumulation
        self.modules_no_intermediate_train = tuple(modules_no_intermediate_train)
        self.group = process_group
        self.distributed = dist.is_available() and dist.is_initialized() if distributed is None else distributed
        self.global_negatives = global_negatives
        # nn.Dropout(config['dropout']) is active iff model.training (reference: model.train() in train_module.py:341).  Every run()
        # draws a fresh counter-based mask: seed = base seed + number of windows run so far (+ rank under data parallelism).
        self.dropout_seed = int(torch.initial_seed() if dropout_seed is None else dropout_seed) & 0xFFFFFFFFFFFFFFFF
        self.windows_run = 0
        # keep the module intermediates of the forward for the backward when they fit in this many bytes (else recompute them)
        self.save_activations_budget = int(save_activations_budget)
        # CriterionByModule(word2id) (train_module.py:33-55): entity name -> class id; ids are re-indexed to 0..O-1 in sorted order
        self.word2index = None
        if word2id is not None:
            ids = sorted(set(word2id.values()))
            id2index = {v: i for i, v in enumerate(ids)}
            self.word2index = {w: id2index[v] for w, v in word2id.items()}
        self._targets = None
        self._cache = {}
        self.last = None
        self.last_launches = 0
        # data parallel: all-reduce the decoder / module gradients while the encoders are still back-propagating (two-phase backward);
        # split_backward forces the two-phase enqueue without a process group (single-GPU equivalence test)
        self.overlap_allreduce = bool(overlap_allreduce)
        self.split_backward = False

    # ---- flat gradient buffer ------------------------------------------------------------------------------------
    def _layout(self):
        if self._targets is None:
            tg = grad_targets(self.model.submodules, self.model.config)
            off, o = {}, 0
            for wid in sorted(tg):
                off[wid] = o
                o += (tg[wid][0] + 63) // 64 * 64                 # 256-byte aligned slots
            self._targets, self._offsets, self._flat_numel = tg, off, o
        return self._targets, self._offsets, self._flat_numel

    def _grad_views(self):
        """(sizes of the consecutive pieces of the flat buffer, [(slot id, piece index, parameter, shares-its-piece)])."""
        if getattr(self, '_views', None) is None:
            tg, offsets, total = self._layout()
            regions = {}                                                     # (start, numel) -> piece index, in address order
            for wid, (_, targets) in tg.items():
                for prm, o in targets:
                    regions.setdefault((offsets[wid] + o, prm.numel()), None)
            cuts, pos = [], 0
            for k, (start, numel) in enumerate(sorted(regions)):
                if start < pos:
                    raise L.StairError('overlapping gradient regions in the flat buffer')
                if start > pos:
                    cuts.append(start - pos)                                 # alignment gap
                regions[(start, numel)] = len(cuts)
                cuts.append(numel)
                pos = start + numel
            if pos < total:
                cuts.append(total - pos)
            entries, seen = [], set()
            for wid, (_, targets) in tg.items():
                for prm, o in targets:
                    key = (offsets[wid] + o, prm.numel())
                    entries.append((wid, regions[key], prm, key in seen))
                    seen.add(key)
            self._views = (cuts, entries)
        return self._views

    def _buf(self, name, numel, dtype, device):
        t = self._cache.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != device:
            t = torch.empty(max(int(numel), 1), dtype=dtype, device=device)
            self._cache[name] = t
        return t

    def _world(self):
        return dist.get_world_size(self.group) if self.distributed else 1

    # ---- the step ----------------------------------------------------------------------------------------------------
    def plan(self, data):
        """Host work of one window: collate, apply the reference's loss inclusion rules, upload the loss-row tables.
        The returned plan can be ``run`` repeatedly (e.g. several epochs over a resident batch)."""
        model = self.model
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise L.StairError('VideoNMN parameters are on %s: stair_b200 trains only on CUDA (sm_100a) devices' % dev)
        batch = data if isinstance(data, LY.NMNBatch) else LY.collate([data] if isinstance(data, dict) else list(data))
        if batch.device is None:
            batch.to(dev)
        if batch.answer is None:
            raise ValueError('training needs data["answer"] for every question (train_module.py:376)')
        cfg = model.config
        world = self._world()
        n_window = batch.B
        if world > 1:
            cnt = torch.tensor([batch.B], device=dev, dtype=torch.int64)
            dist.all_reduce(cnt, group=self.group)
            n_window = int(cnt.item())
        ga = self.gradient_accumulation or n_window
        rows = collate_losses(batch, model.pretrain_modules, batch.T, self.module_loss_weight, ga, self.modules_no_intermediate_train,
                              self.word2index, int(cfg.get('object_types', 0) or 0))
        if rows.ff_node and not cfg['have_pretrain_head']:
            raise L.StairError('FilterFrame supervision needs have_pretrain_head (its criterion reads the [T, O] head)')
        rows.counts['decoder'] = batch.B if self.decoder_loss_weight != 0 else 0          # decoder CE applies to every question (:376-380)
        if rows.bin_node and not cfg['have_pretrain_head']:
            raise L.StairError('Exists/Xor/Equals supervision needs have_pretrain_head (their criterion reads the head logits)')
        if rows.con_node and not cfg['have_pretrain_head']:
            # without the head the reference scores the raw execution_result (module_net.py:110-113); the contrastive kernel always
            # L2-normalises (the contrastive_head), so that configuration is refused rather than silently computed differently
            raise L.StairError('Filter/ToAction/Superlative supervision needs have_pretrain_head (the contrastive loss is defined on the '
                               'L2-normalised head output)')
        # labels index device tables without a range check: validate them here like nn.CrossEntropyLoss does in the reference
        A = int(cfg['answer_vocab_length'])
        ans = batch.answer.reshape(-1)
        if self.decoder_loss_weight != 0 and (int(ans.min()) < 0 or int(ans.max()) >= A):
            bad = int(((ans < 0) | (ans >= A)).nonzero()[0])
            raise IndexError('Target %d is out of bounds (question %d; answer_vocab_length = %d)' % (int(ans[bad]), bad, A))
        if any(l not in (0, 1) for l in rows.bin_label):
            raise ValueError('Exists / Xor / Equals gold must be a bool (train_module.py:92-107), got labels %s' % sorted(set(rows.bin_label)))
        # window-level class names (train_module.py:360-366,388-406): under data parallelism the negatives of the whole
        # window are every rank's classes, so names + word embeddings are exchanged on the host (small)
        class_emb = rows.class_emb
        touched = touched_slots(batch, rows, cfg['have_pretrain_head'], decoder_active=self.decoder_loss_weight != 0)
        if world > 1:
            # one host-side exchange per window: class phrases (contrastive negatives) and the touched-parameter sets (Adam skips
            # parameters no rank's questions used, like the reference) — known from the layouts, so run() needs no device sync
            gathered = [None] * world
            mine = {'classes': {k: v.cpu() for k, v in class_emb.items()} if self.global_negatives else {}, 'touched': sorted(touched)}
            dist.all_gather_object(gathered, mine, group=self.group)
            if self.global_negatives:
                class_emb = {}
                for part in gathered:
                    class_emb.update(part['classes'])
            for part in gathered:
                touched |= set(part['touched'])
        pl = TrainPlan()
        pl.batch, pl.rows, pl.ga, pl.world = batch, rows, float(ga), world
        pl.class_names = sorted(class_emb)
        pl.class_phrases = [class_emb[n] for n in pl.class_names]
        pl.class_packed = model.pack_questions(pl.class_phrases) if pl.class_names else None     # uploaded once per window
        pos_of = {n: i for i, n in enumerate(pl.class_names)}

        def up(x, dtype):
            return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype).reshape(-1))).to(dev, non_blocking=True)

        pl.att = (up(rows.att_node, np.int32), up(rows.att_kind, np.int32), up(rows.att_slot, np.int32),
                  up(np.stack(rows.att_gold) if rows.att_gold else np.zeros(0), np.float32), up(rows.att_w, np.float32))
        pl.bin = (up(rows.bin_node, np.int32), up(rows.bin_which, np.int32), up(rows.bin_label, np.int32), up(rows.bin_w, np.float32))
        pl.con = (up(rows.con_node, np.int32), up([pos_of[c] for c in rows.con_cls], np.int32), up(rows.con_w, np.float32))
        pl.ff = (up(rows.ff_node, np.int32), up(np.stack(rows.ff_gold) if rows.ff_gold else np.zeros(0), np.float32), up(rows.ff_w, np.float32))
        pl.answer = batch.answer.to(torch.int32).to(dev, non_blocking=True)
        pl.touched = touched
        return pl

    def run(self, pl, assign_grads=True, dropout_seed=None):
        """Device work of one window: forward with history, losses, backward, (all-reduce), gradients into ``.grad``.
        ``dropout_seed`` pins the dropout masks of this window (tests); default: a fresh seed per call."""
        model, batch, rows = self.model, pl.batch, pl.rows
        dev = batch.device
        cfg = model.config
        T, H, A = batch.T, cfg['hidden_size'], cfg['answer_vocab_length']
        lib = L.lib()
        # gold text reps of the window's classes (module_net.py:78-89): text encoder without grad + L2Normalize
        cls_rep = None
        if pl.class_names:
            _, sent = model.encode_packed(*pl.class_packed)
            cls_rep = torch.empty((len(pl.class_names), H), dtype=torch.float32, device=dev)
            L.check(lib.stair_l2normalize(L.i32(L.dtype_code(sent.dtype)), L.ptr(sent), L.ptr(cls_rep), L.i32(len(pl.class_names)), L.i32(H),
                                          L.stream_ptr()), 'stair_l2normalize')
        st, ms, sb, bufs = model.prepare(batch, frozenset(['FilterFrame']) if rows.ff_node else frozenset(), training=True)
        tg, offsets, flat_numel = self._layout()
        flat = torch.zeros(flat_numel, dtype=torch.float32, device=dev)
        tr = L.StairTrain()
        for wid in range(L.W_COUNT):
            tr.grad[wid] = flat.data_ptr() + 4 * offsets[wid] if wid in offsets else None

        def p(t):
            return t.data_ptr() if t.numel() else None

        tr.n_att = len(rows.att_node)
        tr.att_node, tr.att_kind, tr.att_slot, tr.att_gold, tr.att_w = (p(t) for t in pl.att)
        tr.n_bin = len(rows.bin_node)
        tr.bin_node, tr.bin_which, tr.bin_label, tr.bin_w = (p(t) for t in pl.bin)
        tr.n_con = len(rows.con_node)
        tr.con_node, tr.con_pos, tr.con_w = (p(t) for t in pl.con)
        tr.n_ff = len(rows.ff_node)
        if tr.n_ff:
            tr.ff_node, tr.ff_gold, tr.ff_w = (p(t) for t in pl.ff)
            dhead = self._buf('dhead_ff', st.head_ff.numel(), torch.float32, dev)
            tr.dhead_ff, tr.dhead_ff_elems = dhead.data_ptr(), st.head_ff.numel()
        tr.n_cls = len(pl.class_names)
        tr.cls_rep = cls_rep.data_ptr() if cls_rep is not None else None
        tr.answer = pl.answer.data_ptr()
        tr.dec_w = self.decoder_loss_weight / pl.ga
        tr.dropout_p = float(cfg.get('dropout', 0.0) or 0.0) if model.training else 0.0
        if dropout_seed is None:
            rank = dist.get_rank(self.group) if self.distributed else 0
            dropout_seed = (self.dropout_seed + 0x9E3779B97F4A7C15 * (self.windows_run * max(1, pl.world) + rank)) & 0xFFFFFFFFFFFFFFFF
        tr.dropout_seed = int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
        self.windows_run += 1
        loss = torch.zeros(8, dtype=torch.float32, device=dev)
        tr.loss = loss.data_ptr()
        sizes = st.sizes
        dvid = self._buf('dvid', sizes['vid'] * T * H, torch.float32, dev)
        dvec = self._buf('dvec', sizes['vec'] * H, torch.float32, dev)
        datt = self._buf('datt', sizes['att'] * T, torch.float32, dev)
        dtok = self._buf('dtok', batch.n_tok * H, torch.float32, dev)
        dq = self._buf('dq', batch.B * H, torch.float32, dev)
        dlogits = self._buf('dlogits', batch.B * A, torch.float32, dev)
        tr.dvid, tr.dvec, tr.datt, tr.dtokfeat, tr.dqfeat, tr.dlogits = (t.data_ptr() for t in (dvid, dvec, datt, dtok, dq, dlogits))
        saved_bytes = int(lib.stair_train_saved_bytes(ctypes.byref(ms), ctypes.byref(sb)))
        saved = self._buf('saved', saved_bytes, torch.uint8, dev)
        tr.saved, tr.saved_bytes = saved.data_ptr(), saved.numel()
        act_bytes = int(lib.stair_train_act_bytes(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs)))
        if 0 < act_bytes <= self.save_activations_budget:
            act = self._buf('act_saved', act_bytes, torch.uint8, dev)
            tr.act_saved, tr.act_saved_bytes = act.data_ptr(), act.numel()
        ws_bytes = int(lib.stair_train_workspace_bytes(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr)))
        if ws_bytes < 0:
            raise L.StairError('stair_train_workspace_bytes failed (unsupported configuration)')
        ws = self._buf('train_ws', ws_bytes + 256, torch.uint8, dev)
        tr.workspace, tr.workspace_bytes = ws.data_ptr(), ws.numel()
        stream = L.stream_ptr()
        L.check(lib.stair_nmn_forward_train(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr), stream), 'stair_nmn_forward_train')
        launches = int(lib.stair_last_launch_count())
        touched = pl.touched                                               # already the union over ranks (plan)
        if (pl.world > 1 or self.split_backward) and self.overlap_allreduce:
            # Two enqueue steps: once losses + decoder + module groups are back-propagated every gradient slot except the encoders' is
            # final, so its NCCL all-reduce (the tail of the flat buffer: slots are laid out in STAIR_W_* order, encoders first) runs on
            # NCCL's stream while BPTT and the encoder weight gradients are still computed; the encoder slots follow.
            enc_end = min(offsets[w] for w in offsets if w >= L.W['DEC0_W'])
            L.check(lib.stair_nmn_backward_phases(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr), L.i32(1), stream),
                    'stair_nmn_backward_phases(modules)')
            launches += int(lib.stair_last_launch_count())
            works = []
            if pl.world > 1:
                works.append(dist.all_reduce(flat[enc_end:], group=self.group, async_op=True))
            L.check(lib.stair_nmn_backward_phases(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr), L.i32(2), stream),
                    'stair_nmn_backward_phases(encoders)')
            self.last_launches = launches + int(lib.stair_last_launch_count())
            if pl.world > 1:
                works.append(dist.all_reduce(flat[:enc_end], group=self.group, async_op=True))
                works.append(dist.all_reduce(loss, group=self.group, async_op=True))
                for w in works:
                    w.wait()                                               # the current stream waits for NCCL's
        else:
            L.check(lib.stair_nmn_backward(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr), stream), 'stair_nmn_backward')
            self.last_launches = launches + int(lib.stair_last_launch_count())
            if pl.world > 1:
                dist.all_reduce(flat, group=self.group)                    # NCCL sum over NVLink: gradients
                dist.all_reduce(loss, group=self.group)
        if assign_grads:
            cuts, entries = self._grad_views()
            pieces = torch.split_with_sizes(flat, cuts)                     # one call: a view per parameter region of the flat buffer
            for wid, piece, prm, dup in entries:
                if wid not in touched or not prm.requires_grad:
                    continue
                g = pieces[piece].view(prm.shape)
                if dup:
                    g = g.clone()                                            # the two LSTM biases of a direction share a slot
                prm.grad = g if prm.grad is None else prm.grad + g
        self.last = dict(state=st, plan=pl, flat=flat, offsets=offsets, touched=touched, train=tr, cls_rep=cls_rep)
        return {'logits': st.logits, 'answers': st.answers, 'loss_terms': loss, 'loss': loss[:8].sum(), 'loss_counts': dict(rows.counts),
                'state': st}

    def __call__(self, data, assign_grads=True, dropout_seed=None):
        return self.run(self.plan(data), assign_grads=assign_grads, dropout_seed=dropout_seed)


class TrainPlan:
    """Host-collated window: batch + loss-row tables on the device (see ``NMNTrainStep.plan``)."""
    pass


class Adam:
    """``torch.optim.Adam(lr, betas, eps, weight_decay=0)`` (train_module.py:326-332) running ``stair_adam_step`` per parameter;
    parameters whose ``grad`` is None are skipped (and their step counter does not advance), like torch."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.state = {}

    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if set_to_none:
                p.grad = None
            elif p.grad is not None:
                p.grad.zero_()

    @torch.no_grad()
    def step(self, lr=None):
        lib = L.lib()
        lr = self.lr if lr is None else lr
        for p in self.params:
            if p.grad is None:
                continue
            L.require_cuda(p, 'parameter')
            s = self.state.get(id(p))
            if s is None:
                s = self.state[id(p)] = {'step': 0, 'm': torch.zeros_like(p, dtype=torch.float32), 'v': torch.zeros_like(p, dtype=torch.float32)}
            s['step'] += 1
            g = p.grad.contiguous()
            L.check(lib.stair_adam_step(L.ptr(p), L.ptr(g), L.ptr(s['m']), L.ptr(s['v']), L.i64(p.numel()), ctypes.c_float(lr),
                                        ctypes.c_double(self.betas[0]), ctypes.c_double(self.betas[1]), ctypes.c_float(self.eps),
                                        L.i32(s['step']), L.stream_ptr()), 'stair_adam_step')
            # the update happened behind torch's back: bump the version counter so PackedWeights.refresh re-packs
            torch.autograd.graph.increment_version(p)


class FusedAdam(torch.optim.Adam):
    """``torch.optim.Adam(model.parameters(), lr, betas, eps, weight_decay=0)`` (train_module.py:326-332) as ONE kernel per step
    (``stair_adam_multi``): every parameter with a gradient is updated (parameters whose ``grad`` is None are skipped and their
    step counter does not advance, like torch) and the copies the CUDA kernels read — bf16 planes, transposed planes, the
    gate-interleaved W_hh, fp32 vectors (``PackedWeights``) — are rewritten in the same pass, so the next forward neither re-packs
    119 tensors with ~400 small torch kernels nor launches one optimizer kernel per tensor.

    State layout (``state[p] = {'step', 'exp_avg', 'exp_avg_sq'}``), ``param_groups`` (so ``LambdaLR`` works) and ``state_dict()``
    are torch.optim.Adam's: checkpoints are interchangeable with the reference's optimizer."""

    def __init__(self, model, lr=2e-4, betas=(0.9, 0.999), eps=1e-8):
        super().__init__(model.parameters(), lr=lr, betas=betas, eps=eps, weight_decay=0.0)
        self.model = model
        self._slots = None
        self._seg, self._seg_key, self._book, self._covered = None, None, {}, set()

    def _build_segments(self, slots, active, pw):
        """ctypes segment array + per-segment (parameter, twin, bookkeeping) rows for the parameters ``active`` (indices into the slot
        table); gradient pointers and bias corrections are filled in by ``step``."""
        arr = (L.StairAdamSeg * len(active))()
        rows, tile0 = [], 0
        for j, i in enumerate(active):
            wid, kind, prm, prm2, off, perm_wid = slots[i]
            L.require_cuda(prm, 'parameter')
            sg = arr[j]
            st = self._state_of(prm)
            book = self._book.setdefault(id(prm), {'state': st, 'k': int(st['step'])})
            sg.p, sg.m, sg.v = prm.data_ptr(), st['exp_avg'].data_ptr(), st['exp_avg_sq'].data_ptr()
            book2 = None
            if prm2 is not None:
                s2 = self._state_of(prm2)
                book2 = self._book.setdefault(id(prm2), {'state': s2, 'k': int(s2['step'])})
                sg.p2, sg.m2, sg.v2 = prm2.data_ptr(), s2['exp_avg'].data_ptr(), s2['exp_avg_sq'].data_ptr()
            packed = pw.tensors[wid]
            if kind == 'V':
                sg.kind, sg.nplanes, sg.rows, sg.cols = 0, 1, 1, prm.numel()
                sg.packed = packed.data_ptr() + 4 * off
                ntiles = (prm.numel() + 1023) // 1024
            else:
                rws = prm.shape[0]
                cols = prm.numel() // rws
                nplanes = packed.shape[0] if packed.dim() == 3 else 1
                n_total, ld = packed.shape[-2], packed.shape[-1]
                row_off = off // cols
                sg.kind, sg.nplanes, sg.rows, sg.cols = 1, nplanes, rws, cols
                sg.packed, sg.packed_ld, sg.packed_plane = packed.data_ptr() + 2 * row_off * ld, ld, n_total * ld
                tt = pw.transposed.get(wid)
                if tt is not None:
                    ld_t = tt.shape[-1]
                    sg.packed_t, sg.packed_t_ld, sg.packed_t_plane = tt.data_ptr() + 2 * row_off, ld_t, tt.shape[-2] * ld_t
                if perm_wid is not None and perm_wid in pw.tensors:
                    sg.packed_perm, sg.perm_hh = pw.tensors[perm_wid].data_ptr(), cols
                ntiles = ((rws + 63) // 64) * ((cols + 63) // 64)
            sg.tile0 = tile0
            tile0 += ntiles
            rows.append((prm, prm2, book, book2))
        return arr, rows, tile0

    def _slot_table(self):
        if self._slots is None:
            from .params import weight_sources
            sub, cfg = self.model.submodules, self.model.config
            kinds = {wid: kind for wid, (kind, _) in weight_sources(sub, cfg).items()}
            perm_of = {}
            for enc in ('VENC', 'TENC'):
                for d in ('F', 'R'):
                    if L.W.get('%s_WHHI_%s' % (enc, d)) in kinds:
                        perm_of[L.W['%s_WHH_%s' % (enc, d)]] = L.W['%s_WHHI_%s' % (enc, d)]
            slots = []
            for wid, (_, targets) in sorted(grad_targets(sub, cfg).items()):
                if wid not in kinds:
                    continue
                if kinds[wid] == 'V':                       # pair parameters that share an offset (b_ih + b_hh of one direction)
                    by_off = {}
                    for prm, o in targets:
                        by_off.setdefault(o, []).append(prm)
                    for o, prms in sorted(by_off.items()):
                        if len(prms) > 2:
                            raise L.StairError('more than two parameters share weight-table slot %d' % wid)
                        slots.append((wid, 'V', prms[0], prms[1] if len(prms) > 1 else None, o, None))
                else:
                    for prm, o in targets:
                        slots.append((wid, 'M', prm, None, o, perm_of.get(wid)))
            self._slots = slots
            self._covered = set()
            for _, _, prm, prm2, _, _ in slots:
                self._covered.add(id(prm))
                if prm2 is not None:
                    self._covered.add(id(prm2))
        return self._slots

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._seg, self._seg_key, self._book = None, None, {}               # the cached table points at the old moment tensors

    def _state_of(self, p):
        s = self.state[p]
        if not s:
            s['step'] = torch.tensor(0.0, dtype=torch.float32)
            s['exp_avg'] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
            s['exp_avg_sq'] = torch.zeros_like(p, dtype=torch.float32, memory_format=torch.preserve_format)
        return s

    @torch.no_grad()
    def step(self, closure=None):
        from .nmn import PRECISIONS
        if closure is not None:
            raise L.StairError('FusedAdam does not take a closure')
        if len(self.param_groups) != 1 or self.param_groups[0].get('weight_decay', 0):
            raise L.StairError('FusedAdam implements the reference optimizer: one parameter group, weight_decay = 0')
        model = self.model
        group = self.param_groups[0]
        lr, (b1, b2), eps = float(group['lr']), group['betas'], float(group['eps'])
        dev = next(model.parameters()).device
        pw = model._packed
        pw.refresh(model.submodules, model.config, PRECISIONS[model.precision], dev, training=True)   # no-op when current
        lib = L.lib()
        slots = self._slot_table()
        active = tuple(i for i, (wid, kind, prm, prm2, off, perm_wid) in enumerate(slots)
                       if prm.grad is not None and (prm2 is None or prm2.grad is not None))
        if active:
            # the segment table of this set of updated parameters: everything but the gradient pointers and the bias corrections is the
            # same from step to step, so the ctypes array is built once per (set of parameters, generation of the weight copies) and patched
            key = (active, pw.version)
            if self._seg_key != key:
                self._seg, self._seg_key = self._build_segments(slots, active, pw), key
            arr, rows, ntiles = self._seg
            steps, keep = [], []
            for j, (prm, prm2, book, book2) in enumerate(rows):
                book['k'] += 1
                k = book['k']
                sg = arr[j]
                sg.bc1, sg.bc2 = 1.0 - b1 ** k, 1.0 - b2 ** k
                g = prm.grad if prm.grad.is_contiguous() else prm.grad.contiguous()
                sg.g = g.data_ptr()
                keep.append(g)
                steps.append(book['state']['step'])
                if prm2 is not None:
                    book2['k'] += 1
                    g2 = prm2.grad if prm2.grad.is_contiguous() else prm2.grad.contiguous()
                    sg.g2 = g2.data_ptr()
                    keep.append(g2)
                    steps.append(book2['state']['step'])
            torch._foreach_add_(steps, 1)                                      # state[p]['step'] (torch.optim.Adam's layout), one call
            table = torch.frombuffer(bytearray(arr), dtype=torch.uint8).to(dev, non_blocking=True)
            L.check(lib.stair_adam_multi(L.ptr(table), L.i32(len(rows)), L.i32(ntiles), ctypes.c_float(lr), ctypes.c_double(b1),
                                         ctypes.c_double(b2), ctypes.c_float(eps), L.stream_ptr()), 'stair_adam_multi')
            del keep
            # the update (and the refresh of the packed copies) happened behind torch's back: keep PackedWeights' signature valid
            pw.mark_current(model.submodules, PRECISIONS[model.precision], dev)
        covered = self._covered
        # parameters outside the weight table (none in the reference model) fall back to the per-tensor kernel + a full re-pack
        for prm in group['params']:
            if id(prm) in covered or prm.grad is None:
                continue
            st = self._state_of(prm)
            st['step'] += 1
            L.check(lib.stair_adam_step(L.ptr(prm), L.ptr(prm.grad.contiguous()), L.ptr(st['exp_avg']), L.ptr(st['exp_avg_sq']), L.i64(prm.numel()),
                                        ctypes.c_float(lr), ctypes.c_double(b1), ctypes.c_double(b2), ctypes.c_float(eps), L.i32(int(st['step'])),
                                        L.stream_ptr()), 'stair_adam_step')
            torch.autograd.graph.increment_version(prm)
        return None


# ---- differentiable forward: the reference's own training loop on top of the CUDA path ------------------------------------------------
class _NMNFunction(torch.autograd.Function):
    """``stair_nmn_forward_train`` / ``stair_nmn_backward`` as ONE autograd node: inputs are the model's parameters, outputs the logits, the
    attention arena and the three pretrain-head buffers of a batch; the backward seeds the CUDA backward with the caller's gradients
    (``StairTrain.ext_*``) instead of the built-in criteria."""

    @staticmethod
    def forward(ctx, owner, batch, dropout_seed, *params):
        model = owner.model
        cfg = model.config
        lib = L.lib()
        dev = batch.device
        ctx.set_materialize_grads(False)                          # unused outputs arrive as None in backward (no zero tensors, no head work)
        heads = frozenset(m for m in model.pretrain_modules if m in LY.HEAD_KIND) if cfg['have_pretrain_head'] else frozenset()
        st, ms, sb, bufs = model.prepare(batch, heads, training=True, private=True)
        T, H, A = batch.T, cfg['hidden_size'], cfg['answer_vocab_length']
        tr = L.StairTrain()
        tr.dropout_p = float(cfg.get('dropout', 0.0) or 0.0) if model.training else 0.0
        tr.dropout_seed = int(dropout_seed) & 0xFFFFFFFFFFFFFFFF
        keep = {}

        def buf(name, numel, dtype):
            keep[name] = torch.empty(max(int(numel), 1), dtype=dtype, device=dev)
            return keep[name]

        saved = buf('saved', int(lib.stair_train_saved_bytes(ctypes.byref(ms), ctypes.byref(sb))), torch.uint8)
        tr.saved, tr.saved_bytes = saved.data_ptr(), saved.numel()
        act_bytes = int(lib.stair_train_act_bytes(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs)))
        if 0 < act_bytes <= owner.save_activations_budget:
            act = buf('act', act_bytes, torch.uint8)
            tr.act_saved, tr.act_saved_bytes = act.data_ptr(), act.numel()
        if st.head_ff.numel() > 1:                               # FilterFrame heads present: their gradient buffer takes part in the workspace plan
            dh = buf('dhead_ff', st.head_ff.numel(), torch.float32)
            tr.dhead_ff, tr.dhead_ff_elems = dh.data_ptr(), st.sizes['ff'] * T * (cfg.get('object_types', 0) or 0)
        dummy = buf('ext_probe', 4, torch.float32)
        tr.ext_dlogits = dummy.data_ptr()                         # external-seed mode from the start (workspace planning depends on it)
        ws_bytes = int(lib.stair_train_workspace_bytes(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr)))
        if ws_bytes < 0:
            raise L.StairError('stair_train_workspace_bytes failed (unsupported configuration)')
        ws = buf('train_ws', ws_bytes + 256, torch.uint8)
        tr.workspace, tr.workspace_bytes = ws.data_ptr(), ws.numel()
        loss = buf('loss', 8, torch.float32)
        tr.loss = loss.data_ptr()
        L.check(lib.stair_nmn_forward_train(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr), L.stream_ptr()),
                'stair_nmn_forward_train')
        ctx.owner, ctx.batch, ctx.state, ctx.structs, ctx.keep, ctx.params = owner, batch, st, (ms, sb, bufs, tr), keep, params
        owner._last_state = st
        O = cfg.get('object_types', 0) or 0
        sizes = st.sizes
        outs = (st.logits.clone(),
                st.att[:sizes['att'] * T].clone().view(-1, T),
                st.head_small[:sizes['small'] * 2].clone().view(-1, 2),
                st.head_vec[:sizes['hvec'] * H].clone().view(-1, H),
                st.head_ff[:sizes['ff'] * T * O].clone().view(-1, T, max(O, 1)))
        return outs

    @staticmethod
    def backward(ctx, dlogits, datt, dsmall, dhvec, dhff):
        owner, batch, st = ctx.owner, ctx.batch, ctx.state
        ms, sb, bufs, tr = ctx.structs
        model = owner.model
        cfg = model.config
        dev = batch.device
        lib = L.lib()
        T, H, A = batch.T, cfg['hidden_size'], cfg['answer_vocab_length']
        tg, offsets, flat_numel = owner._step._layout()
        flat = torch.zeros(flat_numel, dtype=torch.float32, device=dev)
        for wid in range(L.W_COUNT):
            tr.grad[wid] = flat.data_ptr() + 4 * offsets[wid] if wid in offsets else None

        def seed(g, numel):
            if g is None:
                return None
            g = g.detach().to(torch.float32).contiguous()
            if g.numel() != numel:
                raise L.StairError('gradient seed has %d elements, expected %d' % (g.numel(), numel))
            return g

        sizes = st.sizes
        O = cfg.get('object_types', 0) or 0
        gl = seed(dlogits, batch.B * A)
        if gl is None:
            gl = torch.zeros(batch.B * A, dtype=torch.float32, device=dev)
        ga_, gs, gv, gf = seed(datt, sizes['att'] * T), seed(dsmall, sizes['small'] * 2), seed(dhvec, sizes['hvec'] * H), seed(dhff, sizes['ff'] * T * max(O, 1))
        if sizes['ff'] == 0 or not tr.dhead_ff:
            gf = None
        tr.ext_dlogits = gl.data_ptr()
        tr.ext_datt = ga_.data_ptr() if ga_ is not None and ga_.numel() else None
        tr.ext_dhead_small = gs.data_ptr() if gs is not None and gs.numel() else None
        tr.ext_dhead_vec = gv.data_ptr() if gv is not None and gv.numel() else None
        tr.ext_dhead_ff = gf.data_ptr() if gf is not None and gf.numel() else None
        arenas = [torch.empty(max(n, 1), dtype=torch.float32, device=dev) for n in
                  (sizes['vid'] * T * H, sizes['vec'] * H, sizes['att'] * T, batch.n_tok * H, batch.B * H, batch.B * A)]
        tr.dvid, tr.dvec, tr.datt, tr.dtokfeat, tr.dqfeat, tr.dlogits = (t.data_ptr() for t in arenas)
        L.check(lib.stair_nmn_backward(ctypes.byref(ms), ctypes.byref(sb), ctypes.byref(bufs), ctypes.byref(tr), L.stream_ptr()), 'stair_nmn_backward')
        # which parameters the graph reaches (everything the batch's layouts execute, the encoders and — when logits were used — the decoder;
        # head Linears only when their outputs received a gradient): the others get None, like parameters outside an autograd graph
        touched = touched_slots(batch, LossRows(), False, decoder_active=True)
        if dlogits is None:
            touched -= {L.W['DEC0_W'], L.W['DEC0_B'], L.W['DEC1_W'], L.W['DEC1_B']}
        if gs is not None:
            for name, op in (('EQUALS_HEAD', 'Equals'), ('XOR_HEAD', 'Xor'), ('EXISTS_HEAD', 'Exists')):
                if any(op in lay.tokens for lay in batch.layouts):
                    touched |= {L.W[name + '_W'], L.W[name + '_B']}
        if gf is not None:
            touched |= {L.W['FF_HEAD_W'], L.W['FF_HEAD_B']}
        by_param = {}
        cuts, entries = owner._step._grad_views()
        pieces = torch.split_with_sizes(flat, cuts)
        for wid, piece, prm, dup in entries:
            if wid in touched:
                g = pieces[piece].view(prm.shape)
                by_param[id(prm)] = g.clone() if dup else g
        ctx.keep['backward'] = (flat, arenas, gl, ga_, gs, gv, gf)
        return (None, None, None) + tuple(by_param.get(id(p)) if p.requires_grad else None for p in ctx.params)


class DifferentiableNMN:
    """``model(data, return_res_by_step=True)`` of the reference with autograd history: ``logits`` and every ``res_by_step`` tensor
    (``pretrain_head`` outputs, module_net.py:107-113) are differentiable with respect to the model's parameters, so the reference's own
    training loop runs unchanged on the CUDA path — any torch criterion on them (``train_module.CriterionByModule``), ``loss.backward()``
    (once per window over several retained calls, train_module.py:408), ``optimizer.step()``:

    >>> net = DifferentiableNMN(model.train())
    >>> out = net(data)                                  # one data dict (reference shapes) or a list of them (batched shapes)
    >>> loss = criterion('decoder', out['logits'], data['answer']) + sum(criterion(m, r, out['sg_res_by_step'][k]) ...)
    >>> loss.backward(); optimizer.step()

    Forward and backward are the kernels of ``NMNTrainStep`` (forward with history, CUDA backward); only the loss gradients come from
    autograd (``StairTrain.ext_*`` seeds).  Every call owns its buffers, so several calls can be alive until ``backward``.
    ``NMNTrainStep`` (built-in criteria, one launch sequence per window) is the fast path; this is the compatible one."""

    def __init__(self, model, dropout_seed=None, save_activations_budget=24 << 30):
        if not model.config['have_pretrain_head']:
            raise L.StairError('DifferentiableNMN exposes the pretrain_head outputs: it needs have_pretrain_head')
        self.model = model
        self.save_activations_budget = int(save_activations_budget)
        self.dropout_seed = int(torch.initial_seed() if dropout_seed is None else dropout_seed) & 0xFFFFFFFFFFFFFFFF
        self.calls = 0
        self._step = NMNTrainStep(model, distributed=False)
        from .params import all_parameters
        self._params = all_parameters(model)

    def __call__(self, data, return_res_by_step=True):
        from .nmn import OutputViews
        model = self.model
        single = isinstance(data, dict)
        batch = data if isinstance(data, LY.NMNBatch) else LY.collate([data] if single else list(data))
        dev = next(model.parameters()).device
        if dev.type != 'cuda':
            raise L.StairError('VideoNMN parameters are on %s: stair_b200 runs only on CUDA (sm_100a) devices' % dev)
        if batch.device is None:
            batch.to(dev)
        seed = (self.dropout_seed + 0x9E3779B97F4A7C15 * self.calls) & 0xFFFFFFFFFFFFFFFF
        self.calls += 1
        logits, att, small, hvec, hff = _NMNFunction.apply(self, batch, seed, *self._params)
        st = self._last_state
        self._last_state = None
        model.check_status(st)
        ret = {'logits': logits, 'answers': logits.detach().argmax(1).to(torch.int32), 'state': st}
        if return_res_by_step:
            views = _DiffViews(model, st, att, small, hvec, hff)
            ret['res_by_step'] = [views.res_by_step(q) for q in range(batch.B)]
        else:
            ret['res_by_step'] = [dict() for _ in range(batch.B)]
        ret['sg_res_by_step'] = model._encode_gold(batch)
        if single:
            ret['logits'], ret['answers'] = ret['logits'][0], ret['answers'][0]
            ret['res_by_step'], ret['sg_res_by_step'] = ret['res_by_step'][0], ret['sg_res_by_step'][0]
        return ret


def _make_diff_views():
    from .nmn import OutputViews

    class DiffViews(OutputViews):
        """``res_by_step`` views over the DIFFERENTIABLE buffers of one ``_NMNFunction`` call (attention arena + head outputs)."""

        def __init__(self, model, st, att, small, hvec, hff):
            self.model, self.st, self.heads = model, st, None
            b, il = st.batch, st.itab_layout
            itab = st.itab.cpu().numpy()
            n = b.n_nodes
            self.out_slot = itab[il.out_slot:il.out_slot + n].copy()
            self.aux_slot = itab[il.aux_slot:il.aux_slot + n].copy()
            self.att, self.head_small, self.head_vec = att, small, hvec
            self.head_ff = hff if (model.config.get('object_types', 0) or 0) else None
            self.vid = self.vec = None
            self._last_temporal = None

        def node_output(self, q, nd):
            lay = self.st.batch.layouts[q]
            if lay.out_type[nd] in (LY.VID, LY.VEC, LY.VEC2):
                raise L.StairError('only attention maps and pretrain_head outputs are differentiable outputs of DifferentiableNMN')
            return super().node_output(q, nd)

    return DiffViews


def _DiffViews(*args):
    global _DIFF_VIEWS
    try:
        cls = _DIFF_VIEWS
    except NameError:
        cls = _DIFF_VIEWS = _make_diff_views()
    return cls(*args)
