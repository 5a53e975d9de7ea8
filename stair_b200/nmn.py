"""Drop-in ``VideoNMN`` (reference: video_nmn/module_net.py:11-176) executing on the sm_100a CUDA library.

Same constructor, attributes (``submodules``, ``contrastive_head``, ``words_to_keep``, ``encode_question_no_grad``),
``state_dict`` keys and ``forward`` contract as the reference; the difference is that ``forward`` also accepts a *list*
of reference-schema ``data`` dicts (or a pre-collated ``NMNBatch``) and then returns batched outputs:

    out = model(list_of_data, return_res_by_step=False, test_mode=True)
    out['logits']   -> [B, A] float32     (what evaluate.py:41 indexes)
    out['answers']  -> [B] int32          (torch.argmax(logits, dim=1), computed on device)

A single ``data`` dict keeps the reference's unbatched shapes (``logits`` [A], ``res_by_step`` {idx: (name, tensor)}).
There is no CPU path and no PyTorch fallback: the arithmetic runs in ``libstair_b200.so`` (csrc/); torch provides
device memory, streams and parameter storage only.
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import layout as LY
from .params import L2Normalize, PackedWeights, build_submodules

PRECISIONS = {'bf16': L.BF16, 'fp32': L.F32}


class ForwardState:
    """Device buffers of one forward call (arenas, tables).

    ``logits``, ``answers`` and ``status`` are allocated per call and stay valid.  The arenas (``vid``, ``vec``, ``att``, ``tokfeat``,
    ``qfeat``, ``head_*``, ``itab``) are views into the model's grow-only buffer cache: they are valid only until the NEXT
    ``forward`` / ``forward_batch`` / training step on the same model (``generation`` records which call filled them; ``OutputViews``
    refuses a stale state).  Everything ``VideoNMN.forward`` returns to the caller (``res_by_step``, ``result_of_each_step``) is copied out
    of the arenas first, like the independent tensors the reference returns."""
    pass


class VideoNMN(nn.Module):
    def __init__(self, config, debug=False, pretrain_modules=set(), precision='bf16'):
        """``precision``: 'bf16' (bf16 storage, fp32 accumulate — the fast path) or 'fp32' (fp32 storage, every
        contraction as a 6-term bf16x3 split on the tensor cores — the strict-parity path)."""
        super().__init__()
        self.debug = debug
        self.config = config
        self.pretrain_modules = pretrain_modules
        self.contrastive_head = L2Normalize()
        self.words_to_keep = set(LY.WORDS_TO_KEEP)
        self.submodules = build_submodules(config, self.contrastive_head)
        self.precision = precision
        self._packed = PackedWeights()
        self._ws = None                  # grow-only workspace / arena cache (torch caching allocator owns the memory)
        self._cache = {}
        self.last_launches = 0
        self._bind_operators()

    def _bind_operators(self):
        """Give every operator object a weak back-reference to this model (packed weights / precision for ``submodules[name](*params)``)."""
        from .modules import Operator
        for m in self.submodules.values():
            if isinstance(m, Operator):
                m._bind(self)

    def __setstate__(self, state):                      # pickled-Module loading (train_module.py:296-298)
        super().__setstate__(state)
        self._bind_operators()

    def _precision_code(self):
        return PRECISIONS[self.precision]

    # ---- helpers ---------------------------------------------------------------------------------------------------
    def _buf(self, name, numel, dtype, device):
        t = self._cache.get(name)
        if t is None or t.numel() < numel or t.dtype != dtype or t.device != device:
            t = torch.empty(max(int(numel), 1), dtype=dtype, device=device)
            self._cache[name] = t
        return t

    def release_buffers(self):
        self._cache.clear()

    @property
    def act_dtype(self):
        return torch.float32 if PRECISIONS[self.precision] == L.F32 else torch.bfloat16

    def _head_modules(self, return_res_by_step, return_result_of_each_step):
        if not self.config['have_pretrain_head'] or not (return_res_by_step or return_result_of_each_step):
            return frozenset()
        return frozenset(m for m in self.pretrain_modules if m in LY.HEAD_KIND)

    # ---- the batched forward ---------------------------------------------------------------------------------------
    def prepare(self, batch: LY.NMNBatch, head_modules=frozenset(), training=False, private=False):
        """Allocate (grow-only cache) the arenas / tables of one call and fill the C-ABI structs.
        Returns (state, StairModel, StairBatch, StairBuffers).  ``private``: fresh buffers owned by the returned state instead of the
        model's cache (several forwards whose results must stay alive at once, e.g. retained autograd graphs)."""
        if batch.device is None or batch.device.type != 'cuda':
            raise L.StairError('batch is not on a CUDA device: call batch.to("cuda") — stair_b200 has no CPU fallback')
        dev = batch.device
        cfg = self.config
        prec = PRECISIONS[self.precision]
        model = self._packed.refresh(self.submodules, cfg, prec, dev, training=training)
        if batch.V != cfg['video_size'] or batch.text_size != cfg['text_size']:
            raise ValueError('feature sizes %s/%s do not match the config %s/%s' % (batch.V, batch.text_size, cfg['video_size'], cfg['text_size']))
        groups, tab, sizes = LY.build_groups(batch, head_modules)
        B, T, H, A, O = batch.B, batch.T, cfg['hidden_size'], cfg['answer_vocab_length'], cfg.get('object_types', 0) or 0
        adt = self.act_dtype
        st = ForwardState()
        self._generation = getattr(self, '_generation', 0) + 1         # the cached arenas now belong to this call
        st.generation = self._generation
        st.batch, st.groups, st.sizes, st.T, st.H = batch, groups, sizes, T, H
        n, ng = batch.n_nodes, batch.n_groups
        gtab = torch.from_numpy(tab.reshape(-1)).to(dev, non_blocking=True)
        sb = L.StairBatch()
        sb.B, sb.T, sb.n_tok, sb.L_max, sb.n_nodes, sb.n_groups = B, T, batch.n_tok, batch.L_max, n, ng
        sb.video_dtype, sb.question_dtype = L.dtype_code(batch.video_dev.dtype), L.dtype_code(batch.question_dev.dtype)
        sb.video, sb.question = batch.video_dev.data_ptr(), batch.question_dev.data_ptr()
        for name in ('q_off', 'node_gid', 'node_q', 'node_arg', 'node_span', 'root_node', 'q_order', 'q_soff', 'tok_src'):
            setattr(sb, name, batch.tab_ptr(name))
        if getattr(self, 'device_text_sort', False):             # leave the length-sorted text schedule to the library's device sort
            sb.q_order = sb.q_soff = sb.tok_src = None
        sb.groups = ctypes.cast(groups, ctypes.POINTER(L.StairGroup))
        sb.group_tab = gtab.data_ptr()
        sb.group_deps = batch.group_deps.ctypes.data if getattr(batch, 'group_deps', None) is not None and ng else None
        lib = L.lib()
        ws_bytes = int(lib.stair_nmn_workspace_bytes(ctypes.byref(model), ctypes.byref(sb)))
        itab_ints = int(lib.stair_itab_ints(L.i32(n), L.i32(ng)))
        _cached = self._buf
        if private:
            def _cached(name, numel, dtype, device):
                return torch.empty(max(int(numel), 1), dtype=dtype, device=device)
        self_buf = _cached
        st.vid = self_buf('vid', sizes['vid'] * T * H, adt, dev)
        st.vec = self_buf('vec', sizes['vec'] * H, adt, dev)
        st.att = self_buf('att', sizes['att'] * T, torch.float32, dev)
        st.tokfeat = self_buf('tokfeat', batch.n_tok * H, adt, dev)
        st.qfeat = self_buf('qfeat', B * H, adt, dev)
        st.logits = torch.empty((B, A), dtype=torch.float32, device=dev)
        st.answers = torch.empty(B, dtype=torch.int32, device=dev)
        st.head_small = self_buf('head_small', sizes['small'] * 2, torch.float32, dev)
        st.head_vec = self_buf('head_vec', sizes['hvec'] * H, torch.float32, dev)
        st.head_ff = self_buf('head_ff', sizes['ff'] * T * O, torch.float32, dev)
        st.itab = self_buf('itab', itab_ints, torch.int32, dev)
        st.status = torch.zeros(4, dtype=torch.int32, device=dev)
        ws = self_buf('workspace', ws_bytes, torch.uint8, dev)
        bufs = L.StairBuffers()
        bufs.vid, bufs.vid_slots = st.vid.data_ptr(), sizes['vid']
        bufs.vec, bufs.vec_rows = st.vec.data_ptr(), sizes['vec']
        bufs.att, bufs.att_rows = st.att.data_ptr(), sizes['att']
        bufs.tokfeat, bufs.qfeat = st.tokfeat.data_ptr(), st.qfeat.data_ptr()
        bufs.logits, bufs.answers = st.logits.data_ptr(), st.answers.data_ptr()
        bufs.head_small, bufs.head_vec, bufs.head_ff = st.head_small.data_ptr(), st.head_vec.data_ptr(), st.head_ff.data_ptr()
        bufs.itab, bufs.itab_ints = st.itab.data_ptr(), st.itab.numel()
        bufs.workspace, bufs.workspace_bytes = ws.data_ptr(), ws.numel()
        bufs.status = st.status.data_ptr()
        st.keepalive = (gtab, model, sb, bufs, ws)
        il = L.StairItabLayout()
        lib.stair_itab_layout(L.i32(n), L.i32(ng), ctypes.byref(il))
        st.itab_layout = il
        return st, model, sb, bufs

    def forward_batch(self, batch: LY.NMNBatch, head_modules=frozenset(), phases=L.FWD_ALL, stream=None) -> ForwardState:
        """Enqueue VideoNMN.forward for a collated batch on the current stream; returns the device buffers."""
        if self.training and float(self.config.get('dropout', 0.0) or 0.0) > 0.0:
            raise L.StairError('forward() runs the inference path (no dropout); this model is in training mode with dropout=%g: '
                               'train through stair_b200.train.NMNTrainStep (dropout applied there) or call model.eval()' % self.config['dropout'])
        st, model, sb, bufs = self.prepare(batch, head_modules)
        lib = L.lib()
        rc = lib.stair_nmn_forward(ctypes.byref(model), ctypes.byref(sb), ctypes.byref(bufs), L.i32(phases), L.stream_ptr(stream))
        L.check(rc, 'stair_nmn_forward')
        self.last_launches = int(lib.stair_last_launch_count())
        return st

    def forward_pipelined(self, batches, head_modules=frozenset()):
        """Inference over a list of HOST sub-batches (``layout.collate_chunks(..., pin_memory=True)``) with the host->device
        copies on a separate stream: the upload of chunk k+1 overlaps the execution of chunk k, so a step costs about
        max(PCIe time, compute time) instead of their sum.  Returns (answers [B] int32, logits [B, A], per-chunk states).
        The chunks share the model's arena cache: of the returned states only ``logits`` / ``answers`` / ``status`` are valid for every
        chunk, the arenas (intermediates) only for the LAST one (see ``ForwardState``); use ``forward`` per chunk for audit outputs."""
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise L.StairError('VideoNMN parameters are on %s: stair_b200 runs only on CUDA (sm_100a) devices' % dev)
        comp = torch.cuda.current_stream(dev)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        events = []
        for b in batches:
            with torch.cuda.stream(cs):
                b.to(dev)
                for t in (b.video_dev, b.question_dev, b.itab_dev):
                    t.record_stream(comp)
                ev = torch.cuda.Event()
                ev.record(cs)
            events.append(ev)
        states = []
        for b, ev in zip(batches, events):
            comp.wait_event(ev)
            states.append(self.forward_batch(b, head_modules))
        answers = torch.cat([st.answers for st in states]) if len(states) > 1 else states[0].answers
        logits = torch.cat([st.logits for st in states]) if len(states) > 1 else states[0].logits
        return answers, logits, states

    def forward_stream(self, host_batches, depth=2, head_modules=frozenset(), device_hook=None):
        """Streaming inference over an iterable of HOST batches (each an ``NMNBatch`` or a list of chunk ``NMNBatch``es in
        pinned memory, e.g. ``layout.collate_chunks(..., pin_memory=True)``).  Generator: yields the answers of every batch
        (CPU int32 tensor, in order).  Up to ``depth`` batches are in flight: the host->device copies run on a copy stream and
        never wait for the compute stream, the answers come back through an asynchronous device->host copy into pinned
        memory, so in steady state a batch costs max(PCIe time, compute time) with no per-batch synchronisation bubble
        (the per-call ``forward_pipelined`` + ``.cpu()`` pays the last chunk's compute and the first chunk's upload serially).
        ``device_hook(answers) -> tensor`` runs on the compute stream before the read-back (e.g. the NCCL all-gather)."""
        import collections
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise L.StairError('VideoNMN parameters are on %s: stair_b200 runs only on CUDA (sm_100a) devices' % dev)
        comp = torch.cuda.current_stream(dev)
        if getattr(self, '_copy_stream', None) is None:
            self._copy_stream = torch.cuda.Stream(dev)
        cs = self._copy_stream
        inflight = collections.deque()

        def finish():
            host, done, _ = inflight.popleft()
            done.synchronize()
            return host

        for hb in host_batches:
            chunks = [hb] if isinstance(hb, LY.NMNBatch) else list(hb)
            events = []
            for b in chunks:
                with torch.cuda.stream(cs):
                    b.to(dev)
                    for t in (b.video_dev, b.question_dev, b.itab_dev):
                        t.record_stream(comp)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                events.append(ev)
            states = []
            for b, ev in zip(chunks, events):
                comp.wait_event(ev)
                states.append(self.forward_batch(b, head_modules))
            answers = torch.cat([st.answers for st in states]) if len(states) > 1 else states[0].answers
            if device_hook is not None:
                answers = device_hook(answers)
            host = torch.empty(answers.shape, dtype=answers.dtype, pin_memory=True)
            host.copy_(answers, non_blocking=True)
            done = torch.cuda.Event()
            done.record(comp)
            inflight.append((host, done, (states, answers)))
            if len(inflight) >= max(1, depth):
                yield finish()
        while inflight:
            yield finish()

    def check_status(self, st: ForwardState):
        """Synchronising check of the device-side status word (layout grouping mismatch)."""
        code = int(st.status[0].item())
        if code != 0:
            raise L.StairError('device layout grouping disagrees with the host histogram (status %d)' % code)

    # ---- reference surface -----------------------------------------------------------------------------------------
    def forward(self, data, return_res_by_step=True, return_result_of_each_step=False, test_mode=False):
        single = isinstance(data, dict)
        if isinstance(data, LY.NMNBatch):
            batch = data
        else:
            batch = LY.collate([data] if single else list(data))
        dev = next(self.parameters()).device
        if dev.type != 'cuda':
            raise L.StairError('VideoNMN parameters are on %s: stair_b200 runs only on CUDA (sm_100a) devices' % dev)
        if batch.device is None:
            batch.to(dev)
        heads = self._head_modules(return_res_by_step, return_result_of_each_step)
        st = self.forward_batch(batch, heads)
        ret = {'logits': st.logits, 'answers': st.answers, 'state': st}
        views = None
        if return_res_by_step or return_result_of_each_step:
            self.check_status(st)
            views = OutputViews(self, st, heads)
        ret['res_by_step'] = [views.res_by_step(q) for q in range(batch.B)] if return_res_by_step else [dict() for _ in range(batch.B)]
        if return_result_of_each_step:
            ret['result_of_each_step'] = [views.result_of_each_step(q) for q in range(batch.B)]
        if views is not None:
            self.submodules['Temporal'].related_attn = views.last_temporal()
        if not test_mode:
            ret['sg_res_by_step'] = self._encode_gold(batch)
        if single:
            ret['logits'] = ret['logits'][0]
            ret['answers'] = ret['answers'][0]
            for k in ('res_by_step', 'result_of_each_step', 'sg_res_by_step'):
                if k in ret:
                    ret[k] = ret[k][0]
        return ret

    # ---- text encoder alone (module_net.py:147-158) -------------------------------------------------------------------
    def pack_questions(self, questions):
        """Upload a list of [L_i, text] word-embedding phrases once: (packed [n_tok, text] fp32, q_off [n+1] int32 on the
        device, host offsets, L_max).  ``encode_packed`` can then run any number of times without touching the host."""
        dev = next(self.parameters()).device
        lens = np.array([int(q.shape[0]) for q in questions], np.int64)
        q_off = np.zeros(len(questions) + 1, np.int64)
        np.cumsum(lens, out=q_off[1:])
        packed = torch.cat([q.detach().reshape(-1, self.config['text_size']).to(torch.float32) for q in questions]).to(dev)
        qo = torch.from_numpy(q_off.astype(np.int32)).to(dev)
        return packed, qo, q_off, int(lens.max())

    def encode_packed(self, packed, qo, q_off, L_max):
        """Text encoder (module_net.py:147-158) over packed phrases -> (token features [n_tok, H], sentence features [n, H])."""
        dev = packed.device
        cfg = self.config
        H = cfg['hidden_size']
        n, n_tok = len(q_off) - 1, int(q_off[-1])
        model = self._packed.refresh(self.submodules, cfg, PRECISIONS[self.precision], dev)
        sb = L.StairBatch()
        sb.B, sb.T, sb.n_tok, sb.L_max, sb.n_nodes, sb.n_groups = n, cfg['max_video_length'], n_tok, L_max, 0, 0
        sb.video_dtype, sb.question_dtype = L.F32, L.F32
        sb.question, sb.q_off = packed.data_ptr(), qo.data_ptr()
        lib = L.lib()
        adt = self.act_dtype
        ws = torch.empty(int(lib.stair_nmn_workspace_bytes(ctypes.byref(model), ctypes.byref(sb))), dtype=torch.uint8, device=dev)
        tok = torch.empty((n_tok, H), dtype=adt, device=dev)
        sent = torch.empty((n, H), dtype=adt, device=dev)
        itab = torch.empty(int(lib.stair_itab_ints(L.i32(0), L.i32(0))) + 4, dtype=torch.int32, device=dev)
        status = torch.empty(4, dtype=torch.int32, device=dev)
        bufs = L.StairBuffers()
        bufs.tokfeat, bufs.qfeat = tok.data_ptr(), sent.data_ptr()
        bufs.itab, bufs.itab_ints = itab.data_ptr(), itab.numel()
        bufs.workspace, bufs.workspace_bytes = ws.data_ptr(), ws.numel()
        bufs.status = status.data_ptr()
        L.check(lib.stair_nmn_forward(ctypes.byref(model), ctypes.byref(sb), ctypes.byref(bufs), L.i32(L.FWD_ENCODE_TEXT), L.stream_ptr()),
                'stair_nmn_forward(text)')
        return tok, sent

    def encode_questions(self, questions):
        """List of [L_i, text] embeddings -> (list of token features [L_i, H], sentence features [n, H])."""
        packed, qo, q_off, L_max = self.pack_questions(questions)
        tok, sent = self.encode_packed(packed, qo, q_off, L_max)
        return [tok[q_off[i]:q_off[i + 1]] for i in range(len(questions))], sent

    def encode_question(self, question):
        toks, sent = self.encode_questions([question])
        return toks[0], sent[0]

    @torch.no_grad()
    def encode_question_no_grad(self, question):
        return self.encode_question(question)

    def _encode_gold(self, batch):
        """module_net.py:78-89: class-name lists in sg_res_by_step -> L2-normalised text-encoder sentence features."""
        out, phrases, where = [], [], []
        for qi, e in enumerate(batch.examples):
            new = {}
            for key, value in e.get('sg_res_by_step', {}).items():
                if isinstance(value, list) and len(value) and isinstance(value[0][1], torch.Tensor):
                    new[key] = [None] * len(value)
                    for j, (name, emb) in enumerate(value):
                        where.append((qi, key, j, name))
                        phrases.append(emb)
                else:
                    new[key] = value
            out.append(new)
        if phrases:
            _, sent = self.encode_questions(phrases)
            H = sent.shape[1]
            normed = torch.empty((len(phrases), H), dtype=torch.float32, device=sent.device)
            L.check(L.lib().stair_l2normalize(L.i32(L.dtype_code(sent.dtype)), L.ptr(sent), L.ptr(normed), L.i32(len(phrases)), L.i32(H),
                                              L.stream_ptr()), 'stair_l2normalize')
            for i, (qi, key, j, name) in enumerate(where):
                out[qi][key][j] = (name, normed[i])
        return out


class OutputViews:
    """Rebuilds the reference's per-question ``res_by_step`` / ``result_of_each_step`` from the arenas.

    The used part of every arena is copied ONCE into storage private to this object (a handful of device copies per forward), and the
    per-question tensors are views of those copies: what leaves ``forward`` never aliases the model's grow-only buffer cache, so a
    later forward on the same model cannot overwrite results a caller kept (evaluate.py:65-117 accumulates Filter outputs)."""

    def __init__(self, model: VideoNMN, st: ForwardState, heads):
        if getattr(st, 'generation', None) != getattr(model, '_generation', None):
            raise L.StairError('stale ForwardState: its arenas were overwritten by a later forward on the same model')
        self.model, self.st, self.heads = model, st, heads
        b = st.batch
        il = st.itab_layout
        itab = st.itab.cpu().numpy()
        n = b.n_nodes
        self.out_slot = itab[il.out_slot:il.out_slot + n].copy()
        self.aux_slot = itab[il.aux_slot:il.aux_slot + n].copy()
        T, H = st.T, st.H
        self.vid = st.vid[:st.sizes['vid'] * T * H].clone().view(-1, T, H)
        self.vec = st.vec[:st.sizes['vec'] * H].clone().view(-1, H)
        self.att = st.att[:st.sizes['att'] * T].clone().view(-1, T)
        O = model.config.get('object_types', 0) or 0
        self.head_small = st.head_small[:st.sizes['small'] * 2].clone().view(-1, 2)
        self.head_vec = st.head_vec[:st.sizes['hvec'] * H].clone().view(-1, H)
        self.head_ff = st.head_ff[:st.sizes['ff'] * T * O].clone().view(-1, T, max(O, 1)) if O else None
        self._last_temporal = None

    def node_output(self, q, nd):
        b = self.st.batch
        lay = b.layouts[q]
        g = int(b.node_start[q]) + nd
        t, s = lay.out_type[nd], int(self.out_slot[g])
        if t == LY.VID:
            return self.vid[s]
        if t == LY.VEC:
            return self.vec[s]
        if t == LY.VEC2:
            return self.vec[s:s + 2]
        K = lay.out_K[nd]
        name = LY.OP_NAME[int(lay.op[nd])]
        if K > 1 or lay.out_rank2[nd]:
            return self.att[s:s + K]                      # [K, T] (Localize, and whatever keeps a Localize map's rank)
        return self.att[s]                                # [T]

    def head_output(self, q, nd):
        """``submodules[prog].pretrain_head(execution_result)`` for node nd."""
        b = self.st.batch
        lay = b.layouts[q]
        name = LY.OP_NAME[int(lay.op[nd])]
        g = int(b.node_start[q]) + nd
        a = int(self.aux_slot[g])
        if name == 'Temporal':
            return self.att[a]
        kind = LY.HEAD_KIND.get(name)
        if kind is None:
            if name in ('HasItem', 'ExistsFrame', 'Localize'):
                return self.node_output(q, nd)            # nn.Identity heads
            raise AttributeError("'%s' module has no pretrain_head" % name)     # as the reference would
        if kind == 'small':
            return self.head_small[a, :1 if name == 'Equals' else 2]
        if kind == 'vec':
            return self.head_vec[a]
        return self.head_ff[a]

    def res_by_step(self, q):
        """module_net.py:107-113."""
        b = self.st.batch
        lay, e = b.layouts[q], b.examples[q]
        idx = e['nmn_program_idx']
        use_head = self.model.config['have_pretrain_head']
        res = {}
        for i in range(len(lay.tokens) - 1, 0, -1):                 # i != 0: the root is trained by the decoder
            tok = lay.tokens[i]
            if tok in LY.OP_OF and idx[i] is not None and tok in self.model.pretrain_modules:
                nd = lay.node_of_token[i]
                res[idx[i]] = (tok, self.head_output(q, nd) if use_head else self.node_output(q, nd))
        return res

    def result_of_each_step(self, q):
        """module_net.py:115-131, list aligned to token order: (params, result)."""
        b = self.st.batch
        lay = b.layouts[q]
        use_head = self.model.config['have_pretrain_head']
        video = self.vid[q]
        out = []
        for i, tok in enumerate(lay.tokens):
            nd = lay.node_of_token[i]
            if tok in LY.OP_OF:
                params = []
                for pt in lay.param_tokens[i]:
                    ptok = lay.tokens[pt]
                    if ptok == 'video':
                        params.append(video)
                    elif lay.node_of_token[pt] < 0:
                        params.append(ptok)
                    else:
                        params.append(self.node_output(q, lay.node_of_token[pt]))
                if use_head and tok in self.model.pretrain_modules:
                    res = self.head_output(q, nd)
                else:
                    res = self.node_output(q, nd)
                if tok == 'Temporal':
                    self._last_temporal = self.att[int(self.aux_slot[int(b.node_start[q]) + nd])]
                out.append((params, res))
            elif nd < 0:
                out.append(([], tok))
            else:
                out.append(([], self.node_output(q, nd)))
        return out

    def last_temporal(self):
        return self._last_temporal
