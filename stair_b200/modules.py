"""The 18 NMN operators as classes with the reference's per-class ``forward(*params)`` surface (video_nmn/modules.py:7-465).

``NAME_TO_MODULE`` maps the module names to classes exactly like ``video_nmn/modules.py:446-465``; ``VideoNMN.__init__``
instantiates them the way ``module_net.py:27-35`` does (``Superlative`` receives the *same* ``Localize`` object, ``Filter`` /
``Superlative`` / ``ToAction`` the shared ``contrastive_head``).  The classes hold the parameters under the reference's
``state_dict`` names (the torch layers are parameter holders with the reference's default initialisation — their own
``forward`` is never used).

Arithmetic: inside ``VideoNMN.forward`` the operators run batched, grouped by type, in the CUDA interpreter
(csrc/executor.cu).  Calling an operator object directly — ``model.submodules['Localize'](feat, keyword)`` with the
reference's per-question shapes, or ``.forward_batched(...)`` with a leading instance axis on every tensor — packs the operands
into small arenas and runs that one group through ``stair_op_forward`` (the same group code and kernels as the interpreter).
There is no CPU path: operands must live on the CUDA device of the owning ``VideoNMN``.
"""
from __future__ import annotations

import ctypes
import weakref

import numpy as np
import torch
from torch import nn

from . import _lib as L
from . import layout as LY


def _seq(spec, p):
    """spec: list of ('lin', in, out) | 'relu' | 'drop' | 'sigmoid' | 'softmax' -> nn.Sequential with reference indices."""
    layers = []
    for s in spec:
        if isinstance(s, tuple):
            layers.append(nn.Linear(s[1], s[2]))
        elif s == 'relu':
            layers.append(nn.ReLU())
        elif s == 'drop':
            layers.append(nn.Dropout(p))
        elif s == 'sigmoid':
            layers.append(nn.Sigmoid())
        elif s == 'softmax':
            layers.append(nn.Softmax(dim=None))
    return nn.Sequential(*layers)


def _mlp2(i, H):
    return [('lin', i, H), 'relu', 'drop', ('lin', H, H), 'relu', 'drop']


# operand kinds per operator, in the reference's argument order (= pop order, module_net.py:100-106)
_PARAM_KINDS = {
    'And': ('va', 'va'), 'XorFrame': ('va', 'va'),
    'AttnVideo': ('vid', 'att1'),
    'Choose': ('vec', 'vec', 'vec'),
    'Compare': ('vec', 'vec'), 'Equals': ('vec', 'vec'), 'Xor': ('vec', 'vec'), 'ToAction': ('vec', 'vec'), 'Exists': ('vec', 'vec'),
    'Array2': ('vec', 'vec'),
    'ExistsFrame': ('vec', 'vid'),
    'Filter': ('vid', 'kw'), 'FilterFrame': ('vid', 'kw'),
    'HasItem': ('vid',),
    'Localize': ('vid', 'vecs'),
    'Relate': ('str', 'att1'),
    'Superlative': ('str', 'actions', 'vid'),
    'Temporal': ('str', 'vid', 'attK'),
}


class Operator(nn.Module):
    """Base class: parameter holder + the per-operator forward through ``stair_op_forward``."""
    op_name = None

    def __init__(self):
        super().__init__()

    # the owning VideoNMN (packed weights, precision, config) — a weak back-reference that is neither a submodule nor pickled
    def _bind(self, owner):
        self.__dict__['_owner_ref'] = weakref.ref(owner)

    def __getstate__(self):
        st = dict(self.__dict__)
        st.pop('_owner_ref', None)
        return st

    def _owner(self):
        ref = self.__dict__.get('_owner_ref')
        owner = ref() if ref is not None else None
        if owner is None:
            raise L.StairError('%s is not attached to a VideoNMN: operators execute on the CUDA library with the packed weights of '
                               'their model (construct stair_b200.VideoNMN and use model.submodules[name])' % self.op_name)
        return owner

    def forward(self, *params):
        """Reference shapes of ONE instance (video_nmn/modules.py ``forward`` of this class)."""
        out = self._run([p if isinstance(p, str) else p.unsqueeze(0) for p in params])
        return out[0]

    def forward_batched(self, *params):
        """Every tensor operand carries a leading instance axis n (keyword strings are shared by the group)."""
        return self._run(list(params))

    # ------------------------------------------------------------------------------------------------------------------
    def _run(self, params):
        name = self.op_name
        owner = self._owner()
        cfg = owner.config
        H = cfg['hidden_size']
        kinds = _PARAM_KINDS[name]
        if len(params) != len(kinds):
            raise TypeError('%s.forward() takes %d operands (%d given)' % (name, len(kinds), len(params)))
        tensors = [p for p in params if not isinstance(p, str)]
        if not tensors:
            raise TypeError('%s.forward() needs tensor operands' % name)
        for t in tensors:
            L.require_cuda(t, '%s operand' % name)
        dev = tensors[0].device
        n = int(tensors[0].shape[0])
        # frames: from the first [n, T, H] operand, else from an attention-map operand, else the model's max_video_length
        T = None
        for p, k in zip(params, kinds):
            if isinstance(p, str):
                continue
            if k == 'vid':
                T = int(p.shape[1])
                break
            is_vec = k == 'va' and tuple(p.shape[1:]) == (H,) and H != cfg['max_video_length']
            if T is None and k in ('att1', 'attK', 'va') and not is_vec:
                T = int(p.shape[-1])
        if T is None:
            T = int(cfg['max_video_length'])
        adt = owner.act_dtype
        # typed values for the layout compiler's resolver (same typing rules as the interpreter: layout.Layout._resolve)
        vals, packs = [], []                      # packs: (arena, units per instance, tensor [n, units, width])
        counts = {'vid': 0, 'vec': 0, 'att': 0}

        def push(arena, t, units, type_, K=1, rank2=False):
            vals.append(LY._Val(type_, node=len(packs), K=K, rank2=rank2))
            packs.append((arena, units, t.reshape(n, units, -1)))

        for p, k in zip(params, kinds):
            if isinstance(p, str):
                if k not in ('str', 'kw'):
                    raise TypeError('%s: operand must be a tensor, got the string %r' % (name, p))
                vals.append(LY._Val(LY.STR, text=(p, -1)))
                continue
            if p.shape[0] != n:
                raise ValueError('%s: operands disagree on the number of instances' % name)
            inst = tuple(p.shape[1:])
            if k == 'str':
                raise TypeError('%s: operand must be a keyword string' % name)
            if k == 'vid':
                if inst != (T, H):
                    raise ValueError('%s: expected frame features [%d, %d], got %s' % (name, T, H, inst))
                push('vid', p, 1, LY.VID)
            elif k in ('vec', 'kw'):
                if inst != (H,):
                    raise ValueError('%s: expected a [%d] vector, got %s' % (name, H, inst))
                push('vec', p, 1, LY.VEC)
            elif k == 'vecs':
                if inst == (H,):
                    push('vec', p, 1, LY.VEC)
                elif inst == (2, H):
                    push('vec', p, 2, LY.VEC2)
                else:
                    raise ValueError('%s: keyword must be [%d] or [2, %d], got %s' % (name, H, H, inst))
            elif k == 'actions':
                if inst == (H,):
                    push('vec', p, 1, LY.VEC)
                elif inst == (2, H) and T != 2:
                    push('vec', p, 2, LY.VEC2)
                elif inst == (T, H):
                    push('vid', p, 1, LY.VID)
                else:
                    raise ValueError('%s: actions must be [%d], [2, %d] or [%d, %d], got %s' % (name, H, H, T, H, inst))
            elif k == 'att1':
                if inst != (T,):
                    raise ValueError('%s: expected a [%d] attention map, got %s' % (name, T, inst))
                push('att', p, 1, LY.ATT, 1)
            elif k == 'attK':
                if len(inst) != 2 or inst[1] != T or inst[0] not in (1, 2):
                    raise ValueError('%s: expected a [K, %d] attention map (K = 1 or 2), got %s' % (name, T, inst))
                push('att', p, inst[0], LY.ATT, inst[0], rank2=True)
            elif k == 'va':
                if inst == (H,) and H != T:
                    push('vec', p, 1, LY.VEC)
                elif inst == (T,):
                    push('att', p, 1, LY.ATT, 1)
                elif len(inst) == 2 and inst[1] == T:
                    push('att', p, inst[0], LY.ATT, inst[0], rank2=True)
                elif inst == (H,):
                    push('vec', p, 1, LY.VEC)
                else:
                    raise ValueError('%s: operands must be [%d] vectors or attention maps, got %s' % (name, H, inst))
        variant, arg_nodes, out_type, out_K = LY.Layout._resolve(name, vals)
        op = LY.OP_OF[name]
        out_arena, out_mult = LY._out_units(op, variant, T)
        # arenas: inputs first (instance-major per operand), then the group's outputs (+ Temporal's stash)
        base = {}
        for j, (arena, units, _) in enumerate(packs):
            base[j] = counts[arena]
            counts[arena] += n * units
        out_base = counts[out_arena]
        counts[out_arena] += n * out_mult
        aux_base = -1
        if name == 'Temporal':
            aux_base = counts['att']
            counts['att'] += n
        vid = torch.zeros((max(counts['vid'], 1), T, H), dtype=adt, device=dev)
        vec = torch.zeros((max(counts['vec'], 1), H), dtype=adt, device=dev)
        att = torch.zeros((max(counts['att'], 1), T), dtype=torch.float32, device=dev)
        store = {'vid': vid, 'vec': vec, 'att': att}
        for j, (arena, units, t) in enumerate(packs):
            dst = store[arena][base[j]:base[j] + n * units]
            dst.copy_(t.detach().reshape(dst.shape))
        args = np.full((3, n), -1, np.int32)
        for k, node in enumerate(arg_nodes):
            arena, units, _ = packs[node]
            args[k] = base[node] + units * np.arange(n, dtype=np.int32)
        args_dev = torch.from_numpy(args.reshape(-1)).to(dev)
        model = owner._packed.refresh(owner.submodules, cfg, owner._precision_code(), dev)
        g = L.StairGroup()
        g.op, g.variant, g.level, g.count, g.node_off = op, variant, 1, n, 0
        g.out_base, g.out_mult, g.aux_base, g.head = out_base, out_mult, aux_base, 0
        lib = L.lib()
        ws_bytes = int(lib.stair_op_workspace_bytes(ctypes.byref(model), L.i32(T), ctypes.byref(g)))
        if ws_bytes < 0:
            raise L.StairError('%s: unsupported configuration (T = %d, max_video_length = %d)' % (name, T, cfg['max_video_length']))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        bufs = L.StairBuffers()
        bufs.vid, bufs.vid_slots = vid.data_ptr(), vid.shape[0]
        bufs.vec, bufs.vec_rows = vec.data_ptr(), vec.shape[0]
        bufs.att, bufs.att_rows = att.data_ptr(), att.shape[0]
        bufs.workspace, bufs.workspace_bytes = ws.data_ptr(), ws.numel()
        L.check(lib.stair_op_forward(ctypes.byref(model), L.i32(T), ctypes.byref(g), L.ptr(args_dev), ctypes.byref(bufs), L.stream_ptr()),
                'stair_op_forward(%s)' % name)
        if name == 'Temporal':
            rel = att[aux_base:aux_base + n]
            self.related_attn = rel[-1].clone() if n == 1 else rel.clone()       # modules.py:288,321-325 (stash of the last forward)
        out = store[out_arena][out_base:out_base + n * out_mult]
        if out_type == LY.VID:
            return out.clone()                                                   # [n, T, H]
        if out_type == LY.VEC:
            return out.clone()                                                   # [n, H]
        if out_type == LY.VEC2:
            return out.reshape(n, 2, H).clone()
        if out_K > 1 or LY.Layout._att_rank2(name, vals, out_type, out_K):
            return out.reshape(n, out_K, T).clone()                              # [n, K, T]
        return out.clone()                                                       # [n, T]


class AndModule(Operator):                                          # modules.py:7-12
    op_name = 'And'

    def __init__(self, config):
        super().__init__()


class AttnVideoModule(Operator):                                    # modules.py:330-340
    op_name = 'AttnVideo'

    def __init__(self, config):
        super().__init__()


class ChooseModule(Operator):                                       # modules.py:40-56
    op_name = 'Choose'

    def __init__(self, config):
        super().__init__()


class CompareModule(Operator):                                      # modules.py:15-21
    op_name = 'Compare'

    def __init__(self, config):
        super().__init__()
        H = config['hidden_size']
        self.param = _seq([('lin', 2 * H, H), 'relu'], config['dropout'])


class EqualsModule(Operator):                                       # modules.py:24-37
    op_name = 'Equals'

    def __init__(self, config):
        super().__init__()
        H = config['hidden_size']
        self.param = _seq([('lin', 2 * H, H), 'relu'], config['dropout'])
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Linear(H, 1)


class ExistsModule(Operator):                                       # modules.py:141-159
    op_name = 'Exists'

    def __init__(self, config):
        super().__init__()
        H = config['hidden_size']
        self.param = _seq(_mlp2(3 * H, H), config['dropout'])
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Linear(H, 2)


class ExistsFrameModule(Operator):                                  # modules.py:162-178
    op_name = 'ExistsFrame'

    def __init__(self, config):
        super().__init__()
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Identity()


class FilterModule(Operator):                                       # modules.py:343-378
    op_name = 'Filter'

    def __init__(self, config, contrastive_head):
        super().__init__()
        H, p = config['hidden_size'], config['dropout']
        self.param = nn.ModuleDict({kw: _seq(_mlp2(H, H), p) for kw in ['representation', 'actions', 'objects', 'relations']})
        self.attention = _seq([('lin', 2 * H, 1), 'softmax'], p)
        self.dense = _seq([('lin', H, H), 'relu'], p)
        if config['have_pretrain_head']:
            self.pretrain_head = contrastive_head


class FilterFrameModule(Operator):                                  # modules.py:381-414
    op_name = 'FilterFrame'

    def __init__(self, config):
        super().__init__()
        H, p = config['hidden_size'], config['dropout']
        self.param = nn.ModuleDict({kw: _seq(_mlp2(H, H), p) for kw in ['representation', 'relations', 'actions']})
        self.attention = _seq([('lin', 2 * H, 1), 'sigmoid'], p)
        self.dense = _seq([('lin', H, H), 'relu', 'drop'], p)
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Linear(H, config['object_types'])


class HasItemModule(Operator):                                      # modules.py:123-138
    op_name = 'HasItem'

    def __init__(self, config):
        super().__init__()
        H, p = config['hidden_size'], config['dropout']
        self.param = _seq([('lin', H, H), 'relu', 'drop', ('lin', H, 1), 'sigmoid', 'drop'], p)
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Identity()


class LocalizeModule(Operator):                                     # modules.py:181-217
    op_name = 'Localize'

    def __init__(self, config):
        super().__init__()
        H, p = config['hidden_size'], config['dropout']
        self.video_linear = _seq([('lin', H, H), 'relu', 'drop', ('lin', H, H)], p)
        self.keyword_linear = _seq([('lin', H, H)], p)
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Identity()


class RelateModule(Operator):                                       # modules.py:417-435
    op_name = 'Relate'

    def __init__(self, config):
        super().__init__()
        self.beta = nn.Parameter(torch.rand(config['max_video_length']))


class SuperlativeModule(Operator):                                  # modules.py:220-248 (shares the Localize object, module_net.py:31-32)
    op_name = 'Superlative'

    def __init__(self, config, localize_module, contrastive_head):
        super().__init__()
        H = config['hidden_size']
        self.localize_module = localize_module
        self.dense = _seq([('lin', H, H), 'relu'], config['dropout'])
        if config['have_pretrain_head']:
            self.pretrain_head = contrastive_head


class TemporalModule(Operator):                                     # modules.py:251-327
    """``pretrain_head()`` returns the related attention stashed by the last forward that executed a Temporal module
    (video_nmn/modules.py:287-288, 321-325)."""
    op_name = 'Temporal'

    def __init__(self, config):
        super().__init__()
        H, p, T = config['hidden_size'], config['dropout'], config['max_video_length']
        if T > 32:                                                   # modules.py:255-266
            k = round(T / 4)
            relate = {m: nn.Sequential(nn.Conv1d(1, 1, k, padding='same'), nn.ReLU(), nn.Conv1d(1, 1, k, padding='same'), nn.ReLU(),
                                       nn.Conv1d(1, 1, 2 * k + 1, padding='same'), nn.Sigmoid()) for m in ['before', 'after', 'between']}
        else:                                                        # modules.py:267-277
            relate = {m: _seq([('lin', T, T), 'relu', ('lin', T, T), 'relu', ('lin', T, T), 'sigmoid'], p)
                      for m in ['before', 'after', 'between']}
        self.relate = nn.ModuleDict(relate)
        self.relate['while'] = nn.Identity()
        self.dense = _seq([('lin', H, H), 'relu', 'drop'], p)
        self.layer_norm = nn.LayerNorm(H)
        self.related_attn = None

    def pretrain_head(self, *args):
        return self.related_attn

    def relate_(self, attention_scores, mode):
        """modules.py:290-308 — the cumsum before / after / between masks (dead code in the reference forward), on the device
        through ``stair_relate_scan``.  attention_scores: [T] / [1, T], or [2, T] for 'between'; returns [T]."""
        L.require_cuda(attention_scores, 'attention_scores')
        a = attention_scores.detach().to(torch.float32).contiguous()
        T = int(a.shape[-1])
        K = 2 if mode == 'between' else 1
        if a.numel() != K * T:
            raise ValueError("relate_: mode %r expects %d x %d scores, got %s" % (mode, K, T, tuple(a.shape)))
        out = torch.empty(T, dtype=torch.float32, device=a.device)
        L.check(L.lib().stair_relate_scan(L.ptr(a), L.i32(LY.TEMPORAL_MODES[mode]), L.ptr(out), L.i32(1), L.i32(T), L.stream_ptr()),
                'stair_relate_scan')
        return out


class ToActionModule(Operator):                                     # modules.py:102-120
    op_name = 'ToAction'

    def __init__(self, config, contrastive_head):
        super().__init__()
        H, p = config['hidden_size'], config['dropout']
        self.param = _seq([('lin', 2 * H, H), 'relu', 'drop', ('lin', H, H), 'relu'], p)
        if config['have_pretrain_head']:
            self.pretrain_head = contrastive_head


class XorModule(Operator):                                          # modules.py:59-72
    op_name = 'Xor'

    def __init__(self, config):
        super().__init__()
        H = config['hidden_size']
        self.param = _seq([('lin', 3 * H, H), 'relu'], config['dropout'])
        if config['have_pretrain_head']:
            self.pretrain_head = nn.Linear(H, 2)


class XorFrameModule(Operator):                                     # modules.py:75-80
    op_name = 'XorFrame'

    def __init__(self, config):
        super().__init__()


class Array2Module(Operator):                                       # modules.py:438-443
    op_name = 'Array2'

    def __init__(self, config):
        super().__init__()


# video_nmn/modules.py:446-465 — same names, same order
NAME_TO_MODULE = {
    'And': AndModule,
    'AttnVideo': AttnVideoModule,
    'Choose': ChooseModule,
    'Compare': CompareModule,
    'Equals': EqualsModule,
    'Exists': ExistsModule,
    'ExistsFrame': ExistsFrameModule,
    'Filter': FilterModule,
    'FilterFrame': FilterFrameModule,
    'HasItem': HasItemModule,
    'Localize': LocalizeModule,
    'Relate': RelateModule,
    'Superlative': SuperlativeModule,
    'Temporal': TemporalModule,
    'ToAction': ToActionModule,
    'Xor': XorModule,
    'XorFrame': XorFrameModule,
    'Array2': Array2Module,
}
