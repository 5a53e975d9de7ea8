"""Parameter tree of the drop-in ``VideoNMN`` and its packing into the C-ABI weight table.

The tree reproduces the reference ``state_dict`` exactly (119 keys such as
``submodules.Localize.video_linear.{0,3}.weight``, ``submodules.Temporal.relate.before.{0,2,4}.*``,
``submodules.video_encoder.weight_ih_l0_reverse``), so ``load_state_dict`` (evaluate.py:139) and pickled-module
loading (train_module.py:296-298) work unchanged.  The operator classes live in ``stair_b200/modules.py`` (one class per
module + ``NAME_TO_MODULE``, like video_nmn/modules.py); the torch layers are *parameter holders with the reference's default
initialisation* — their ``forward`` is never used; all arithmetic runs in the CUDA library (csrc/).

Reference for the layer shapes: video_nmn/modules.py:15-443 and video_nmn/module_net.py:39-53.
"""
from __future__ import annotations

import torch
from torch import nn

from . import _lib as L


class L2Normalize(nn.Module):
    """``contrastive_head`` (video_nmn/module_net.py:211-216): x / max(|x|_2, 1e-12) over dim 0."""

    def forward(self, feat):
        L.require_cuda(feat, 'feat')
        x = feat.detach().contiguous()
        out = torch.empty(x.shape, device=x.device, dtype=torch.float32)
        L.check(L.lib().stair_l2normalize(L.i32(L.dtype_code(x.dtype)), L.ptr(x), L.ptr(out), L.i32(1), L.i32(x.numel()),
                                          L.stream_ptr()), 'stair_l2normalize')
        return out


def build_submodules(config, contrastive_head):
    """nn.ModuleDict in NAME_TO_MODULE order + encoders + decoder, constructed like module_net.py:27-53: ``Superlative`` receives the
    same ``Localize`` object, ``Filter`` / ``Superlative`` / ``ToAction`` the shared ``contrastive_head``."""
    from .modules import NAME_TO_MODULE, _seq
    H, p = config['hidden_size'], config['dropout']
    sub = nn.ModuleDict()
    for name, cls in NAME_TO_MODULE.items():
        init_params = [config]
        if name in ['Superlative']:
            init_params.append(sub['Localize'])
        if name in ['Filter', 'Superlative', 'ToAction']:
            init_params.append(contrastive_head)
        sub[name] = cls(*init_params)
    sub['video_encoder'] = nn.LSTM(input_size=config['video_size'], hidden_size=H // 2, batch_first=True, bidirectional=True)
    sub['text_encoder'] = nn.LSTM(input_size=config['text_size'], hidden_size=H // 2, batch_first=True, bidirectional=True)
    sub['decoder'] = _seq([('lin', 2 * H, 2 * H), 'relu', 'drop', ('lin', 2 * H, config['answer_vocab_length'])], p)
    return sub


# ---------------------------------------------------------------------------------------------------------------------
# packing: state_dict tensors -> the StairWeight table (include/stair_b200.h)
# ---------------------------------------------------------------------------------------------------------------------
def _pad8(n):
    return (n + 7) // 8 * 8


def _matrix(w, nplanes):
    """fp32 [N,K] -> bf16 [nplanes][N, K_ld] (K zero-padded to a multiple of 8; 3 planes = exact bf16x3 split)."""
    N, K = w.shape
    ld = _pad8(K)
    x = torch.zeros((N, ld), device=w.device, dtype=torch.float32)
    x[:, :K] = w.detach().float()
    if nplanes == 1:
        return x.to(torch.bfloat16).contiguous()
    p0 = x.to(torch.bfloat16)
    r = x - p0.float()
    p1 = r.to(torch.bfloat16)
    r = r - p1.float()
    return torch.stack([p0, p1, r.to(torch.bfloat16)]).contiguous()


def _vector(v):
    return v.detach().float().contiguous().reshape(-1)


def _matrix_t(w, nplanes):
    """fp32 [N,K] -> transposed bf16 planes [nplanes][K, N_ld] (B operand of the dX = dZ.W GEMM of the backward pass)."""
    return _matrix(w.detach().float().t().contiguous(), nplanes)


# GEMM matrices whose layer input needs no gradient (encoder input projections): no transposed copy
NO_DX = ('VENC_WIH', 'TENC_WIH')


def weight_sources(sub, config):
    """{STAIR_W_* name: (kind, callable -> fp32 tensor)}; kind 'M' = GEMM matrix (bf16 planes), 'V' = fp32 vector."""
    s = {}

    def lin(prefix, layer, mat='M'):
        s[prefix + '_W'] = (mat, lambda: layer.weight)
        s[prefix + '_B'] = ('V', lambda: layer.bias)

    def lstm(prefix, m):
        s[prefix + '_WIH'] = ('M', lambda: torch.cat([m.weight_ih_l0, m.weight_ih_l0_reverse], 0))
        s[prefix + '_B'] = ('V', lambda: torch.cat([m.bias_ih_l0 + m.bias_hh_l0, m.bias_ih_l0_reverse + m.bias_hh_l0_reverse]))
        s[prefix + '_WHH_F'] = ('M', lambda: m.weight_hh_l0)
        s[prefix + '_WHH_R'] = ('M', lambda: m.weight_hh_l0_reverse)
        hh = m.hidden_size
        if hh % 64 == 0:                        # gate-interleaved copy for the fused recurrence kernel (csrc/lstm_fused.cu)
            il = lambda w: w.reshape(4, hh // 64, 64, hh).permute(1, 0, 2, 3).reshape(4 * hh, hh)        # noqa: E731
            s[prefix + '_WHHI_F'] = ('M1', lambda: il(m.weight_hh_l0))
            s[prefix + '_WHHI_R'] = ('M1', lambda: il(m.weight_hh_l0_reverse))

    def mlp4(prefix, seq):                      # 4 consecutive slots w0,b0,w1,b1 (Sequential indices 0 and 3)
        base = L.W[prefix]
        for j, (kind, layer, attr) in enumerate((('M', seq[0], 'weight'), ('V', seq[0], 'bias'), ('M', seq[3], 'weight'), ('V', seq[3], 'bias'))):
            s[base + j] = (kind, (lambda l=layer, a=attr: getattr(l, a)))

    lstm('VENC', sub['video_encoder'])
    lstm('TENC', sub['text_encoder'])
    lin('DEC0', sub['decoder'][0]); lin('DEC1', sub['decoder'][3])
    loc = sub['Localize']
    lin('LOC_V0', loc.video_linear[0]); lin('LOC_V1', loc.video_linear[3]); lin('LOC_K', loc.keyword_linear[0])
    tmp = sub['Temporal']
    lin('TEMP_D', tmp.dense[0])
    s['TEMP_LN_G'] = ('V', lambda: tmp.layer_norm.weight)
    s['TEMP_LN_B'] = ('V', lambda: tmp.layer_norm.bias)
    for mode in ('BEFORE', 'AFTER', 'BETWEEN'):
        seq = tmp.relate[mode.lower()]
        base = L.W['TEMP_REL_' + mode]
        for j, layer in enumerate((seq[0], seq[2], seq[4])):
            s[base + 2 * j] = ('V', (lambda l=layer: l.weight))
            s[base + 2 * j + 1] = ('V', (lambda l=layer: l.bias))
    flt = sub['Filter']
    for kind, slot in (('representation', 'FILT_REPR'), ('actions', 'FILT_ACTIONS'), ('objects', 'FILT_OBJECTS'), ('relations', 'FILT_RELATIONS')):
        mlp4(slot, flt.param[kind])
    lin('FILT_D', flt.dense[0])
    ff = sub['FilterFrame']
    for kind, slot in (('representation', 'FF_REPR'), ('relations', 'FF_RELATIONS'), ('actions', 'FF_ACTIONS')):
        mlp4(slot, ff.param[kind])
    lin('FF_ATT', ff.attention[0], mat='V')
    lin('FF_D', ff.dense[0])
    if config['have_pretrain_head']:
        lin('FF_HEAD', ff.pretrain_head)
        lin('EQUALS_HEAD', sub['Equals'].pretrain_head, mat='V')
        lin('XOR_HEAD', sub['Xor'].pretrain_head, mat='V')
        lin('EXISTS_HEAD', sub['Exists'].pretrain_head, mat='V')
    has = sub['HasItem']
    lin('HAS0', has.param[0]); lin('HAS1', has.param[3], mat='V')
    s['REL_BETA'] = ('V', lambda: sub['Relate'].beta)
    lin('SUP_D', sub['Superlative'].dense[0])
    lin('COMPARE', sub['Compare'].param[0])
    lin('EQUALS', sub['Equals'].param[0])
    lin('XOR', sub['Xor'].param[0])
    lin('EXISTS0', sub['Exists'].param[0]); lin('EXISTS1', sub['Exists'].param[3])
    lin('TOACT0', sub['ToAction'].param[0]); lin('TOACT1', sub['ToAction'].param[3])
    return {(L.W[k] if isinstance(k, str) else k): v for k, v in s.items()}


def grad_targets(sub, config):
    """{STAIR_W_* id: (numel, [(parameter, element offset inside the slot), ...])} — where the fp32 gradient accumulator of
    each weight-table slot (StairTrain.grad, logical [N, K] / vector shapes) lands in the reference's parameter tree.
    The two LSTM biases of a direction share one slot (the kernels see b_ih + b_hh), so both receive the same gradient."""
    t = {}

    def one(key, p):
        t[L.W[key] if isinstance(key, str) else key] = (p.numel(), [(p, 0)])

    def lin(prefix, layer):
        one(prefix + '_W', layer.weight)
        one(prefix + '_B', layer.bias)

    def lstm(prefix, m):
        n = m.weight_ih_l0.numel()
        t[L.W[prefix + '_WIH']] = (2 * n, [(m.weight_ih_l0, 0), (m.weight_ih_l0_reverse, n)])
        nb = m.bias_ih_l0.numel()
        t[L.W[prefix + '_B']] = (2 * nb, [(m.bias_ih_l0, 0), (m.bias_hh_l0, 0), (m.bias_ih_l0_reverse, nb), (m.bias_hh_l0_reverse, nb)])
        one(prefix + '_WHH_F', m.weight_hh_l0)
        one(prefix + '_WHH_R', m.weight_hh_l0_reverse)

    def mlp4(prefix, seq):
        base = L.W[prefix]
        for j, p in enumerate((seq[0].weight, seq[0].bias, seq[3].weight, seq[3].bias)):
            one(base + j, p)

    lstm('VENC', sub['video_encoder'])
    lstm('TENC', sub['text_encoder'])
    lin('DEC0', sub['decoder'][0]); lin('DEC1', sub['decoder'][3])
    loc = sub['Localize']
    lin('LOC_V0', loc.video_linear[0]); lin('LOC_V1', loc.video_linear[3]); lin('LOC_K', loc.keyword_linear[0])
    tmp = sub['Temporal']
    lin('TEMP_D', tmp.dense[0])
    one('TEMP_LN_G', tmp.layer_norm.weight); one('TEMP_LN_B', tmp.layer_norm.bias)
    for mode in ('BEFORE', 'AFTER', 'BETWEEN'):
        seq = tmp.relate[mode.lower()]
        base = L.W['TEMP_REL_' + mode]
        for j, layer in enumerate((seq[0], seq[2], seq[4])):
            one(base + 2 * j, layer.weight)
            one(base + 2 * j + 1, layer.bias)
    flt = sub['Filter']
    for kind, slot in (('representation', 'FILT_REPR'), ('actions', 'FILT_ACTIONS'), ('objects', 'FILT_OBJECTS'), ('relations', 'FILT_RELATIONS')):
        mlp4(slot, flt.param[kind])
    lin('FILT_D', flt.dense[0])
    ff = sub['FilterFrame']
    for kind, slot in (('representation', 'FF_REPR'), ('relations', 'FF_RELATIONS'), ('actions', 'FF_ACTIONS')):
        mlp4(slot, ff.param[kind])
    lin('FF_ATT', ff.attention[0])
    lin('FF_D', ff.dense[0])
    if config['have_pretrain_head']:
        lin('FF_HEAD', ff.pretrain_head)                    # only receives a gradient when FilterFrame is supervised (off by default)
        lin('EQUALS_HEAD', sub['Equals'].pretrain_head)
        lin('XOR_HEAD', sub['Xor'].pretrain_head)
        lin('EXISTS_HEAD', sub['Exists'].pretrain_head)
    has = sub['HasItem']
    lin('HAS0', has.param[0]); lin('HAS1', has.param[3])
    one('REL_BETA', sub['Relate'].beta)
    lin('SUP_D', sub['Superlative'].dense[0])
    lin('COMPARE', sub['Compare'].param[0])
    lin('EQUALS', sub['Equals'].param[0])
    lin('XOR', sub['Xor'].param[0])
    lin('EXISTS0', sub['Exists'].param[0]); lin('EXISTS1', sub['Exists'].param[3])
    lin('TOACT0', sub['ToAction'].param[0]); lin('TOACT1', sub['ToAction'].param[3])
    return t


def all_parameters(root):
    """``list(root.parameters())`` without the name bookkeeping of ``nn.Module.named_parameters`` (a 10x cheaper walk: this runs
    several times per step to detect changed weights).  Same order, shared modules / parameters listed once.  The module tree is
    walked once and its ``_parameters`` dicts are remembered (they are live: a replaced Parameter is seen; submodules added to an
    existing model afterwards are not — call ``all_parameters(root, rescan=True)``)."""
    return _walk_parameters(root, False)


def _walk_parameters(root, rescan):
    dicts = None if rescan else root.__dict__.get('_stair_param_dicts')
    if dicts is None:
        dicts, seen_m, stack = [], set(), [root]
        while stack:
            m = stack.pop()
            if id(m) in seen_m:
                continue
            seen_m.add(id(m))
            if m._parameters:
                dicts.append(m._parameters)
            stack.extend(reversed([c for c in m._modules.values() if c is not None]))
        root.__dict__['_stair_param_dicts'] = dicts
    out = [p for d in dicts for p in d.values() if p is not None]
    if len(set(map(id, out))) != len(out):                  # tied parameters: list each once, first occurrence wins
        seen, uniq = set(), []
        for p in out:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        out = uniq
    return out


class PackedWeights:
    """Device copies of the weights in the layout the kernels read, rebuilt when a parameter changes."""

    def __init__(self):
        self.signature = None
        self.tensors = {}
        self.transposed = {}
        self.want_transposed = False
        self.model_struct = None
        self.version = 0

    def mark_current(self, sub, precision, device):
        """The packed copies were just rewritten from the parameters by ``stair_adam_multi`` (train.FusedAdam)."""
        self.signature = (precision, str(device)) + tuple((p.data_ptr(), p._version) for p in all_parameters(sub))

    def refresh(self, sub, config, precision, device, training=False):
        params = all_parameters(sub)
        sig = (precision, str(device)) + tuple((p.data_ptr(), p._version) for p in params)
        training = training or self.want_transposed          # once a model trains, keep the transposed copies current
        if sig == self.signature and (not training or self.transposed):
            return self.model_struct
        self.want_transposed = training
        nplanes = 3 if precision == L.F32 else 1
        srcs = weight_sources(sub, config)
        tensors, transposed = {}, {}
        no_dx = {L.W[k] for k in NO_DX}
        with torch.no_grad():
            for wid, (kind, get) in srcs.items():
                w = get()
                if w.device != device:
                    raise L.StairError('model parameters live on %s but the batch is on %s; call model.to(device)' % (w.device, device))
                if kind == 'M':
                    tensors[wid] = _matrix(w.reshape(w.shape[0], -1), nplanes)
                    if training and wid not in no_dx:
                        transposed[wid] = _matrix_t(w.reshape(w.shape[0], -1), nplanes)
                elif kind == 'M1':
                    tensors[wid] = _matrix(w.reshape(w.shape[0], -1), 1)
                else:
                    tensors[wid] = _vector(w)
        m = L.StairModel()
        T = config['max_video_length']
        m.T_max, m.V, m.V_ld, m.H = T, config['video_size'], _pad8(config['video_size']), config['hidden_size']
        m.text_size, m.text_ld = config['text_size'], _pad8(config['text_size'])
        m.A, m.O = config['answer_vocab_length'], config.get('object_types', 0) or 0
        m.conv_k = round(T / 4) if T > 32 else 0
        m.precision = precision
        for wid in range(L.W_COUNT):
            m.w[wid] = tensors[wid].data_ptr() if wid in tensors else None
            m.wt[wid] = transposed[wid].data_ptr() if wid in transposed else None
        self.tensors, self.transposed, self.model_struct, self.signature = tensors, transposed, m, sig
        self.version += 1                                   # generation of the packed copies (consumers cache raw pointers into them)
        return m
