"""ORACLE — CPU restatement of STAIR's video_nmn ModuleNet hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped path
(``stair_b200``) never routes through it and has no CPU fallback.

What it restates (plain PyTorch fp32 on CPU, functional, one question at a time exactly like the
reference — the reference has no batch axis, SURVEY.md "three things" #2):

  * ``OracleNMN.forward``         <- video_nmn/module_net.py:65-145 (reverse-prefix stack interpreter)
  * ``OracleNMN.encode_video``    <- video_nmn/module_net.py:160-163 + nn.LSTM(bidirectional) :39-42
  * ``OracleNMN.encode_question`` <- video_nmn/module_net.py:151-158 + :44-47
  * the 18 operators              <- video_nmn/modules.py:7-465 (each method cites its lines)
  * layout helpers                <- utils/program_parser.py:16-23,182-200,307-333
  * ``OracleCriterion``           <- train_module.py:33-194 (intermediate-supervision losses)
  * ``window_loss``               <- train_module.py:341-406 (one gradient-accumulation window)

Third-party arithmetic: everything the reference computes goes through PyTorch ATen (reference pins
torch==1.13, requirements.txt; this image has 2.11).  The ATen semantics this file restates by hand:
LSTM gate order i,f,g,o with b_ih+b_hh; cosine_similarity = x·y / (max(|x|,eps)·max(|y|,eps)), eps=1e-8;
F.normalize eps=1e-12; legacy implicit softmax dim (ndim in {0,1,3} -> 0 else 1);
Conv1d(padding='same') even-kernel split left=(k-1)//2, right=k-1-left; LayerNorm eps=1e-5, biased var.

PARITY PIN: the reference ships no tests or golden vectors for this path, so the pin is
the reference itself: ``tests/golden/make_golden.py`` imports the unmodified reference from
``/root/reference`` in the build container, runs it on seeded synthetic inputs and commits inputs, weights
and outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file against those
fixtures (logits, every intermediate, res_by_step, losses, gradients, layout helpers).
"""
from __future__ import annotations

import math
from typing import Dict, List

import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------------------
# layout semantics (integer side) — utils/program_parser.py
# --------------------------------------------------------------------------------------------------
# utils/program_parser.py:16-23 (nary_mappings)
NARY = {**{n: 1 for n in ['Array1', 'HasItem', 'OnlyItem', 'Query']},
        **{n: 2 for n in ['Array2', 'AND', 'XOR', 'And', 'Xor', 'Compare', 'Equals', 'Exists', 'Filter', 'Iterate',
                          'Localize', 'ToAction', 'Relate', 'AttnVideo', 'FilterFrame', 'ExistsFrame', 'XorFrame']},
        **{n: 3 for n in ['Array3', 'Superlative', 'Choose', 'Temporal']},
        **{n: 4 for n in ['IterateUntil']}}

# video_nmn/modules.py:446-465 (NAME_TO_MODULE keys, in order)
MODULE_NAMES = ['And', 'AttnVideo', 'Choose', 'Compare', 'Equals', 'Exists', 'ExistsFrame', 'Filter', 'FilterFrame',
                'HasItem', 'Localize', 'Relate', 'Superlative', 'Temporal', 'ToAction', 'Xor', 'XorFrame', 'Array2']
# video_nmn/dataset.py:23 | video_nmn/module_net.py:23-25
WORDS_TO_KEEP = {'forward', 'backward', 'while', 'between', 'before', 'after', 'max', 'min', 'start', 'end', 'video',
                 'actions', 'objects', 'relations'}


def children_and_parents(tokens: List[str]):
    """utils/program_parser.py:182-200 — children (in pop order) and parent index of every token."""
    children = [[] for _ in tokens]
    parents = [0 for _ in tokens]
    stack = []
    for i in range(len(tokens) - 1, -1, -1):
        if tokens[i] in NARY:
            for _ in range(NARY[tokens[i]]):
                children[i].append(stack.pop())
        stack.append(i)
    for i, chs in enumerate(children):
        for c in chs:
            parents[c] = i
    return children, parents


def module_levels(tokens: List[str]):
    """utils/program_parser.py:307-321 — leaf = 0, module = 1 + max(children levels)."""
    levels, stack = [], []
    for tok in reversed(tokens):
        if tok not in NARY:
            stack.append(0)
            levels.append(0)
        else:
            n = NARY[tok]
            params, stack = stack[-n:], stack[:-n]
            lvl = max(params) + 1
            stack.append(lvl)
            levels.append(lvl)
    return levels[::-1]


def program_is_valid(tokens: List[str]) -> bool:
    """utils/program_parser.py:324-333."""
    depth = 0
    for tok in reversed(tokens):
        depth = depth - NARY[tok] + 1 if tok in NARY else depth + 1
        if depth < 0:
            return False
    return depth == 1


# --------------------------------------------------------------------------------------------------
# small ATen restatements
# --------------------------------------------------------------------------------------------------
def _linear(x, w, b):
    return x @ w.t() + b


def _cos(x, y, eps=1e-8):
    """nn.CosineSimilarity(dim=-1): normalise each side (norm clamped at eps), then dot."""
    xn = x / x.norm(dim=-1, keepdim=True).clamp_min(eps)
    yn = y / y.norm(dim=-1, keepdim=True).clamp_min(eps)
    return (xn * yn).sum(-1)


def _legacy_softmax(x):
    """nn.Softmax() with dim=None: dim = 0 if ndim in (0,1,3) else 1 (torch.nn.functional._get_softmax_dim)."""
    dim = 0 if x.dim() in (0, 1, 3) else 1
    return torch.softmax(x, dim=dim)


def _conv1d_same(x, w, b):
    """Conv1d(1,1,k,padding='same',zeros) on a [T] signal; even k pads (k-1)//2 left, k-1-(k-1)//2 right."""
    k = w.numel()
    left = (k - 1) // 2
    xp = F.pad(x.view(1, 1, -1), (left, k - 1 - left))
    return F.conv1d(xp, w.view(1, 1, k), b.view(1)).view(-1)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM (1 layer): gates i,f,g,o; returns outputs [L,h] in input order and final h."""
    L = x.size(0)
    h = x.new_zeros(w_hh.size(1))
    c = x.new_zeros(w_hh.size(1))
    pre = x @ w_ih.t() + (b_ih + b_hh)
    outs = [None] * L
    order = range(L - 1, -1, -1) if reverse else range(L)
    for t in order:
        g = pre[t] + w_hh @ h
        i, f, gg, o = g.chunk(4)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs[t] = h
    return torch.stack(outs), h


class OracleNMN:
    """Functional fp32 restatement of ``VideoNMN`` over a reference-keyed ``state_dict``.

    ``weights`` uses the reference's state_dict key names, e.g.
    ``submodules.Localize.video_linear.0.weight``.  ``Superlative.localize_module.*`` aliases ``Localize.*``
    (module_net.py:31-32) and is read from the ``Localize`` keys.
    """

    def __init__(self, config: dict, weights: Dict[str, torch.Tensor], pretrain_modules=frozenset(), aten_lstm=False, dropout=None):
        """``aten_lstm=True`` runs the two encoders through ``torch.nn.LSTM`` itself — the very ATen call the reference
        makes (module_net.py:39-47,151-163; mkldnn/cuDNN kernels) — instead of the explicit per-step restatement.  Used by
        the CPU-baseline timer so the baseline is not handicapped by a Python time loop; tested equal to the loop."""
        self.aten_lstm = aten_lstm
        # training-mode nn.Dropout sites: ``dropout(site, x, (question_index, token_index)) -> tensor`` is called wherever the
        # reference has an nn.Dropout (site = the Sequential entry it follows, e.g. 'Localize.video_linear.0'); None = eval
        self.dropout = dropout
        self._q, self._tok = 0, None
        self._lstm_cache = {}
        self.config = config
        self.W = weights
        self.pretrain_modules = set(pretrain_modules)
        self.T_max = config['max_video_length']
        self.conv_mode = self.T_max > 32          # modules.py:255
        self._temporal_related = None             # Temporal.related_attn stash, modules.py:288,321-325

    def p(self, name):
        return self.W['submodules.' + name]

    # ---- encoders (module_net.py:147-163) ---------------------------------------------------------
    def _bilstm(self, name, x):
        if self.aten_lstm:
            m = self._lstm_cache.get(name)
            if m is None:
                m = torch.nn.LSTM(input_size=x.size(1), hidden_size=self.p(name + '.weight_hh_l0').size(1), batch_first=True,
                                  bidirectional=True)
                with torch.no_grad():
                    for k, v in m.named_parameters():
                        v.copy_(self.p(name + '.' + k))
                self._lstm_cache[name] = m
            out, (hn, _) = m(x.unsqueeze(0))
            return out[0], hn[:, 0, :].reshape(-1)
        f, hf = _lstm_dir(x, self.p(name + '.weight_ih_l0'), self.p(name + '.weight_hh_l0'),
                          self.p(name + '.bias_ih_l0'), self.p(name + '.bias_hh_l0'), False)
        b, hb = _lstm_dir(x, self.p(name + '.weight_ih_l0_reverse'), self.p(name + '.weight_hh_l0_reverse'),
                          self.p(name + '.bias_ih_l0_reverse'), self.p(name + '.bias_hh_l0_reverse'), True)
        return torch.cat([f, b], dim=-1), torch.cat([hf, hb])

    def encode_video(self, video):                       # module_net.py:160-163
        return self._bilstm('video_encoder', video)[0]

    def encode_question(self, question):                 # module_net.py:151-158
        return self._bilstm('text_encoder', question)    # (token_feature [L,H], sent [H] = [h_fwd ; h_bwd])

    @staticmethod
    def l2normalize(x):                                  # module_net.py:211-216, F.normalize(dim=0), eps 1e-12
        return x / x.norm().clamp_min(1e-12)

    # ---- operators (video_nmn/modules.py) ---------------------------------------------------------
    def _drop(self, site, x):
        return x if self.dropout is None else self.dropout(site, x, (self._q, self._tok))

    def _mlp2(self, prefix, x, last_relu=True, drop_last=False):
        """Linear-ReLU-Dropout-Linear[-ReLU[-Dropout]] with Sequential indices 0 and 3."""
        x = self._drop(prefix + '.0', torch.relu(_linear(x, self.p(prefix + '.0.weight'), self.p(prefix + '.0.bias'))))
        x = _linear(x, self.p(prefix + '.3.weight'), self.p(prefix + '.3.bias'))
        if not last_relu:
            return x
        x = torch.relu(x)
        return self._drop(prefix + '.3', x) if drop_last else x

    def And(self, a, b):                                 # modules.py:7-12
        return torch.min(a, b)

    def AttnVideo(self, feat, attn):                     # modules.py:330-340
        return attn.unsqueeze(1) * feat

    def Choose(self, k1, k2, q):                         # modules.py:40-56 (strict '>' : tie -> k2)
        return k1 if bool(_cos(k1, q) > _cos(k2, q)) else k2

    def Compare(self, f1, f2):                           # modules.py:15-21
        return torch.relu(_linear(torch.cat([f1, f2]), self.p('Compare.param.0.weight'), self.p('Compare.param.0.bias')))

    def Equals(self, f1, f2):                            # modules.py:24-37
        return torch.relu(_linear(torch.cat([f1, f2]), self.p('Equals.param.0.weight'), self.p('Equals.param.0.bias')))

    def Exists(self, keyword, feat):                     # modules.py:141-159
        return self._mlp2('Exists.param', torch.cat([feat, keyword, feat * keyword]), drop_last=True)

    def ExistsFrame(self, keyword, feat):                # modules.py:162-178
        return (_cos(feat, keyword.unsqueeze(0)) + 1) * 0.49

    def Filter(self, feat, keyword):                     # modules.py:343-378
        if isinstance(keyword, torch.Tensor):
            x = self._mlp2('Filter.param.representation', feat, drop_last=True)
            fk = torch.cat([x, keyword.unsqueeze(0).expand(x.size(0), -1)], dim=1)
            # nn.Softmax() on [T,1] -> implicit dim=1 -> attention == 1.0 exactly (SURVEY §8a Filter)
            a = _legacy_softmax(_linear(fk, self.p('Filter.attention.0.weight'), self.p('Filter.attention.0.bias')))
            agg = torch.sum(a * x, dim=0)
        else:
            agg = torch.sum(self._mlp2('Filter.param.' + keyword, feat, drop_last=True), dim=0)
        return torch.relu(_linear(agg, self.p('Filter.dense.0.weight'), self.p('Filter.dense.0.bias')))

    def FilterFrame(self, feat, keyword):                # modules.py:381-414 (no 'objects' key: KeyError as in ref)
        if isinstance(keyword, torch.Tensor):
            x = self._mlp2('FilterFrame.param.representation', feat, drop_last=True)
            fk = torch.cat([x, keyword.unsqueeze(0).expand(x.size(0), -1)], dim=1)
            a = torch.sigmoid(_linear(fk, self.p('FilterFrame.attention.0.weight'), self.p('FilterFrame.attention.0.bias')))
            agg = a * x
        else:
            if keyword not in ('relations', 'actions'):
                raise KeyError(keyword)
            agg = self._mlp2('FilterFrame.param.' + keyword, feat, drop_last=True)
        return self._drop('FilterFrame.dense.0', torch.relu(_linear(agg, self.p('FilterFrame.dense.0.weight'), self.p('FilterFrame.dense.0.bias'))))

    def HasItem(self, feat):                             # modules.py:123-138
        x = self._drop('HasItem.param.0', torch.relu(_linear(feat, self.p('HasItem.param.0.weight'), self.p('HasItem.param.0.bias'))))
        return self._drop('HasItem.param.3', torch.sigmoid(_linear(x, self.p('HasItem.param.3.weight'), self.p('HasItem.param.3.bias')))).squeeze()

    def Localize(self, feat, keyword):                   # modules.py:181-217
        f = self._mlp2('Localize.video_linear', feat, last_relu=False)             # [T,H]
        if keyword.dim() == 1:
            keyword = keyword.unsqueeze(0)
        k = _linear(keyword, self.p('Localize.keyword_linear.0.weight'), self.p('Localize.keyword_linear.0.bias'))
        att = _cos(f.unsqueeze(0), k.unsqueeze(1))                                 # [K,T]
        return (att + 1) * 0.49

    def Relate(self, mode, attn):                        # modules.py:417-435
        beta = self.p('Relate.beta')[:attn.size(0)]
        return _legacy_softmax(attn + beta if mode == 'forward' else attn - beta)

    def Superlative(self, mode, actions, feat):          # modules.py:220-248 (shares Localize weights)
        att = self.Localize(feat, actions)
        w = torch.softmax(att.sum(dim=1), dim=0)
        if mode == 'min':
            w = 1 - w
        v = torch.sum(w.unsqueeze(1) * actions, dim=0)
        return torch.relu(_linear(v, self.p('Superlative.dense.0.weight'), self.p('Superlative.dense.0.bias')))

    def temporal_relate(self, mode, a):                  # modules.py:255-278 (learned relate[mode])
        if mode == 'while':
            return a
        pre = 'Temporal.relate.%s.' % mode
        if self.conv_mode:
            a = torch.relu(_conv1d_same(a, self.p(pre + '0.weight'), self.p(pre + '0.bias')))
            a = torch.relu(_conv1d_same(a, self.p(pre + '2.weight'), self.p(pre + '2.bias')))
            return torch.sigmoid(_conv1d_same(a, self.p(pre + '4.weight'), self.p(pre + '4.bias')))
        a = torch.relu(_linear(a, self.p(pre + '0.weight'), self.p(pre + '0.bias')))
        a = torch.relu(_linear(a, self.p(pre + '2.weight'), self.p(pre + '2.bias')))
        return torch.sigmoid(_linear(a, self.p(pre + '4.weight'), self.p(pre + '4.bias')))

    def Temporal(self, mode, feat, attention):           # modules.py:310-327
        a = attention.mean(dim=0)
        r = self.temporal_relate(mode, a)
        self._temporal_related = r
        x = self._drop('Temporal.dense.0', torch.relu(_linear(r.unsqueeze(-1) * feat, self.p('Temporal.dense.0.weight'), self.p('Temporal.dense.0.bias'))))
        return F.layer_norm(x, (x.size(-1),), self.p('Temporal.layer_norm.weight'), self.p('Temporal.layer_norm.bias'), 1e-5)

    @staticmethod
    def relate_scan(attention, mode):
        """modules.py:290-308 ``TemporalModule.relate_`` — cumsum before/after/between masks.
        DEAD CODE in the reference forward (never called); restated because the north star names the scans."""
        if mode == 'while':
            return attention.squeeze()
        a = torch.relu(attention).squeeze()
        if mode == 'before':
            return torch.cumsum(a, dim=-1)
        if mode == 'after':
            return torch.cumsum(a.flip([-1]), dim=-1).flip([-1])
        if mode == 'between':
            r = OracleNMN.relate_scan
            return torch.max(torch.min(r(a[0], 'before'), r(a[0], 'after')), torch.min(r(a[1], 'before'), r(a[1], 'after')))
        raise KeyError(mode)

    def ToAction(self, action, keyword):                 # modules.py:102-120
        return self._mlp2('ToAction.param', torch.cat([action, keyword]))

    def Xor(self, f1, f2):                               # modules.py:59-72
        return torch.relu(_linear(torch.cat([torch.abs(f1 - f2), f1, f2]), self.p('Xor.param.0.weight'), self.p('Xor.param.0.bias')))

    def XorFrame(self, a, b):                            # modules.py:75-80
        return torch.abs(a - b)

    def Array2(self, f1, f2):                            # modules.py:438-443
        return torch.stack([f1, f2])

    def pretrain_head(self, name, out):
        """Per-module ``pretrain_head`` (modules.py: Equals :29, Xor :63, Exists :150, FilterFrame :396,
        Identity for HasItem/ExistsFrame/Localize, contrastive L2Normalize for Filter/Superlative/ToAction
        module_net.py:33-34, Temporal returns the stashed related_attn modules.py:287-288)."""
        if name in ('Filter', 'Superlative', 'ToAction'):
            return self.l2normalize(out)
        if name in ('Equals', 'Xor', 'Exists', 'FilterFrame'):
            return _linear(out, self.p(name + '.pretrain_head.weight'), self.p(name + '.pretrain_head.bias'))
        if name == 'Temporal':
            return self._temporal_related
        if name in ('HasItem', 'ExistsFrame', 'Localize'):
            return out
        raise AttributeError('%s has no pretrain_head' % name)      # as the reference would

    # ---- interpreter (module_net.py:65-145) -------------------------------------------------------
    def forward(self, data, return_res_by_step=True, return_result_of_each_step=False, test_mode=False):
        question, video = data['question'], data['video_features']
        spans, tokens, prog_idx = data['prog_str_to_question_tokens'], data['nmn_program_list'], data['nmn_program_idx']
        video_feat = self.encode_video(video)
        token_feature, question_feature = self.encode_question(question)

        new_gold = {}
        for key, value in data.get('sg_res_by_step', {}).items():          # module_net.py:78-89
            if isinstance(value, list) and len(value) and isinstance(value[0][1], torch.Tensor):
                with torch.no_grad():
                    new_gold[key] = [(n, self.l2normalize(self.encode_question(e)[1])) for n, e in value]
            else:
                new_gold[key] = value

        stack, res_by_step, each = [], {}, []
        head = self.config['have_pretrain_head']
        for i in range(len(tokens) - 1, -1, -1):
            tok = tokens[i]
            params = []
            if tok in MODULE_NAMES:
                for _ in range(NARY[tok]):
                    p = stack.pop()
                    params.append(video_feat if isinstance(p, str) and p == 'video' else p)
                self._tok = i
                out = getattr(self, tok)(*params)
                if return_res_by_step and prog_idx[i] is not None and tok in self.pretrain_modules and i != 0:
                    res_by_step[prog_idx[i]] = (tok, self.pretrain_head(tok, out) if head else out)
                if return_result_of_each_step:
                    each.append((params, self.pretrain_head(tok, out) if head and tok in self.pretrain_modules else out))
            elif tok in WORDS_TO_KEEP:
                out = tok
                if return_result_of_each_step:
                    each.append((params, out))
            else:
                s, e = spans[i]
                out = torch.mean(token_feature[s:e, :], dim=0)
                if return_result_of_each_step:
                    each.append((params, out))
            stack.append(out)
        assert len(stack) == 1
        hid = torch.cat([stack[0], question_feature])
        self._tok = None
        x = self._drop('decoder.0', torch.relu(_linear(hid, self.p('decoder.0.weight'), self.p('decoder.0.bias'))))
        logits = _linear(x, self.p('decoder.3.weight'), self.p('decoder.3.bias'))
        ret = {'logits': logits, 'res_by_step': res_by_step}
        if return_result_of_each_step:
            ret['result_of_each_step'] = list(reversed(each))
        if not test_mode:
            ret['sg_res_by_step'] = new_gold
        return ret

    __call__ = forward


# --------------------------------------------------------------------------------------------------
# losses — train_module.py:33-194
# --------------------------------------------------------------------------------------------------
def span_to_attention(gold, T):
    """train_module.py:67-81 — soft [T] mask from a float interval."""
    g = torch.zeros(T)
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


class OracleCriterion:
    def __init__(self, word2id: Dict[str, int] | None = None, module_loss_type='cont_nolinear'):
        self.module_loss_type = module_loss_type
        word2id = word2id or {}
        ids = sorted(set(word2id.values()))                       # train_module.py:50-55
        id2index = {v: i for i, v in enumerate(ids)}
        self.word2id = {w: id2index[v] for w, v in word2id.items()}
        self.names = ['Exists', 'Xor', 'Equals', 'Filter', 'ToAction', 'FilterFrame', 'ExistsFrame', 'Superlative',
                      'Localize', 'Temporal', 'decoder']

    @staticmethod
    def attention_score(pred, gold):                              # train_module.py:83-90
        g = torch.stack([gold, 1 - gold], dim=-1)
        p = torch.stack([pred, 1 - pred], dim=-1)
        return torch.mean(-torch.sum(torch.log(p) * g, dim=-1))

    def __call__(self, name, pred, gold):
        if name in ('Exists', 'Xor'):                             # :92-99
            return F.cross_entropy(pred.unsqueeze(0), torch.tensor([int(gold)]))
        if name == 'Equals':                                      # :101-107
            return torch.mean(torch.square(pred - int(gold)))
        if name in ('Filter', 'ToAction', 'Superlative'):         # :113-139,166-171
            if self.module_loss_type in ('cont', 'cont_nolinear'):
                return F.cross_entropy(torch.matmul(pred, gold.t()).unsqueeze(0), torch.tensor([0]))
            if gold == []:
                return torch.tensor(0.)
            gt = torch.stack([g[1] for g in gold]).mean(dim=0)
            return F.cosine_similarity(pred, gt, dim=0)
        if name == 'FilterFrame':                                 # :141-155
            T = pred.size(0)
            g = torch.zeros_like(pred)
            for key, val in gold.items():
                g[:, self.word2id[key]] = span_to_attention(val, T)
            g = g / g.sum(dim=1, keepdim=True)
            g = torch.where(g.isinf() | g.isnan(), torch.zeros_like(g), g)
            return F.binary_cross_entropy(torch.softmax(pred, dim=1), g)
        if name in ('ExistsFrame', 'Temporal'):                   # :157-164,184-191
            return self.attention_score(pred, span_to_attention(gold, pred.size(0)))
        if name == 'Localize':                                    # :173-182
            g = torch.stack([span_to_attention(gold[i], pred.size(1)) for i in range(pred.size(0))])
            return self.attention_score(pred, g)
        if name == 'decoder':                                     # :193-194
            return F.cross_entropy(pred.unsqueeze(0), data_answer(gold).unsqueeze(0))
        raise KeyError(name)


# --------------------------------------------------------------------------------------------------
# dropout masks — restatement of the counter-based hash of csrc/stair_common.cuh (lowbias32 / make_drop / drop_keep).
# torch's Philox stream is not reproducible in a batched executor, so parity under dropout is checked with the SAME masks
# injected into this oracle (tests/test_train_gpu.py::test_dropout_*).
# --------------------------------------------------------------------------------------------------
def _lowbias32(x):
    import numpy as np
    x = np.asarray(x, dtype=np.uint64) & 0xffffffff
    x ^= x >> 16; x = (x * 0x7feb352d) & 0xffffffff
    x ^= x >> 15; x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    return x


def dropout_keep(seed: int, site: int, row0: int, rows: int, cols: int, p: float):
    """bool [rows, cols]: element (row0 + r, c) of dropout site ``site`` is kept."""
    import numpy as np
    thresh = min(int(float(np.float32(p)) * 4294967296.0), 0xffffffff)
    key_lo = int(_lowbias32((seed & 0xffffffff) ^ ((site * 0x9E3779B1) & 0xffffffff)))
    key_hi = int(_lowbias32((((seed >> 32) & 0xffffffff) + site * 0x85EBCA77 + 1) & 0xffffffff))
    r = (np.arange(row0, row0 + rows, dtype=np.uint64) & 0xffffffff) ^ key_lo
    rh = (_lowbias32(r) + key_hi) & 0xffffffff
    c = (np.arange(cols, dtype=np.uint64) * 0x9E3779B1) & 0xffffffff
    h = _lowbias32((rh[:, None] + c[None, :]) & 0xffffffff)
    return h >= thresh


def data_answer(a):
    return a if isinstance(a, torch.Tensor) else torch.tensor(int(a))


def window_loss(model: OracleNMN, crit: OracleCriterion, batch: List[dict], module_loss_weight=1.0,
                decoder_loss_weight=1.0, gradient_accumulation=None,
                modules_no_intermediate_train=('FilterFrame',)):
    """One gradient-accumulation window of train_module.py:341-406: per-question module losses, decoder CE,
    then the window-level contrastive losses (class_reps / neg_reps, last writer wins).  Returns the scalar
    that the reference calls ``batch_loss`` before ``backward()`` plus per-module loss lists."""
    ga = gradient_accumulation or len(batch)
    total = 0.
    logs = {n: [] for n in crit.names}
    class_reps, neg_reps = {}, {}
    outs = []
    for it, data in enumerate(batch):
        model._q = it
        out = model.forward(data, return_res_by_step=module_loss_weight != 0)
        outs.append(out)
        gold_by_step = out['sg_res_by_step']
        for step, (name, res) in out['res_by_step'].items():
            if step not in gold_by_step or name in modules_no_intermediate_train or name not in crit.names:
                continue
            gold = gold_by_step[step]
            if gold is None:
                continue
            if name in ('Filter', 'Superlative', 'ToAction'):
                for cname, crep in gold:
                    class_reps.setdefault(cname, []).append((it, name, res))
                    neg_reps[cname] = crep
            else:
                l = crit(name, res, gold)
                logs[name].append(float(l.detach()))
                total = total + l * module_loss_weight / ga
        l = crit('decoder', out['logits'], data['answer'])
        logs["decoder"].append(float(l.detach()))
        total = total + l * decoder_loss_weight / ga
    for cname, vals in class_reps.items():                        # train_module.py:388-406
        for _, name, res in vals:
            pos = neg_reps[cname]
            neg = [v for k, v in neg_reps.items() if k != cname]
            gold = torch.cat([pos.unsqueeze(0), torch.stack(neg)]) if neg else pos.unsqueeze(0)
            l = crit(name, res, gold)
            logs[name].append(float(l.detach()))
            total = total + l * module_loss_weight / ga
    return total, logs, outs


# --------------------------------------------------------------------------------------------------
# Filter audit — evaluate.py:65-117 (get_filter_text_results), one question at a time like the reference
# --------------------------------------------------------------------------------------------------
def filter_text_results(model: OracleNMN, batch: List[dict], filter_vocab: List[str], embed_sent, top_k: int = 10):
    """-> ({qa_id: {prog_idx: (level, keyword, top-k phrases)}}, {qa_id: {prog_idx: sims [len(vocab)]}})."""
    reps = []
    for answer in filter_vocab:                                           # evaluate.py:66-75
        _, rep = model.encode_question(embed_sent(answer))
        reps.append(model.l2normalize(rep.squeeze()))
    reps = torch.stack(reps)
    out, sims_out = {}, {}
    for data in batch:                                                    # evaluate.py:83-111
        res = model.forward(data, return_res_by_step=False, return_result_of_each_step=True, test_mode=True)
        prog = data['nmn_program_list']
        idx = data.get('nmn_program_idx') or list(range(len(prog)))
        levels = module_levels(prog)
        children, _ = children_and_parents(prog)
        entry, sims_entry = {}, {}
        for i, (p, p_idx, lvl, ch) in enumerate(zip(prog, idx, levels, children)):
            if p != 'Filter':
                continue
            _, result = res['result_of_each_step'][i]
            sims = F.cosine_similarity(result.unsqueeze(0), reps)        # nn.CosineSimilarity() defaults: dim=1, eps=1e-8
            ranks = torch.argsort(sims, descending=True)
            entry[p_idx] = (lvl, prog[ch[1]].replace('_', ' '), [filter_vocab[int(j)] for j in ranks[:top_k]])
            sims_entry[p_idx] = sims
        out[data['qa_id']] = entry
        sims_out[data['qa_id']] = sims_entry
    return out, sims_out
