"""Filter-audit head — the step right after the hot path (reference: ``evaluate.py:65-117`` ``get_filter_text_results``).

For every ``Filter`` call of every question the reference takes the module output (through ``pretrain_head`` = L2Normalize when
heads are on), scores it against the text-encoder representations of an audit vocabulary (``data/AGQA/filter_answers.json``, 214
phrases) with ``nn.CosineSimilarity`` and keeps the 10 best phrases; the result feeds the LLM prompt files.  Here the phrase
representations are encoded once in one batched text-encoder call and all Filter outputs of a batch are ranked by one CUDA kernel
(``stair_cosine_topk``, csrc/audit.cu).  No CPU path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib as L
from . import layout as LY


class FilterAudit:
    def __init__(self, model, filter_vocab, embed_sent, top_k=10):
        """``filter_vocab``: list of phrases; ``embed_sent(phrase) -> [n_words, text_size]`` word embeddings
        (``dataloader.dataset.embed_sent`` in the reference, evaluate.py:69)."""
        self.model, self.vocab, self.top_k = model, list(filter_vocab), min(top_k, len(filter_vocab))
        _, sent = model.encode_questions([embed_sent(a) for a in self.vocab])               # evaluate.py:68-72
        H = sent.shape[1]
        self.reps = torch.empty((len(self.vocab), H), dtype=torch.float32, device=sent.device)
        L.check(L.lib().stair_l2normalize(L.i32(L.dtype_code(sent.dtype)), L.ptr(sent), L.ptr(self.reps), L.i32(len(self.vocab)), L.i32(H),
                                          L.stream_ptr()), 'stair_l2normalize')                # model.contrastive_head (evaluate.py:73)

    @torch.no_grad()
    def __call__(self, data):
        """-> {qa_id: {prog_idx: (level, keyword text, [top-k phrases])}} exactly like ``filter_results_text_list`` (evaluate.py:100-111)."""
        model = self.model
        batch = data if isinstance(data, LY.NMNBatch) else LY.collate([data] if isinstance(data, dict) else list(data))
        dev = next(model.parameters()).device
        if batch.device is None:
            batch.to(dev)
        use_head = bool(model.config['have_pretrain_head']) and 'Filter' in model.pretrain_modules
        st = model.forward_batch(batch, frozenset(['Filter']) if use_head else frozenset())
        il = st.itab_layout
        itab = st.itab.cpu().numpy()
        n = batch.n_nodes
        out_slot, aux_slot = itab[il.out_slot:il.out_slot + n], itab[il.aux_slot:il.aux_slot + n]
        model.check_status(st)
        rows, where = [], []
        for q, (lay, e) in enumerate(zip(batch.layouts, batch.examples)):
            idx = e.get('nmn_program_idx') or list(range(len(lay.tokens)))                    # evaluate.py:95-96
            base = int(batch.node_start[q])
            for i, tok in enumerate(lay.tokens):
                if tok != 'Filter':
                    continue
                node = base + lay.node_of_token[i]
                rows.append(int(aux_slot[node] if use_head else out_slot[node]))
                kw_tok = lay.tokens[lay.param_tokens[i][1]]                                     # nmn_program[childrens[1]] (evaluate.py:110)
                where.append((q, idx[i], int(lay.level[lay.node_of_token[i]]), kw_tok.replace('_', ' ')))
        results = {e.get('qa_id', qi): {} for qi, e in enumerate(batch.examples)}
        if not rows:
            return results
        H = st.H
        src = st.head_vec if use_head else st.vec
        k = self.top_k
        ridx = torch.tensor(rows, dtype=torch.int32, device=dev)
        out_idx = torch.empty((len(rows), k), dtype=torch.int32, device=dev)
        out_sim = torch.empty((len(rows), k), dtype=torch.float32, device=dev)
        L.check(L.lib().stair_cosine_topk(L.i32(L.dtype_code(src.dtype)), L.ptr(src), L.i64(H), L.ptr(ridx), L.ptr(self.reps), L.i32(len(self.vocab)),
                                          L.i32(H), L.i32(k), L.ptr(out_idx), L.ptr(out_sim), L.i32(len(rows)), L.stream_ptr()), 'stair_cosine_topk')
        top = out_idx.cpu().numpy()
        self.last_sims = out_sim
        for (q, prog_idx, level, kw), ids in zip(where, top):
            qa = batch.examples[q].get('qa_id', q)
            results[qa][prog_idx] = (level, kw, [self.vocab[int(j)] for j in ids])
        return results
