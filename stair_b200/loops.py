"""Batched drivers of the callers on either side of the hot path.

  * ``evaluate``        <- evaluate.py:28-62 (test accuracy; the reference indexes ``logits`` as [B, A] there, which its own
                           1-question model does not produce — the batched model does)
  * ``train``           <- train_module.py:326-439 (Adam + LambdaLR linear decay, one optimizer step per accumulation window,
                           per-criterion loss logging without a device->host sync per loss)
  * ``save_checkpoint`` / ``load_checkpoint`` <- train_module.py:212-216, 292-298 and evaluate.py:137-139: ``pytorch_model.bin``
                           holds a *state_dict* with the reference's 119 keys (what evaluate.py loads), ``config.json`` the model
                           config; optimizer / scheduler / step state go to ``trainer_state.pt`` so a run can really resume
                           (the reference saves neither).
The arithmetic is the CUDA library's; these loops only move batches and bookkeeping.
"""
from __future__ import annotations

import collections
import json
import os

import torch

from . import layout as LY
from .train import NMNTrainStep, FusedAdam, LOSS_SLOTS


@torch.no_grad()
def evaluate(batches, model, unk_token_id=None, id2word=None, preds_file=None, pipelined_chunks=0):
    """``batches``: iterable of lists of data dicts / ``NMNBatch``.  Returns (accuracy, preds_golds) like evaluate.py:28-60:
    a question counts as correct iff pred == gold and gold != <UNK> (evaluate.py:46)."""
    correct, total = 0, 0
    out = {'preds': [], 'golds': [], 'qa_ids': []}

    def account(pred, gold, examples):
        nonlocal correct, total
        pred, gold = pred.long(), gold.long()
        ok = pred == gold
        if unk_token_id is not None:
            ok &= gold != unk_token_id
        correct += int(ok.sum())
        total += int(gold.numel())
        conv = (lambda i: id2word[i]) if id2word is not None else (lambda i: i)
        out['preds'] += [conv(int(p)) for p in pred]
        out['golds'] += [conv(int(g)) for g in gold]
        out['qa_ids'] += [e.get('qa_id') for e in examples]

    if pipelined_chunks > 1:
        # streaming: uploads of batch k+1 overlap the execution of batch k, answers return through pinned memory
        pending = collections.deque()

        def host_batches():
            for data in batches:
                chunks = [data] if isinstance(data, LY.NMNBatch) else LY.collate_chunks(list(data), pipelined_chunks, pin_memory=True)
                pending.append(chunks)
                yield chunks

        for pred in model.forward_stream(host_batches(), depth=2):
            chunks = pending.popleft()
            account(pred, torch.cat([c.answer for c in chunks]), [e for c in chunks for e in c.examples])
    else:
        for data in batches:
            batch = data if isinstance(data, LY.NMNBatch) else LY.collate(list(data), pin_memory=True)
            pred = model(batch, return_res_by_step=False, test_mode=True)['answers']
            account(pred.cpu(), batch.answer, batch.examples)   # one D2H per batch (the reference syncs three times per question)
    if preds_file is not None:
        json.dump(out, open(preds_file, 'w'))
    return (correct / total if total else 0.0), out


def make_scheduler(optimizer, start_factor=1.0, end_factor=0.1, total_iters=200000):
    """train_module.py:328-332 — linear LambdaLR from start_factor to end_factor over total_iters, constant afterwards."""
    def lr_lambda(it):
        if it > total_iters:
            return end_factor
        return start_factor + (end_factor - start_factor) / total_iters * it
    return torch.optim.lr_scheduler.LambdaLR(optimizer, lr_lambda=lr_lambda)


def _make_optimizer(model, lr, weight_decay):
    """Reference optimizer (Adam, weight_decay 0, train_module.py:326): one fused kernel per step that also refreshes the kernels'
    weight copies; a non-zero weight decay is outside what ``FusedAdam`` implements and runs on ``torch.optim.Adam``."""
    if weight_decay == 0:
        return FusedAdam(model, lr=lr), 'fused_adam'
    return torch.optim.Adam(model.parameters(), lr, weight_decay=weight_decay), 'torch_adam'


def train(windows, model, lr=2e-4, weight_decay=0.0, module_loss_weight=1.0, decoder_loss_weight=1.0,
          modules_no_intermediate_train=('FilterFrame',), scheduler_kwargs=None, report_interval=0, log=None, state=None,
          train_module_before_iters=None, train_decoder_after_iters=0, questions_per_iter=1):
    """One pass over ``windows`` (each a list of data dicts = one gradient-accumulation window of train_module.py:386-412):
    step = forward + losses + backward on the GPU, Adam (skips parameters the window did not touch, like the reference),
    scheduler step.  Returns the trainer state (optimizer, scheduler, global_steps, loss history) for ``save_checkpoint``.

    Staged schedules (video_nmn/args.py:44-45, train_module.py:349,376): module losses apply while the reference's iteration counter is
    ``< train_module_before_iters`` (None = always), the decoder loss once it is ``>= train_decoder_after_iters``.  The reference counts
    questions (one per iteration); here a window is the unit, so the gates are evaluated per window at the iteration index of its
    first question (questions seen so far x ``questions_per_iter``; pass windows of 32 for the reference's granularity)."""
    if state is None:
        opt, kind = _make_optimizer(model, lr, weight_decay)
        state = {'optimizer': opt, 'optimizer_kind': kind, 'scheduler': make_scheduler(opt, **(scheduler_kwargs or {})), 'global_steps': 0,
                 'losses': [], 'scheduler_kwargs': dict(scheduler_kwargs or {}), 'questions_seen': 0}
    steps = {}

    def step_for(mlw, dlw):
        if (mlw, dlw) not in steps:
            steps[(mlw, dlw)] = NMNTrainStep(model, module_loss_weight=mlw, decoder_loss_weight=dlw,
                                             modules_no_intermediate_train=modules_no_intermediate_train)
        return steps[(mlw, dlw)]

    model.train()
    pending = []
    for window in windows:
        it = state.get('questions_seen', 0) * questions_per_iter
        # the reference's counter is 1-based (global_steps += 1 before the tests): modules while g < before, decoder once g > after
        mlw = module_loss_weight if (train_module_before_iters is None or it + 1 < train_module_before_iters) else 0.0
        dlw = decoder_loss_weight if it + 1 > train_decoder_after_iters else 0.0
        out = step_for(mlw, dlw)(window)
        state['questions_seen'] = state.get('questions_seen', 0) + (window.B if isinstance(window, LY.NMNBatch) else len(window))
        state['optimizer'].step()
        state['scheduler'].step()
        state['optimizer'].zero_grad(set_to_none=True)
        state['global_steps'] += 1
        pending.append(out['loss_terms'])
        if report_interval and state['global_steps'] % report_interval == 0:
            terms = torch.stack(pending).cpu()                     # one sync per report, not one per loss (train_module.py:369,378)
            pending = []
            for row in terms:
                state['losses'].append({k: float(v) for k, v in zip(LOSS_SLOTS, row[:len(LOSS_SLOTS)])})
            if log is not None:
                log(state['global_steps'], state['losses'][-1], state['scheduler'].get_last_lr()[0])
    if pending:
        for row in torch.stack(pending).cpu():
            state['losses'].append({k: float(v) for k, v in zip(LOSS_SLOTS, row[:len(LOSS_SLOTS)])})
    return state


def save_checkpoint(output_dir, model, state=None):
    os.makedirs(output_dir, exist_ok=True)
    torch.save({k: v.detach().cpu() for k, v in model.state_dict().items()}, os.path.join(output_dir, 'pytorch_model.bin'))
    json.dump(model.config, open(os.path.join(output_dir, 'config.json'), 'w'))
    if state is not None:
        opt = state['optimizer']
        kind = state.get('optimizer_kind') or ('fused_adam' if isinstance(opt, FusedAdam) else 'torch_adam')
        torch.save({'optimizer': opt.state_dict(), 'optimizer_kind': kind, 'scheduler': state['scheduler'].state_dict(),
                    'global_steps': state['global_steps'], 'questions_seen': state.get('questions_seen', 0),
                    'scheduler_kwargs': state.get('scheduler_kwargs', {})},
                   os.path.join(output_dir, 'trainer_state.pt'))


def load_checkpoint(ckpt_dir, model_cls, device='cuda', precision='bf16', pretrain_modules=frozenset(), with_trainer_state=False, lr=2e-4,
                    allow_pickle=False):
    """Accepts what the reference writes / reads: ``pytorch_model.bin`` as a state_dict (evaluate.py:139) plus ``config.json``.
    Everything ``save_checkpoint`` writes is plain tensors / numbers, so files are read with ``weights_only=True`` (no code runs on
    load).  The reference's *training* script saves a whole pickled ``nn.Module`` instead (train_module.py:214, loaded at :296); unpickling
    executes arbitrary code from the file, so that form is only accepted with an explicit ``allow_pickle=True`` for trusted files."""
    config = json.load(open(os.path.join(ckpt_dir, 'config.json')))
    path = os.path.join(ckpt_dir, 'pytorch_model.bin')
    try:
        blob = torch.load(path, map_location='cpu', weights_only=True)
    except Exception as e:
        if not allow_pickle:
            raise RuntimeError('%s is not a plain state_dict (a pickled nn.Module as train_module.py:214 writes it?): pass '
                               'allow_pickle=True to unpickle it — only for files you trust, unpickling runs code' % path) from e
        blob = torch.load(path, map_location='cpu', weights_only=False)
    sd = blob.state_dict() if hasattr(blob, 'state_dict') else blob
    model = model_cls(config, pretrain_modules=set(pretrain_modules), precision=precision)
    model.load_state_dict(sd)
    model = model.to(device)
    if not with_trainer_state:
        return model
    p = os.path.join(ckpt_dir, 'trainer_state.pt')
    ts = torch.load(p, map_location='cpu', weights_only=True) if os.path.exists(p) else None
    # the optimizer class follows the run that wrote the state: FusedAdam implements weight_decay = 0 only
    wd = 0.0
    if ts:
        wd = max([float(g.get('weight_decay', 0.0) or 0.0) for g in ts['optimizer'].get('param_groups', [])] or [0.0])
    opt, kind = _make_optimizer(model, lr, wd)
    kw = dict(ts.get('scheduler_kwargs', {})) if ts else {}          # the LambdaLR schedule itself is not in its state_dict
    state = {'optimizer': opt, 'optimizer_kind': kind, 'scheduler': make_scheduler(opt, **kw), 'global_steps': 0, 'losses': [],
             'scheduler_kwargs': kw, 'questions_seen': 0}
    if ts:
        opt.load_state_dict(ts['optimizer'])
        state['scheduler'].load_state_dict(ts['scheduler'])
        state['global_steps'] = ts['global_steps']
        state['questions_seen'] = ts.get('questions_seen', 0)
    return model, state
