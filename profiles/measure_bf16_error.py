"""Measure what the bf16 path's error against the fp32 CPU oracle actually IS at full size (H = 512), per intermediate type, so that the
test bars (tests/test_forward_gpu.py BF16_REL / BF16_ABS) can be set from data instead of guessed.  Prints one table per shape.

    python profiles/measure_bf16_error.py [n_questions]
"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import nmn_oracle as orc  # noqa: E402
from stair_b200 import VideoNMN, synthetic as syn  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
torch.set_num_threads(os.cpu_count() or 1)
for shape, (T, V) in (('rx', (8, 4096)), ('i3d', (64, 1024))):
    cfg = syn.model_config(T=T, V=V)
    torch.manual_seed(0)
    m = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16')
    weights = {k: v.detach().clone() for k, v in m.state_dict().items()}
    oracle = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES, aten_lstm=True)
    qs = syn.make_questions(n, T, V, seed=99, templates=list(syn.ALL_TEMPLATES))
    with torch.no_grad():
        want = [oracle(d, return_res_by_step=False, return_result_of_each_step=True, test_mode=True) for d in qs]
    m = m.cuda().eval()
    out = m(qs, return_res_by_step=False, return_result_of_each_step=True, test_mode=True)
    torch.cuda.synchronize()
    worst = collections.defaultdict(lambda: [0.0, 0.0, 0])       # kind -> [max err / max|ref| of the tensor, max abs err, count]
    arg_total = arg_bad = 0
    arg_bad_margin = 0.0
    for qi, w in enumerate(want):
        for j, ((_, got), (_, exp)) in enumerate(zip(out['result_of_each_step'][qi], w['result_of_each_step'])):
            if isinstance(exp, str):
                continue
            tok = qs[qi]['nmn_program_list'][j]
            kind = tok if tok in syn.MODULE_ARITY else '<word>'
            g = got.float().cpu()
            err = float((g - exp).abs().max())
            rel = err / max(float(exp.abs().max()), 1e-3)
            r = worst[kind]
            r[0], r[1], r[2] = max(r[0], rel), max(r[1], err), r[2] + 1
            if exp.dim() >= 1 and exp.size(-1) == T and exp.numel() <= 2 * T:
                ga, ea = g.argmax(-1).reshape(-1), exp.argmax(-1).reshape(-1)
                top = exp.reshape(-1, T).topk(2, dim=-1).values
                for a, b, t2 in zip(ga.tolist(), ea.tolist(), top):
                    arg_total += 1
                    if a != b:
                        arg_bad += 1
                        arg_bad_margin = max(arg_bad_margin, float(t2[0] - t2[1]))
    logits = out['logits'].cpu()
    ref = torch.stack([w['logits'] for w in want])
    lerr = (logits - ref).abs().max(1).values
    lscale = ref.abs().max(1).values
    top = ref.topk(2, dim=1).values
    margin = top[:, 0] - top[:, 1]
    mism = out['answers'].cpu().long() != ref.argmax(1)
    print('== %s: %d questions, H=512' % (shape, n))
    for kind in sorted(worst):
        r = worst[kind]
        print('  %-12s tensors %5d   max err/max|ref| %.3e   max abs err %.3e' % (kind, r[2], r[0], r[1]))
    print('  logits: max |d| %.3e, max |d|/max|logit| %.3e (median %.3e); max|logit| %.3f' %
          (float(lerr.max()), float((lerr / lscale).max()), float((lerr / lscale).median()), float(lscale.max())))
    print('  answers: %d / %d equal; largest oracle top-2 margin among mismatches %.3e (relative to max|logit|: %.3e); '
          'median margin %.3e' % (int((~mism).sum()), n, float(margin[mism].max()) if mism.any() else 0.0,
                                  float((margin[mism] / lscale[mism]).max()) if mism.any() else 0.0, float(margin.median())))
    print('  attention argmax: %d / %d equal; largest top-2 margin among mismatches %.3e' % (arg_total - arg_bad, arg_total, arg_bad_margin))
