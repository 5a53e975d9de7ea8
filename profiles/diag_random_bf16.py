"""Per-step bf16 error of the CUDA path vs the CPU oracle on random layouts at full width (diagnostic for the test bars)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import nmn_oracle as orc
from stair_b200 import VideoNMN, synthetic as syn

shape = sys.argv[1] if len(sys.argv) > 1 else 'i3d'
T, V = (8, 4096) if shape == 'rx' else (64, 1024)
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
ref_model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32')
weights = {k: v.detach().clone() for k, v in ref_model.state_dict().items()}
oracle = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES, aten_lstm=True)
qs = syn.make_random_questions(120, T, V, seed=4242)
with torch.no_grad():
    want = [oracle(d, return_res_by_step=False, return_result_of_each_step=True, test_mode=True) for d in qs]
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16')
model.load_state_dict(weights)
model = model.cuda().eval()
out = model(qs, return_res_by_step=False, return_result_of_each_step=True, test_mode=True)
torch.cuda.synchronize()
worst = []
for qi, w in enumerate(want):
    for j, ((_, got), (_, exp)) in enumerate(zip(out['result_of_each_step'][qi], w['result_of_each_step'])):
        if isinstance(exp, str):
            continue
        g = got.float().cpu(); e = exp.float()
        scale = max(float(e.abs().max()), 1e-3)
        err = float((g - e).abs().max())
        worst.append((err / scale, err, scale, qi, j, qs[qi]['nmn_program_list'][j]))
worst.sort(reverse=True)
for r in worst[:12]:
    print('rel %.3f err %.4f scale %.3f  q%d step %d %s' % r)
    print('      ', ' '.join(qs[r[3]]['nmn_program_list']))
