"""Small end-to-end forward + training step for compute-sanitizer (memcheck / racecheck): every kernel of the path at tiny sizes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate
from stair_b200.train import NMNTrainStep, Adam

for T, V, H, prec in ((8, 128, 128, 'bf16'), (8, 64, 64, 'fp32'), (64, 64, 64, 'bf16')):
    cfg = syn.model_config(T=T, V=V, hidden=H, object_types=16)
    torch.manual_seed(0)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision=prec).cuda().eval()
    qs = syn.make_questions(32, T, V, seed=3, templates=list(syn.ALL_TEMPLATES), with_gold=True, object_types=16)
    out = model(qs, return_res_by_step=True, return_result_of_each_step=True)
    torch.cuda.synchronize()
    model.device_text_sort = True                          # the library's device counting sort instead of collate's schedule
    model.forward_batch(collate(qs[:29]).to('cuda'))
    model.device_text_sort = False
    torch.cuda.synchronize()
    model.train()
    step, opt = NMNTrainStep(model), Adam(model.parameters())
    o = step(qs); opt.step(); opt.zero_grad()
    torch.cuda.synchronize()
    print('ok', T, V, H, prec, float(o['loss']), flush=True)
