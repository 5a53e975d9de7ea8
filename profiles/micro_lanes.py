"""Forward step time vs number of concurrent lanes of the module phase and vs the wave schedule (merged / ASAP levels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L

B, T, V = 4096, 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
lib = L.lib()
for merge in (True, False):
    batch = collate(qs, video_dtype=torch.bfloat16, merge_waves=merge).to('cuda')
    for lanes in (1, 2, 3, 4, 6, 8):
        lib.stair_set_lanes(lanes)
        for _ in range(5):
            model.forward_batch(batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            model.forward_batch(batch)
        e1.record(); torch.cuda.synchronize()
        print('merge_waves=%s groups=%d lanes=%d: %.3f ms/step (%d launches)' % (merge, batch.n_groups, lanes, e0.elapsed_time(e1) / 20, model.last_launches), flush=True)
