"""Whole inference forward, B = 4096 RX questions (the bench workload), 3 x 20 back-to-back forwards; A/B through the library's
environment switches (STAIR_LANE_PRIO, STAIR_TEXT_SORT, ...).  argv: [B] [i3d]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
i3d = len(sys.argv) > 2 and sys.argv[2] == 'i3d'
T, V = (64, 1024) if i3d else (8, 4096)
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234, templates=['xor_between', 'and_between_until', 'compare_between'] if i3d else None)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
for _ in range(5):
    st = model.forward_batch(batch)
torch.cuda.synchronize()
out = []
for rep in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        st = model.forward_batch(batch)
    e1.record(); torch.cuda.synchronize()
    out.append(e0.elapsed_time(e1) / 20)
print('env %s: forward %s ms, %d launches, checksum %.6f' % ({k: v for k, v in os.environ.items() if k.startswith('STAIR_')},
      ' / '.join('%.3f' % t for t in out), model.last_launches, float(st.logits.float().abs().sum())), flush=True)
