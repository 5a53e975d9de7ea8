"""Gantt table of the module phase (dependency scheduling): start / end of every module group relative to the start of the phase, its lane,
operator, variant and instance count.  argv: [B] [full|modules]   ('full' = inside a whole forward, default; 'modules' = the phase alone)"""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mode = sys.argv[2] if len(sys.argv) > 2 else 'full'
T, V = 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
lib = L.lib()
if 'TEXT_SORT' in os.environ:
    lib.stair_set_text_sort(int(os.environ['TEXT_SORT']))
ph = L.FWD_ALL if mode == 'full' else L.FWD_MODULES
for _ in range(5):
    model.forward_batch(batch, phases=L.FWD_ALL)
torch.cuda.synchronize()
lib.stair_debug_timeline(1)
for _ in range(3):
    model.forward_batch(batch, phases=ph)
torch.cuda.synchronize()
cap = 96
t0 = np.zeros(cap, np.float32); t1 = np.zeros(cap, np.float32)
lane = np.zeros(cap, np.int32); op = np.zeros(cap, np.int32); cnt = np.zeros(cap, np.int32); var = np.zeros(cap, np.int32)
n = lib.stair_debug_timeline_read(*(a.ctypes.data_as(ctypes.c_void_p) for a in (t0, t1, lane, op, cnt, var)), cap)
lib.stair_debug_timeline(0)
names = {v: k for k, v in L.OP.items()}
print('module phase timeline (%s), B=%d: %d groups, last group ends at %.1f us' % (mode, B, n, 1e3 * t1[:n].max()))
print('%3s %-12s %3s %6s %4s %8s %8s %7s' % ('g', 'op', 'var', 'count', 'lane', 'start', 'end', 'dur'))
for g in range(n):
    print('%3d %-12s %3d %6d %4d %8.1f %8.1f %7.1f' % (g, names.get(int(op[g]), op[g]), var[g], cnt[g], lane[g], 1e3 * t0[g], 1e3 * t1[g], 1e3 * (t1[g] - t0[g])))
busy = sum(float(t1[g] - t0[g]) for g in range(n))
print('sum of group durations %.1f us over %d lanes' % (1e3 * busy, len(set(lane[:n].tolist()))))
