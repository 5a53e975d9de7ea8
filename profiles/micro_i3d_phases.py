"""Phase split and module-group timeline of the I3D stress configuration (BASELINE configs[4]: T = 64, V = 1024, only layouts of >= 12 modules)."""
import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
T, V = 64, 1024
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=777, templates=list(syn.LONG_TEMPLATES))
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
lib = L.lib()


def run(ph, n=10):
    for _ in range(3):
        model.forward_batch(batch, phases=ph)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        model.forward_batch(batch, phases=ph)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


model.forward_batch(batch)
print('I3D B=%d: whole forward %.3f ms (%d launches)' % (B, run(L.FWD_ALL), model.last_launches))
for name, ph in (('group', L.FWD_GROUP), ('encode_video', L.FWD_ENCODE_VIDEO), ('encode_text', L.FWD_ENCODE_TEXT), ('modules', L.FWD_MODULES), ('decode', L.FWD_DECODE)):
    print('  %-14s %.3f ms' % (name, run(ph)))
lib.stair_debug_timeline(1)
for _ in range(2):
    model.forward_batch(batch)
torch.cuda.synchronize()
cap = 96
t0 = np.zeros(cap, np.float32); t1 = np.zeros(cap, np.float32)
lane, op, cnt, var = (np.zeros(cap, np.int32) for _ in range(4))
n = lib.stair_debug_timeline_read(*(a.ctypes.data_as(ctypes.c_void_p) for a in (t0, t1, lane, op, cnt, var)), cap)
lib.stair_debug_timeline(0)
names = {v: k for k, v in L.OP.items()}
print('module phase inside the forward: %d groups, ends at %.1f us' % (n, 1e3 * t1[:n].max()))
for g in range(n):
    print('%3d %-12s %3d %6d %4d %8.1f %8.1f %7.1f' % (g, names.get(int(op[g]), op[g]), var[g], cnt[g], lane[g], 1e3 * t0[g], 1e3 * t1[g], 1e3 * (t1[g] - t0[g])))
