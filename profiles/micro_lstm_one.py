"""Run the text (or video) encoder phase alone a few times — target for `ncu -k regex:lstm_` captures.  argv: text|video [B]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L

which = sys.argv[1] if len(sys.argv) > 1 else 'text'
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
T, V = 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
ph = L.FWD_ENCODE_TEXT if which == 'text' else L.FWD_ENCODE_VIDEO
for i in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); model.forward_batch(batch, phases=ph); e1.record(); torch.cuda.synchronize()
    print('%s encoder phase %d: %.1f us' % (which, i, e0.elapsed_time(e1) * 1e3), flush=True)
