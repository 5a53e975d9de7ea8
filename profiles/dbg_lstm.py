import sys; sys.path.insert(0,'/root/repo')
import torch
from stair_b200 import VideoNMN, collate, synthetic as syn, _lib as L
T,V,hidden=8,256,128
cfg = syn.model_config(T=T, V=V, hidden=hidden)
torch.manual_seed(1)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(40, T, V, seed=11)
batch = collate(qs).to('cuda')
import ctypes
dbg = torch.zeros(16, dtype=torch.int32).pin_memory()
L.lib().stair_lstm_debug(ctypes.c_void_p(dbg.data_ptr()))
try:
    st = model.forward_batch(batch, phases=L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
    torch.cuda.synchronize()
    print('ok')
except Exception as e:
    print('ERR', str(e)[:200])
print('dbg', [hex(int(x) & 0xffffffff) for x in dbg.tolist()])
print('flag', L.lib().stair_gemm_error_flag())
