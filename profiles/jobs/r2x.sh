#!/bin/bash
timeout 600 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x -k "frame_sum" 2>&1 | tail -15
timeout 900 python -m pytest tests/test_forward_gpu.py tests/test_operators_gpu.py tests/test_edge_gpu.py -m gpu -q -x 2>&1 | tail -4
python - <<'PY'
import sys
sys.path.insert(0, '.')
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
lib = L.lib()
for (T, V, tmpl, seed, name) in ((8, 4096, None, 1234, 'RX'), (64, 1024, list(syn.LONG_TEMPLATES), 777, 'I3D')):
    cfg = syn.model_config(T=T, V=V)
    torch.manual_seed(0)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
    qs = syn.make_questions(4096, T, V, seed=seed, templates=tmpl) if tmpl else syn.make_questions(4096, T, V, seed=seed)
    batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
    res = {}
    for rep in range(2):
        for fuse in (0, 1):
            lib.stair_set_fuse_sum(fuse)
            for _ in range(3):
                st = model.forward_batch(batch)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                st = model.forward_batch(batch)
            e1.record(); torch.cuda.synchronize()
            res[fuse] = st.logits.clone()
            print('%s fuse_sum %d: %.3f ms per forward, %d launches' % (name, fuse, e0.elapsed_time(e1) / 10, model.last_launches), flush=True)
    d = (res[0] - res[1]).abs().max().item()
    print('%s max |dlogit| fused vs unfused %.3g (max |logit| %.3g), answers equal %d / 4096' % (name, d, res[0].abs().max().item(), int((res[0].argmax(1) == res[1].argmax(1)).sum())))
    del model, batch
    torch.cuda.empty_cache()
lib.stair_set_fuse_sum(1)
PY
