"""Host-side cost of collate (no GPU work): 4096 RX questions as reference-schema dicts -> pinned bf16 NMNBatch chunks, the call the bench's
e2e.from_dicts leg times.  Prints the wall time and a cProfile split (the native staging call shows up as _stage_rows' own time)."""
import sys, os, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import synthetic as syn
from stair_b200.layout import collate_chunks
B, T, V = 4096, 8, 4096
qs = syn.make_questions(B, T, V, seed=1234)
pin = torch.cuda.is_available()
for _ in range(3):
    t0 = time.perf_counter()
    ch = collate_chunks(qs, 2, pin_memory=pin, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
    print('collate_chunks(4096 questions, 2 chunks, pinned=%s): %.1f ms' % (pin, 1e3 * (time.perf_counter() - t0)), flush=True)
    del ch
pr = cProfile.Profile(); pr.enable()
ch = collate_chunks(qs, 2, pin_memory=pin, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
