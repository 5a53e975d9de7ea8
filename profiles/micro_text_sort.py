"""A/B of the length-sorted inference text recurrence (stair_set_text_sort): whole forward and the encoder phases alone, B = 4096 RX
questions of 8-24 words (the bench workload), L2-flushing not needed (inputs 309 MB)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from stair_b200 import VideoNMN, synthetic as syn, collate, _lib as L
B, T, V = int(os.environ.get('B', 4096)), 8, 4096
cfg = syn.model_config(T=T, V=V)
torch.manual_seed(0)
model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
qs = syn.make_questions(B, T, V, seed=1234)
batch = collate(qs, video_dtype=torch.bfloat16).to('cuda')
lib = L.lib()


N = int(os.environ.get('N', 20))


def timed(phases, n=N):
    for _ in range(min(4, n)):
        st = model.forward_batch(batch, phases=phases)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        st = model.forward_batch(batch, phases=phases)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, st


ref = None
for rep in range(2):
    for on in (0, 1):
        lib.stair_set_text_sort(on)
        full, st = timed(L.FWD_ALL)
        lg = st.logits.clone()
        enc, _ = timed(L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT)
        txt, _ = timed(L.FWD_ENCODE_TEXT)
        if ref is None:
            ref = lg
        print('text_sort %d: forward %.3f ms, both encoders %.3f ms, text encoder alone %.3f ms, launches %d, logits equal to first run: %s'
              % (on, full, enc, txt, model.last_launches, bool(torch.equal(lg, ref))), flush=True)
lib.stair_set_text_sort(1)

# split forward: (encoders + grouping) | (modules + decoder) as two calls with an event between them, back to back (no host sync inside
# the loop) — where does the time go?
for on in (0, 1, 0, 1):
    lib.stair_set_text_sort(on)
    n = N + 4
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(n)]
    for it in range(n):
        ev[it][0].record()
        model.forward_batch(batch, phases=L.FWD_ENCODE_VIDEO | L.FWD_ENCODE_TEXT | L.FWD_GROUP)
        ev[it][1].record()
        model.forward_batch(batch, phases=L.FWD_MODULES | L.FWD_DECODE)
        ev[it][2].record()
    torch.cuda.synchronize()
    a = sum(ev[it][0].elapsed_time(ev[it][1]) for it in range(4, n)) / N
    b = sum(ev[it][1].elapsed_time(ev[it][2]) for it in range(4, n)) / N
    tot = ev[4][0].elapsed_time(ev[n - 1][2]) / N
    print('text_sort %d, split forward back to back: encoders + grouping %.3f ms, modules + decoder %.3f ms, per iteration %.3f ms' % (on, a, b, tot), flush=True)
lib.stair_set_text_sort(1)

# phase marks inside an unsplit forward (stair_debug_timeline / stair_debug_phase_marks), forwards back to back
import ctypes, numpy as np
for on in (0, 1, 0, 1):
    lib.stair_set_text_sort(on)
    for _ in range(3):
        model.forward_batch(batch)
    lib.stair_debug_timeline(1)
    acc = np.zeros(7)
    for _ in range(5):
        for _ in range(3):
            model.forward_batch(batch)
        ms = np.zeros(8, np.float32)
        lib.stair_debug_phase_marks(ms.ctypes.data_as(ctypes.c_void_p), 8)
        acc += ms[:7]
    lib.stair_debug_timeline(0)
    print('text_sort %d, marks (ms from forward start): video proj %.3f, text proj %.3f, recurrence %.3f, grouping joined %.3f, modules %.3f, decoder %.3f'
          % ((on,) + tuple(acc[1:] / 5)), flush=True)
lib.stair_set_text_sort(1)

# host time to enqueue 20 forwards back to back vs the device time of the same loop
import time
for on in (0, 1):
    lib.stair_set_text_sort(on)
    for _ in range(3):
        model.forward_batch(batch)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(20):
        model.forward_batch(batch)
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print('text_sort %d: host enqueue %.3f ms per forward, device %.3f ms per forward, wall incl. final sync %.3f ms' %
          (on, (t1 - t0) * 50, e0.elapsed_time(e1) / 20, (t2 - t0) * 50), flush=True)
    # the same with the C call only (prepare() once): how much of the host time is Python
    st, ms_, sb, bufs = model.prepare(batch, frozenset())
    torch.cuda.synchronize()
    t0 = time.perf_counter(); e0.record()
    for _ in range(20):
        lib.stair_nmn_forward(ctypes.byref(ms_), ctypes.byref(sb), ctypes.byref(bufs), L.i32(L.FWD_ALL), L.stream_ptr(None))
    e1.record(); t1 = time.perf_counter()
    torch.cuda.synchronize()
    print('text_sort %d: C call only: host enqueue %.3f ms per forward, device %.3f ms per forward' % (on, (t1 - t0) * 50, e0.elapsed_time(e1) / 20), flush=True)
lib.stair_set_text_sort(1)
