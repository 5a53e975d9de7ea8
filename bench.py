#!/usr/bin/env python
"""Benchmark of the NMN hot path: questions/sec of the batched ModuleNet forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our arm (CUDA library through the C ABI)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path on the host cores

Workload (config.workload): BASELINE.json configs[1] — 4096 mixed-program questions per GPU (the 10 AGQA layout
templates of SURVEY.md App. B), RX/TGIF-QA-shaped features [8, 4096] (appearance 8x16x2048 mean-pooled + motion 8x2048,
video_nmn/dataset.py:150-172), questions of 8-24 GloVe-sized words, random-init weights, bf16 storage / fp32 accumulate,
inference (test_mode=True, return_res_by_step=False).  N > 1: every rank runs its own 4096 questions (weak scaling,
configs[2]: 32768 questions at 8 GPUs) and the int32 answers are all-gathered with NCCL inside the timed region (one step behind the compute).

One JSON line on stdout (rank 0): the contract keys plus
  roofline      dominant kernel (video input projection GEMM) against the measured bf16 peak
  roofline_hbm  the memory-bound module kernels against the measured HBM peak, at the step's own group size and at a streaming size
  phases_ms / host_enqueue_ms / strict (fp32 strict mode, timed) / audit / i3d (configs[4]) / train (configs[3]) / e2e (+ from_dicts)
  parity        ALL answers of the step against the CPU oracle (bf16 and fp32 strict), attention argmax, error statistics
  cpu_baseline  the reference's CPU path restated (oracle port) timed on the box's host cores
The oracle (oracle/nmn_oracle.py) is used ONLY as the cpu_baseline / --impl reference timer and as the checker — never on the
measured path.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402,F401
import torch  # noqa: E402

METRIC, UNIT = 'nmn_questions_per_sec', 'questions/s'
PER_GPU_B = 4096
T, V = 8, 4096
CPU_SAMPLE = 32
TRAIN_DROPOUT = 0.25
E2E_CHUNKS = 2
PARITY_PER_RANK_MULTI = 512        # N > 1: questions of every rank's shard checked against the oracle (the host cores are shared by N ranks)
ATT_CHECK_Q = 512                  # questions whose Localize attention maps are pulled back and argmax-compared


def algorithmic_step_flops(batch, cfg):
    """GEMM flops of one inference step over ``batch`` (2·M·K·N per Linear): input projections and recurrences of
    both encoders, decoder, and every module instance of the batch's layouts."""
    from stair_b200 import layout as LY, _lib as L
    H, h, T, V, A = cfg['hidden_size'], cfg['hidden_size'] // 2, batch.T, cfg['video_size'], cfg['answer_vocab_length']
    B, n_tok, txt = batch.B, batch.n_tok, cfg['text_size']
    f = B * T * (2.0 * V * 4 * H + 2.0 * h * 4 * H) + n_tok * (2.0 * txt * 4 * H + 2.0 * h * 4 * H)          # 2 directions x 4h gates = 4H
    f += B * (2.0 * 2 * H * 2 * H + 2.0 * 2 * H * A)
    lin, frame = 2.0 * H * H, 2.0 * T * H * H
    groups, _, _ = LY.build_groups(batch, frozenset())
    names = {v: k for k, v in L.OP.items()}
    for g in range(batch.n_groups):
        op, var, n = names[groups[g].op], groups[g].variant, groups[g].count
        if op == 'LOCALIZE':
            per = 2 * frame + (var + 1) * lin
        elif op in ('TEMPORAL', 'HASITEM'):
            per = frame
        elif op == 'FILTER':
            per = 2 * frame + lin
        elif op == 'FILTERFRAME':
            per = 3 * frame
        elif op == 'SUPERLATIVE':
            kind = var >> 1
            per = 2 * frame + (1 if kind == 0 else (2 if kind == 1 else T)) * lin + lin
        elif op in ('COMPARE', 'EQUALS'):
            per = 2 * lin
        elif op == 'XOR':
            per = 3 * lin
        elif op == 'EXISTS':
            per = 4 * lin
        elif op == 'TOACTION':
            per = 3 * lin
        else:
            per = 0.0
        f += n * per
    return f


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'src': 'measured (MEASURED_PEAKS.json)'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback (B200_PROFILING.md)'}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace('.', '').isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith('active')})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons,
                'samples': len(sm)}


TEMPLATES = None          # None = the ten AGQA templates


def workload_string(args, B):
    if args.workload == 'rx':
        return ('ModuleNet batched inference, %d mixed-program questions per GPU (10 AGQA layout templates, 2-12 modules), RX/TGIF-QA features '
                '[8,4096], questions 8-24 words x 300, random init' % B)
    return ('I3D stress test (BASELINE configs[4]): %d questions per GPU, layouts of >= 12 modules only (xor_between, and_between_until, '
            'compare_between), features [64,1024], conv-mode Temporal, random init' % B)


def build_inputs(rank, B):
    from stair_b200 import synthetic as syn, collate
    qs = syn.make_questions(B, T, V, seed=1234 + rank, with_gold=True, templates=TEMPLATES)      # gold is only read by the training leg
    batch = collate(qs, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.float32)
    return qs, batch


def cpu_reference_timer(qs, weights, cfg, min_seconds=10.0, max_passes=50, threads=None):
    """The reference's CPU path restated (oracle port, encoders through torch.nn.LSTM exactly like the reference):
    per-question Python loop, eval, no_grad (train_module.py:229-232 / evaluate.py:33-38)."""
    from oracle import nmn_oracle as orc
    from stair_b200 import synthetic as syn
    torch.set_num_threads(threads or os.cpu_count() or 1)
    model = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES, aten_lstm=True)
    sample = qs[:CPU_SAMPLE]
    with torch.no_grad():
        for d in sample[:4]:
            model(d, return_res_by_step=False, test_mode=True)
        times = []
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            for d in sample:
                model(d, return_res_by_step=False, test_mode=True)
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all >= min_seconds or len(times) >= max_passes:
                break
    med = statistics.median(times)
    return len(sample) / med, med, len(times)


def oracle_outputs(qs, weights, cfg, threads, with_attention):
    """CPU oracle over ``qs`` (fp32, one question at a time like the reference): logits [n, A] and, for the first ``with_attention``
    questions, every Localize attention map of the layout (list per question of [K, T] tensors in token order)."""
    from oracle import nmn_oracle as orc
    from stair_b200 import synthetic as syn
    torch.set_num_threads(max(1, threads))
    model = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES, aten_lstm=True)
    logits, maps = [], []
    t0 = time.perf_counter()
    with torch.no_grad():
        for i, d in enumerate(qs):
            att = i < with_attention
            out = model(d, return_res_by_step=False, return_result_of_each_step=att, test_mode=True)
            logits.append(out['logits'])
            if att:
                maps.append([r for tok, (_, r) in zip(d['nmn_program_list'], out['result_of_each_step']) if tok == 'Localize'])
    return torch.stack(logits), maps, time.perf_counter() - t0


def parity_report(model_bf16, strict_model, qs, weights, cfg, answers_bf16, logits_bf16, threads):
    """Answers / logits of the measured arm against the CPU oracle on the SAME questions (all of ``qs``)."""
    n = len(qs)
    n_att = min(ATT_CHECK_Q, n)
    ref_logits, ref_maps, secs = oracle_outputs(qs, weights, cfg, threads, n_att)
    ref_ans = ref_logits.argmax(1)
    top2 = ref_logits.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    lmax = ref_logits.abs().max(1).values
    got = answers_bf16[:n].long().cpu()
    lg = logits_bf16[:n].float().cpu()
    mism = got != ref_ans
    rel = (lg - ref_logits).abs().max(1).values / lmax
    clear = margin > 1e-2 * lmax                                   # the bar of tests/test_forward_gpu.py (3x the largest margin ever seen to flip)
    rep = {'questions_checked': n, 'oracle_seconds': secs,
           'bf16_answers_equal': int((~mism).sum()),
           'bf16_answers_equal_where_margin_clear': int(((~mism) & clear).sum()), 'bf16_clear_margin': int(clear.sum()),
           'bf16_margin_rule': 'reference top-2 logit margin > 1e-2 * max|logit|',
           'bf16_mismatch_max_margin_rel': float((margin[mism] / lmax[mism]).max()) if bool(mism.any()) else 0.0,
           'bf16_logit_err_rel_median': float(rel.median()), 'bf16_logit_err_rel_p99': float(rel.kthvalue(max(1, int(0.99 * n))).values),
           'bf16_logit_err_rel_max': float(rel.max())}
    # attention argmax of every Localize map of the first n_att questions (audit outputs of the measured model)
    out = model_bf16(qs[:n_att], return_res_by_step=False, return_result_of_each_step=True, test_mode=True)
    rows = eq = clear_rows = clear_eq = 0
    worst = 0.0
    for qi in range(n_att):
        toks = qs[qi]['nmn_program_list']
        mine = [r for tok, (_, r) in zip(toks, out['result_of_each_step'][qi]) if tok == 'Localize']
        for g, w in zip(mine, ref_maps[qi]):
            g = g.float().cpu()
            same = g.argmax(-1) == w.argmax(-1)
            t2 = w.topk(2, dim=-1).values
            ok = (t2[:, 0] - t2[:, 1]) > 6e-4
            rows += same.numel(); eq += int(same.sum()); clear_rows += int(ok.sum()); clear_eq += int((same & ok).sum())
            if bool((~same).any()):
                worst = max(worst, float((t2[:, 0] - t2[:, 1])[~same].max()))
    rep.update({'localize_rows': rows, 'localize_argmax_equal': eq, 'localize_rows_margin_clear': clear_rows,
                'localize_argmax_equal_where_margin_clear': clear_eq, 'localize_mismatch_max_margin': worst,
                'localize_margin_rule': 'reference top-2 attention margin > 6e-4'})
    if strict_model is not None:
        s_out = strict_model(qs, return_res_by_step=False, test_mode=True)
        s_ans = s_out['answers'].long().cpu()
        rep.update({'fp32_strict_answers_equal': int((s_ans == ref_ans).sum()),
                    'fp32_strict_max_logit_err': float((s_out['logits'].float().cpu() - ref_logits).abs().max())})
    return rep


def make_weights(cfg):
    """Random-init weights with the reference's default initialisers under torch.manual_seed(0) (CPU fp32 state_dict)."""
    from stair_b200 import VideoNMN, synthetic as syn
    torch.manual_seed(0)
    m = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES)
    return {k: v.detach().clone() for k, v in m.state_dict().items()}


def run_reference(args, rank, world):
    if rank != 0:
        return
    from stair_b200 import synthetic as syn
    cfg = syn.model_config(T=T, V=V)
    weights = make_weights(cfg)
    qs = syn.make_questions(CPU_SAMPLE, T, V, seed=1234, templates=TEMPLATES)
    from oracle import nmn_oracle as orc
    torch.set_num_threads(os.cpu_count() or 1)
    model = orc.OracleNMN(cfg, weights, syn.PRETRAIN_MODULES, aten_lstm=True)
    step = lambda: [model(d, return_res_by_step=False, test_mode=True)['logits'] for d in qs]      # noqa: E731
    with torch.no_grad():
        for _ in range(args.warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()
        dt = time.perf_counter() - t0
    qps = CPU_SAMPLE * args.steps / dt
    cores = torch.get_num_threads()
    line = {'impl': 'reference', 'metric': METRIC, 'value': qps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * dt / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic',
            # the measured arm's workload; the reference computes it in fp32 on the host (its only arithmetic type)
            'config': {'workload': workload_string(args, args.batch) + ', fp32 on the host cores (the reference\'s arithmetic)',
                       'questions_per_gpu': args.batch, 'global_questions': args.gpus * args.batch, 'frames': T, 'video_size': V,
                       'hidden_size': cfg['hidden_size'], 'parallelism': 'reference arm: host cores of rank 0 only',
                       'reference_step': 'CPU reference path (fp32): per-question loop, eval, no_grad; one step = the first %d questions of '
                                         'that workload (bounded sample)' % CPU_SAMPLE},
            'cpu_baseline': {'value': qps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                             'sample': '%d questions per step x %d steps, torch threads=%d; oracle port of module_net.py/modules.py with '
                                       'the encoders through torch.nn.LSTM (the reference cannot travel to the GPU box: its import needs '
                                       '/root/reference)' % (CPU_SAMPLE, args.steps, cores)},
            'e2e': {'value': qps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def device_timed(fn, steps, barrier, dev, dist, world, finish=None):
    """CUDA-event time (ms, max over ranks) of ``steps`` calls of ``fn`` on the current stream; ``finish`` (work the steps left on other
    streams, e.g. the last answer gathers) is joined into the stream before the closing event."""
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = fn()
    if finish is not None:
        finish()
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), out


def main():
    # the driver reads ONE JSON line from stdout: libraries that print there (NCCL's version banner) go to stderr instead
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, 'w')
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='stair_b200', choices=['stair_b200', 'reference'])
    ap.add_argument('--batch', type=int, default=PER_GPU_B, help='questions per GPU (default: the BASELINE config)')
    ap.add_argument('--no-cpu-baseline', action='store_true', help='skip the CPU legs (cpu_baseline and the oracle parity check)')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step leg (BASELINE configs[3])')
    ap.add_argument('--no-extras', action='store_true', help='skip the strict / i3d / roofline_hbm / e2e-from-dicts legs (quick runs)')
    ap.add_argument('--no-overlap-allreduce', action='store_true',
                    help='training leg at N > 1: one gradient all-reduce after the whole backward instead of overlapping it with BPTT (comparison)')
    ap.add_argument('--workload', default='rx', choices=['rx', 'i3d'],
                    help="rx (default, BASELINE configs[1]): T=8, V=4096, the 10 AGQA templates; i3d (configs[4] stress test): T=64, V=1024, "
                         "conv-mode Temporal, only layouts of >= 12 modules")
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    global T, V, TEMPLATES
    from stair_b200 import synthetic as syn
    if args.workload == 'i3d':
        T, V, TEMPLATES = 64, 1024, list(syn.LONG_TEMPLATES)
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    if args.warmup < 3:
        args.warmup = 3
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device (sm_100a); there is no CPU fallback for the measured arm')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from stair_b200 import VideoNMN, _lib as L
    from stair_b200.distributed import bind_to_gpu_numa_node
    numa = bind_to_gpu_numa_node(local) if world > 1 else {'numa_node': None}     # before any pinned allocation
    if world > 1:
        sys.stderr.write('rank %d: %s\n' % (rank, numa))

    cfg = syn.model_config(T=T, V=V)
    weights = make_weights(cfg)
    model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16')
    model.load_state_dict(weights)
    model = model.to(dev).eval()
    B = args.batch
    qs, batch = build_inputs(rank, B)
    gathered = torch.empty(world * B, dtype=torch.int32, device=dev) if world > 1 else None

    # the answers of step k are gathered (NCCL) while step k+1 computes; every gather completes inside the timed region
    from stair_b200.distributed import AnswerGather
    gather = AnswerGather(B, dev, depth=2) if world > 1 else None

    def step_device():
        st = model.forward_batch(batch)
        if world > 1:
            gather.submit(st.answers)
        return st

    def finish_device():
        if world > 1:
            gathered.copy_(gather.finish())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed throughput: inputs resident in HBM ---------------------------------------------------------
    batch.to(dev)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        st = step_device()
    finish_device()
    model.check_status(st)
    launches_per_step = model.last_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, st = device_timed(step_device, args.steps, barrier, dev, dist, world, finish=finish_device)
    value = world * B * args.steps / (ms * 1e-3)
    answers_dev, logits_dev = st.answers.clone(), st.logits.clone()
    gathered_ok = None
    if world > 1:                                                    # the gathered answers are every rank's own answers, in rank order
        gathered_ok = bool(torch.equal(gathered[rank * B:(rank + 1) * B], answers_dev))

    # host time to ENQUEUE one forward (ctypes call: ~120 cuTensorMapEncodeTiled + ~80 launches), the device idle
    torch.cuda.synchronize()
    enq = []
    for _ in range(5):
        t0 = time.perf_counter()
        model.forward_batch(batch)
        enq.append(time.perf_counter() - t0)
        torch.cuda.synchronize()
    host_enqueue_ms = 1e3 * statistics.median(enq)

    # ---- end to end through the public API: pinned host batch -> H2D -> forward -> answers D2H ----------------------
    # The pinned host batch is collated as E2E_CHUNKS sub-batches (bf16 features and word embeddings, the storage type of the
    # bf16 path); VideoNMN.forward_pipelined uploads chunk k+1 on a copy stream while chunk k executes.
    from stair_b200 import collate_chunks
    host_chunks = collate_chunks(qs, E2E_CHUNKS, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
    h2d_bytes = sum(c.h2d_bytes() for c in host_chunks)

    def step_e2e():
        answers, _, _ = model.forward_pipelined(host_chunks)
        if world > 1:
            dist.all_gather_into_tensor(gathered, answers)
            return gathered.cpu()
        return answers.cpu()

    for _ in range(3):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    percall_s = time.perf_counter() - t0

    # streaming form of the same call (VideoNMN.forward_stream, what loops.evaluate uses): every step still uploads its own
    # batch from pinned host memory and reads its answers back; two steps are in flight so the PCIe link never idles.
    def gather_hook(answers):
        if world > 1:
            dist.all_gather_into_tensor(gathered, answers)
            return gathered
        return answers

    def run_stream(n):
        got = 0
        for ans in model.forward_stream((host_chunks for _ in range(n)), depth=2, device_hook=gather_hook):
            got += int(ans.numel())
        return got

    run_stream(3)
    barrier()
    t0 = time.perf_counter()
    got = run_stream(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    assert got == args.steps * B * world
    tt = torch.tensor([e2e_s, percall_s], device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    e2e_s, percall_s = float(tt[0].item()), float(tt[1].item())
    e2e_value = world * B * args.steps / e2e_s
    clock_info = clocks.stop() if rank == 0 else None

    # the same call from reference-schema data dicts (video_nmn/dataset.py:189-233): collate (layout compile + fp32 -> bf16 staging into
    # pinned memory) INSIDE the timer.  This is what `model(list_of_dicts)` costs a caller who does not collate in DataLoader workers.
    from_dicts = None
    if not args.no_extras:
        def step_dicts():
            chunks = collate_chunks(qs, E2E_CHUNKS, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
            answers, _, _ = model.forward_pipelined(chunks)
            return answers.cpu()
        step_dicts()
        barrier()
        t0 = time.perf_counter()
        nd = 3
        for _ in range(nd):
            step_dicts()
        barrier()
        ds = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(ds, op=dist.ReduceOp.MAX)
        t0 = time.perf_counter()
        collate_chunks(qs, E2E_CHUNKS, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.bfloat16)
        collate_ms = 1e3 * (time.perf_counter() - t0)
        from_dicts = {'value': world * B * nd / float(ds.item()), 'unit': UNIT, 'ms_per_step': 1e3 * float(ds.item()) / nd, 'collate_ms': collate_ms,
                      'host_threads': os.cpu_count(),
                      'what': 'collate_chunks(list of reference-schema dicts, fp32 tensors) + forward_pipelined + answers.cpu() per step, one process'}

    # ---- per-phase split and the dominant kernel (video input projection GEMM), CUDA events on the launch stream ---
    phases = [('group', L.FWD_GROUP), ('encode_video', L.FWD_ENCODE_VIDEO), ('encode_text', L.FWD_ENCODE_TEXT),
              ('modules', L.FWD_MODULES), ('decode', L.FWD_DECODE)]
    batch.to(dev)
    M, N, K = B * T, 4 * cfg['hidden_size'], V
    xw = model._packed.tensors[L.W['VENC_WIH']]
    xb = model._packed.tensors[L.W['VENC_B']]
    xout = torch.empty((M, N), dtype=torch.bfloat16, device=dev)
    vid2d = batch.video_dev.view(M, K)
    ph_ms = {n: 0.0 for n, _ in phases}
    gemm_ms = 0.0
    nrep = max(3, min(args.steps, 10))
    for _ in range(nrep):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(len(phases) + 3)]
        evs[0].record()
        for i, (_, ph) in enumerate(phases):
            model.forward_batch(batch, phases=ph)
            evs[i + 1].record()
        L.gemm(vid2d, xw, bias=xb, out=xout)
        evs[len(phases) + 1].record()
        torch.cuda.synchronize()
        for i, (n, _) in enumerate(phases):
            ph_ms[n] += evs[i].elapsed_time(evs[i + 1]) / nrep
        gemm_ms += evs[len(phases)].elapsed_time(evs[len(phases) + 1]) / nrep
    del xout

    # ---- audit mode: the same forward with every pretrain head computed (res_by_step / result_of_each_step:
    # FilterFrame [T, O] head GEMM, L2-normalised Filter / ToAction / Superlative outputs, Exists / Xor / Equals heads) -------------
    heads = model._head_modules(True, True)
    for _ in range(3):
        model.forward_batch(batch, heads)
    a_ms, _ = device_timed(lambda: model.forward_batch(batch, heads), nrep, barrier, dev, dist, world)
    audit_ms = a_ms / nrep
    audit = {'value': world * B / (audit_ms * 1e-3), 'unit': UNIT, 'ms_per_step': audit_ms, 'launches_per_step': model.last_launches,
             'what': 'forward with all pretrain heads (return_res_by_step / return_result_of_each_step device work)'}

    # ---- fp32 strict mode, timed: fp32 storage, every contraction as six bf16-plane products (the mode whose answers are guaranteed
    # bit-identical to the reference) on the same questions -------------------------------------------------------------------------
    strict, strict_model = None, None
    if not args.no_extras:
        strict_model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32')
        strict_model.load_state_dict(weights)
        strict_model = strict_model.to(dev).eval()
        for _ in range(2):
            strict_model.forward_batch(batch)
        ks = max(2, min(args.steps, 5))
        s_ms, s_st = device_timed(lambda: strict_model.forward_batch(batch), ks, barrier, dev, dist, world)
        strict = {'value': world * B * ks / (s_ms * 1e-3), 'unit': UNIT, 'ms_per_step': s_ms / ks, 'launches_per_step': strict_model.last_launches,
                  'answers_equal_to_bf16_path': int((s_st.answers == answers_dev).sum()), 'questions': B,
                  'what': 'precision="fp32": fp32 activations, bf16x3 split contractions (6 plane products per GEMM) on the tensor cores'}
        strict_model.release_buffers()

    # ---- training step (BASELINE configs[3]): forward with history + intermediate-supervision losses + backward +
    # gradient all-reduce (N > 1) + Adam, one window = the rank's 4096 questions; device-timed, max over ranks -------------
    train = None
    if not args.no_train:
        from stair_b200.train import NMNTrainStep, FusedAdam
        from stair_b200 import collate
        # throughput run: the reference's default training dropout (video_nmn/args.py:31); parity runs (tests) use 0 or injected masks.
        # The window mixes in the two layouts with a NON-root Equals / Xor so that every criterion of train_module.py:92-107 runs.
        tmpl = None if args.workload == 'i3d' else list(syn.TEMPLATES) + ['and_equals_xor', 'compare_xor_equals']
        tqs = qs if tmpl is None else syn.make_questions(B, T, V, seed=4321 + rank, with_gold=True, templates=tmpl)
        tbatch = batch if tmpl is None else collate(tqs, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.float32).to(dev)
        tmodel = VideoNMN(dict(cfg, dropout=TRAIN_DROPOUT), pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16')
        tmodel.load_state_dict(weights)
        tmodel = tmodel.to(dev).train()
        tstep = NMNTrainStep(tmodel, overlap_allreduce=not args.no_overlap_allreduce)
        opt = FusedAdam(tmodel, lr=2e-4)
        plan = tstep.plan(tbatch)

        def train_step():
            out = tstep.run(plan)
            opt.step()
            opt.zero_grad()
            return out

        for _ in range(2):
            out = train_step()
        tmodel.check_status(out['state'])
        ksteps = max(2, min(args.steps, 5))
        tms, out = device_timed(train_step, ksteps, barrier, dev, dist, world)
        train = {'value': world * B * ksteps / (tms * 1e-3), 'unit': UNIT, 'ms_per_step': tms / ksteps, 'steps': ksteps,
                 'launches_per_step': tstep.last_launches, 'window_questions': world * B, 'loss': float(out['loss']),
                 'loss_rows': out['loss_counts'], 'dropout': TRAIN_DROPOUT,
                 'layouts': 'the 10 AGQA templates + and_equals_xor + compare_xor_equals (supervised non-root Equals / Xor)' if tmpl else 'as the inference leg',
                 'what': 'forward with encoder history + losses (train_module.py:83-194) + backward + %sAdam (one fused multi-tensor kernel that also refreshes the bf16 weight copies); bf16 storage, fp32 gradients'
                         % ('NCCL gradient all-reduce + ' if world > 1 else '')}
        # the reference-faithful window: 32 questions per optimizer step (train_module.py gradient_accumulation = 32), latency-bound
        plan32 = tstep.plan(tqs[:32])

        def train_step32():
            o = tstep.run(plan32)
            opt.step()
            opt.zero_grad()
            return o

        for _ in range(3):
            train_step32()
        w_ms, _ = device_timed(train_step32, 10, barrier, dev, dist, world)
        w32 = w_ms / 10
        train['window32'] = {'ms_per_step': w32, 'value': 32 / (w32 * 1e-3), 'unit': UNIT, 'launches_per_step': tstep.last_launches,
                             'what': 'one 32-question window per optimizer step on this rank (reference accumulation window; launch-latency bound)'}
        del tmodel, tstep, opt, plan, plan32, tbatch
        torch.cuda.empty_cache()

    # ---- I3D stress configuration (BASELINE configs[4]): T = 64, V = 1024, conv-mode Temporal, ONLY layouts of >= 12 modules --------
    i3d = None
    if not args.no_extras and args.workload == 'rx':
        from stair_b200 import collate
        Ti, Vi = 64, 1024
        icfg = syn.model_config(T=Ti, V=Vi)
        imodel = VideoNMN(icfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16')
        imodel.load_state_dict(make_weights(icfg))
        imodel = imodel.to(dev).eval()
        iqs = syn.make_questions(B, Ti, Vi, seed=777 + rank, templates=list(syn.LONG_TEMPLATES))
        ibatch = collate(iqs, pin_memory=True, video_dtype=torch.bfloat16, question_dtype=torch.float32).to(dev)
        for _ in range(3):
            ist = imodel.forward_batch(ibatch)
        imodel.check_status(ist)
        ki = max(3, min(args.steps, 10))
        i_ms, _ = device_timed(lambda: imodel.forward_batch(ibatch), ki, barrier, dev, dist, world)
        n_mod = [int(lay.is_module.sum()) for lay in {id(l): l for l in ibatch.layouts}.values()]
        i3d = {'value': world * B * ki / (i_ms * 1e-3), 'unit': UNIT, 'ms_per_step': i_ms / ki, 'launches_per_step': imodel.last_launches,
               'questions_per_gpu': B, 'frames': Ti, 'video_size': Vi, 'layouts': list(syn.LONG_TEMPLATES), 'modules_per_layout': sorted(n_mod),
               'what': 'BASELINE configs[4]: I3D features [64, 1024], conv-mode Temporal (k = 16, 16, 33), every question a layout of >= 12 modules; '
                       'inference, bf16 storage'}
        del imodel, ibatch, iqs
        torch.cuda.empty_cache()

    # ---- memory-bound module kernels against the HBM roofline, at the step's own group size and at a streaming size ----------------
    roofline_hbm = None
    if not args.no_extras and rank == 0:
        sys.path.insert(0, os.path.join(ROOT, 'profiles'))
        import module_roofline as mr
        from stair_b200 import layout as LY
        vid_ops = {LY.OP_OF[n] for n in ('Localize', 'Temporal', 'Filter', 'FilterFrame', 'HasItem', 'AttnVideo', 'ExistsFrame')}
        counts = [int(c) for k, c in zip(batch.group_keys, batch.group_counts) if (int(k) // 8) % 32 in vid_ops]
        n_step = int(statistics.median(counts)) if counts else 2048
        names = ('cos_att', 'layernorm', 'sum_frames', 'attn_video', 'exists_frame', 'hasitem_tail', 'l2normalize')
        roofline_hbm = {'peak_gbs': mr.hbm_peak()[0], 'peak_source': mr.hbm_peak()[1],
                        'bytes': 'algorithmic: every input read once + every exposed output written once (SURVEY 8d), bf16 activations',
                        'in_step': {'n': n_step, 'what': 'median instance count of the frame-sized module groups of this step; L2 flushed between '
                                                         'repetitions; 17-34 MB per launch = launch-latency bound',
                                    'kernels': [{k: r[k] for k in ('kernel', 'n', 'bytes', 'ms', 'frac')} for r in mr.measure(n_step, T=T, kernels=names)]},
                        'streaming': {'n': 32768, 'what': 'inputs larger than the 126 MB L2, no flush (l2normalize: 262144 rows)',
                                      'kernels': [{k: r[k] for k in ('kernel', 'n', 'bytes', 'ms', 'frac')} for r in mr.measure(32768, T=T, kernels=names)]}}
        torch.cuda.empty_cache()

    pk = peaks()
    flops = 2.0 * M * N * K
    achieved_tf = flops / (gemm_ms * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at the default shape, from the committed `ncu --set full` capture;
    # algorithmic bytes: A 268 MB + W 17 MB + C 134 MB = 419 MB
    traffic_file = os.path.join(ROOT, 'profiles', 'r2_gemm_pair_ncu_traffic.json')
    traffic, traffic_src = None, None
    if args.workload == 'rx' and B == PER_GPU_B and os.path.exists(traffic_file):
        tj = json.load(open(traffic_file))
        traffic, traffic_src = tj.get('dram_bytes'), tj.get('source')
    roofline = {'bound': 'tensor', 'kernel': 'gemm_tcgen05_pair_kernel<6,1> (video input projection [%d,%d]x[%d,%d], tcgen05 cta_group::2)' % (M, K, K, N),
                'achieved': achieved_tf, 'peak': pk['tf_burst'], 'unit': 'TFLOP/s', 'frac': achieved_tf / pk['tf_burst'],
                'frac_of_sustained_peak': achieved_tf / pk['tf_sustained'],
                'peak_source': pk['src'] + ': burst bf16 (the kernel is timed alone, one launch between two events)',
                'traffic': traffic, 'traffic_source': traffic_src, 'ms': gemm_ms,
                'flops_per_launch': flops, 'algorithmic_bytes_per_launch': 2.0 * (M * K + N * K + M * N),
                'step_share': gemm_ms / (ms / args.steps)}

    # whole step against the tensor roof: algorithmic GEMM flops of this rank's batch (encoders, decoder, every module
    # instance's Linears by the per-instance column of SURVEY 8a) over the device-timed step
    step_flops = algorithmic_step_flops(batch, cfg)
    step_tf = step_flops / (ms / args.steps * 1e-3) / 1e12
    roofline['whole_step'] = {'flops': step_flops, 'mflop_per_question': step_flops / B / 1e6, 'achieved': step_tf, 'unit': 'TFLOP/s',
                              'frac_of_sustained_peak': step_tf / pk['tf_sustained'], 'frac_of_burst_peak': step_tf / pk['tf_burst'],
                              'what': 'the step is a chain of 81 launches; the recurrence (latency-bound, 0.33 ms) and the module phase '
                                      '(0.44 ms of small GEMMs) are not tensor-bound, see DESIGN.md section 4'}

    # ---- CPU legs: baseline timing (rank 0, N = 1) and the parity check of this rank's answers against the oracle -------------------
    cpu = None
    parity = None
    if not args.no_cpu_baseline:
        ncpu = os.cpu_count() or 1
        if world == 1:
            qps1, med1, _ = cpu_reference_timer(qs, weights, cfg, min_seconds=5.0, max_passes=20, threads=1)
            qps, med, npass = cpu_reference_timer(qs, weights, cfg)
            cpu = {'value': qps, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                   'single_thread': {'value': qps1, 'cores': 1, 'median_s_per_pass': med1},
                   'sample': 'first %d questions of the GPU batch, %d passes, median %.3f s/pass, torch threads=%d; oracle port with the '
                             'encoders through torch.nn.LSTM' % (CPU_SAMPLE, npass, med, torch.get_num_threads())}
        n_check = B if world == 1 else min(B, PARITY_PER_RANK_MULTI)
        mine = parity_report(model, strict_model if world == 1 else None, qs[:n_check], weights, cfg, answers_dev, logits_dev,
                             threads=max(1, ncpu // world))
        if world > 1:
            keys = [k for k, v in mine.items() if isinstance(v, int) and not isinstance(v, bool)]
            sums = torch.tensor([mine[k] for k in keys], device=dev, dtype=torch.int64)
            dist.all_reduce(sums)
            mx = torch.tensor([mine['bf16_mismatch_max_margin_rel'], mine['bf16_logit_err_rel_max'], mine['localize_mismatch_max_margin'],
                               mine['oracle_seconds']], device=dev)
            dist.all_reduce(mx, op=dist.ReduceOp.MAX)
            ok = torch.tensor([1 if gathered_ok else 0], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            parity = dict(mine)
            parity.update({k: int(v) for k, v in zip(keys, sums.tolist())})
            parity.update({'bf16_mismatch_max_margin_rel': float(mx[0]), 'bf16_logit_err_rel_max': float(mx[1]),
                           'localize_mismatch_max_margin': float(mx[2]), 'oracle_seconds': float(mx[3]),
                           'ranks': world, 'per_rank': n_check, 'gathered_answers_are_each_ranks_answers': bool(int(ok.item())),
                           'what': 'every rank checks the first %d questions of ITS shard (answers as gathered) against the CPU oracle; '
                                   'counts summed over ranks (medians / p99 are rank 0\'s)' % n_check})
        else:
            parity = mine

    if rank == 0:
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
                'data': 'synthetic',
                'config': {'workload': workload_string(args, B) + ', bf16 storage / fp32 accumulate', 'questions_per_gpu': B,
                           'global_questions': world * B, 'frames': T, 'video_size': V,
                           'hidden_size': cfg['hidden_size'], 'parallelism': 'question-sharded x%d, answers all-gathered (NCCL, one step behind the compute)' % world,
                           'l2': 'inputs larger than L2 (video %.0f MB per step)' % (B * T * V * 2 / 1e6)},
                'clocks': clock_info,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4 * B * world, 'chunks': E2E_CHUNKS,
                        'ms_per_step': 1e3 * e2e_s / args.steps,
                        'per_call': {'value': world * B * args.steps / percall_s, 'ms_per_step': 1e3 * percall_s / args.steps,
                                     'what': 'forward_pipelined(host chunks) + answers.cpu() per step, synchronising every step'},
                        'from_dicts': from_dicts,
                        'host_numa_binding': numa,
                        'timer': 'wall clock between synchronize()s over all steps; VideoNMN.forward_stream: every step uploads its pinned host batch '
                                 '(%d chunks, copy stream) and reads its answers back (async D2H into pinned memory), 2 steps in flight' % E2E_CHUNKS},
                'gpu_launches': launches_per_step * args.steps, 'launches_per_step': launches_per_step, 'host_enqueue_ms': host_enqueue_ms,
                'roofline': roofline, 'roofline_hbm': roofline_hbm, 'phases_ms': ph_ms, 'strict': strict, 'audit': audit, 'i3d': i3d, 'train': train,
                'cpu_baseline': cpu, 'parity': parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
