#!/usr/bin/env python
"""Benchmark of the NMN hot path: questions/sec of the batched ModuleNet forward on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]             # our arm (CUDA library through the C ABI)
    python bench.py --impl reference [--gpus N] [--steps K] ...     # the reference's CPU path on the host cores

Workload (config.workload): BASELINE.json configs[1] — 4096 mixed-program questions per GPU (the 10 AGQA layout
templates of SURVEY.md App. B), RX/TGIF-QA-shaped features [8, 4096] (appearance 8x16x2048 mean-pooled + motion 8x2048,
video_nmn/dataset.py:150-172), questions of 8-24 GloVe-sized words, random-init weights, bf16 storage / fp32 accumulate,
inference (test_mode=True, return_res_by_step=False).  N > 1: every rank runs its own 4096 questions (weak scaling,
configs[2]: 32768 questions at 8 GPUs) and the int32 answers are all-gathered with NCCL inside the timed step.

One JSON line on stdout (rank 0): see the contract in the task statement; extra keys: roofline, cpu_baseline, phases_ms,
parity.  The oracle (oracle/nmn_oracle.py) is used ONLY as the cpu_baseline / --impl reference timer and as the checker of
the first 32 answers — never on the measured path.
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

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC, UNIT = 'nmn_questions_per_sec', 'questions/s'
PER_GPU_B = 4096
T, V = 8, 4096
CPU_SAMPLE = 32
TRAIN_DROPOUT = 0.25
E2E_CHUNKS = 2


def peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return {'hbm_gbs': d['hbm_gbs'], 'tf_burst': d['bf16_tflops'], 'tf_sustained': d.get('bf16_tflops_sustained', d['bf16_tflops']),
                'src': 'measured'}
    return {'hbm_gbs': 6650.0, 'tf_burst': 1590.0, 'tf_sustained': 1400.0, 'src': 'fallback'}


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
    logits = None
    with torch.no_grad():
        for d in sample[:4]:
            model(d, return_res_by_step=False, test_mode=True)
        times = []
        t_all = time.perf_counter()
        while True:
            t0 = time.perf_counter()
            logits = [model(d, return_res_by_step=False, test_mode=True)['logits'] for d in sample]
            times.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all >= min_seconds or len(times) >= max_passes:
                break
    med = statistics.median(times)
    return len(sample) / med, med, len(times), torch.stack(logits)


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
            # the measured arm's config (same workload, same keys), plus what one reference step is
            'config': {'workload': ('ModuleNet batched inference, %d mixed-program questions per GPU (10 AGQA layout templates, 2-12 modules), '
                                    'RX/TGIF-QA features [8,4096] bf16, questions 8-24 words x 300, random init, bf16 storage / fp32 accumulate'
                                    % args.batch) if args.workload == 'rx' else
                                   ('I3D stress test (BASELINE configs[4]): %d questions per GPU, compare / xor_between layouts (9-12 modules), '
                                    'features [64,1024] bf16, conv-mode Temporal, random init, bf16 storage / fp32 accumulate' % args.batch),
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
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-train', action='store_true', help='skip the training-step leg (BASELINE configs[3])')
    ap.add_argument('--no-overlap-allreduce', action='store_true',
                    help='training leg at N > 1: one gradient all-reduce after the whole backward instead of overlapping it with BPTT (comparison)')
    ap.add_argument('--workload', default='rx', choices=['rx', 'i3d'],
                    help="rx (default, BASELINE configs[1]): T=8, V=4096, the 10 AGQA templates; i3d (configs[4] stress test): T=64, V=1024, "
                         "conv-mode Temporal, only the >= 9-module layouts (compare, xor_between)")
    args = ap.parse_args()
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    global T, V, TEMPLATES
    if args.workload == 'i3d':
        T, V, TEMPLATES = 64, 1024, ['compare', 'xor_between']
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
    from stair_b200 import VideoNMN, synthetic as syn, _lib as L
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

    def step_device():
        st = model.forward_batch(batch)
        if world > 1:
            dist.all_gather_into_tensor(gathered, st.answers)
        return st

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed throughput: inputs resident in HBM ---------------------------------------------------------
    batch.to(dev)
    torch.cuda.synchronize()
    for _ in range(args.warmup):
        st = step_device()
    model.check_status(st)
    launches_per_step = model.last_launches
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st = step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = world * B * args.steps / (ms * 1e-3)
    answers_dev = st.answers.cpu()

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
    # ---- audit mode: the same forward with every pretrain head computed (res_by_step / result_of_each_step:
    # FilterFrame [T, O] head GEMM, L2-normalised Filter / ToAction / Superlative outputs, Exists / Xor / Equals heads) -------------
    heads = model._head_modules(True, True)
    for _ in range(3):
        model.forward_batch(batch, heads)
    barrier()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(nrep):
        model.forward_batch(batch, heads)
    a1.record()
    barrier()
    ta = torch.tensor([a0.elapsed_time(a1)], device=dev)
    if world > 1:
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
    audit_ms = float(ta.item()) / nrep
    audit = {'value': world * B / (audit_ms * 1e-3), 'unit': UNIT, 'ms_per_step': audit_ms, 'launches_per_step': model.last_launches,
             'what': 'forward with all pretrain heads (return_res_by_step / return_result_of_each_step device work)'}

    # ---- training step (BASELINE configs[3]): forward with history + intermediate-supervision losses + backward +
    # gradient all-reduce (N > 1) + Adam, one window = the rank's 4096 questions; device-timed, max over ranks -------------
    train = None
    if not args.no_train:
        from stair_b200.train import NMNTrainStep, FusedAdam
        # throughput run: the reference's default training dropout (video_nmn/args.py:31); parity runs (tests) use 0 or injected masks
        tmodel = VideoNMN(dict(cfg, dropout=TRAIN_DROPOUT), pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16')
        tmodel.load_state_dict(weights)
        tmodel = tmodel.to(dev).train()
        tstep = NMNTrainStep(tmodel, overlap_allreduce=not args.no_overlap_allreduce)
        opt = FusedAdam(tmodel, lr=2e-4)
        plan = tstep.plan(batch)

        def train_step():
            out = tstep.run(plan)
            opt.step()
            opt.zero_grad()
            return out

        for _ in range(2):
            out = train_step()
        tmodel.check_status(out['state'])
        ksteps = max(2, min(args.steps, 5))
        barrier()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0e.record()
        for _ in range(ksteps):
            out = train_step()
        t1e.record()
        barrier()
        tms = torch.tensor([t0e.elapsed_time(t1e)], device=dev)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        tms = float(tms.item())
        train = {'value': world * B * ksteps / (tms * 1e-3), 'unit': UNIT, 'ms_per_step': tms / ksteps, 'steps': ksteps,
                 'launches_per_step': tstep.last_launches, 'window_questions': world * B, 'loss': float(out['loss']),
                 'loss_rows': out['loss_counts'], 'dropout': TRAIN_DROPOUT,
                 'what': 'forward with encoder history + losses (train_module.py:83-194) + backward + %sAdam (one fused multi-tensor kernel that also refreshes the bf16 weight copies); bf16 storage, fp32 gradients'
                         % ('NCCL gradient all-reduce + ' if world > 1 else '')}
        # the reference-faithful window: 32 questions per optimizer step (train_module.py gradient_accumulation = 32), latency-bound
        plan32 = tstep.plan(qs[:32])

        def train_step32():
            o = tstep.run(plan32)
            opt.step()
            opt.zero_grad()
            return o

        for _ in range(3):
            train_step32()
        barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(10):
            train_step32()
        w1.record()
        barrier()
        w32 = w0.elapsed_time(w1) / 10
        train['window32'] = {'ms_per_step': w32, 'value': 32 / (w32 * 1e-3), 'unit': UNIT,
                             'what': 'one 32-question window per optimizer step on this rank (reference accumulation window; launch-latency bound)'}
        del tmodel, tstep, opt, plan, plan32
        torch.cuda.empty_cache()

    pk = peaks()
    flops = 2.0 * M * N * K
    achieved_tf = flops / (gemm_ms * 1e-3) / 1e12
    # dram__bytes_read.sum + dram__bytes_write.sum of this kernel at the default shape, from the committed `ncu --set full` capture
    # (profiles/r1_gemm_xproj_ncu_summary_v3.txt); algorithmic bytes: A 268 MB + W 17 MB + C 134 MB = 419 MB
    traffic = 398.9e6 if (args.workload == 'rx' and B == PER_GPU_B) else None
    roofline = {'bound': 'tensor', 'kernel': 'gemm_tcgen05_kernel<256,4> (video input projection [%d,%d]x[%d,%d])' % (M, K, K, N),
                'achieved': achieved_tf, 'peak': pk['tf_sustained'], 'unit': 'TFLOP/s', 'frac': achieved_tf / pk['tf_sustained'],
                'frac_of_burst_peak': achieved_tf / pk['tf_burst'], 'peak_source': pk['src'] + ' (sustained; kernel timed inside the step loop)',
                'traffic': traffic, 'traffic_source': 'profiles/r1_gemm_xproj_ncu_summary_v3.txt' if traffic else None, 'ms': gemm_ms,
                'flops_per_launch': flops, 'algorithmic_bytes_per_launch': 2.0 * (M * K + N * K + M * N)}

    line = None
    if rank == 0:
        cpu = None
        parity = None
        if not args.no_cpu_baseline and world == 1:
            qps1, med1, _, _ = cpu_reference_timer(qs, weights, cfg, min_seconds=5.0, max_passes=20, threads=1)
            qps, med, npass, ref_logits = cpu_reference_timer(qs, weights, cfg)
            cpu = {'value': qps, 'unit': UNIT, 'cores': torch.get_num_threads(), 'kind': 'port',
                   'single_thread': {'value': qps1, 'cores': 1, 'median_s_per_pass': med1},
                   'sample': 'first %d questions of the GPU batch, %d passes, median %.3f s/pass, torch threads=%d; oracle port with the '
                             'encoders through torch.nn.LSTM' % (CPU_SAMPLE, npass, med, torch.get_num_threads())}
            ref_ans = ref_logits.argmax(1)
            top2 = ref_logits.topk(2, dim=1).values
            margin = (top2[:, 0] - top2[:, 1])
            clear = margin > 2 * (3e-2 * ref_logits.abs().max(1).values + 2e-3)
            got = answers_dev[:CPU_SAMPLE].long()
            # strict mode on the same questions: answers must be bit-identical
            strict = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32')
            strict.load_state_dict(weights)
            strict = strict.to(dev).eval()
            s_out = strict(qs[:CPU_SAMPLE], return_res_by_step=False, test_mode=True)
            s_ans = s_out['answers'].cpu().long()
            parity = {'questions': CPU_SAMPLE, 'fp32_strict_answers_equal': int((s_ans == ref_ans).sum()),
                      'fp32_strict_max_logit_err': float((s_out['logits'].cpu() - ref_logits).abs().max()),
                      'bf16_answers_equal': int((got == ref_ans).sum()), 'bf16_clear_margin': int(clear.sum()),
                      'bf16_answers_equal_where_margin_clear': int(((got == ref_ans) & clear).sum())}
        line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
                'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16',
                'data': 'synthetic',
                'config': {'workload': ('ModuleNet batched inference, %d mixed-program questions per GPU (10 AGQA layout templates, 2-12 modules), '
                                        'RX/TGIF-QA features [8,4096] bf16, questions 8-24 words x 300, random init, bf16 storage / fp32 accumulate'
                                        % B) if args.workload == 'rx' else
                                       ('I3D stress test (BASELINE configs[4]): %d questions per GPU, compare / xor_between layouts (9-12 modules), '
                                        'features [64,1024] bf16, conv-mode Temporal, random init, bf16 storage / fp32 accumulate' % B), 'questions_per_gpu': B, 'global_questions': world * B, 'frames': T, 'video_size': V,
                           'hidden_size': cfg['hidden_size'], 'parallelism': 'question-sharded x%d, answers all-gathered (NCCL)' % world,
                           'l2': 'inputs larger than L2 (video %.0f MB per step)' % (B * T * V * 2 / 1e6)},
                'clocks': clock_info,
                'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': h2d_bytes, 'd2h_bytes_per_step': 4 * B * world, 'chunks': E2E_CHUNKS,
                        'ms_per_step': 1e3 * e2e_s / args.steps,
                        'per_call': {'value': world * B * args.steps / percall_s, 'ms_per_step': 1e3 * percall_s / args.steps,
                                     'what': 'forward_pipelined(host chunks) + answers.cpu() per step, synchronising every step'},
                        'host_numa_binding': numa,
                        'timer': 'wall clock between synchronize()s over all steps; VideoNMN.forward_stream: every step uploads its pinned host batch '
                                 '(%d chunks, copy stream) and reads its answers back (async D2H into pinned memory), 2 steps in flight' % E2E_CHUNKS},
                'gpu_launches': launches_per_step * args.steps, 'launches_per_step': launches_per_step,
                'roofline': roofline, 'phases_ms': ph_ms, 'audit': audit, 'train': train, 'cpu_baseline': cpu, 'parity': parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
