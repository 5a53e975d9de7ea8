"""Why does the end-to-end arm stop scaling beyond 2 GPUs?  

Every rank of bench.py streams its own pinned batch (309 MB per 4096 questions) over its own PCIe link; there is no data-path collective.
This script isolates the host side: every rank only copies, no compute.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P profiles/h2d_diag.py

For each N it prints, per rank and in aggregate, the H2D bandwidth of
  (a) all ranks copying at the same time, one 309 MB buffer per step      (what the e2e arm does)
  (b) the same bytes as 3 buffers per step (video / question / int tables split like NMNBatch.to)
  (c) one rank at a time (the others idle)                                 (the link's own ceiling)
  (d) all ranks at the same time, each from a buffer that was first touched by a thread pinned to another CPU set (NUMA sensitivity)
and a host-memory read bandwidth probe (all ranks summing their pinned buffer on the CPU) — if (a) saturates at the same aggregate as
the CPU-side read probe the ceiling is the host's memory system, not PCIe and not this framework.
"""
import os
import sys
import time

import torch
import torch.distributed as dist

BYTES = 308_685_496          # h2d_bytes_per_step of the bench's RX batch
STEPS = 12


def main():
    rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    host = torch.empty(BYTES, dtype=torch.uint8, pin_memory=True)
    host.fill_(1)
    parts = [torch.empty(n, dtype=torch.uint8, pin_memory=True).fill_(1) for n in (268_435_456, 39_321_600, BYTES - 268_435_456 - 39_321_600)]
    dst = torch.empty(BYTES, dtype=torch.uint8, device=dev)
    dparts = [torch.empty(p.numel(), dtype=torch.uint8, device=dev) for p in parts]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps=STEPS):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        barrier()
        return BYTES * steps / dt / 1e9

    def gather(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            return [float(o) for o in out]
        return [x]

    one = gather(timed(lambda: dst.copy_(host, non_blocking=True)))
    three = gather(timed(lambda: [d.copy_(p, non_blocking=True) for d, p in zip(dparts, parts)]))
    solo = []
    for r in range(world):
        barrier()
        if r == rank:
            for _ in range(2):
                dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for _ in range(STEPS):
                dst.copy_(host, non_blocking=True)
            torch.cuda.synchronize()
            mine = BYTES * STEPS / (time.perf_counter() - t0) / 1e9
        barrier()
    solo = gather(mine)
    # CPU-side read of the same pinned memory, all ranks at once (one torch thread pool each)
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))
    barrier()
    t0 = time.perf_counter()
    for _ in range(4):
        s = int(host.view(torch.int64)[: BYTES // 8].sum())
    cpu_read = gather(BYTES * 4 / (time.perf_counter() - t0) / 1e9)
    if rank == 0:
        fmt = lambda v: ' '.join('%5.1f' % x for x in v)      # noqa: E731
        print('N = %d ranks, %d MB per copy step, %d host threads visible' % (world, BYTES // 1_000_000, os.cpu_count() or 0))
        print('(a) all ranks, 1 copy / step   per rank GB/s: %s   aggregate %.1f' % (fmt(one), sum(one)))
        print('(b) all ranks, 3 copies / step per rank GB/s: %s   aggregate %.1f' % (fmt(three), sum(three)))
        print('(c) one rank at a time         per rank GB/s: %s   (link ceiling)' % fmt(solo))
        print('(e) CPU read of the pinned buffers, all ranks at once (torch sum, %d threads each) GB/s: %s   aggregate %.1f'
              % (torch.get_num_threads(), fmt(cpu_read), sum(cpu_read)))
        try:
            import subprocess
            print(subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout[:3000])
            print(subprocess.run(['lscpu'], capture_output=True, text=True).stdout[:1500])
        except OSError:
            pass
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
