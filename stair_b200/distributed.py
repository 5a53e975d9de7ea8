"""Question sharding across the GPUs of one box.

Questions are independent in the forward (video_nmn/module_net.py:65-145 has no cross-question op), so the batch is
split into contiguous slices, one per rank; weights are replicated; every rank groups and executes its own slice.  The
only data-path collective of inference is the all-gather of the int32 answers (4 B/question), plus optionally the fp32
logits.  One process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch; gloo in the CPU tests).
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        lo, _, hi = part.partition('-')
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int, sysfs: str = '/sys') -> dict:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is allocated.

    Every rank streams its own batches over its own PCIe link (no data-path collective); with 8 ranks the host side
    only keeps up if each rank's pinned buffers are first-touched on the memory controller next to its GPU (measured without
    binding: end-to-end 2.4 M q/s at 8 GPUs vs 8 x 0.72 M).  Returns what was found / done; never raises (containers may hide
    sysfs or restrict the CPU set — then the process is left as it is)."""
    import os
    info = {'device': device_index, 'numa_node': None, 'bound_cpus': 0}
    try:
        bus = torch.cuda.get_device_properties(device_index)
        bus_id = '%04x:%02x:%02x.0' % (bus.pci_domain_id, bus.pci_bus_id, bus.pci_device_id)
        info['pci'] = bus_id
        with open(os.path.join(sysfs, 'bus/pci/devices', bus_id, 'numa_node')) as f:
            node = int(f.read().strip())
        info['numa_node'] = node
        if node < 0:
            return info
        with open(os.path.join(sysfs, 'devices/system/node/node%d/cpulist' % node)) as f:
            cpus = _parse_cpulist(f.read())
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            info['bound_cpus'] = len(allowed)
    except Exception as e:                                            # noqa: BLE001 — diagnostics only
        info['error'] = '%s: %s' % (type(e).__name__, e)
    return info


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous slice [lo, hi) of ``n`` questions owned by ``rank``; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(items, rank=None, world=None):
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    lo, hi = shard_bounds(len(items), rank, world)
    return items[lo:hi]


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (sizes given by ``shard_bounds``) into the full [n_total, ...] tensor on every
    rank.  Ragged slices are padded to the largest slice for the collective and trimmed afterwards."""
    world = dist.get_world_size(group)
    sizes = [shard_bounds(n_total, r, world)[1] - shard_bounds(n_total, r, world)[0] for r in range(world)]
    mx = max(sizes)
    if local.shape[0] != sizes[dist.get_rank(group)]:
        raise ValueError('local block has %d rows, expected %d' % (local.shape[0], sizes[dist.get_rank(group)]))
    pad = local
    if local.shape[0] < mx:
        pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        pad[:local.shape[0]] = local
    out = torch.empty((world * mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    if all(s == mx for s in sizes):
        return out
    return torch.cat([out[r * mx:r * mx + sizes[r]] for r in range(world)])


class AnswerGather:
    """NCCL all-gather of every step's answers, ``depth`` steps deep: step k's gather runs on NCCL's stream while step k+1 computes, and
    the compute stream only waits for a gather when its buffer comes round again (or at ``finish``).  A synchronous gather makes every step
    a rendez-vous of all ranks — each step then lasts as long as ITS slowest rank, which at 8 ranks costs more than the 131 KB collective
    itself; one step of slack absorbs that jitter.  Equal shard sizes (``n_local`` answers per rank)."""

    def __init__(self, n_local: int, device, group=None, depth: int = 2, dtype=torch.int32):
        self.group, self.depth = group, max(1, int(depth))
        world = dist.get_world_size(group)
        self.bufs = [torch.empty(world * n_local, dtype=dtype, device=device) for _ in range(self.depth)]
        self.pending = []                                        # (work, buffer, input kept alive), oldest first
        self.k = 0

    def submit(self, answers: torch.Tensor) -> torch.Tensor:
        """Start gathering ``answers`` (this rank's [n_local]); returns the buffer the result lands in (valid after ``finish`` or after
        ``depth`` further submits)."""
        if len(self.pending) >= self.depth:
            self.pending.pop(0)[0].wait()                        # stream-side wait: the buffer about to be reused is complete
        buf = self.bufs[self.k % self.depth]
        self.k += 1
        work = dist.all_gather_into_tensor(buf, answers.contiguous(), group=self.group, async_op=True)
        self.pending.append((work, buf, answers))
        return buf

    def finish(self):
        """Make the current stream wait for every outstanding gather; returns the most recent result (None before the first submit)."""
        last = self.pending[-1][1] if self.pending else (self.bufs[(self.k - 1) % self.depth] if self.k else None)
        for work, _, _ in self.pending:
            work.wait()
        self.pending.clear()
        return last


class ShardedNMN:
    """Batch-sharded inference: ``model`` is a (replicated) ``stair_b200.VideoNMN`` on this rank's GPU."""

    def __init__(self, model, group=None):
        self.model, self.group = model, group

    @torch.no_grad()
    def answer(self, questions, gather_logits=False):
        """questions: the FULL list of data dicts (every rank passes the same list).  Returns (answers [B] int32,
        logits [B, A] or None) on every rank."""
        n = len(questions)
        mine = shard(questions, dist.get_rank(self.group), dist.get_world_size(self.group))
        out = self.model(mine, return_res_by_step=False, test_mode=True)
        answers = all_gather_rows(out['answers'], n, self.group)
        logits = all_gather_rows(out['logits'], n, self.group) if gather_logits else None
        return answers, logits
