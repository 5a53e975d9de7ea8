"""CPU, world_size 2 over gloo: question sharding + answer/logit gather (the N>1 host path of bench.py / ShardedNMN)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stair_b200 import distributed as D


def test_shard_bounds_cover_everything():
    for n in (0, 1, 7, 8, 4096, 32769):
        for world in (1, 2, 3, 8):
            spans = [D.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        lo, hi = D.shard_bounds(n, rank, world)
        items = list(range(n))
        assert D.shard(items) == items[lo:hi]
        local_ans = (torch.arange(lo, hi, dtype=torch.int32) * 7) % 172           # stand-in for the rank's argmax answers
        local_logits = torch.arange(lo, hi, dtype=torch.float32)[:, None] + torch.arange(5)[None, :] * 0.5
        ans = D.all_gather_rows(local_ans, n)
        logits = D.all_gather_rows(local_logits, n)
        ok = torch.equal(ans, (torch.arange(n, dtype=torch.int32) * 7) % 172) and \
            torch.equal(logits, torch.arange(n, dtype=torch.float32)[:, None] + torch.arange(5)[None, :] * 0.5)
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('n', [8, 11])
def test_gather_world2_gloo(n):
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def test_numa_binding_helper_is_tolerant(tmp_path):
    """bind_to_gpu_numa_node: cpulist parsing, and no exception when there is no GPU / no sysfs entry (containers)."""
    from stair_b200.distributed import _parse_cpulist, bind_to_gpu_numa_node
    assert _parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist('') == set()
    info = bind_to_gpu_numa_node(0, sysfs=str(tmp_path))
    assert info['bound_cpus'] == 0 and info['numa_node'] is None


def _gather_worker(rank, world, port, out_q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    try:
        n = 6
        g = D.AnswerGather(n, 'cpu', depth=2)
        assert g.finish() is None
        ok, bufs = True, []
        for step in range(5):                                            # five steps through two buffers: results checked one step late
            local = torch.arange(n, dtype=torch.int32) + 100 * rank + 1000 * step
            bufs.append((step, g.submit(local)))
        last = g.finish()
        want = torch.cat([torch.arange(n, dtype=torch.int32) + 100 * r + 1000 * 4 for r in range(world)])
        ok = ok and torch.equal(last, want) and last is bufs[-1][1]
        prev = bufs[-2][1]                                                # step 3's result sits in the other buffer
        ok = ok and torch.equal(prev, torch.cat([torch.arange(n, dtype=torch.int32) + 100 * r + 1000 * 3 for r in range(world)]))
        out_q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_pipelined_answer_gather_world2_gloo():
    """distributed.AnswerGather (the bench's per-step NCCL gather, one step behind the compute): order and contents over gloo."""
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
