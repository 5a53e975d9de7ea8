"""GPU (>= 2 devices): question-sharded inference over NCCL == unsharded inference (answers and logits bit-identical), and the
data-parallel training step (gradient all-reduce overlapped with BPTT, window-global contrastive negatives, union of the touched-parameter
sets) == one process running the whole window.  Parametrised over world sizes 2 / 4 / 8 (skipped beyond the box's GPU count); run on
hardware with `gpurun --gpus N`, logs under profiles/ (SURVEY.md §8e)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from stair_b200 import VideoNMN, synthetic as syn
    from stair_b200.distributed import ShardedNMN
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        cfg = syn.model_config(T=8, V=256, hidden=128)
        torch.manual_seed(0)
        model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='bf16').cuda().eval()
        qs = syn.make_questions(101, 8, 256, seed=5, templates=list(syn.ALL_TEMPLATES))
        ans, logits = ShardedNMN(model).answer(qs, gather_logits=True)
        full = model(qs, return_res_by_step=False, test_mode=True)
        torch.cuda.synchronize()
        ok = torch.equal(ans, full['answers']) and torch.equal(logits, full['logits'])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def _spawn(target, world):
    if torch.cuda.device_count() < world:
        pytest.skip('needs %d GPUs' % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=target, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    return res


@pytest.mark.parametrize('world', [2, 4, 8])
def test_sharded_equals_unsharded(world):
    res = _spawn(_worker, world)
    assert res == [(r, True) for r in range(world)]


def _train_worker(rank, world, port, q):
    """Data-parallel training step: each rank owns half of the window; after the NCCL all-reduce every rank must hold the
    gradients of the whole window (== one process running the full window), incl. the window-wide contrastive negatives."""
    import torch.distributed as dist
    from stair_b200 import VideoNMN, synthetic as syn
    from stair_b200.distributed import shard
    from stair_b200.train import NMNTrainStep
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    try:
        cfg = syn.model_config(T=8, V=256, hidden=128, object_types=16)
        torch.manual_seed(0)
        model = VideoNMN(cfg, pretrain_modules=syn.PRETRAIN_MODULES, precision='fp32').cuda().train()
        qs = syn.make_questions(68, 8, 256, seed=5, templates=list(syn.ALL_TEMPLATES), with_gold=True, object_types=16)
        dp = NMNTrainStep(model)
        out = dp(shard(qs, rank, world))
        torch.cuda.synchronize()
        g_dp = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        loss_dp = float(out['loss'])
        model.zero_grad(set_to_none=True)
        single = NMNTrainStep(model, distributed=False)
        out1 = single(qs)
        torch.cuda.synchronize()
        g_1 = {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}
        ok = set(g_dp) == set(g_1) and abs(loss_dp - float(out1['loss'])) <= 1e-5 * abs(loss_dp)
        worst = 0.0
        for k in g_1:
            scale = float(g_1[k].abs().max())
            worst = max(worst, float((g_dp[k] - g_1[k]).abs().max()) / max(scale, 1e-12) if scale > 1e-9 else 0.0)
        q.put((rank, bool(ok and worst <= 1e-4), worst))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 4, 8])
def test_data_parallel_gradients_equal_single_process(world):
    res = _spawn(_train_worker, world)
    assert [r[:2] for r in res] == [(r, True) for r in range(world)], res
    print('world %d: worst relative gradient difference sharded vs single process %.3g' % (world, max(r[2] for r in res)))
