"""world_size-2 test of the data-parallel plumbing on CPU (gloo): flat-gradient all-reduce + scaler, per-rank shard
seeds, max-over-ranks timing reduction, and that identical optimizer steps keep replicas in lock-step."""
import os
import socket

import numpy as np
import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from coach_b200 import parallel
    r, w = parallel.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    assert parallel.shard_seed(100) == 100 + rank
    # each rank has its own shard-local gradient; after the all-reduce everyone holds the sum
    g = torch.full((1000,), float(rank + 1))
    scaler = parallel.allreduce_gradients(g, scale_down=True)
    theta = torch.zeros(1000)
    theta -= 0.1 * g * scaler                   # identical "optimizer step" on every rank
    g2 = torch.full((10,), float(rank + 1))
    scaler2 = parallel.allreduce_gradients(g2, scale_down=False)
    t = parallel.max_over_ranks(10.0 * (rank + 1))
    # shared observation statistics: increments of both shards merged (SharedRunningStats semantics)
    rng = np.random.RandomState(rank)
    x = rng.randn(50 + 10 * rank, 17)
    d_sum, d_sq = torch.from_numpy(x.sum(0)), torch.from_numpy((x * x).sum(0))
    rows = parallel.allreduce_running_stats(d_sum, d_sq, x.shape[0])
    # advantage standardisation over the whole distributed rollout
    adv = torch.from_numpy(np.concatenate([rng.randn(40 + rank), [7.0, 7.0]]))
    mean, std = parallel.global_standardize_(adv, 40 + rank)
    out_q.put((rank, float(g[0]), scaler, float(theta[0]), float(g2[0]), scaler2, t, rows, d_sum.numpy(), d_sq.numpy(),
               adv.numpy(), mean, std))
    torch.distributed.destroy_process_group()


def test_allreduce_world2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    xs = [np.random.RandomState(r).randn(50 + 10 * r, 17) for r in range(world)]
    advs = []
    for r in range(world):
        rng = np.random.RandomState(r)
        rng.randn(50 + 10 * r, 17)
        advs.append(rng.randn(40 + r))
    all_adv = np.concatenate(advs)
    for rank, gsum, scaler, theta0, g2, scaler2, t, rows, d_sum, d_sq, adv, mean, std in res:
        assert rows == sum(x.shape[0] for x in xs)
        np.testing.assert_allclose(d_sum, sum(x.sum(0) for x in xs), rtol=1e-13)
        np.testing.assert_allclose(d_sq, sum((x * x).sum(0) for x in xs), rtol=1e-13)
        np.testing.assert_allclose([mean, std], [all_adv.mean(), all_adv.std()], rtol=1e-13)
        np.testing.assert_allclose(adv[:40 + rank], (advs[rank] - all_adv.mean()) / all_adv.std(), rtol=1e-12)
        assert np.all(np.isnan(adv[40 + rank:]))
        assert gsum == 3.0 and scaler == 0.5            # sum over ranks, mean applied by the optimizer step
        assert np.isclose(theta0, -0.1 * 3.0 * 0.5)     # replicas stay identical
        assert g2 == 3.0 and scaler2 == 1.0             # DDPG/TD3 semantics: sum, no scale-down
        assert t == 20.0                                # slowest rank


def test_single_process_is_a_noop():
    from coach_b200 import parallel
    g = torch.ones(4)
    assert parallel.allreduce_gradients(g) == 1.0 and parallel.world() == (0, 1)
    assert parallel.max_over_ranks(3.5) == 3.5
