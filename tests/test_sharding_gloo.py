"""world_size-2 gloo test of the env-sharding host logic: shard-concat == single run, statistics all-reduce.
The per-shard stepper here is the CPU oracle (the CUDA kernel cannot run in this tier); the sharding code
under test (quadruped_gym_b200/sharding.py) is the one bench.py and the envs use.  CPU only."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from quadruped_gym_b200.sharding import reduce_rollout_stats, max_over_ranks, shard_range


def test_shard_range_partitions():
    for n, w in ((65536, 8), (10, 3), (7, 2), (5, 8)):
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _actions(n_total, T):
    return np.random.default_rng(99).uniform(-1, 1, (T, n_total, 12))


def _run_shard(start, stop, T):
    from oracle.oracle import OracleBatch, OracleModel
    from quadruped_gym_b200.model import DEFAULT_BLOB
    om = OracleModel(open(DEFAULT_BLOB, "rb").read())
    ob = OracleBatch(om, stop - start)
    acts = _actions(6, T)[:, start:stop]      # actions keyed on the GLOBAL env id
    return ob.rollout(acts, 4, 10.0, True, want_obs=True)


def _worker(rank, world, port, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = shard_range(6, rank, world)
    obs = _run_shard(a, b, T)
    stats = reduce_rollout_stats({"env_steps": float((b - a) * T), "reward_sum": float(obs.sum())})
    tmax = max_over_ranks(1.0 + rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, obs)
    if rank == 0:
        q.put((np.concatenate(gathered, axis=1), stats, tmax))
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_run():
    T = 25
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, T, q)) for r in range(2)]
    for p in procs:
        p.start()
    obs2, stats, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    obs1 = _run_shard(0, 6, T)
    assert np.array_equal(obs1, obs2)                    # no exchange on the physics path: bit-identical
    assert stats["env_steps"] == 6 * T and stats["reward_sum"] == pytest.approx(float(obs1.sum()), rel=1e-12)
    assert tmax == 2.0                                   # timing = max over ranks
