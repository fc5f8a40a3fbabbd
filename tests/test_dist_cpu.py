"""Host-side sharding logic over two `gloo` ranks on the CPU (the N > 1 path of bench.py):
the partition is exact, and the all-reduced episode statistics of two shards equal the
statistics of the unsharded batch.  The per-shard sums come from the CPU oracle here (the
test infrastructure stand-in for `BatchedEnviron.shard_stats()`)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ris_vec_marl_b200._lib import NSTAT, STAT_COLUMNS
from ris_vec_marl_b200.dist import global_episode_stats, shard_range


def test_shard_range_is_an_exact_partition():
    for n in (1, 7, 4096, 4097, 65536):
        for world in (1, 2, 3, 4, 8):
            if world > n:
                continue
            seen = []
            for r in range(world):
                s, c = shard_range(n, r, world)
                seen.extend(range(s, s + c))
            assert seen == list(range(n))
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _oracle_shard_sums(e0, count, E_total):
    """stats column sums of envs [e0, e0 + count) after one MARL step of a seeded batch."""
    from oracle.env_oracle import EnvOracle, InjectedDraws, OracleParams, encode_groups

    rng = np.random.default_rng(123)  # the SAME global inputs on every rank
    V, M = 8, 40
    reset = np.zeros((E_total, 19), dtype=np.int64)
    pattern_lo = [0, 220, 10, 170, 10, 220, 10, 170, 10] * 2 + [5]
    pattern_hi = [4, 230, 15, 180, 15, 230, 15, 180, 15] * 2 + [9]
    for j, (lo, hi) in enumerate(zip(pattern_lo, pattern_hi)):
        reset[:, j] = rng.integers(lo, hi, E_total)
    acts = rng.random((E_total, 2, V))
    arr = rng.poisson(1.0, (1, E_total, V))
    sl = slice(e0, e0 + count)
    d = InjectedDraws(reset_ints=reset[sl], arrivals=arr[:, sl])
    env = EnvOracle("marl", V, M, 3, E=count, params=OracleParams.marl_yaml(), draws=d)
    env.make_new_game()
    d.set_mobility_uniforms(np.full((count, 64), 0.9))
    env.renew_positions(); env.compute_parms(); env.get_next_phase(np.zeros((count, M))); env.update_channel_gains()
    part, ng = encode_groups([[0, 1], [2, 3], [4], [5], [6, 7]], V)
    _, r_glob, _ = env.step_marl(acts[sl], np.tile(part, (count, 1)), np.full(count, ng))
    sums = np.zeros(NSTAT + 1)
    for i, name in enumerate(STAT_COLUMNS):
        sums[i] = env.last[name].sum()
    sums[NSTAT] = r_glob.sum()
    return sums


def _worker(rank, world, port, E_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s, c = shard_range(E_total, rank, world)
    local = torch.from_numpy(_oracle_shard_sums(s, c, E_total))
    stats = global_episode_stats(local, c)
    q.put((rank, stats))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_stats_equal_unsharded_stats():
    E_total, world = 13, 2
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, E_total, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    whole = _oracle_shard_sums(0, E_total, E_total)
    for r in range(world):
        assert got[r]["n_envs"] == E_total
        for i, name in enumerate(STAT_COLUMNS):
            assert abs(got[r][name] - whole[i] / E_total) <= 1e-12 * max(1.0, abs(whole[i])), name
        assert abs(got[r]["reward"] - whole[NSTAT] / E_total) < 1e-12
    assert got[0] == got[1]
