"""Multi-GPU plumbing: env instances are sharded over ranks (one process per GPU, torchrun),
nothing on the data path crosses GPUs; the only collective is an all-reduce (sum) of the
episode-statistics vector, NCCL on GPUs / gloo in the CPU tests."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from ._lib import NSTAT, STAT_COLUMNS


def rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_envs_total: int, rank: int, world: int):
    """Contiguous block partition of the global env index range; the first `n % world` ranks
    hold one extra env.  Returns (first_global_index, count)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_envs_total), int(world))
    count = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, count


def init_process_group(backend: str | None = None, device=None):
    """Join the job torchrun described (no-op for a single process)."""
    rank, world, local = rank_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {"device_id": device} if (backend == "nccl" and device is not None) else {}
        dist.init_process_group(backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def all_reduce_sum(vec: torch.Tensor) -> torch.Tensor:
    """In-place sum over ranks of a small statistics vector (identity for one process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def global_episode_stats(local_sums: torch.Tensor, n_envs_local: int):
    """`local_sums` = f64 [NSTAT + 1] from `BatchedEnviron.shard_stats()` (column sums over the
    shard's envs; last entry = sum of global rewards).  Returns {name: mean over ALL envs of
    the job}, i.e. what the reference driver logs per step for its single env
    (marl_train_bcd.py:1627-1662), averaged over the batch."""
    v = torch.cat([local_sums.double().reshape(-1), torch.tensor([float(n_envs_local)], dtype=torch.float64,
                                                                 device=local_sums.device)])
    all_reduce_sum(v)
    n = v[-1].item()
    out = {name: v[i].item() / n for i, name in enumerate(STAT_COLUMNS)}
    out["reward"] = v[NSTAT].item() / n
    out["n_envs"] = int(n)
    return out
