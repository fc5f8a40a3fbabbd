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


class SharedStats:
    """The episode-statistics reduction of a one-node job without a collective: rank 0 owns a float64
    [NSTAT + 1] accumulator in its HBM, every other rank maps it (CUDA IPC, NVLink peer access) and all ranks either
    call `env.shard_stats(out=shared.tensor, accumulate=True)` (k_shard_stats adds the shard's sums with float64
    atomics straight into the owner's memory) or sum locally and `add_()` once per statistics interval.  No rank ever waits for another one; `read()` (barrier + copy)
    returns the totals.  Raises if the peer mapping is not available (the caller then all-reduces instead)."""

    def __init__(self, local_device: int, rank: int, world: int):
        import ctypes as C

        from . import _lib
        from .batched import _DevView

        self._lib, self.rank, self.world, self.device = _lib.load_library(), rank, world, int(local_device)
        n_bytes = 8 * (NSTAT + 1)
        ptr = C.c_void_p()
        payload = [None]
        if rank == 0:
            handle = C.create_string_buffer(64)
            try:
                _lib.check(self._lib.risvec_shared_buffer_create(self.device, n_bytes, C.byref(ptr), handle))
                payload = [bytes(handle.raw)]
            except RuntimeError as exc:  # the other ranks are waiting in the broadcast: tell them
                payload = [repr(exc)]
        if world > 1:
            dist.broadcast_object_list(payload, src=0)
        if not isinstance(payload[0], bytes):
            raise RuntimeError(f"rank 0 could not create the shared buffer: {payload[0]}")
        if rank != 0:
            _lib.check(self._lib.risvec_shared_buffer_open(self.device, payload[0], C.byref(ptr)))
        self._ptr, self.owner = ptr, rank == 0
        self.tensor = torch.as_tensor(_DevView(ptr.value, (NSTAT + 1,), "<f8"), device=torch.device("cuda", self.device))

    def add_(self, local_sums: torch.Tensor):
        """accumulator += local_sums (one launch of float64 atomics on torch's current stream): the per-interval
        push of a rank that summed its per-rollout statistics in its own HBM."""
        from . import _lib

        assert local_sums.dtype == torch.float64 and local_sums.is_cuda and local_sums.numel() == NSTAT + 1
        _lib.check(self._lib.risvec_shared_buffer_add(self.device, self.tensor.data_ptr(), local_sums.data_ptr(),
                                                      NSTAT + 1, torch.cuda.current_stream(self.device).cuda_stream))

    def zero_(self):
        """Owner clears the accumulator; everybody leaves together (call outside timed regions)."""
        if self.world > 1:
            dist.barrier()
        if self.owner:
            self.tensor.zero_()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()

    def read(self) -> torch.Tensor:
        """Totals over all ranks (host tensor), valid once every rank's kernels have completed."""
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier()
        out = self.tensor.cpu().clone()
        if self.world > 1:
            dist.barrier()
        return out

    def close(self):
        if self._ptr is not None:
            self.tensor = None
            self._lib.risvec_shared_buffer_close(self.device, self._ptr, int(self.owner))
            self._ptr = None
