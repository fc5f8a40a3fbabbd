#!/usr/bin/env python
"""torchrun check (N >= 2 GPUs of one node): the peer-memory statistics accumulator (dist.SharedStats: float64
atomics into rank 0's HBM through a CUDA IPC mapping) holds exactly the sums an NCCL all-reduce of the per-rank
vectors gives (the sums are float64 of float32 inputs: order-independent to ~1e-16 relative).

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_shared_stats.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from ris_vec_marl_b200 import BatchedEnviron  # noqa: E402
from ris_vec_marl_b200.dist import SharedStats  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    E, V, M = 1024, 8, 40
    env = BatchedEnviron("sarl", E, V, M, device=local, seed=7, env_index_base=rank * E)
    env.make_new_game(); env.renew_positions(); env.compute_parms()
    shared = SharedStats(local, rank, world)
    shared.zero_()
    local_sum = torch.zeros_like(shared.tensor)
    g = torch.Generator(device=dev).manual_seed(rank)
    for _ in range(6):
        act = torch.rand(E, 2, V, device=dev, generator=g)
        ph = torch.rand(E, M, device=dev, generator=g) * 6.2831853
        env.step_sarl(act, ph)
        env.shard_stats(out=shared.tensor, accumulate=True)
        env.shard_stats(out=local_sum, accumulate=True)
    shared.add_(local_sum)           # the per-interval push: the accumulator now holds every sum twice
    total = shared.read() * 0.5
    dist.all_reduce(local_sum)
    ref = local_sum.cpu()
    err = float(((total - ref).abs() / ref.abs().clamp_min(1e-30)).max())
    if rank == 0:
        print(f"shared-stats check: world {world}, max relative difference {err:.3e}, reward sum {float(total[-1]):.6f}")
    shared.close()
    dist.destroy_process_group()
    assert err < 1e-12, err


if __name__ == "__main__":
    main()
