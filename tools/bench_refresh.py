#!/usr/bin/env python
"""Per-episode refresh kernels: renew_positions + compute_parms +
BCD phase optimiser + cascaded gains, CUDA-event timed per launch, for a given shape."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ris_vec_marl_b200 import BatchedEnviron, marl_yaml_overrides  # noqa: E402


def timed(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3  # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=1024)
    ap.add_argument("--V", type=int, default=32)
    ap.add_argument("--M", type=int, default=256)
    a = ap.parse_args()
    env = BatchedEnviron("marl", a.envs, a.V, a.M, **marl_yaml_overrides())
    env.make_new_game()
    res = {"config": {"envs": a.envs, "V": a.V, "M": a.M}, "unit": "us per launch (all envs)"}
    res["renew_positions"] = timed(env.renew_positions)
    res["compute_parms"] = timed(env.compute_parms)
    res["optimize_phase_shift_bcd"] = timed(env.optimize_phase_shift)
    res["update_channel_gains"] = timed(env.update_channel_gains)
    # flops of the float64 kernels: BCD = V*M sincospi + M*ncand complex MACs; gains = V*M (sincospi + complex MAC)
    res["env_refreshes_per_s"] = a.envs / (sum(res[k] for k in ("renew_positions", "compute_parms",
                                                                  "optimize_phase_shift_bcd",
                                                                  "update_channel_gains")) * 1e-6)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
