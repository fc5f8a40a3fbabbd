#!/usr/bin/env python
"""A/B timing of the SARL rollout kernels on one GPU (reference layout [T,E,2,V] / [T,E,M] / [T,E,V]):
RISVEC_SARL_PATH = mma (tensor-core cascade) | v8 (FP32-pipe cascade) | generic, plus the packed records.
Prints one JSON line per path: kernel, ms per launch (CUDA events, median of --reps), env-steps/s and
the fraction of the measured HBM peak the algorithmic bytes correspond to."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import algorithmic_bytes, measured_peak_gbs  # noqa: E402


def make_env(path, E, V, M):
    from ris_vec_marl_b200 import BatchedEnviron

    os.environ["RISVEC_SARL_PATH"] = "v8" if path == "packed" else path
    env = BatchedEnviron("sarl", E, V, M, seed=1234)
    env.make_new_game(); env.renew_positions(); env.compute_parms()
    return env


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--T", type=int, default=256)
    ap.add_argument("--V", type=int, default=8)
    ap.add_argument("--M", type=int, default=40)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--paths", default="mma,mma-ldg,v8,packed")
    args = ap.parse_args()
    E, V, M, T = args.envs, args.V, args.M, args.T
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(0)
    actions = torch.rand(T, E, 2, V, device=dev, generator=gen)
    phases = torch.rand(T, E, M, device=dev, generator=gen) * 6.283185307179586
    arrivals = torch.poisson(torch.full((T, E, V), 3.0, device=dev), generator=gen).to(torch.int32)
    peak, _ = measured_peak_gbs()
    ref = None
    for path in args.paths.split(","):
        env = make_env(path, E, V, M)
        names = ("reward", "DataBuf", "data_t", "data_p", "over_power", "over_data", "rate")
        if path == "packed":
            rec = env.pack_inputs(actions, arrivals, phases)
            out_rec = torch.empty(T, E // 4, 4 * env.packed_out_words(), dtype=torch.float32, device=dev)
            rew = torch.empty(T, E, dtype=torch.float32, device=dev)
            run = lambda: env.rollout_packed(rec, out_rec=out_rec, reward=rew)
        else:
            out = env._alloc_traces(names, T, names)
            run = lambda: env.rollout_sarl(actions, phases, arrivals, out=out)
        for _ in range(5):
            run()
        torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.reps)]
        for a, b in ev:
            a.record(); run(); b.record()
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        med = ms[len(ms) // 2]
        alg = algorithmic_bytes("sarl", V, M, T, E)
        line = {"path": path, "kernel": env.last_kernel(), "ms_median": med, "ms_min": ms[0],
                "env_steps_per_s": E * T / (med * 1e-3), "hbm_frac": alg / (med * 1e-3) / 1e9 / peak}
        if path != "packed":
            cur = {k: v.clone() for k, v in out.items()}
            if ref is None:
                ref = cur
            else:
                line["max_abs_diff_vs_first"] = {k: float((cur[k] - ref[k]).abs().max()) for k in cur}
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
