#!/usr/bin/env python
"""NOMA pairing stage: `risvec_pair_noma` timed with CUDA events on real BCD gains,
next to the CPU oracle (pure-python port, one env at a time, all host cores) on a bounded sample."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ris_vec_marl_b200 import BatchedEnviron, marl_yaml_overrides, mask_schedule  # noqa: E402


def _cpu_worker(args):
    from oracle import pairing_oracle as po

    gains, p01, noise, pmax, rmin = args
    cfg = po.PairingConfig.marl_yaml(gains.shape[1])
    K, q = po.mask_schedule(10, gains.shape[1], cfg)
    t0 = time.perf_counter()
    for e in range(gains.shape[0]):
        st = po.PairingState(gains.shape[1])
        po.pair_step(st, gains[e], p01[e], cfg, noise, pmax, rmin, K, q, recalc_mask=True)
    return time.perf_counter() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--V", type=int, default=8)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--cpu-sample", type=int, default=2048)
    a = ap.parse_args()
    E, V = a.envs, a.V
    env = BatchedEnviron("marl", E, V, 40, seed=1234, **marl_yaml_overrides())
    env.set_pairing(yaml=True)
    env.make_new_game(); env.renew_positions(); env.compute_parms()
    env.optimize_phase_shift(); env.update_channel_gains()
    act = torch.rand(E, 2, V, device=env.device)
    K, q = mask_schedule(10, V, 7, 7, 0.10, 0.25, 200)
    frozen = torch.ones(E, dtype=torch.int32, device=env.device)
    res = {"config": {"envs": E, "V": V, "pairing": "config.yaml"}, "unit": "us per launch (all envs)"}

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(a.reps):
            fn()
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / a.reps * 1e3

    def solve():
        env.pair_reset()
        env.pair_noma(act, K, q, recalc_mask=True)

    res["pair_reset"] = timed(env.pair_reset)
    res["reset_plus_solve"] = timed(solve)
    res["solve_recalc"] = res["reset_plus_solve"] - res["pair_reset"]
    res["solve_no_recalc"] = timed(lambda: env.pair_noma(act, K, q, recalc_mask=False))
    res["frozen_step"] = timed(lambda: env.pair_noma(act, K, q, recalc_mask=False, reuse=frozen))
    res["rounds_mean"] = float(env.pair_rounds.float().mean())
    res["pairs_mean"] = float(env.noma_npairs.float().mean())
    res["env_solves_per_s"] = E / (res["solve_recalc"] * 1e-6)
    # algorithmic bytes of a solve: gains 8V + p01 4V in; hist 4V^2 in+out; mask V^2, partner/pairs/streak 4V each,
    # 5 per-env scalars out
    nbytes = 8 * V + 4 * V + 8 * V * V + V * V + 12 * V + 4 * V + 28
    res["algorithmic_bytes_per_env"] = nbytes
    res["achieved_GBps"] = nbytes * E / (res["solve_recalc"] * 1e-6) / 1e9
    # CPU oracle port on all cores, bounded sample
    import multiprocessing as mp

    n = min(a.cpu_sample, E)
    cores = os.cpu_count() or 1
    g = env.gains[:n].cpu().numpy()
    p = act[:n, 0].cpu().numpy().astype(np.float64)
    chunks = [(g[i::cores], p[i::cores], env.params.noise_power, env.params.P_max, env.params.R_min_bpsHz)
              for i in range(cores) if len(g[i::cores])]
    t0 = time.perf_counter()
    with mp.get_context("fork").Pool(len(chunks)) as pool:
        pool.map(_cpu_worker, chunks)
    wall = time.perf_counter() - t0
    res["cpu_baseline"] = {"value": n / wall, "unit": "env-solves/s", "cores": len(chunks), "kind": "port",
                           "sample": f"{n} envs, one solve each"}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
