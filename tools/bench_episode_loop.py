#!/usr/bin/env python
"""MARL episode loop in the shape of the reference drivers (BASELINE config 1 on the GPU):
per episode [renew_positions + compute_parms every 5th episode] + optimize_phase_shift +
update_channel_gains + NOMA pairing (mask + solve) + one fused rollout of 100 steps with pre-staged
actions.  Reports env-steps/s for the steps alone and amortised over the per-episode work, per kernel
time shares (CUDA events), and the same loop replayed as one CUDA graph."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from ris_vec_marl_b200 import BatchedEnviron, marl_yaml_overrides, mask_schedule  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--episodes", type=int, default=50)
    ap.add_argument("--T", type=int, default=100)
    a = ap.parse_args()
    E, V, M, T = a.envs, 8, 40, a.T
    dev = torch.device("cuda", 0)
    env = BatchedEnviron("marl", E, V, M, seed=1234, **marl_yaml_overrides())
    env.set_pairing(yaml=True)
    env.make_new_game()
    gen = torch.Generator(device=dev).manual_seed(0)
    actions = torch.rand(T, E, 2, V, device=dev, generator=gen)
    actions[:, :, 1].clamp_(min=0.10)
    arrivals = torch.poisson(torch.full((T, E, V), 1.0, device=dev), generator=gen).to(torch.int32)
    names = ("reward_user", "reward", "DataBuf", "data_t", "data_p", "rate")
    out = env._alloc_traces(names, T, names)
    K, q = mask_schedule(10, V, 7, 7, 0.10, 0.25, 200)

    def refresh_positions():
        env.renew_positions()
        env.compute_parms()

    def refresh_channel():
        env.optimize_phase_shift()
        env.update_channel_gains()

    def pairing():
        env.pair_noma(actions[0], K, q, recalc_mask=True, new_episode=True)

    def steps():
        env.rollout_marl(actions, env.noma_partner, env.noma_ngroups, arrivals, out=out)

    def episode(i):
        if i % 5 == 0:
            refresh_positions()
        refresh_channel()
        pairing()
        steps()

    def timed(fn, n):
        for i in range(3):
            fn(i)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for i in range(n):
            fn(i)
        t1.record()
        torch.cuda.synchronize()
        return t0.elapsed_time(t1) / n * 1e3  # us

    res = {"config": {"envs": E, "V": V, "M": M, "steps_per_episode": T, "positions_every": 5},
           "unit": "us per episode (all envs)"}
    res["positions"] = timed(lambda i: refresh_positions(), 20)
    res["bcd_plus_gains"] = timed(lambda i: refresh_channel(), 20)
    res["pairing"] = timed(lambda i: pairing(), 20)
    res["steps"] = timed(lambda i: steps(), 20)
    res["episode"] = timed(episode, a.episodes)
    res["env_steps_per_s_steps_only"] = E * T / (res["steps"] * 1e-6)
    res["env_steps_per_s_amortised"] = E * T / (res["episode"] * 1e-6)
    try:  # five episodes (one position refresh) as one CUDA graph
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for i in range(5):
                episode(i)
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                for i in range(5):
                    episode(i)
        us = timed(lambda i: g.replay(), max(4, a.episodes // 5)) / 5
        res["episode_graph"] = us
        res["env_steps_per_s_amortised_graph"] = E * T / (us * 1e-6)
    except Exception as exc:
        res["graph_error"] = repr(exc)[:200]
    res["mean_reward"] = float(env.reward.mean())
    print(json.dumps(res))


if __name__ == "__main__":
    main()
