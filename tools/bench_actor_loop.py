#!/usr/bin/env python
"""Actor-in-the-loop rollout (BASELINE config 5 shape, one GPU's share): per step
observe -> V per-agent policy MLPs in torch (5 -> 512 -> LN -> 256 -> LN -> 2, tanh; the power head
of Simulation-MARL-BCD/sac_agent.py:23-34) -> map_actions -> Environ.step (T = 1 launch).
Random-init weights, no learner update.  Reports env-steps/s eager and replayed as a CUDA graph
(the launch-bound regime the fused T-step rollout of bench.py avoids).

`--driver` adds the rest of the driver's per-step device work (marl_train_bcd.py:1304-1799): the NOMA
pairing stage (mask + solve on the first step of each 100-step episode, frozen groups afterwards), an
intent head (softmax over the V partners, masked by the feasibility mask), the second observation and
the replay-memory write of the assembled transition."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from ris_vec_marl_b200 import (BatchedEnviron, ReplayBuffer, encode_groups, marl_yaml_overrides,  # noqa: E402
                               mask_schedule)


class Actors(torch.nn.Module):
    """V independent policy networks evaluated as batched matmuls ([V, E, .] x [V, ., .])."""

    def __init__(self, V, dev):
        super().__init__()
        g = torch.Generator(device="cpu").manual_seed(0)
        mk = lambda i, o: torch.nn.Parameter((torch.randn(V, i, o, generator=g) / i ** 0.5).to(dev))
        self.w1, self.w2, self.w3 = mk(5, 512), mk(512, 256), mk(256, 2)
        self.b1 = torch.nn.Parameter(torch.zeros(V, 1, 512, device=dev))
        self.b2 = torch.nn.Parameter(torch.zeros(V, 1, 256, device=dev))
        self.b3 = torch.nn.Parameter(torch.zeros(V, 1, 2, device=dev))
        self.w4 = mk(256, V)      # intent head (sac_agent.py: logits over partners)

    @torch.no_grad()
    def forward(self, obs):  # obs [E, V, 5] -> raw actions [E, V, 2] in (-1, 1)
        x = obs.transpose(0, 1)
        x = torch.relu(torch.nn.functional.layer_norm(torch.baddbmm(self.b1, x, self.w1), (512,)))
        x = torch.relu(torch.nn.functional.layer_norm(torch.baddbmm(self.b2, x, self.w2), (256,)))
        self.hidden = x
        return torch.tanh(torch.baddbmm(self.b3, x, self.w3)).transpose(0, 1).contiguous()

    @torch.no_grad()
    def intent(self, mask):  # mask [E, V, V] u8 -> probs [E, V, V]
        logits = torch.bmm(self.hidden, self.w4).transpose(0, 1).float()
        return torch.softmax(logits.masked_fill(mask == 0, -1e9), dim=-1)


def driver_loop(envs, steps, driver, actor_dtype, rank, world, local, ranks=None):
    """One GPU's share of the loop (under torchrun: every rank its own shard, replicated actor weights);
    returns the result dict (whole-job env-steps/s, device-timed, max over ranks)."""
    torch.backends.cuda.matmul.allow_tf32 = actor_dtype == "tf32"
    dev = torch.device("cuda", local)
    E, V, M = envs, 8, 40
    env = BatchedEnviron("marl", E, V, M, device=local, seed=1234, env_index_base=rank * E, **marl_yaml_overrides())
    env.make_new_game(); env.renew_positions(); env.compute_parms(); env.optimize_phase_shift(); env.update_channel_gains()
    part, ng = encode_groups([[0, 1], [2, 3], [4, 5], [6], [7]], V)
    partner = torch.as_tensor(np.tile(part, (E, 1))).to(dev)
    ngroups = torch.full((E,), ng, dtype=torch.int32, device=dev)
    actors = Actors(V, dev)
    if actor_dtype == "bf16":
        actors = actors.to(torch.bfloat16)
    obs = torch.empty(E, V, 5, device=dev)
    obs2 = torch.empty(E, V, 5, device=dev)
    if driver:
        env.set_pairing(yaml=True)
        env.pair_reset()
        rb = ReplayBuffer(min(1_000_000, max(4 * E, 65536)), 5, V + 2, V, device=local)
        K, q = mask_schedule(10, V, 7, 7, 0.10, 0.25, 200)
        frozen = torch.ones(E, dtype=torch.int32, device=dev)
        ctr = [0]

    bufs = [obs, obs2]  # the fused step writes the next observation: the two buffers alternate
    env.observe(out=bufs[0])

    def driver_step():
        first = ctr[0] % 100 == 0
        ctr[0] += 1
        cur, nxt = bufs[0], bufs[1]
        x = cur.to(torch.bfloat16) if actor_dtype == "bf16" else cur
        raw = actors(x).float()
        act = env.map_actions(raw)      # the pairing stage reads the mapped offload power
        part_v, ng_v = env.pair_noma(act, K, q, recalc_mask=first, reuse=None if first else frozen, new_episode=first)
        probs = actors.intent(env.pair_mask)
        env.step_marl_fused(raw, part_v, ng_v, obs_out=nxt)   # mapping + Environ.step + marl_get_state: one launch
        rb.store_marl(cur, probs, raw, env.reward, env.reward_user, nxt, done=(ctr[0] % 100 == 0),
                      mask_u8=env.pair_mask)
        bufs[0], bufs[1] = nxt, cur

    def step():
        if driver:
            return driver_step()
        cur, nxt = bufs[0], bufs[1]
        x = cur.to(torch.bfloat16) if actor_dtype == "bf16" else cur
        env.step_marl_fused(actors(x).float(), partner, ngroups, obs_out=nxt)  # arrivals: on-device Philox
        bufs[0], bufs[1] = nxt, cur

    for _ in range(5):
        step()
    torch.cuda.synchronize()
    if ranks is not None:
        ranks.barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        step()
    t1.record()
    torch.cuda.synchronize()
    sec = t0.elapsed_time(t1) * 1e-3
    if ranks is not None:
        sec = ranks.max(sec)
    res = {"envs_per_gpu": E, "V": V, "M": M, "steps": steps, "actor": "8 x MLP 5-512-256-{2, V} (bmm), " + actor_dtype,
           "loop": ("actors + intent head (torch), map_actions (for the pairing), NOMA pairing (solve on step 0 of each "
                    "100-step episode, frozen after), fused [action mapping + Environ.step + observation] (" +
                    env.last_kernel() + "), replay write")
                   if driver else "actors (torch), fused [action mapping + Environ.step + observation] (" + env.last_kernel() + ")",
           "value": world * E * steps / sec, "unit": "env-steps/s", "us_per_step": sec / steps * 1e6}
    if driver:
        res["replay_rows"] = int(rb.mem_cntr)
        res["pairs_mean"] = float(env.noma_npairs.float().mean())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--actor-dtype", default="fp32", choices=["fp32", "tf32", "bf16"])
    ap.add_argument("--driver", action="store_true", help="pairing + intent head + replay write in the loop")
    ap.add_argument("--replay-size", type=int, default=1_000_000)
    a = ap.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = a.actor_dtype == "tf32"
    # config 5 of BASELINE: under torchrun every rank steps its own shard of the envs (8192 per GPU x 8 =
    # 65 536) with replicated actor weights; the only collective is the statistics all-reduce at the end
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    E, V, M = a.envs, 8, 40
    env = BatchedEnviron("marl", E, V, M, device=local, seed=1234, env_index_base=rank * E, **marl_yaml_overrides())
    env.make_new_game(); env.renew_positions(); env.compute_parms(); env.optimize_phase_shift(); env.update_channel_gains()
    part, ng = encode_groups([[0, 1], [2, 3], [4, 5], [6], [7]], V)
    partner = torch.as_tensor(np.tile(part, (E, 1))).to(dev)
    ngroups = torch.full((E,), ng, dtype=torch.int32, device=dev)
    actors = Actors(V, dev)
    if a.actor_dtype == "bf16":
        actors = actors.to(torch.bfloat16)
    obs = torch.empty(E, V, 5, device=dev)

    obs2 = torch.empty(E, V, 5, device=dev)
    if a.driver:
        env.set_pairing(yaml=True)
        env.pair_reset()
        rb = ReplayBuffer(a.replay_size, 5, V + 2, V, device=local)
        K, q = mask_schedule(10, V, 7, 7, 0.10, 0.25, 200)
        frozen = torch.ones(E, dtype=torch.int32, device=dev)
        ctr = [0]

    bufs = [obs, obs2]  # the fused step writes the next observation: the two buffers alternate
    env.observe(out=bufs[0])

    def driver_step():
        first = ctr[0] % 100 == 0
        ctr[0] += 1
        cur, nxt = bufs[0], bufs[1]
        x = cur.to(torch.bfloat16) if a.actor_dtype == "bf16" else cur
        raw = actors(x).float()
        act = env.map_actions(raw)      # the pairing stage reads the mapped offload power
        part_v, ng_v = env.pair_noma(act, K, q, recalc_mask=first, reuse=None if first else frozen, new_episode=first)
        probs = actors.intent(env.pair_mask)
        env.step_marl_fused(raw, part_v, ng_v, obs_out=nxt)   # mapping + Environ.step + marl_get_state: one launch
        rb.store_marl(cur, probs, raw, env.reward, env.reward_user, nxt, done=(ctr[0] % 100 == 0),
                      mask_u8=env.pair_mask)
        bufs[0], bufs[1] = nxt, cur

    def step():
        if a.driver:
            return driver_step()
        cur, nxt = bufs[0], bufs[1]
        x = cur.to(torch.bfloat16) if a.actor_dtype == "bf16" else cur
        env.step_marl_fused(actors(x).float(), partner, ngroups, obs_out=nxt)  # arrivals: on-device Philox
        bufs[0], bufs[1] = nxt, cur

    def timed(fn, n):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(n):
            fn()
        t1.record()
        torch.cuda.synchronize()
        sec = torch.tensor([t0.elapsed_time(t1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)     # device-timed, max over ranks
        return float(sec.item())

    for _ in range(5):
        step()
    res = {"config": {"envs": E, "V": V, "M": M, "steps": a.steps, "actor": "8 x MLP 5-512-256-2 (bmm), " + a.actor_dtype,
                      "driver_loop": bool(a.driver), "envs_per_gpu": E}}
    sec = timed(step, a.steps)
    res["eager_env_steps_per_s"] = world * E * a.steps / sec
    res["eager_us_per_step"] = sec / a.steps * 1e6
    try:
        if a.driver:
            raise RuntimeError("graph replay skipped: the driver loop branches on the step counter")
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            step(); step()
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=s):
                step(); step()   # the two observation buffers alternate: one replay = two steps
        sec = timed(g.replay, a.steps // 2)
        res["graph_env_steps_per_s"] = world * E * (a.steps // 2) * 2 / sec
        res["graph_us_per_step"] = sec / ((a.steps // 2) * 2) * 1e6
    except Exception as exc:  # capture is best-effort
        res["graph_error"] = repr(exc)[:200]
    if a.driver:
        res["replay_rows"] = rb.mem_cntr
        res["pairs_mean"] = float(env.noma_npairs.float().mean())
    stats = env.shard_stats()
    if world > 1:
        dist.all_reduce(stats)
    res["n_gpus"] = world
    res["mean_reward"] = float(stats[-1].item()) / (world * E)
    if rank == 0:
        print(json.dumps(res))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
