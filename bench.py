#!/usr/bin/env python
"""Benchmark of the batched Environ.step hot path (contract: task prompt, section 4).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload sarl|marl]

One bench "step" = one fused T-step rollout launch over E envs per GPU (T*E env-steps) with
pre-staged synthetic actions; the metric is env-steps/s (BASELINE.json).  For N > 1 launch with
torchrun (one rank per GPU); envs are sharded, the only collective is the NCCL all-reduce of the
episode-statistics vector after every rollout.
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec"
UNIT = "env-steps/s"


def algorithmic_bytes(workload, V, M, T, E):
    """Bytes one rollout launch must move ('Algorithmic bytes').
    MARL: per step 28V + 4, per env per rollout 2(4V + 8) + 4V + 4V + 4.
    SARL: per step 36V + 4M + 4 (actions 8V, phases 4M, arrivals 4V, six traces 24V, reward 4),
          per env per rollout 2 * 4V (DataBuf in/out) + 8V (angle f64) + 4V (amplitude)."""
    if workload == "marl":
        return E * (T * (28 * V + 4) + 2 * (4 * V + 8) + 4 * V + 4 * V + 4)
    return E * (T * (36 * V + 4 * M + 4) + 2 * 4 * V + 8 * V + 4 * V)


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def recorded_traffic(workload):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's per-env step loop, one env object per process
# ------------------------------------------------------------------------------------------
def _cpu_worker(args):
    workload, V, M, seconds, seed = args
    import numpy as np

    from oracle.env_oracle import EnvOracle, GlobalNumpyDraws, OracleParams, encode_groups

    np.random.seed(seed)
    p = OracleParams.marl_yaml() if workload == "marl" else OracleParams()
    env = EnvOracle(workload, V, M, 3, E=1, params=p, draws=GlobalNumpyDraws())
    env.make_new_game()
    rng = np.random.default_rng(seed)
    groups = [[i, i + 1] for i in range(0, 3 * V // 4, 2)] + [[i] for i in range(3 * V // 4, V)]
    partner, ng = encode_groups(groups, V)
    steps, t0 = 0, time.perf_counter()
    episode = 0
    while time.perf_counter() - t0 < seconds:
        # loop shape of the reference drivers (marl_train_bcd.py:1268-1309, ddpg_train.py:120-162)
        if workload == "marl":
            if episode % 5 == 0:
                env.renew_positions(); env.compute_parms()
            env.optimize_phase_shift(); env.update_channel_gains()
        elif episode % 100 == 0:
            env.renew_positions(); env.compute_parms()
        for _ in range(100):
            a = rng.random((1, 2, V))
            if workload == "marl":
                env.step_marl(a, partner[None], np.array([ng]))
            else:
                env.step_sarl(a, rng.random((1, M)) * 2 * np.pi)
            steps += 1
        episode += 1
    return steps, time.perf_counter() - t0


def cpu_port_throughput(workload, V, M, seconds, cores):
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(workload, V, M, seconds, 100 + i) for i in range(cores)])
    total = sum(r[0] for r in res)
    wall = max(r[1] for r in res)
    return total / wall, total


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args, rank):
    if rank != 0:
        return
    cores = host_cores()
    V, M = args.V, args.M
    # every "step" of this arm is a bounded sample of the same workload: `sample_s` seconds of
    # the per-env python loop on every host core
    per_step_s = max(1.0, min(10.0, 40.0 / max(1, args.steps + args.warmup)))
    # at most 40 samples are actually run (~1 min), whatever K the caller asks for: the per-sample
    # throughput of the CPU loop does not depend on how many samples are averaged
    warm = min(args.warmup, 3)
    n_run = warm + min(args.steps, 40 - warm)
    vals = []
    for i in range(n_run):
        v, n = cpu_port_throughput(args.workload, V, M, per_step_s, cores)
        if i >= warm:
            vals.append(v)
    value = sum(vals) / len(vals)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step_s * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"oracle/env_oracle.py per-env step loop (E=1 per process, reference loop shape), "
                                   f"{cores} processes x {per_step_s:.1f} s per bench step, {len(vals)} timed samples"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    name = {"sarl": "Simulation-SARL Environment (DDPG single-agent variant)",
            "marl": "Simulation-MARL-BCD Environment (config.yaml parameters)"}[args.workload]
    return {"workload": f"{name}, V={args.V} vehicles, M={args.M} RIS elements, {args.envs} batched envs per GPU, "
                        f"fused T={args.T}-step rollout per launch with pre-staged actions",
            "envs_per_gpu": args.envs, "V": args.V, "M": args.M, "T": args.T,
            "l2_policy": "inputs+outputs of one launch exceed the 126 MB L2; nothing is re-read between launches"}


# ------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sarl", choices=["sarl", "marl"])
    ap.add_argument("--envs", type=int, default=4096, help="env instances per GPU")
    ap.add_argument("--T", type=int, default=256, help="env steps fused per launch")
    ap.add_argument("--V", type=int, default=8)
    ap.add_argument("--M", type=int, default=40)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch
    import torch.distributed as dist

    from ris_vec_marl_b200 import BatchedEnviron, load_library, marl_yaml_overrides

    load_library()  # no CUDA extension => hard failure, never a fallback
    assert torch.cuda.is_available(), "bench.py needs a GPU for --impl ours"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    E, V, M, T = args.envs, args.V, args.M, args.T
    wl = args.workload
    over = marl_yaml_overrides() if wl == "marl" else {}
    env = BatchedEnviron(wl, E, V, M, 3, device=local, seed=1234, env_index_base=rank * E, **over)
    env.make_new_game()
    env.renew_positions()
    env.compute_parms()
    lam = float(env.get_param("rate"))
    gen = torch.Generator(device=dev).manual_seed(rank)
    actions = torch.rand(T, E, 2, V, device=dev, generator=gen)
    arrivals = torch.poisson(torch.full((T, E, V), lam, device=dev), generator=gen).to(torch.int32)
    if wl == "marl":
        env.optimize_phase_shift()
        env.update_channel_gains()
        actions[:, :, 1, :].clamp_(min=float(env.get_param("cpu_share_floor")))
        # 3V/8 pairs + V/4 singletons (SURVEY.md 8d synthetic inputs)
        import numpy as np

        from ris_vec_marl_b200 import encode_groups

        groups = [[i, i + 1] for i in range(0, 3 * V // 4, 2)] + [[i] for i in range(3 * V // 4, V)]
        part, ng = encode_groups(groups, V)
        partner = torch.as_tensor(np.tile(part, (E, 1))).to(dev)
        ngroups = torch.full((E,), ng, dtype=torch.int32, device=dev)
        trace_names = ("reward_user", "reward", "data_t", "data_p", "rate", "DataBuf")
        phases = None
    else:
        phases = torch.rand(T, E, M, device=dev, generator=gen) * 6.283185307179586
        trace_names = ("reward", "DataBuf", "data_t", "data_p", "over_power", "over_data", "rate")
    # SARL streams through the library's packed (tiled) records; MARL through the per-array entry
    # points, which measure faster for its 3-in / 6-out streams
    packed = wl == "sarl" and V == 8 and M in (16, 40) and E % 4 == 0
    if os.environ.get("RISVEC_BENCH_MARL_PACKED") == "1" and wl == "marl" and V == 8 and E % 4 == 0:
        packed = True   # A/B switch (profiles/r1_summary.md)
    stats_sum = torch.zeros(17, dtype=torch.float64, device=dev)
    if packed:
        in_rec = env.pack_inputs(actions, arrivals, phases)
        out_rec = torch.empty(T, E // 4, 4 * env.packed_out_words(), dtype=torch.float32, device=dev)
        reward = torch.empty(T, E, dtype=torch.float32, device=dev)
        del actions, arrivals, phases

        pk_groups = (partner, ngroups) if wl == "marl" else (None, None)

        def one_step():
            env.rollout_packed(in_rec, *pk_groups, out_rec=out_rec, reward=reward)
    else:
        out = env._alloc_traces(trace_names, T, trace_names)

        def one_step():
            if wl == "marl":
                env.rollout_marl(actions, partner, ngroups, arrivals, out=out)
            else:
                env.rollout_sarl(actions, phases, arrivals, out=out)

    pending = []
    ring = torch.zeros(max(args.steps, args.warmup, 20) + 1, 17, dtype=torch.float64, device=dev) if world > 1 else None

    def episode_stats():
        # the only collective on the path: NCCL sum of a 17-entry f64 vector.  It is
        # issued asynchronously (NCCL's own stream) into a slot of a preallocated ring so the next
        # rollout overlaps it; `drain_stats` waits for the last one (NCCL runs them in order) inside
        # the timed region and folds the ring with one reduction.
        if world > 1:
            sv = ring[len(pending)]
            env.shard_stats(out=sv)
            pending.append(dist.all_reduce(sv, async_op=True))
        else:
            env.shard_stats(out=stats_sum, accumulate=True)

    def drain_stats():
        if pending:
            pending[-1].wait()
            stats_sum.add_(ring[:len(pending)].sum(0))
        pending.clear()

    for _ in range(args.warmup):
        one_step(); episode_stats()
    drain_stats()
    torch.cuda.synchronize()

    # ---- timed region: K steps, device-timed, max over ranks
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_spin = time.time()
        while True:  # keep the GPU under load until nvidia-smi delivers its first sample (<= 3 s)
            for _ in range(max(args.warmup, 20)):
                one_step()
            torch.cuda.synchronize()
            if sampler.lines or sampler.proc is None or time.time() - t_spin > 3.0:
                break
        sampler.lines.clear()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = env.launch_count
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev[0].record()
    for i in range(args.steps):
        kev[i][0].record()
        one_step()
        kev[i][1].record()
        episode_stats()
    drain_stats()
    ev[1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = env.launch_count - l0
    ms_total = ev[0].elapsed_time(ev[1])
    kern_ms = sorted(a.elapsed_time(b) for a, b in kev)
    kern_ms_avg = sum(kern_ms) / len(kern_ms)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = world * E * T * args.steps / (ms_total * 1e-3)

    # ---- end to end through the public host-buffer API (pinned host in, pinned host out)
    if packed:
        h_in = in_rec.cpu().pin_memory()
        h_out = torch.empty(out_rec.shape, dtype=torch.float32).pin_memory()
        h_rew = torch.empty(reward.shape, dtype=torch.float32).pin_memory()
        if wl == "marl":
            h_pt, h_ng = partner.cpu().pin_memory(), ngroups.cpu().pin_memory()
            run_host = lambda: env.rollout_packed_host(h_in, h_out, h_rew, h_pt, h_ng)
        else:
            run_host = lambda: env.rollout_packed_host(h_in, h_out, h_rew)
        h2d, d2h = h_in.numel() * 4, h_out.numel() * 4 + h_rew.numel() * 4
    else:
        h_act, h_arr = actions.cpu().pin_memory(), arrivals.cpu().pin_memory()
        h_o = {k: torch.empty(v.shape, dtype=torch.float32).pin_memory() for k, v in out.items()}
        h_rew = h_o["reward"]
        if wl == "marl":
            h_part, h_ng = partner.cpu().pin_memory(), ngroups.cpu().pin_memory()
            run_host = lambda: env.rollout_marl_host(h_act, h_part, h_ng, h_arr, h_o)
            h2d = (h_act.numel() + h_arr.numel() + h_part.numel() + h_ng.numel()) * 4
        else:
            h_ph = phases.cpu().pin_memory()
            run_host = lambda: env.rollout_sarl_host(h_act, h_ph, h_arr, h_o)
            h2d = (h_act.numel() + h_arr.numel() + h_ph.numel()) * 4
        d2h = sum(v.numel() * 4 for v in h_o.values())
    run_host(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.e2e_steps):
        run_host()
        torch.cuda.current_stream().synchronize()  # the caller reads the step's result on the host
        _ = float(h_rew[-1, 0])
    e1.record()
    torch.cuda.synchronize()
    te = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * E * T * args.e2e_steps / (float(te.item()) * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        if wl == "sarl":
            kernel_name = ("k_sarl_v8<%d,true,true,true> (packed records)" % (M // 8)) if packed else (
                "k_sarl_v8" if (V <= 8 and M <= 40) else
                ("k_sarl_rollout" if M <= 40 else "k_sarl_cascade2 + k_sarl_scan (timed together)"))
        else:
            kernel_name = "k_marl_v8<true,false>" if V <= 8 else "k_marl_rollout"
        alg = algorithmic_bytes(wl, V, M, T, E)
        achieved = alg / (kern_ms_avg * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args),
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": args.e2e_steps},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(wl) if (E, V, M, T) == (4096, 8, 40, 256) else None,
                         "kernel": kernel_name, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg, "kernel_ms_avg": kern_ms_avg,
                         "kernel_ms_min": kern_ms[0], "env_steps_per_launch": E * T},
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            v, n = cpu_port_throughput(wl, V, M, args.cpu_seconds, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle/env_oracle.py per-env step loop (E=1 per process, reference "
                                              f"loop shape), {cores} processes x {args.cpu_seconds:.0f} s = {n} env-steps"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
