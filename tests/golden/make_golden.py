"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Runs only where the reference tree is mounted (the build container):

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Each fixture is one "scenario": E independent reference `Environ` objects (one per seed)
driven through the call sequence of the reference drivers
(Simulation-MARL-BCD/marl_train_bcd.py:543-545,1268-1309,1611 and
Simulation-SARL/ddpg_train.py:40-42,120-123,162) with injected float32-representable
actions, while `oracle/ref_harness.py` records every random draw the reference makes.
Inputs (actions, groups, draws) and outputs (per-step and per-episode state) are stored
stacked over E so the oracle and the CUDA library can replay the scenario as one batch.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import ref_harness as rh  # noqa: E402
from oracle.env_oracle import (DIR_CHARS, PARTNER_NONE, encode_groups)  # noqa: E402

MARL_LAST = ("off_kbit_sum", "local_kbit_sum", "mec_queue_cycles", "delay_local_mean", "delay_edge_q_mean",
             "delay_edge_c_mean", "t_tx_mean", "backlog_kbit_mean", "mec_utilization", "local_util_mean",
             "qos_violation", "delay_mean", "energy_mean")

SPECS = {
    # name: dict(variant, V, M, seeds, episodes, T, refresh_every, params)
    "marl_v8_m40_yaml": dict(variant="marl", V=8, M=40, seeds=[11, 12, 13, 14], episodes=4, T=25,
                             refresh_every=2, params="yaml"),
    "marl_v8_m40_default": dict(variant="marl", V=8, M=40, seeds=[21, 22, 23], episodes=3, T=20,
                                refresh_every=1, params="default"),
    "marl_v6_m7_ragged": dict(variant="marl", V=6, M=7, seeds=[31, 32, 33], episodes=3, T=10,
                              refresh_every=1, params="yaml"),
    "marl_v32_m256": dict(variant="marl", V=32, M=256, seeds=[41, 42], episodes=2, T=8,
                          refresh_every=1, params="yaml"),
    "sarl_v8_m40": dict(variant="sarl", V=8, M=40, seeds=[51, 52, 53, 54], episodes=4, T=25,
                        refresh_every=2, params="default"),
    "sarl_v5_m33_ragged": dict(variant="sarl", V=5, M=33, seeds=[61, 62, 63], episodes=3, T=10,
                               refresh_every=1, params="default"),
    "sarl_v32_m256": dict(variant="sarl", V=32, M=256, seeds=[71, 72], episodes=2, T=8,
                          refresh_every=1, params="default"),
}


def random_groups(rng, V):
    """A random disjoint grouping in the reference's ragged list form."""
    users = list(rng.permutation(V))
    groups = []
    mode = rng.integers(0, 10)
    if mode == 0:
        return []  # empty list: G = max(1, 0), every rate 0
    if mode == 1:
        return [[int(u)] for u in users]  # all singletons
    while users:
        r = rng.random()
        if r < 0.55 and len(users) >= 2:
            groups.append([int(users.pop()), int(users.pop())])
        elif r < 0.85:
            groups.append([int(users.pop())])
        elif r < 0.92 and len(users) >= 3:
            groups.append([int(users.pop()), int(users.pop()), int(users.pop())])  # ignored size
        else:
            users.pop()  # unscheduled user
    return groups


def marl_actions(rng, T, V, floor):
    a = rng.random((T, 2, V)).astype(np.float32)
    a[:, 1, :] = np.maximum(a[:, 1, :], np.float32(floor))  # marl_train_bcd.py:1606-1608
    # edge cases: exact zeros, values that need clipping / projection
    mask = rng.random((T, 2, V))
    a[mask < 0.04] = 0.0
    a[(mask >= 0.04) & (mask < 0.07)] *= -1.0
    a[(mask >= 0.07) & (mask < 0.10)] += 1.0
    return a


def run_one(spec, seed):
    variant, V, M = spec["variant"], spec["V"], spec["M"]
    log = rh.DrawLog()
    rs = np.random.RandomState(seed)
    import random as _pyr

    _pyr.seed(seed)
    mod, env = rh.make_reference_env(variant, V, M, 3, log=log, mode="record", rs=rs)
    if variant == "marl" and spec["params"] == "yaml":
        rh.apply_marl_yaml_params(env)
    env.make_new_game()
    n_reset_ints = len(log.q["randint"])
    reset_ints = np.array(log.q["randint"], dtype=np.int32)
    reset_dirs = np.array([DIR_CHARS.index(c) for c in log.q["choice"]], dtype=np.int32)
    out = dict(reset_ints=reset_ints, reset_dirs=reset_dirs)
    out["reset_pos"] = np.array([v.position for v in env.vehicles], dtype=float)
    out["reset_dir"] = np.array([DIR_CHARS.index(v.direction) for v in env.vehicles], dtype=np.int32)
    out["reset_vel"] = np.array([v.velocity for v in env.vehicles], dtype=np.int32)
    out["reset_DataBuf"] = np.array(env.DataBuf, dtype=float)

    rng = np.random.default_rng(1000 + seed)
    EP, T = spec["episodes"], spec["T"]
    per_ep = {k: [] for k in ("mob_uniforms", "pos", "dir", "dist", "angle", "gains", "theta", "partner", "ngroups")}
    per_step = {}

    def push(name, val):
        per_step.setdefault(name, []).append(np.array(val, dtype=float).copy())

    actions, phases = [], []
    for ep in range(EP):
        u0 = len(log.q["uniform"])
        if ep % spec["refresh_every"] == 0:
            env.renew_positions()
            env.compute_parms()
        per_ep["mob_uniforms"].append(np.array(log.q["uniform"][u0:], dtype=float))
        per_ep["pos"].append(np.array([v.position for v in env.vehicles], dtype=float))
        per_ep["dir"].append(np.array([DIR_CHARS.index(v.direction) for v in env.vehicles], dtype=np.int32))
        per_ep["dist"].append(np.array(env.distances_R_i, dtype=float))
        per_ep["angle"].append(np.array(env.angles_R_i, dtype=float))
        if variant == "marl":
            env.optimize_phase_shift()
            env.update_channel_gains()
            per_ep["gains"].append(np.array(env.get_channel_gains(), dtype=float))
            per_ep["theta"].append(np.array(env.elements_phase_shift_complex, dtype=complex))
            groups = random_groups(rng, V)
            partner, ng = encode_groups(groups, V)
            per_ep["partner"].append(partner)
            per_ep["ngroups"].append(ng)
            acts = marl_actions(rng, T, V, env.cpu_share_floor)
            for t in range(T):
                a = acts[t].astype(np.float64)
                r_user, r_glob, buf, d_t, d_p, over_p, over_d = env.step(a, groups)
                push("reward_user", r_user); push("reward", r_glob); push("DataBuf", buf)
                push("data_t", d_t); push("data_p", d_p); push("over_power", over_p)
                push("rate", env.vehicle_rate); push("last_power_W", env.last_power_W)
                for k in MARL_LAST:
                    push("last_" + k, getattr(env, "last_" + k))
            actions.append(acts)
        else:
            acts = rng.random((T, 2, V)).astype(np.float32)
            acts[rng.random((T, 2, V)) < 0.03] = 0.0
            ph = (rng.random((T, M)) * 2 * np.pi).astype(np.float32)
            for t in range(T):
                r, buf, d_t, d_p, over_p, over_d = env.step(acts[t].astype(np.float64), ph[t].astype(np.float64))
                push("reward", r); push("DataBuf", buf); push("data_t", d_t); push("data_p", d_p)
                push("over_power", over_p); push("over_data", over_d); push("rate", env.vehicle_rate)
            actions.append(acts)
            phases.append(ph)
    out["arrivals"] = np.array(log.q["poisson"], dtype=np.int32).reshape(EP * T, V)
    out["actions"] = np.concatenate(actions, axis=0)
    if phases:
        out["phases"] = np.concatenate(phases, axis=0)
    assert len(log.q["randint"]) == n_reset_ints
    for k, v in per_ep.items():
        if v:
            out["ep_" + k] = v
    for k, v in per_step.items():
        out["step_" + k] = np.stack(v, axis=0)
    if variant == "marl":
        out["final_mec_queue_cycles"] = float(env.mec_queue_cycles)
    return out


def run_spec(spec):
    """Run all seeds and stack over E.  Shapes: per-step [T_total, E, ...], per-episode
    [EP, E, ...]; mobility uniforms are padded to a common width with 0.99 (never read)."""
    runs = [run_one(spec, s) for s in spec["seeds"]]
    E = len(runs)
    g = {}
    for k in ("reset_ints", "reset_dirs", "reset_pos", "reset_dir", "reset_vel", "reset_DataBuf"):
        g[k] = np.stack([r[k] for r in runs], axis=0)
    for k in runs[0]:
        if k.startswith("step_") or k in ("arrivals", "actions", "phases"):
            g[k] = np.stack([r[k] for r in runs], axis=1)
    EP = spec["episodes"]
    width = max(1, max(len(u) for r in runs for u in r["ep_mob_uniforms"]))
    mu = np.full((EP, E, width), 0.99)
    used = np.zeros((EP, E), dtype=np.int32)
    for e, r in enumerate(runs):
        for ep, u in enumerate(r["ep_mob_uniforms"]):
            mu[ep, e, :len(u)] = u
            used[ep, e] = len(u)
    g["ep_mob_uniforms"], g["ep_mob_used"] = mu, used
    for k in runs[0]:
        if k.startswith("ep_") and k != "ep_mob_uniforms":
            g[k] = np.stack([np.stack(r[k], axis=0) for r in runs], axis=1)
    if "final_mec_queue_cycles" in runs[0]:
        g["final_mec_queue_cycles"] = np.array([r["final_mec_queue_cycles"] for r in runs])
    g["meta"] = np.array([spec["variant"], str(spec["V"]), str(spec["M"]), str(spec["episodes"]), str(spec["T"]),
                          str(spec["refresh_every"]), spec["params"]])
    return g


def main():
    if not rh.reference_available():
        raise SystemExit("reference tree not mounted; fixtures can only be regenerated in the build container")
    for name, spec in SPECS.items():
        g = run_spec(spec)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **g)
        print(f"{name}: E={len(spec['seeds'])} -> {os.path.getsize(path) / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
