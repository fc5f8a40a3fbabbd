"""Generate the NOMA-pairing golden fixtures (`pairing_*.npz`) from the UNMODIFIED reference driver.

    python tests/golden/make_golden_pairing.py

Each scenario is E independent "episodes" of the pairing stage of
`Simulation-MARL-BCD/marl_train_bcd.py` (:1315-1561), executed by `oracle/ref_harness.PairingReference`
(the driver's own helper functions and call-site statements, compiled from the reference source by
AST, with `np.argsort` forced stable -- see oracle/pairing_oracle.py).  Gains are a mix of real BCD
gains produced by the reference `Environ` (seeded) and log-uniform synthetic ones incl. values below
the 1e-15 / 1e-12 clips and exact ties; `p01` is float32-representable.
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

YAML_CFG = dict(mask_topk_start=7, mask_topk_end=7, mask_tau_q_start=0.10, mask_tau_q_end=0.25, min_pair_target=3,
                mwm_accept_quantile=0.10, mwm_backoff_rounds=3, mwm_accept_q_step=0.05, abs_gain_min_db=-120.0)
ENV_YAML = dict(noise_power=10 ** ((-174 - 30) / 10) * 5e6, P_max=2.0, R_min=0.15)
ENV_DEFAULT = dict(noise_power=10 ** ((-174 - 30) / 10) * 1e6, P_max=1.0, R_min=0.20)

SPECS = {
    # name: V, cfg overrides, env params, episodes, steps, share of real (reference Environ) gains
    "pairing_v8_yaml": dict(V=8, cfg=YAML_CFG, env=ENV_YAML, E=48, T=6, real=24),
    "pairing_v8_default": dict(V=8, cfg={}, env=ENV_DEFAULT, E=40, T=6, real=12),
    "pairing_v5_yaml": dict(V=5, cfg=dict(YAML_CFG, mask_topk_start=4, mask_topk_end=3, min_pair_target=2),
                            env=ENV_YAML, E=24, T=5, real=0),
    "pairing_v6_default": dict(V=6, cfg=dict(mask_topk_end=2), env=ENV_DEFAULT, E=24, T=5, real=0),
    "pairing_v12_default": dict(V=12, cfg=dict(mask_topk_end=5), env=ENV_DEFAULT, E=8, T=4, real=4),
}


def real_gains(V, seed, yaml_env):
    """BCD gains of a seeded reference env (make_new_game -> renew -> compute -> BCD -> gains)."""
    log = rh.DrawLog()
    _, env = rh.make_reference_env("marl", V, 40, 3, log, "record", np.random.RandomState(seed))
    if yaml_env:
        rh.apply_marl_yaml_params(env)
    env.make_new_game()
    for _ in range(1 + seed % 3):
        env.renew_positions()
    env.compute_parms()
    env.optimize_phase_shift()
    env.update_channel_gains()
    return np.array(env.get_channel_gains(), dtype=np.float64)


def synthetic_gains(rng, V, kind):
    g = 10.0 ** rng.uniform(-15.6, -10.7, V)
    if kind == 1:       # everything under the 1e-12 clip: the score degenerates to history + QoS
        g = 10.0 ** rng.uniform(-14.5, -12.1, V)
    elif kind == 2:     # strong spread, all above the clip
        g = 10.0 ** rng.uniform(-11.9, -9.5, V)
    elif kind == 3:     # exact ties
        g[rng.integers(0, V)] = g[rng.integers(0, V)]
        g[rng.integers(0, V)] = g[rng.integers(0, V)]
    return g


def build(name, spec):
    V, E, T = spec["V"], spec["E"], spec["T"]
    rng = np.random.default_rng(sum(map(ord, name)))
    ref = rh.PairingReference(V, dict(spec["cfg"], qos_R_min_bpsHz=spec["env"]["R_min"]))
    out = dict(gains=np.zeros((E, V)), p01=np.zeros((E, T, V), np.float32), freeze=np.zeros((E, T), np.int32),
               i_episode=np.zeros(E, np.int32), pairs=np.full((E, T, V), -1, np.int32),
               npairs=np.zeros((E, T), np.int32), ngroups=np.zeros((E, T), np.int32),
               hist=np.zeros((E, T, V, V), np.float32), streak=np.zeros((E, T, V), np.int32), tau=np.zeros(E),
               K=np.zeros(E, np.int32), q=np.zeros(E), rounds=np.zeros((E, T), np.int32),
               mask=np.zeros((E, V, V), np.uint8))
    groups_all = []
    for e in range(E):
        if e < spec["real"]:
            g = real_gains(V, 1000 + e, spec["env"] is ENV_YAML)
        else:
            g = synthetic_gains(rng, V, e % 5)
        i_ep = int(rng.integers(0, 260))
        ref.new_episode(i_ep)
        out["gains"][e], out["i_episode"][e] = g, i_ep
        glist = []
        for t in range(T):
            p = rng.uniform(0, 1, V).astype(np.float32)
            if rng.random() < 0.15:
                p[rng.integers(0, V)] = 0.0
            fr = int(t > 0 and rng.random() < 0.5)
            r = ref.step(t, g, p.astype(np.float64), spec["env"]["noise_power"], spec["env"]["P_max"], freeze=bool(fr))
            out["p01"][e, t], out["freeze"][e, t] = p, fr
            flat = [u for ab in r["pairs"] for u in ab]
            out["pairs"][e, t, :len(flat)] = flat
            out["npairs"][e, t], out["ngroups"][e, t] = len(r["pairs"]), len(r["groups"])
            out["hist"][e, t], out["streak"][e, t], out["rounds"][e, t] = r["hist"], r["streak"], r["rounds"]
            if t == 0:
                out["tau"][e], out["K"][e], out["q"][e], out["mask"][e] = r["tau"], r["K"], ref.ns["last_q_now"], r["mask"]
            glist.append(r["groups"])
        groups_all.append(glist)
    meta = dict(V=V, E=E, T=T, **{"cfg_" + k: v for k, v in spec["cfg"].items()},
                **{"env_" + k: v for k, v in spec["env"].items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out,
                        meta_keys=np.array(list(meta.keys())), meta_vals=np.array([float(v) for v in meta.values()]))
    print(name, "pairs/solve:", out["npairs"].mean(), "rounds total:", out["rounds"].sum())


if __name__ == "__main__":
    if not rh.reference_available():
        raise SystemExit("reference tree not mounted")
    for n, s in SPECS.items():
        build(n, s)
