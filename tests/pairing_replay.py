"""Replay helpers for the NOMA-pairing fixtures (`tests/golden/pairing_*.npz`, written by
`tests/golden/make_golden_pairing.py` from the unmodified reference driver)."""
from __future__ import annotations

import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
PAIRING_FIXTURES = ("pairing_v8_yaml", "pairing_v8_default", "pairing_v5_yaml", "pairing_v6_default",
                    "pairing_v12_default")
OUT_KEYS = ("pairs", "npairs", "ngroups", "hist", "streak", "rounds")


def load_pairing(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    g = {k: z[k] for k in z.files if not k.startswith("meta_")}
    meta = dict(zip([str(k) for k in z["meta_keys"]], [float(v) for v in z["meta_vals"]]))
    g["V"], g["E"], g["T"] = int(meta["V"]), int(meta["E"]), int(meta["T"])
    g["cfg"] = {k[4:]: v for k, v in meta.items() if k.startswith("cfg_")}
    g["env"] = {k[4:]: v for k, v in meta.items() if k.startswith("env_")}
    return g


def oracle_config(g):
    from oracle.pairing_oracle import PairingConfig

    kw = {}
    for k, v in g["cfg"].items():
        kw[k] = int(v) if k in ("mask_topk_start", "mask_topk_end", "min_pair_target", "mwm_backoff_rounds") else v
    return PairingConfig.for_n_veh(g["V"], **kw)


def replay_oracle(g):
    """Run the oracle over a fixture's inputs -> dict of arrays shaped like the fixture's outputs."""
    from oracle import pairing_oracle as po

    V, E, T = g["V"], g["E"], g["T"]
    cfg = oracle_config(g)
    out = dict(pairs=np.full((E, T, V), -1, np.int32), npairs=np.zeros((E, T), np.int32),
               ngroups=np.zeros((E, T), np.int32), hist=np.zeros((E, T, V, V), np.float32),
               streak=np.zeros((E, T, V), np.int32), rounds=np.zeros((E, T), np.int32), tau=np.zeros(E),
               mask=np.zeros((E, V, V), np.uint8), K=np.zeros(E, np.int32), q=np.zeros(E))
    for e in range(E):
        st = po.PairingState(V)
        K, q = po.mask_schedule(int(g["i_episode"][e]), V, cfg)
        for t in range(T):
            info = {}
            pairs, groups = po.pair_step(st, g["gains"][e], g["p01"][e, t], cfg, g["env"]["noise_power"],
                                         g["env"]["P_max"], g["env"]["R_min"], K, q, recalc_mask=(t == 0),
                                         reuse=bool(g["freeze"][e, t]), info=info)
            flat = [u for ab in pairs for u in ab]
            out["pairs"][e, t, :len(flat)] = flat
            out["npairs"][e, t], out["ngroups"][e, t] = len(pairs), len(groups)
            out["hist"][e, t], out["streak"][e, t], out["rounds"][e, t] = st.hist, st.streak, info["rounds"]
            if t == 0:
                out["tau"][e], out["mask"][e], out["K"][e], out["q"][e] = st.tau, st.mask, K, q
    return out


def replay_gpu(g, device=0):
    """Run `BatchedEnviron.pair_noma` over a fixture's inputs (one env per fixture episode; the K / q
    schedule differs per episode, so envs that share (K, q) are solved together and scattered back)."""
    import torch

    from ris_vec_marl_b200 import BatchedEnviron

    V, E, T = g["V"], g["E"], g["T"]
    cfg = oracle_config(g)
    from oracle import pairing_oracle as po

    sched = [po.mask_schedule(int(i), V, cfg) for i in g["i_episode"]]
    out = dict(pairs=np.full((E, T, V), -1, np.int32), npairs=np.zeros((E, T), np.int32),
               ngroups=np.zeros((E, T), np.int32), hist=np.zeros((E, T, V, V), np.float32),
               streak=np.zeros((E, T, V), np.int32), rounds=np.zeros((E, T), np.int32), tau=np.zeros(E),
               mask=np.zeros((E, V, V), np.uint8), partner=np.zeros((E, T, V), np.int32))
    for kq in sorted(set(sched)):
        idx = np.array([i for i, s in enumerate(sched) if s == kq])
        env = BatchedEnviron("marl", n_envs=len(idx), n_veh=V, M=8, device=device, seed=1,
                             noise_power=g["env"]["noise_power"], P_max=g["env"]["P_max"],
                             R_min_bpsHz=g["env"]["R_min"])
        env.set_pairing(**{k: v for k, v in g["cfg"].items() if not k.startswith("mask_")})
        env.gains.copy_(torch.as_tensor(g["gains"][idx], device=env.device))
        fold = (kq[0] + int(round(kq[1] * 1000))) % 2 == 1   # exercise both ways of starting an episode
        if fold:     # stale state from a "previous episode" that new_episode=True must ignore
            env.pair_hist.fill_(3.0); env.unpaired_streak.fill_(5); env.noma_ngroups.fill_(4)
        else:
            env.pair_reset()
        for t in range(T):
            p01 = torch.as_tensor(g["p01"][idx, t], device=env.device)
            reuse = torch.as_tensor(g["freeze"][idx, t], device=env.device)
            env.pair_noma(p01, kq[0], kq[1], recalc_mask=(t == 0), reuse=reuse, new_episode=(fold and t == 0))
            out["pairs"][idx, t] = env.noma_pairs.cpu().numpy()
            out["npairs"][idx, t] = env.noma_npairs.cpu().numpy()
            out["ngroups"][idx, t] = env.noma_ngroups.cpu().numpy()
            out["partner"][idx, t] = env.noma_partner.cpu().numpy()
            out["hist"][idx, t] = env.pair_hist.cpu().numpy()
            out["streak"][idx, t] = env.unpaired_streak.cpu().numpy()
            out["rounds"][idx, t] = env.pair_rounds.cpu().numpy()
            if t == 0:
                out["tau"][idx] = env.pair_tau.cpu().numpy()
                out["mask"][idx] = env.pair_mask.cpu().numpy()
        env.close()
    return out


def assert_pairing_equal(got, want, g, what, tau_rtol=0.0):
    for k in OUT_KEYS:
        a, b = got[k], want[k]
        if k == "rounds":       # the reference leaves round_id from the discarded solve on frozen steps
            sel = g["freeze"] == 0
            sel[:, 0] = True
            a, b = a[sel], b[sel]
        assert np.array_equal(a, b), f"{what}: {k} differs at {np.argwhere(a != b)[:5].tolist()}"
    assert np.array_equal(got["mask"], want["mask"]), f"{what}: mask differs"
    if tau_rtol == 0.0:
        assert np.array_equal(got["tau"], want["tau"]), f"{what}: tau differs"
    else:
        np.testing.assert_allclose(got["tau"], want["tau"], rtol=tau_rtol, atol=0, err_msg=what)
