"""The drop-in `Environment.Environ` objects (E = 1, numpy in / numpy out, global numpy RNG
stream) against the oracle consuming the same seeded stream: a reference driver sees the same
trajectory it would see from the reference module, within the float32 tolerance."""
import importlib.util
import os
import random

import numpy as np
import pytest

from oracle.env_oracle import EnvOracle, GlobalNumpyDraws, Lanes, OracleParams, encode_groups

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_compat(variant):
    path = os.path.join(ROOT, "ris_vec_marl_b200", "compat", variant, "Environment.py")
    spec = importlib.util.spec_from_file_location(f"compat_{variant}_Environment", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def ctor_args(V, M):
    ln = Lanes.default()
    return (ln.down, ln.up, ln.left, ln.right, 400, 400, V, M, 3)


def test_marl_driver_sequence_matches_oracle_stream():
    mod = load_compat("marl")
    V, M = 8, 40
    groups = [[0, 3], [5, 2], [1], [4], [6, 7]]
    rng = np.random.default_rng(0)
    acts = rng.random((3, 30, 2, V)).astype(np.float32)
    acts[:, :, 1, :] = np.maximum(acts[:, :, 1, :], 0.1)

    def drive(make_env, step):
        np.random.seed(42); random.seed(42)
        env = make_env()
        env.make_new_game()
        out = []
        for ep in range(3):
            env.renew_positions(); env.compute_parms()
            env.optimize_phase_shift(); env.update_channel_gains()
            for t in range(30):
                out.append(step(env, acts[ep, t].astype(np.float64)))
        return env, np.array(out)

    def make_gpu():
        env = mod.Environ(*ctor_args(V, M))
        # attribute overlay exactly as marl_train_bcd.py:563-594 does it
        env.rate = 1.0; env.f_local_max = 3e9; env.f_edge_max = 2e9; env.cycles_per_bit = 300.0
        env.P_max = 2.0; env.bandwidth = 5.0; env.bandwidth_hz = env.bandwidth * 1e6
        env.noise_power = env.N0_W_per_Hz * env.bandwidth_hz
        env.w_d = 1.0; env.w_e = 1.0; env.R_min_bpsHz = 0.15; env.D_max_s = 0.12; env.qos_penalty = 1.5
        env.power_scale = 0.7; env.cpu_share_floor = 0.10
        return env

    def gpu_step(env, a):
        r_user, r_glob, buf, d_t, d_p, over_p, over_d = env.step(a, groups)
        assert abs(r_glob - np.mean(r_user)) < 1e-5 and over_d.shape == (V,)
        return np.concatenate([r_user, [r_glob], buf, d_t, d_p, env.vehicle_rate, [env.last_delay_mean,
                               env.last_energy_mean, env.last_qos_violation, env.mec_queue_cycles * 1e-9]])

    def make_or():
        return EnvOracle("marl", V, M, 3, E=1, params=OracleParams.marl_yaml(), draws=GlobalNumpyDraws())

    part, ng = encode_groups(groups, V)

    def or_step(env, a):
        r_user, r_glob, _ = env.step_marl(a[None], part[None], np.array([ng]))
        L = env.last
        return np.concatenate([r_user[0], r_glob, env.DataBuf[0], env.data_t[0], env.data_p[0], env.vehicle_rate[0],
                               [L["delay_mean"][0], L["energy_mean"][0], L["qos_violation"][0],
                                env.mec_queue_cycles[0] * 1e-9]])

    g_env, got = drive(make_gpu, gpu_step)
    o_env, want = drive(make_or, or_step)
    for i, v in enumerate(g_env.vehicles):
        assert v.position == list(o_env.pos[0, i]) and v.velocity == o_env.vel[0, i]
        assert v.direction == "udlr"[o_env.dir[0, i]]
    np.testing.assert_allclose(g_env.get_channel_gains(), o_env.channel_gains[0], rtol=1e-9)
    np.testing.assert_allclose(got, want, rtol=1e-5, atol=2e-5)


def test_sarl_driver_sequence_matches_oracle_stream():
    mod = load_compat("sarl")
    V, M = 8, 40
    rng = np.random.default_rng(1)
    acts = rng.random((40, 2, V)).astype(np.float32)
    phs = (rng.random((40, M)) * 2 * np.pi).astype(np.float32)

    np.random.seed(1234); random.seed(1)
    env = mod.Environ(*ctor_args(V, M))
    env.make_new_game(); env.renew_positions(); env.compute_parms()
    got = []
    for t in range(40):
        r, buf, d_t, d_p, over_p, over_d = env.step(acts[t].astype(np.float64), phs[t].astype(np.float64))
        got.append(np.concatenate([[r], buf, d_t, d_p, over_p, over_d, env.vehicle_rate]))
    tail_gpu = np.random.randint(0, 1 << 30)
    assert np.allclose(env.elements_phase_shift_real, phs[-1])

    np.random.seed(1234); random.seed(1)
    o = EnvOracle("sarl", V, M, 3, E=1, draws=GlobalNumpyDraws())
    o.make_new_game(); o.renew_positions(); o.compute_parms()
    want = []
    for t in range(40):
        r, over_p = o.step_sarl(acts[t][None].astype(np.float64), phs[t][None].astype(np.float64))
        want.append(np.concatenate([r, o.DataBuf[0], o.data_t[0], o.data_p[0], over_p[0], o.over_data[0],
                                    o.vehicle_rate[0]]))
    tail_or = np.random.randint(0, 1 << 30)
    np.testing.assert_allclose(np.array(got), np.array(want), rtol=1e-5, atol=2e-5)
    assert tail_gpu == tail_or  # same number of draws consumed from the global stream
