"""Parity rules between the CUDA path (float32 arithmetic, float64 where the reference's
rounding residues matter) and the float64 oracle / reference fixtures.

Tolerance (north star): rtol 1e-5, plus the per-quantity absolute floors of SURVEY.md 8a-notes
(float64 leaves +-1e-16 residues where float32 leaves +-1e-7 ones).  Samples sitting on a
discontinuity of the reference (QoS thresholds, SARL penalties, the edge-queue share
amplification of MARL/Environment.py:629) are band-excluded and counted.
"""
import numpy as np

RTOL = 1e-5
ATOL = {
    "rate": 1e-6, "data_t": 1e-5, "data_p": 1e-5, "DataBuf": 1e-5, "over_data": 1e-5, "over_power": 2e-6,
    "reward_user": 2e-6, "reward": 2e-6, "last_power_W": 1e-6,
    "last_delay_mean": 1e-9, "last_energy_mean": 1e-9, "last_delay_local_mean": 1e-9, "last_delay_edge_q_mean": 1e-9,
    "last_delay_edge_c_mean": 1e-9, "last_t_tx_mean": 1e-9, "last_backlog_kbit_mean": 1e-5,
    "last_mec_utilization": 1e-6, "last_local_util_mean": 1e-6, "last_qos_violation": 1e-6,
    "last_off_kbit_sum": 1e-5, "last_local_kbit_sum": 1e-5, "last_mec_queue_cycles": 1.0,
}
EXACT = ("reset_pos", "reset_dir", "reset_vel", "reset_DataBuf", "ep_pos", "ep_dir", "ep_mob_used")


def _close(got, want, atol, mask=None, what=""):
    got, want = np.asarray(got, float), np.asarray(want, float)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    bad = np.abs(got - want) > (atol + RTOL * np.abs(want))
    if mask is not None:
        bad &= ~mask
    if bad.any():
        i = np.argwhere(bad)[0]
        raise AssertionError(f"{what}: {bad.sum()} / {bad.size} outside rtol={RTOL} atol={atol}; first at {tuple(i)}: "
                             f"got {got[tuple(i)]!r} want {want[tuple(i)]!r}")


def theta_equal_up_to_rotation(got, want, tol=1e-9):
    """BCD solutions are defined up to a global rotation by a multiple of 2*pi/2^cb
    (hard part 4): theta_got = r * theta_want with one unit r per env."""
    got, want = np.asarray(got), np.asarray(want)
    ratio = got * np.conj(want)  # |theta| = 1
    return np.max(np.abs(ratio - ratio[..., :1])) < tol and np.allclose(np.abs(ratio), 1.0, atol=tol)


def marl_exclusions(g, want, p):
    """Masks of samples on a reference discontinuity, from the ORACLE's float64 values."""
    rate, delay = want["step_rate"], want["step_x_delay"]
    user = np.zeros(rate.shape, dtype=bool)
    if p.qos_enable:
        user |= np.abs(rate - p.R_min_bpsHz) < 1e-5 * p.R_min_bpsHz
        user |= np.abs(delay - p.D_max_s) < 1e-5 * p.D_max_s
    # share amplification: every offload is a rounding residue yet the queue is non-empty
    amp = (want["step_x_edge_in_sum"] < 1e-3) & (want["step_x_q_before"] > 0) & (want["step_x_edge_in_sum"] > 0)
    env = user.any(axis=-1) | amp
    user = user | amp[..., None]
    return user, env


MAX_EXCLUDED_FRACTION = 0.05  # a band may never swallow a quantity: at most 5 % of the env-steps


def sarl_reward_band(buf_signed, over_data, M):
    """Samples on a discontinuity of the SARL reward (SARL/Environment.py:343-352): the per-user
    reward drops by penalty1 where the SIGNED pre-clamp buffer `DataBuf - (data_t + data_p)` crosses 0
    and by penalty2 where `over_data` crosses 2.  The band is built from the oracle's float64 values
    and is as wide as the error the CUDA path may carry there (DataBuf / over_data tolerance).  A
    buffer that drained (clamped to exactly 0) is NOT on the discontinuity: its signed value is
    well below 0."""
    tol = max(5e-5, 8 * sarl_rate_atol(M))
    return (np.abs(buf_signed) < tol) | (np.abs(over_data - 2.0) < tol)


def assert_band_small(mask, what):
    frac = float(np.mean(mask)) if np.size(mask) else 0.0
    assert frac <= MAX_EXCLUDED_FRACTION, f"{what}: {frac:.1%} of the samples band-excluded (> {MAX_EXCLUDED_FRACTION:.0%})"
    return frac


def sarl_exclusions(g, want):
    user = sarl_reward_band(want["step_x_buf_signed"], want["step_over_data"], g["M"])
    return user, user.any(axis=-1)


def sarl_rate_atol(M):
    """SARL recomputes the cascaded sum S_v = sum_m theta_m w_vm every step from float32 phasors:
    2M components carrying ~6e-8 representation error each bound |S_v| to ~1e-7*sqrt(2M) and
    the float32 accumulation adds about as much again.  Where the M phasors interfere
    destructively (|S| << sqrt(M)) that absolute error dominates rate = ln(1 + c |S|^2), so
    the SARL rate (and data_t = rate in kbit) gets an absolute floor growing with sqrt(M)
    (hard part 2: "the tail needs ... an atol tied to M*ulp")."""
    return 6e-7 * np.sqrt(M)


def compare_replays(g, got, want, params=None):
    """`got` = CUDA replay, `want` = oracle replay (with its `step_x_*` extras); both are also
    checked against the reference fixture `g` where it holds the key."""
    from tests.replay import oracle_params

    p = params if params is not None else oracle_params(g)
    checked, excluded = 0, 0
    for k in EXACT:
        if k in got:
            for ref in (g.get(k), want.get(k)):
                if ref is not None:
                    assert np.array_equal(np.asarray(got[k], float), np.asarray(ref, float)), f"{k} not bit-exact"
            checked += 1
    for k in ("ep_dist", "ep_angle"):
        np.testing.assert_allclose(got[k], g[k], rtol=1e-14, atol=0, err_msg=k)
        checked += 1
    if g["variant"] == "marl":
        np.testing.assert_allclose(got["ep_gains"], g["ep_gains"], rtol=1e-9, err_msg="ep_gains")
        assert theta_equal_up_to_rotation(got["ep_theta"], g["ep_theta"]), "BCD theta differs beyond a global rotation"
        np.testing.assert_allclose(got["final_mec_queue_cycles"], g["final_mec_queue_cycles"], rtol=1e-6, atol=1.0)
        user_x, env_x = marl_exclusions(g, want, p)
        checked += 3
    else:
        user_x, env_x = sarl_exclusions(g, want)
    excluded = int(user_x.sum())
    assert_band_small(env_x, f"{g['variant']} reward band (env-steps)")
    per_user_masked = ("reward_user",)
    per_env_masked = ("reward", "last_qos_violation", "last_delay_mean", "last_delay_edge_q_mean")
    for k, v in got.items():
        if not k.startswith("step_"):
            continue
        name = k[5:]
        atol = ATOL[name]
        if g["variant"] == "sarl" and name in ("rate", "data_t"):
            atol = max(atol, sarl_rate_atol(g["M"]))
        if g["variant"] == "sarl" and name in ("DataBuf", "over_data"):
            atol = max(atol, 4 * sarl_rate_atol(g["M"]))  # accumulates a few steps of data_t error
        mask = None
        if name in per_user_masked:
            mask = user_x
        elif name in per_env_masked:
            mask = env_x
        for ref, tag in ((want.get(k), "oracle"), (g.get(k), "fixture")):
            if ref is not None:
                _close(v, ref, atol, mask, f"{k} vs {tag}")
        checked += 1
    return dict(checked=checked, excluded=excluded, reward_env_steps=int(env_x.size),
                reward_env_steps_excluded=int(env_x.sum()))
