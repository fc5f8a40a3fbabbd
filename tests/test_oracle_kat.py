"""Known-answer vectors recorded from the unmodified reference during the survey
(SURVEY.md section 8c, KAT-1 .. KAT-5), replayed through the CPU oracle."""
import random

import numpy as np

from oracle.env_oracle import (DIR_DOWN, DIR_LEFT, DIR_RIGHT, DIR_UP, EnvOracle, GlobalNumpyDraws, InjectedDraws,
                               OracleParams, encode_groups)

KAT_POS = [[197.375, 225], [200.875, 173], [223, 200.875], [173, 197.375], [199.125, 300], [202.625, 50],
           [350, 202.625], [20, 199.125]]
KAT_DIR = [DIR_DOWN, DIR_UP, DIR_LEFT, DIR_RIGHT] * 2
KAT_VEL = [10] * 8
ACTION = [[0.9, 0.1, 0.5, 0, 1, 0.3, 0.7, 0.2], [0.8, 0.05, 0.5, 1, 1, 0.3, 0, 0.6]]
BUF0 = [4, 2.5, 3, 3.5, 40, 0, 7, 1]


def _env(variant, params=None, arrivals=None):
    d = InjectedDraws(arrivals=arrivals)
    env = EnvOracle(variant, 8, 40, 3, E=1, params=params, draws=d)
    env.set_vehicles(KAT_POS, KAT_DIR, KAT_VEL)
    env.compute_parms()
    return env


def test_kat1_geometry():
    env = _env("marl")
    np.testing.assert_allclose(env.distances_R_i[0], [33.002130613038, 55.919724829437, 30.446931290362,
                               57.211367970011, 85.953566679923, 172.493885761206, 133.244664527327,
                               202.454971845593], rtol=1e-12)
    np.testing.assert_allclose(env.angles_R_i[0], [-0.685561797973, -0.342008120718, 0.098532097419,
                               -0.821515053173, -0.242863685666, -0.100728207979, 0.975648822121,
                               -0.987873985888], rtol=1e-11)


def test_kat2_gains_and_bcd():
    env = _env("marl")
    env.get_next_phase(np.zeros((1, 40)))
    env.update_channel_gains()
    np.testing.assert_allclose(env.channel_gains[0], [7.935340236287e-15, 2.603125100658e-17, 1.258710661723e-14,
                               1.661112193481e-14, 1.127038648740e-20, 1.721890925918e-16, 6.244788460403e-15,
                               2.236236314878e-16], rtol=2e-9)
    env.elements_phase_shift_complex[:] = 0
    env.optimize_phase_shift()
    env.update_channel_gains()
    np.testing.assert_allclose(env.channel_gains[0], [2.008644004386e-12, 2.101820649079e-12, 3.136446554924e-12,
                               1.463046013554e-12, 4.935593085126e-13, 8.177164010538e-14, 1.556329791806e-13,
                               7.485749768441e-14], rtol=1e-11)
    np.testing.assert_allclose(env._objective()[0], 70298.8152284385, rtol=1e-12)


def test_kat3_marl_step():
    env = _env("marl", OracleParams.marl_yaml(), arrivals=np.full((1, 1, 8), 2))
    env.elements_phase_shift_complex[:] = 0
    env.optimize_phase_shift()
    env.update_channel_gains()
    env.DataBuf[0] = BUF0
    env.mec_queue_cycles[0] = 5e6
    partner, ng = encode_groups([[0, 3], [5, 2], [1], [4], [6, 7]], 8)
    r_user, r_glob, over_p = env.step_marl(np.array(ACTION)[None], partner[None], np.array([ng]))
    tol = dict(rtol=1e-9, atol=1e-15)
    np.testing.assert_allclose(env.vehicle_rate[0], [1.35056464209, 0.796054377274, 1.359655645435, 0, 0.937807248395,
                               0.106251009234, 0.62294917573, 0.058494002393], **tol)
    np.testing.assert_allclose(env.data_t[0], [6.752823210452, 3.980271886368, 6.798278227174, 0, 4.689036241976,
                               0.53125504617, 3.11474587865, 0.292470011966], **tol)
    np.testing.assert_allclose(env.data_p[0], [4, 1, 3, 3.5, 10, 0, 1, 1], **tol)
    np.testing.assert_allclose(env.DataBuf[0], [2, 2, 2, 2, 27.310963758024, 2, 4.88525412135, 2], **tol)
    np.testing.assert_allclose(r_user[0], [-1.1912e-03, -2.060380781972e-03, -8.025e-04, -1.501295,
                               -1.019443301683e-02, -1.5, -7.172122909024e-03, -1.500263866667], **tol)
    np.testing.assert_allclose(r_glob[0], -0.5653724379218119, rtol=1e-12)
    np.testing.assert_allclose(env.mec_queue_cycles[0], 5791134.636187739, rtol=1e-12)
    L = env.last
    np.testing.assert_allclose(L["delay_mean"][0], 0.002038180394952203, rtol=1e-11)
    np.testing.assert_allclose(L["energy_mean"][0], 0.0008342575268597349, rtol=1e-11)
    assert L["qos_violation"][0] == 0.375 and abs(L["mec_utilization"][0] - 1.0) < 1e-12
    np.testing.assert_allclose(L["local_util_mean"][0], 0.5770833333333334, rtol=1e-12)
    np.testing.assert_allclose(L["off_kbit_sum"][0], 9.303782120625797, rtol=1e-11)
    assert L["local_kbit_sum"][0] == 23.5 and L["backlog_kbit_mean"][0] == 7.625
    np.testing.assert_allclose(L["power_W"][0], [[0, 0.052760214879, 0, 0, 0.999999999999, 0, 0.98, 0],
                               [0.6912, 0.0027, 0.2025, 0.945, 2.7, 0, 0.0027, 0.0972]], rtol=1e-9, atol=1e-13)
    assert np.all(over_p == 0)


def test_kat4_sarl_step():
    env = _env("sarl", arrivals=np.full((1, 1, 8), 3))
    env.DataBuf[0] = BUF0
    phase = ((np.arange(40) * 0.61803398875) % 1.0) * 2 * np.pi
    reward, over_p = env.step_sarl(np.array(ACTION)[None], phase[None])
    tol = dict(rtol=1e-9, atol=1e-12)
    rate = [1.450414007044, 0.252553865123, 1.381525729301, 0, 1.199257328148, 0.482594598979, 0.066915207739,
            0.008584234067]
    np.testing.assert_allclose(env.vehicle_rate[0], rate, **tol)
    np.testing.assert_allclose(env.data_t[0], rate, **tol)
    np.testing.assert_allclose(env.data_p[0], [4, 1.587401051968, 3.419951893353, 4.308869380064, 4.308869380064,
                               2.884499140615, 0, 3.634241185664], **tol)
    np.testing.assert_allclose(env.DataBuf[0], [3, 3.660045082909, 3, 3, 37.491873291788, 3, 9.933084792261, 3], **tol)
    np.testing.assert_allclose(over_p[0], [0.592833749141, 0, 0.447005913277, 0.4640625, 0, 0.3, 0, 0.587819153344],
                               **tol)
    np.testing.assert_allclose(env.over_data[0], [1.450414007044, 0, 1.801477622654, 0.808869380064, 0,
                               3.367093739594, 0, 2.642825419731], **tol)
    np.testing.assert_allclose(reward[0], -5.4001252375219, rtol=1e-12)


def test_kat5_seeded_reset_stream():
    np.random.seed(0)
    random.seed(0)
    env = EnvOracle("marl", 8, 40, 3, E=1, draws=GlobalNumpyDraws())
    env.make_new_game()
    want = [([197.375, 225], DIR_DOWN, 10), ([200.875, 173], DIR_UP, 13), ([223, 200.875], DIR_LEFT, 11),
            ([173, 197.375], DIR_RIGHT, 12), ([197.375, 227], DIR_DOWN, 10), ([200.875, 178], DIR_UP, 14),
            ([221, 200.875], DIR_LEFT, 10), ([171, 197.375], DIR_RIGHT, 11)]
    for i, (p, d, v) in enumerate(want):
        assert list(env.pos[0, i]) == p and env.dir[0, i] == d and env.vel[0, i] == v
    assert np.all(env.DataBuf == 4.0)
    env.renew_positions(); env.compute_parms(); env.optimize_phase_shift(); env.update_channel_gains()
    np.testing.assert_allclose(env.channel_gains[0], [1.13472945e-11, 7.87616722e-13, 5.63733778e-12, 2.58672591e-12,
                               1.09960322e-11, 1.64846012e-12, 6.63244725e-12, 2.48785096e-12], rtol=1e-8)
