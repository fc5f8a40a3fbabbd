"""CPU: the pairing oracle against the golden fixtures and (when mounted) the live reference driver."""
import numpy as np
import pytest

from oracle import pairing_oracle as po
from oracle import ref_harness as rh
from tests.pairing_replay import PAIRING_FIXTURES, assert_pairing_equal, load_pairing, oracle_config, replay_oracle


@pytest.mark.parametrize("name", PAIRING_FIXTURES)
def test_oracle_reproduces_reference_fixture(name):
    g = load_pairing(name)
    got = replay_oracle(g)
    assert_pairing_equal(got, g, g, name)
    assert np.array_equal(got["K"], g["K"]) and np.allclose(got["q"], g["q"], rtol=0, atol=0)


def test_quantile_linear_is_numpy_quantile():
    rng = np.random.default_rng(5)
    for _ in range(400):
        n = int(rng.integers(1, 60))
        v = rng.normal(size=n)
        if n > 3:
            v[1] = v[2]
        q = float(rng.choice([0.0, 0.05, 0.1, 0.25, 0.3, 0.5, 0.8, 0.9, 0.95, 1.0, rng.random()]))
        assert po.quantile_linear(v, q) == float(np.quantile(v, q))


def test_matching_is_maximum_weight():
    """The DP result must weigh at least as much as every matching found by brute force (V = 6)."""
    import itertools

    rng = np.random.default_rng(9)
    for _ in range(30):
        N = 6
        S = rng.normal(size=(N, N))
        S = (S + S.T) / 2
        np.fill_diagonal(S, -np.inf)
        feas = np.ones((N, N), np.uint8) - np.eye(N, dtype=np.uint8)
        pairs = po.mwm_primary(S, feas, 1.0)      # accept every edge
        w = sum(S[i, j] for i, j in pairs)
        best = 0.0
        for perm in itertools.permutations(range(N)):
            for k in range(N // 2 + 1):
                cand = [(perm[2 * a], perm[2 * a + 1]) for a in range(k)]
                best = max(best, sum(S[i, j] for i, j in cand))
        assert w >= best - 1e-12


def test_mask_schedule_matches_driver_defaults():
    cfg = po.PairingConfig.marl_yaml(8)
    assert po.mask_schedule(0, 8, cfg) == (7, 0.10)
    k, q = po.mask_schedule(400, 8, cfg)
    assert k == 7 and abs(q - 0.25) < 1e-15
    cfg = po.PairingConfig.for_n_veh(8)
    assert po.mask_schedule(0, 8, cfg)[0] == 7 and po.mask_schedule(200, 8, cfg)[0] == 4


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")
def test_oracle_vs_live_reference_driver():
    """Fresh random scenarios through the driver's own statements (AST-compiled) vs the oracle."""
    rng = np.random.default_rng(77)
    for V, cfgkw in ((8, dict(min_pair_target=3, mwm_backoff_rounds=3, abs_gain_min_db=-120.0, mask_topk_end=7,
                              mask_tau_q_start=0.1, mask_tau_q_end=0.25)), (7, dict(mask_topk_end=3)), (4, {})):
        ref = rh.PairingReference(V, dict(cfgkw, qos_R_min_bpsHz=0.15))
        cfg = po.PairingConfig.for_n_veh(V, **cfgkw)
        for ep in range(40):
            g = 10.0 ** rng.uniform(-15.6, -10.7, V)
            i_ep = int(rng.integers(0, 300))
            ref.new_episode(i_ep)
            st = po.PairingState(V)
            K, q = po.mask_schedule(i_ep, V, cfg)
            for t in range(4):
                p = rng.uniform(0, 1, V).astype(np.float32).astype(np.float64)
                fr = bool(t and rng.random() < 0.4)
                r = ref.step(t, g, p, 2e-14, 2.0, freeze=fr)
                pairs, groups = po.pair_step(st, g, p, cfg, 2e-14, 2.0, 0.15, K, q, recalc_mask=(t == 0), reuse=fr)
                assert pairs == r["pairs"] and groups == r["groups"]
                assert np.array_equal(st.hist, r["hist"]) and np.array_equal(st.streak, r["streak"])
                assert st.tau == r["tau"] and st.K == r["K"]


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")
def test_helpers_match_reference_functions():
    ref = rh.PairingReference(8, {})
    h = ref.helpers
    rng = np.random.default_rng(3)
    for _ in range(50):
        g = 10.0 ** rng.uniform(-15.6, -10.7, 8)
        q = float(rng.uniform(0.05, 0.6))
        tau = po.adaptive_threshold(g, q)
        assert tau == h["_adaptive_threshold_from_delta_g"](g, q)
        K = int(rng.integers(1, 8))
        m = po.build_feasible_mask(g, tau, K)
        assert np.array_equal(m, h["_build_feasible_mask_from_delta_g"](g, tau, K).astype(np.uint8))
        hist = rng.integers(0, 4, (8, 8)).astype(np.float32)
        hist = hist + hist.T
        S = po.score_matrix(g, m, hist, 1.0, 0.3, -120.0, None, 6.0)
        S_ref = h["_score_matrix_from_gain_and_history"](g, m, hist, 1.0, 0.3, -120.0)
        assert np.array_equal(S, S_ref)
        m2 = po.relax_mask_once(m, g, 3.0, K)
        assert np.array_equal(m2, h["_relax_mask_once"](m, g, 3.0, K))
        assert po.mwm_primary(S, m, 0.3) == h["_mwm_primary"](S, m, 0.3, True)
