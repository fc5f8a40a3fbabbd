"""GPU parity of the batched NOMA pairing kernel (`risvec_pair_noma`) through the C ABI:
against the reference-generated fixtures and against the oracle on fresh random scenarios.
Pairs / groups / masks / history are index or small-integer data -> compared exactly."""
import numpy as np
import pytest

from tests.pairing_replay import PAIRING_FIXTURES, assert_pairing_equal, load_pairing, replay_gpu, replay_oracle

pytestmark = pytest.mark.gpu

TAU_RTOL = 1e-13    # float64: CUDA log10 vs glibc log10 may differ in the last bit


@pytest.mark.parametrize("name", PAIRING_FIXTURES)
def test_fixture_parity(name):
    g = load_pairing(name)
    got = replay_gpu(g)
    assert_pairing_equal(got, g, g, name, tau_rtol=TAU_RTOL)


def _random_scenario(V, E, T, seed, cfg, env):
    rng = np.random.default_rng(seed)
    kind = rng.integers(0, 4, E)
    lo = np.where(kind == 1, -14.5, np.where(kind == 2, -11.9, -15.6))
    hi = np.where(kind == 1, -12.1, np.where(kind == 2, -9.5, -10.7))
    gains = 10.0 ** (lo[:, None] + (hi - lo)[:, None] * rng.random((E, V)))
    ties = np.where(kind == 3)[0]
    gains[ties, 0] = gains[ties, V - 1]
    return dict(V=V, E=E, T=T, gains=gains, p01=rng.random((E, T, V)).astype(np.float32),
                freeze=(rng.random((E, T)) < 0.4).astype(np.int32) * (np.arange(T)[None, :] > 0),
                i_episode=np.full(E, int(rng.integers(0, 250)), np.int32), cfg=cfg, env=env)


YAML = dict(mask_topk_start=7, mask_topk_end=7, mask_tau_q_start=0.10, mask_tau_q_end=0.25, min_pair_target=3,
            mwm_accept_quantile=0.10, mwm_backoff_rounds=3, mwm_accept_q_step=0.05, abs_gain_min_db=-120.0)
ENV_YAML = dict(noise_power=10 ** ((-174 - 30) / 10) * 5e6, P_max=2.0, R_min=0.15)
ENV_DEF = dict(noise_power=10 ** ((-174 - 30) / 10) * 1e6, P_max=1.0, R_min=0.20)


@pytest.mark.parametrize("V,E,T,cfg,env", [
    (8, 1536, 3, YAML, ENV_YAML),
    (8, 512, 3, dict(mask_topk_end=4), ENV_DEF),
    (4, 256, 3, {}, ENV_DEF),
    (7, 256, 3, dict(mask_topk_end=3, min_pair_target=3), ENV_YAML),
    (10, 96, 2, dict(mask_topk_end=5, min_pair_target=4), ENV_YAML),
    (12, 24, 2, dict(mask_topk_end=6), ENV_DEF),
    (2, 64, 2, {}, ENV_DEF),
    (1, 8, 2, {}, ENV_DEF),
])
def test_random_scenarios_vs_oracle(V, E, T, cfg, env):
    g = _random_scenario(V, E, T, 100 + V, cfg, env)
    want = replay_oracle(g)
    got = replay_gpu(g)
    assert_pairing_equal(got, want, g, f"V={V}", tau_rtol=TAU_RTOL)


def test_partner_encoding_and_rollout_feed():
    """noma_partner / noma_ngroups are the rollout's group encoding of pairs + singles."""
    import torch

    from oracle.env_oracle import encode_groups
    from ris_vec_marl_b200 import BatchedEnviron, marl_yaml_overrides, mask_schedule

    E, V = 257, 8
    env = BatchedEnviron("marl", n_envs=E, n_veh=V, M=40, device=0, seed=5, **marl_yaml_overrides())
    env.set_pairing(yaml=True)
    env.make_new_game(); env.renew_positions(); env.compute_parms()
    env.optimize_phase_shift(); env.update_channel_gains()
    env.pair_reset()
    act = torch.rand(4, E, 2, V, device=env.device)
    K, q = mask_schedule(10, V, 7, 7, 0.10, 0.25, 200)
    partner, ngroups = env.pair_noma(act[0], K, q, recalc_mask=True)
    pairs = env.noma_pairs.cpu().numpy()
    npairs = env.noma_npairs.cpu().numpy()
    for e in range(E):
        pl = [[int(pairs[e, 2 * k]), int(pairs[e, 2 * k + 1])] for k in range(npairs[e])]
        used = {u for ab in pl for u in ab}
        groups = pl + [[k] for k in range(V) if k not in used]
        pe, ng = encode_groups(groups, V)
        assert np.array_equal(pe, partner[e].cpu().numpy()) and ng == int(ngroups[e])
        assert all(a < b for a, b in pl)
    a = env.state_dict()
    tr_view = env.rollout_marl(act, partner, ngroups)
    env.load_state_dict(a)
    tr_copy = env.rollout_marl(act, partner.clone(), ngroups.clone())
    for k in tr_view:
        assert torch.equal(tr_view[k], tr_copy[k]), k
    assert float(tr_view["rate"].abs().sum()) > 0
    env.close()


def test_pairing_rejects_large_v():
    import torch

    from ris_vec_marl_b200 import BatchedEnviron, RisvecError

    env = BatchedEnviron("marl", n_envs=4, n_veh=16, M=8, device=0, seed=1)
    with pytest.raises(RisvecError) as ei:
        env.pair_noma(torch.zeros(4, 16, device=env.device), 15, 0.2)
    assert ei.value.code == -2
    env.close()
