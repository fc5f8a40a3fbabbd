"""Oracle against the LIVE unmodified reference (only where /root/reference is mounted;
skipped elsewhere, e.g. on the GPU box).  Complements the committed fixtures with fresh
seeds and with the optional 3GPP channel branch (MARL/Environment.py:275-327)."""
import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle.env_oracle import EnvOracle, InjectedDraws, OracleParams

pytestmark = pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")


@pytest.mark.parametrize("variant,V,M,seed", [("marl", 8, 40, 101), ("sarl", 8, 40, 102), ("marl", 7, 12, 103),
                                              ("sarl", 4, 9, 104)])
def test_fresh_seed_scenarios(variant, V, M, seed):
    from tests.golden.make_golden import run_spec
    from tests.replay import OracleBackend, replay

    spec = dict(variant=variant, V=V, M=M, seeds=[seed, seed + 1000], episodes=3, T=12, refresh_every=1,
                params="yaml" if variant == "marl" else "default")
    g = run_spec(spec)
    meta = g.pop("meta")
    g.update(variant=variant, V=V, M=M, episodes=3, T=12, refresh_every=1, params=spec["params"], E=2)
    r = replay(g, OracleBackend(g))
    for k, v in r.items():
        if k.startswith("step_x_"):
            continue
        a, b = np.asarray(g[k]), np.asarray(v)
        if k in ("reset_pos", "reset_dir", "reset_vel", "ep_pos", "ep_dir", "ep_mob_used"):
            assert np.array_equal(a.astype(float), b.astype(float)), k
        elif np.iscomplexobj(a):
            np.testing.assert_allclose(b, a, rtol=0, atol=1e-14, err_msg=k)
        else:
            np.testing.assert_allclose(b.astype(float), a.astype(float), rtol=1e-11, atol=1e-300, err_msg=k)


@pytest.mark.parametrize("model,K_dB", [("3gpp_umi", 0.0), ("3gpp_uma", 0.0), ("3gpp_umi", 6.0)])
def test_3gpp_channel_branch(model, K_dB):
    log = rh.DrawLog()
    rs = np.random.RandomState(7)
    mod, ref = rh.make_reference_env("marl", 8, 40, 3, log=log, mode="record", rs=rs)
    ref.make_new_game()
    ref.channel_model = model
    ref.rician_K_dB = K_dB
    ref.update_channel_gains()
    p = OracleParams()
    p.channel_model, p.rician_K_dB = model, K_dB
    d = InjectedDraws(reset_ints=np.array(log.q["randint"])[None])
    env = EnvOracle("marl", 8, 40, 3, E=1, params=p, draws=d)
    env.make_new_game()
    d.set_channel_draws(np.array(log.q["rand"], float)[None], np.array(log.q["normal"][1:], float)[None],
                        np.array(log.q["exponential"] or [0.0], float)[None])
    env.update_channel_gains()
    np.testing.assert_allclose(env.channel_gains[0], ref.channel_gains, rtol=1e-12)
