"""The CPU oracle against the golden fixtures generated from the unmodified reference
(tests/golden/make_golden.py).  Integer / index / position state must be bit-exact; float64
quantities must agree to a few ulp (the oracle is batched numpy, the reference is scalar
python + numpy, so libm `cexp` vs `cmath.exp` may differ in the last bit)."""
import numpy as np
import pytest

from tests.replay import OracleBackend, golden_names, load_golden, replay

EXACT = ("reset_pos", "reset_dir", "reset_vel", "reset_DataBuf", "ep_pos", "ep_dir", "ep_mob_used")


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_fixture(name):
    g = load_golden(name)
    r = replay(g, OracleBackend(g))
    checked = 0
    for k, v in r.items():
        if k.startswith("step_x_"):
            continue
        assert k in g, k
        a, b = np.asarray(g[k]), np.asarray(v)
        assert a.shape == b.shape, (k, a.shape, b.shape)
        if k in EXACT:
            assert np.array_equal(a.astype(np.float64), b.astype(np.float64)), k
        elif np.iscomplexobj(a):
            np.testing.assert_allclose(b, a, rtol=0, atol=1e-14, err_msg=k)
        else:
            np.testing.assert_allclose(b.astype(float), a.astype(float), rtol=1e-11, atol=1e-300, err_msg=k)
        checked += 1
    assert checked >= 12


def test_fixture_inventory():
    names = golden_names()
    for need in ("marl_v8_m40_yaml", "sarl_v8_m40", "marl_v32_m256", "sarl_v32_m256", "marl_v6_m7_ragged",
                 "sarl_v5_m33_ragged"):
        assert need in names
