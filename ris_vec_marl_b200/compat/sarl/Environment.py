"""Drop-in for Simulation-SARL/Environment.py (ddpg_train.py:5, ddpg_test.py).  Like the
reference module it seeds the global numpy stream at import (Environment.py:7)."""
import numpy as np

from ris_vec_marl_b200.compat_env import SarlEnviron as Environ, Vehicle  # noqa: F401

np.random.seed(1234)

n_veh = 8
RIS_x, RIS_y, RIS_z = 220, 220, 25
BS_x, BS_y, BS_z = 0, 0, 25
ro = 10 ** -2
lamb = 1
d = 0.5
sigma = 10 ** (-7)
alpha1 = 2.2
alpha2 = 2.5
