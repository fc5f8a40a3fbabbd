"""Drop-in for Simulation-MARL-BCD/Environment.py: put this directory in front of the
reference directory on sys.path and `import Environment` resolves here
(marl_train_bcd.py:2, marl_test.py:3).  Same class name and constructor signature
(Environment.py:56-57); all arithmetic runs in the sm_100a kernels."""
from ris_vec_marl_b200.compat_env import MarlEnviron as Environ, Vehicle  # noqa: F401

# module constants the reference exposes (Environment.py:29-42)
RIS_x, RIS_y, RIS_z = 220, 220, 25
BS_x, BS_y, BS_z = 0, 0, 25
ro = 10 ** -2
lamb = 1
d = 0.5
sigma = 10 ** (-7)
alpha1 = 2.2
alpha2 = 2.5
