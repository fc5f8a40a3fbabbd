"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Drives the *unmodified* reference `Environment.py` modules (when `/root/reference`
is mounted, i.e. in the build container, never on the GPU box) with controlled
randomness, so that

  * the numpy restatement in `oracle/env_oracle.py` can be pinned against the real
    reference code, and
  * golden fixtures under `tests/golden/` can be (re)generated
    (`tests/golden/make_golden.py`).

Mechanism: the reference modules reach randomness only through
the module global `np` (`np.random.randint/uniform/poisson/normal/rand/exponential`)
and python's `random.choice`.  After loading a module by path under a private name we
swap those two globals for proxies that either RECORD every draw in call order or
REPLAY pre-drawn values.  No reference file is edited or copied.
"""
from __future__ import annotations

import importlib.util
import os
import random as _pyrandom
import sys

import numpy as _np

REFERENCE_ROOT = os.environ.get("RISVEC_REFERENCE_ROOT", "/root/reference")
_VARIANT_DIR = {"marl": "Simulation-MARL-BCD", "sarl": "Simulation-SARL"}

# Lane constants used by both reference drivers
# (Simulation-MARL-BCD/marl_train_bcd.py:446-449, Simulation-SARL/ddpg_train.py:22-29).
UP_LANES = [i / 2.0 for i in [400 + 3.5 / 2, 400 + 3.5 + 3.5 / 2, 800 + 3.5 / 2, 800 + 3.5 + 3.5 / 2]]
DOWN_LANES = [i / 2.0 for i in [400 - 3.5 - 3.5 / 2, 400 - 3.5 / 2, 800 - 3.5 - 3.5 / 2, 800 - 3.5 / 2]]
LEFT_LANES = list(UP_LANES)
RIGHT_LANES = list(DOWN_LANES)
WIDTH = 400
HEIGHT = 400


def reference_available() -> bool:
    return all(
        os.path.isfile(os.path.join(REFERENCE_ROOT, d, "Environment.py")) for d in _VARIANT_DIR.values()
    )


class DrawLog:
    """Per-kind FIFO of random draws, in the order the reference consumed them."""

    KINDS = ("randint", "uniform", "poisson", "normal", "rand", "exponential", "choice")

    def __init__(self):
        self.q = {k: [] for k in self.KINDS}
        self.pos = {k: 0 for k in self.KINDS}

    def push(self, kind, value):
        self.q[kind].append(value)

    def pop(self, kind):
        i = self.pos[kind]
        if i >= len(self.q[kind]):
            raise IndexError(f"draw log exhausted for kind={kind!r}")
        self.pos[kind] = i + 1
        return self.q[kind][i]

    def remaining(self, kind):
        return len(self.q[kind]) - self.pos[kind]

    def rewind(self):
        self.pos = {k: 0 for k in self.KINDS}


class _RandomProxy:
    """Stands in for `numpy.random` inside a loaded reference module."""

    def __init__(self, log: DrawLog, mode: str, rs=None):
        assert mode in ("record", "replay")
        self._log, self._mode = log, mode
        self._rs = rs if rs is not None else _np.random  # RandomState or the global module

    def _draw(self, kind, fn):
        if self._mode == "record":
            v = fn()
            self._log.push(kind, _np.array(v).copy() if isinstance(v, _np.ndarray) else v)
            return v
        return self._log.pop(kind)

    def seed(self, s=None):
        if self._mode == "record":
            self._rs.seed(s)

    def randint(self, low, high=None, size=None):
        return self._draw("randint", lambda: self._rs.randint(low, high, size))

    def uniform(self, low=0.0, high=1.0, size=None):
        return self._draw("uniform", lambda: self._rs.uniform(low, high, size))

    def poisson(self, lam=1.0, size=None):
        return self._draw("poisson", lambda: self._rs.poisson(lam, size))

    def normal(self, loc=0.0, scale=1.0, size=None):
        if self._mode == "record":
            v = self._rs.normal(loc, scale, size)
            # log the *standardised* value so replays are independent of loc/scale
            self._log.push("normal", (_np.asarray(v, dtype=float) - loc) / scale)
            return v
        z = self._log.pop("normal")
        z = _np.asarray(z, dtype=float) * scale + loc
        return z if size is not None else float(z)

    def rand(self, *shape):
        return self._draw("rand", lambda: self._rs.rand(*shape))

    def exponential(self, scale=1.0, size=None):
        if self._mode == "record":
            v = self._rs.exponential(scale, size)
            self._log.push("exponential", _np.asarray(v, dtype=float) / scale)
            return v
        z = _np.asarray(self._log.pop("exponential"), dtype=float) * scale
        return z if size is not None else float(z)


class _NumpyProxy:
    """Forwards everything to numpy except `.random`."""

    def __init__(self, random_proxy):
        object.__setattr__(self, "random", random_proxy)

    def __getattr__(self, name):
        return getattr(_np, name)


class _PyRandomProxy:
    """Stands in for python's `random` module (only `choice` is used by the reference)."""

    def __init__(self, log: DrawLog, mode: str):
        self._log, self._mode = log, mode

    def choice(self, seq):
        if self._mode == "record":
            v = _pyrandom.choice(seq)
            self._log.push("choice", v)
            return v
        return self._log.pop("choice")

    def __getattr__(self, name):
        return getattr(_pyrandom, name)


_LOAD_COUNTER = [0]


def load_reference_module(variant: str, log: DrawLog | None = None, mode: str = "record", rs=None):
    """Load `<variant>/Environment.py` from the reference tree under a fresh private name.

    With `log` given, the module's `np`/`random` globals are replaced by proxies.
    Loading the SARL module calls `np.random.seed(1234)` on the GLOBAL numpy RNG
    (Simulation-SARL/Environment.py:7); callers that rely on the global stream must
    reseed afterwards.
    """
    if not reference_available():
        raise FileNotFoundError(f"reference tree not found under {REFERENCE_ROOT}")
    path = os.path.join(REFERENCE_ROOT, _VARIANT_DIR[variant], "Environment.py")
    _LOAD_COUNTER[0] += 1
    name = f"_risvec_ref_{variant}_{_LOAD_COUNTER[0]}"
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    dont = sys.dont_write_bytecode
    sys.dont_write_bytecode = True  # the reference tree is read-only
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.dont_write_bytecode = dont
    if log is not None:
        mod.np = _NumpyProxy(_RandomProxy(log, mode, rs))
        mod.random = _PyRandomProxy(log, mode)
    return mod


def make_reference_env(variant: str, n_veh: int, M: int, control_bit: int = 3, log: DrawLog | None = None,
                       mode: str = "record", rs=None):
    mod = load_reference_module(variant, log, mode, rs)
    env = mod.Environ(DOWN_LANES, UP_LANES, LEFT_LANES, RIGHT_LANES, WIDTH, HEIGHT, n_veh, M, control_bit)
    return mod, env


# YAML-effective MARL parameters after the train script's overlay (SURVEY.md section 8a header;
# Simulation-MARL-BCD/marl_train_bcd.py:547-779 applied to config.yaml).
MARL_YAML_PARAMS = dict(
    rate=1.0, f_local_max=3e9, f_edge_max=2e9, cycles_per_bit=300.0, k=1e-28, cpu_share_floor=0.10,
    P_max=2.0, bandwidth=5.0, power_scale=0.7, w_d=1.0, w_e=1.0, qos_enable=True, R_min_bpsHz=0.15,
    D_max_s=0.12, qos_penalty=1.5, reward_clip=50.0,
)


def apply_marl_yaml_params(env, params=None):
    """Attribute overlay exactly as the reference train script performs it."""
    p = dict(MARL_YAML_PARAMS if params is None else params)
    for k, v in p.items():
        setattr(env, k, v)
    env.bandwidth_hz = env.bandwidth * 1e6
    env.noise_power = env.N0_W_per_Hz * env.bandwidth_hz  # marl_train_bcd.py:589-590
    return env
