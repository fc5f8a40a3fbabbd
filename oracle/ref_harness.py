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

def _find_reference_root() -> str:
    """RISVEC_REFERENCE_ROOT, else the mounted tree (build container), else the git-ignored copy that
    `tools/stage_reference.py` leaves under baseline/_ref (it travels to the GPU box with the snapshot)."""
    env = os.environ.get("RISVEC_REFERENCE_ROOT")
    if env:
        return env
    staged = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "baseline", "_ref")
    for cand in ("/root/reference", staged):
        if os.path.isfile(os.path.join(cand, "Simulation-MARL-BCD", "Environment.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()
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


# ---------------------------------------------------------------------------------------------
# NOMA pairing stage of the MARL driver (module-level code of `marl_train_bcd.py`)
# ---------------------------------------------------------------------------------------------
PAIRING_HELPERS = ("_anneal_topk", "_build_feasible_mask_from_delta_g", "_score_matrix_from_gain_and_history",
                   "_relax_mask_once", "_mwm_completion", "_mwm_primary", "_adaptive_threshold_from_delta_g",
                   "_qos_pair_feasible")


class _StableArgsortNumpy:
    """numpy with `argsort` forced to `kind='stable'` (see the header of oracle/pairing_oracle.py: at
    exact ties the default kind depends on the SIMD kernel numpy picks for the host CPU)."""

    def __getattr__(self, name):
        return getattr(_np, name)

    @staticmethod
    def argsort(a, axis=-1, kind=None, order=None):
        return _np.argsort(a, axis=axis, kind="stable", order=order)


class PairingReference:
    """The reference's pairing helpers and the driver's own call-site statements, compiled from the
    unmodified `Simulation-MARL-BCD/marl_train_bcd.py` source (selected by AST, nothing is copied).

    `helpers[name]` are the module-level functions; `step(...)` executes, in one persistent namespace,
    the statements the driver runs per step: the mask block (`if config.mask_enable:` holding
    `need_recalc`) and everything from `gamma_decay = ...` to the `unpaired_streak` update."""

    def __init__(self, n_veh: int, config_overrides: dict | None = None, stable_argsort: bool = True):
        import ast
        import math
        import types
        import typing

        path = os.path.join(REFERENCE_ROOT, "Simulation-MARL-BCD", "marl_train_bcd.py")
        src = open(path, "r", encoding="utf-8").read()
        tree = ast.parse(src)
        npx = _StableArgsortNumpy() if stable_argsort else _np
        cfg = types.SimpleNamespace(
            n_veh=n_veh, mask_enable=True, mask_topk_start=n_veh - 1, mask_topk_end=max(4, n_veh // 2),
            mask_tau_q_start=0.2, mask_tau_q_end=0.4, mask_warmup_episodes=200,
            min_pair_target=max(1, n_veh // 4), use_mwm_primary=True, mwm_allow_singles=True,
            mwm_accept_quantile=0.10, mwm_backoff_rounds=5, mwm_accept_q_step=0.05, qos_enable=True,
            qos_R_min_bpsHz=0.15, pairing_threshold_quantile=0.5)
        for k, v in (config_overrides or {}).items():
            setattr(cfg, k, v)
        self.config = cfg
        ns = {"np": npx, "math": math, "random": _pyrandom, "config": cfg, "Optional": typing.Optional,
              "Set": typing.Set, "Tuple": typing.Tuple, "List": typing.List, "__name__": "ref_pairing"}
        # the *last* definition of each name wins, as at import time in the driver
        fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in PAIRING_HELPERS]
        exec(compile(ast.Module(body=fns, type_ignores=[]), path, "exec"), ns)
        self.helpers = {n: ns[n] for n in PAIRING_HELPERS}

        loop = None
        for node in ast.walk(tree):
            if isinstance(node, ast.For) and isinstance(node.target, ast.Name) and node.target.id == "i_step":
                seg = ast.get_source_segment(src, node) or ""
                if "pair_affinity_hist" in seg and "mwm_backoff_rounds" in seg:
                    loop = node
                    break
        if loop is None:
            raise RuntimeError("pairing call site not found in " + path)
        segs = [(ast.get_source_segment(src, s) or "") for s in loop.body]
        i_mask = next(i for i, t in enumerate(segs) if t.startswith("if config.mask_enable") and "need_recalc" in t)
        i_lo = next(i for i, t in enumerate(segs) if t.startswith("gamma_decay"))
        i_hi = next(i for i, t in enumerate(segs) if t.startswith("for u in range(config.n_veh)")
                    and "unpaired_streak" in t)
        self._mask_code = compile(ast.Module(body=[loop.body[i_mask]], type_ignores=[]), path, "exec")
        self._pair_code = compile(ast.Module(body=loop.body[i_lo:i_hi + 1], type_ignores=[]), path, "exec")
        import torch
        ns.update(torch=torch, agents=[types.SimpleNamespace(policy=types.SimpleNamespace(device="cpu"))],
                  K_STEPS_FOR_RIS_OPTIMIZATION=100)
        self.ns = ns
        self.new_episode(0)

    def new_episode(self, i_episode: int):
        """Per-episode initialisation the driver does at `marl_train_bcd.py:1282-1300`."""
        n = self.config.n_veh
        self.ns.update(
            i_episode=int(i_episode), last_mask_mat=None, last_tau_now=None, last_K_now=None, last_q_now=None,
            last_feasible_mask_gpu=None, pair_affinity_hist=_np.zeros((n, n), dtype=_np.float32),
            prev_pairs=set(), unpaired_streak=_np.zeros((n,), dtype=_np.int32),
            freeze_group_in_episode=True, freeze_recalc_every=0, freeze_unstick_prob=0.0,
            freeze_reward_drop_ratio=0.05, episode_groups=None, last_env_global=None, ep_env_best=-1e18,
            unstick_used_flag=False, ep_mask_zero_ratio_sum=0.0)

    def step(self, i_step: int, gains, p01, env_noise_power: float, env_P_max: float, freeze: bool = True):
        import contextlib
        import io
        import types

        ns = self.ns
        ns.update(i_step=int(i_step), current_channel_gains=_np.asarray(gains, dtype=float),
                  offload_power_for_pairing=_np.asarray(p01, dtype=float), freeze_group_in_episode=bool(freeze),
                  env=types.SimpleNamespace(noise_power=float(env_noise_power), P_max=float(env_P_max)),
                  feasible_mask_gpu=None, mask_mat=None)
        with contextlib.redirect_stdout(io.StringIO()):
            exec(self._mask_code, ns)
            exec(self._pair_code, ns)
        return dict(pairs=[(int(a), int(b)) for a, b in ns["pairs"]],
                    groups=[[int(u) for u in g] for g in ns["noma_groups"]],
                    hist=_np.array(ns["pair_affinity_hist"]), streak=_np.array(ns["unpaired_streak"]),
                    mask=None if ns["mask_mat"] is None else _np.array(ns["mask_mat"]).astype(_np.uint8),
                    tau=ns["last_tau_now"], K=ns["last_K_now"], rounds=int(ns["round_id"]))
