"""CPU-side checks of the C-ABI boundary: the shared library builds for sm_100a, loads, and
exports every symbol include/risvec.h declares; the ctypes mirror of `risvec_params_t`
matches the C layout.  No CUDA call is made."""
import ctypes as C
import os
import re

import pytest

from ris_vec_marl_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build_library()
    return _lib.load_library()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "risvec.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(risvec_[a-z_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/risvec.h but not exported"
        assert n in _lib.EXPORTS, f"{n} has no ctypes prototype"


def test_abi_version_and_default_params(lib):
    assert lib.risvec_abi_version() == 1
    p = _lib.Params()
    assert lib.risvec_default_params(0, C.byref(p)) == 0
    # reference class defaults (MARL/Environment.py:70-143)
    assert (p.bandwidth, p.P_max, p.f_local_max, p.cycles_per_bit, p.w_d, p.w_e) == (1.0, 1.0, 1e9, 500.0, 0.5, 3.0)
    assert (p.R_min_bpsHz, p.D_max_s, p.qos_penalty, p.rate, p.qos_enable) == (0.20, 0.10, 5.0, 3.0, 1)
    assert abs(p.noise_power - 10 ** ((-174 - 30) / 10) * 1e6) < 1e-28
    assert list(p.up_lanes)[:4] == [200.875, 202.625, 400.875, 402.625]
    assert list(p.down_lanes)[:4] == [197.375, 199.125, 397.375, 399.125]
    # last member intact => the python struct has the C layout
    assert (p.t_factor1, p.t_factor2, p.penalty1, p.penalty2) == (1.0, 0.6, 2.0, 2.0)


def test_errors_are_reported_not_thrown(lib):
    h = C.c_void_p()
    rc = lib.risvec_create(None, 0, 4, 8, 40, 3, 0, 1, 0, C.byref(h))
    assert rc == -1 and b"params" in lib.risvec_last_error()
    p = _lib.Params()
    lib.risvec_default_params(0, C.byref(p))
    assert lib.risvec_create(C.byref(p), 0, 4, 64, 40, 3, 0, 1, 0, C.byref(h)) == -2  # V > 32 unsupported
    assert lib.risvec_create(C.byref(p), 7, 4, 8, 40, 3, 0, 1, 0, C.byref(h)) == -1
    assert lib.risvec_destroy(None) == 0


def test_no_cpu_fallback_without_gpu(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ris_vec_marl_b200 import BatchedEnviron, RisvecLibraryError

    with pytest.raises(RisvecLibraryError):
        BatchedEnviron("marl", 4)
    p = _lib.Params()
    lib.risvec_default_params(0, C.byref(p))
    h = C.c_void_p()
    assert lib.risvec_create(C.byref(p), 0, 4, 8, 40, 3, 0, 1, 0, C.byref(h)) == -4  # RISVEC_ERR_NODEVICE


def test_product_path_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under the product package (python or CUDA)
    may import, include or execute it, and the package has no CPU implementation to fall back to."""
    pkg = os.path.join(ROOT, "ris_vec_marl_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8").read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "env_oracle" in text:
                    offenders.append(os.path.join(dirpath, f))
    assert not offenders, offenders


def test_header_and_ctypes_agree_on_record_sizes(lib):
    src = open(os.path.join(ROOT, "include", "risvec.h")).read()
    get = lambda name: int(re.search(rf"#define {name} (\d+)", src).group(1))
    assert get("RISVEC_SARL_OUT_WORDS") == _lib.SARL_OUT_WORDS
    assert get("RISVEC_MARL_IN_WORDS") == _lib.MARL_IN_WORDS
    assert get("RISVEC_MARL_OUT_WORDS") == _lib.MARL_OUT_WORDS
    assert get("RISVEC_ABI_VERSION") == lib.risvec_abi_version()
    n_stat = int(re.search(r"RISVEC_NSTAT = (\d+)", src).group(1))
    assert n_stat == _lib.NSTAT and len(_lib.STAT_COLUMNS) == 13
    fields = re.findall(r"^\s+RISVEC_F_([A-Z_]+)", src, flags=re.M)
    assert len([f for f in fields if f != "COUNT"]) == len(_lib.FIELDS)


def test_pairing_defaults_round_trip_through_the_struct(lib):
    """`risvec_default_pairing` fills every field of `risvec_pairing_t`; reading them back through the
    ctypes mirror checks that both sides agree on the layout (marl_train_bcd.py:435-441,489,1404-1498)."""
    p = _lib.Pairing()
    assert lib.risvec_default_pairing(8, 0, C.byref(p)) == 0
    assert (p.min_pair_target, p.mwm_backoff_rounds, p.relax_topk_step, p.qos_enable) == (2, 5, 1, 1)
    assert (p.mwm_accept_quantile, p.mwm_accept_q_step, p.completion_min_quantile) == (0.10, 0.05, 0.30)
    assert (p.relax_tau_factor_per_round, p.tau_back_floor_db) == (0.95, 3.0)
    assert (p.score_w_delta_db, p.score_w_history, p.qos_soft_penalty_dbscore, p.pair_hist_decay) == (1.0, 0.3, 6.0, 0.97)
    assert p.abs_gain_min_db == float("-inf")
    assert lib.risvec_default_pairing(8, 1, C.byref(p)) == 0
    assert (p.min_pair_target, p.mwm_backoff_rounds, p.abs_gain_min_db) == (3, 3, -120.0)
    assert lib.risvec_default_pairing(3, 0, C.byref(p)) == 0 and p.min_pair_target == 1
    assert lib.risvec_default_pairing(0, 0, C.byref(p)) == -1
    assert lib.risvec_pair_noma(None, C.byref(p), None, 8, 7, 0.2, 1, None, 1, 0, None) == -1
    assert lib.risvec_pair_reset(None, None) == -1
    src = open(os.path.join(ROOT, "include", "risvec.h")).read()
    assert int(re.search(r"#define RISVEC_PAIR_MAX_V (\d+)", src).group(1)) == _lib.PAIR_MAX_V


def test_replay_entry_points_validate_arguments(lib):
    h = C.c_void_p()
    assert lib.risvec_replay_create(0, 0, 5, 10, 8, C.byref(h)) == -1
    assert lib.risvec_replay_create(0, 16, 5, 10, 8, None) == -1
    assert lib.risvec_replay_count(None) == 0
    assert lib.risvec_replay_destroy(None) == 0
    assert lib.risvec_replay_store(None, 1, None, None, None, None, None, None, 0, None, None) == -1
    assert lib.risvec_replay_sample(None, 1, None, None, None, None, None, None, None, None, None) == -1
    import torch

    if not torch.cuda.is_available():
        rc = lib.risvec_replay_create(0, 16, 5, 10, 8, C.byref(h))
        assert rc == -4 and b"no CPU fallback" in lib.risvec_last_error()
