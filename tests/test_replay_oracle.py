"""CPU: replay oracle vs the reference `ReplayBuffer` class (loaded from the unmodified file)."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import ref_harness as rh
from oracle.replay_oracle import ReplayOracle, assemble_marl_action, assemble_marl_mask


def _transitions(rng, E, N, S1, A1):
    return dict(state=rng.normal(size=(E, S1 * N)).astype(np.float32), action=rng.normal(size=(E, A1 * N)).astype(np.float32),
                reward_g=rng.normal(size=E).astype(np.float32), reward_l=rng.normal(size=(E, N)).astype(np.float32),
                state_=rng.normal(size=(E, S1 * N)).astype(np.float32), done=rng.random(E) < 0.3,
                mask_flat=(rng.random((E, N * N)) < 0.5).astype(np.float32))


@pytest.mark.skipif(not rh.reference_available(), reason="reference tree not mounted")
def test_oracle_matches_reference_buffer_with_wraparound():
    spec = importlib.util.spec_from_file_location(
        "_ref_buffer", os.path.join(rh.REFERENCE_ROOT, "Simulation-MARL-BCD", "buffer.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    rng = np.random.default_rng(0)
    N, S1, A1, cap = 8, 5, 10, 37
    ref = mod.ReplayBuffer(cap, S1, A1, N)
    orc = ReplayOracle(cap, S1, A1, N)
    for _ in range(5):
        t = _transitions(rng, 11, N, S1, A1)
        for e in range(11):
            ref.store_transition(t["state"][e], t["action"][e], t["reward_g"][e], t["reward_l"][e], t["state_"][e],
                                 t["done"][e], t["mask_flat"][e])
        orc.store_transitions(**t)
    assert ref.mem_cntr == orc.mem_cntr == 55
    for name in ("state_memory", "action_memory", "reward_global_memory", "reward_local_memory", "new_state_memory",
                 "terminal_memory", "mask_memory"):
        assert np.array_equal(getattr(ref, name), getattr(orc, name)), name
    idx = rng.integers(0, cap, 16)
    np.random.seed(3)
    want = ref.sample_buffer(16)
    np.random.seed(3)
    idx = np.random.choice(cap, 16)
    for a, b in zip(orc.sample(idx), want):
        assert np.array_equal(a, b)


def test_marl_assembly_layout():
    rng = np.random.default_rng(1)
    N = 4
    probs, power = rng.random((N, N)), rng.uniform(-1, 1, (N, 2))
    row = assemble_marl_action(probs, power)
    assert row.shape == (N * (N + 2),) and row.dtype == np.float32
    for i in range(N):
        seg = row[i * (N + 2):(i + 1) * (N + 2)]
        assert seg[i] == 0 and np.array_equal(seg[N:], power[i].astype(np.float32))
        assert np.array_equal(np.delete(seg[:N], i), np.delete(probs[i], i).astype(np.float32))
    assert np.array_equal(assemble_marl_mask(None, N), np.ones(N * N, np.float32))
