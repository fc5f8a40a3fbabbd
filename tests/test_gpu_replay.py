"""GPU: device replay memory (`risvec_replay_*`) vs the oracle -- pure float32 copies, compared exactly."""
import numpy as np
import pytest

from oracle.replay_oracle import ReplayOracle, assemble_marl_action, assemble_marl_mask

pytestmark = pytest.mark.gpu

FIELDS = ("state_memory", "action_memory", "reward_global_memory", "reward_local_memory", "new_state_memory",
          "terminal_memory", "mask_memory")


def _same(rb, orc):
    assert rb.mem_cntr == orc.mem_cntr
    for name in FIELDS:
        assert np.array_equal(getattr(rb, name).cpu().numpy(), getattr(orc, name)), name


def test_store_and_wraparound_and_sample():
    import torch

    from ris_vec_marl_b200 import ReplayBuffer

    rng = np.random.default_rng(0)
    N, S1, A1, cap = 8, 5, 10, 1000
    rb = ReplayBuffer(cap, S1, A1, N)
    orc = ReplayOracle(cap, S1, A1, N)
    for k, E in enumerate((1, 257, 600, 333)):          # 1191 rows: wraps once
        t = dict(state=rng.normal(size=(E, S1 * N)).astype(np.float32),
                 action=rng.normal(size=(E, A1 * N)).astype(np.float32),
                 reward_g=rng.normal(size=E).astype(np.float32), reward_l=rng.normal(size=(E, N)).astype(np.float32),
                 state_=rng.normal(size=(E, S1 * N)).astype(np.float32))
        done = [rng.random(E) < 0.3, True, False, rng.random(E) < 0.5][k]
        mask = None if k == 2 else (rng.random((E, N * N)) < 0.5).astype(np.float32)
        rb.store_transitions(**{n: torch.as_tensor(v) for n, v in t.items()},
                             done=done if isinstance(done, bool) else torch.as_tensor(done),
                             mask_flat=None if mask is None else torch.as_tensor(mask))
        orc.store_transitions(**t, done=done, mask_flat=mask)
        _same(rb, orc)
    idx = rng.integers(0, cap, 512)
    got = rb.sample_buffer(512, idx=torch.as_tensor(idx))
    for a, b in zip(got, orc.sample(idx)):
        assert np.array_equal(a.cpu().numpy(), b)
    s = rb.sample_buffer(64)
    assert s[0].shape == (64, S1 * N) and s[5].dtype == torch.bool
    rb.store_transition(t["state"][0], t["action"][0], 1.5, t["reward_l"][0], t["state_"][0], True, np.ones(N * N))
    orc.store_transitions(t["state"][:1], t["action"][:1], [1.5], t["reward_l"][:1], t["state_"][:1], True,
                          np.ones((1, N * N), np.float32))
    _same(rb, orc)
    rb.close()


@pytest.mark.parametrize("N", [8, 5])
def test_store_marl_fused_assembly(N):
    import torch

    from ris_vec_marl_b200 import ReplayBuffer

    rng = np.random.default_rng(N)
    E, cap = 300, 512
    rb = ReplayBuffer(cap, 5, N + 2, N)
    orc = ReplayOracle(cap, 5, N + 2, N)
    for with_mask in (True, False):
        state = rng.normal(size=(E, 5 * N)).astype(np.float32)
        state_ = rng.normal(size=(E, 5 * N)).astype(np.float32)
        probs = rng.random((E, N, N)).astype(np.float32)
        power = rng.uniform(-1, 1, (E, N, 2)).astype(np.float32)
        rg, rl = rng.normal(size=E).astype(np.float32), rng.normal(size=(E, N)).astype(np.float32)
        mask = (rng.random((E, N, N)) < 0.6).astype(np.uint8) if with_mask else None
        rb.store_marl(*(torch.as_tensor(x) for x in (state, probs, power, rg, rl, state_)), done=False,
                      mask_u8=None if mask is None else torch.as_tensor(mask))
        action = np.stack([assemble_marl_action(probs[e], power[e]) for e in range(E)])
        mflat = np.stack([assemble_marl_mask(None if mask is None else mask[e], N) for e in range(E)])
        orc.store_transitions(state, action, rg, rl, state_, False, mflat)
        _same(rb, orc)
    rb.close()
