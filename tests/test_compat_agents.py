"""CPU: the agent shims of the stale `marl_test.py` load the reference's shipped MADDPG checkpoints
(skipped where the reference tree is absent)."""
import os
import sys

import numpy as np
import pytest

from oracle import ref_harness as rh

COMPAT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ris_vec_marl_b200", "compat", "marl")
MODELS = os.path.join(rh.REFERENCE_ROOT, "Simulation-MARL-BCD", "model2", "3-BCD_RIS_marl_ddpg-8")


@pytest.mark.skipif(not os.path.isdir(MODELS), reason="reference checkpoints not mounted")
def test_agent_and_global_critic_load_the_shipped_checkpoints(monkeypatch):
    import torch

    monkeypatch.setenv("RISVEC_MARL_MODEL_DIR", MODELS)
    monkeypatch.syspath_prepend(COMPAT)
    for m in ("ddpg_torch", "global_critic"):
        sys.modules.pop(m, None)
    from ddpg_torch import Agent
    from global_critic import Global_Critic

    # argument lists of marl_test.py:101-110
    agents = [Agent(1e-4, 1e-3, 5, 0.005, 2, 0.99, 1024, 512, 256, 512, 256, 64, 8, i, 0.2) for i in range(8)]
    gc = Global_Critic(1e-3, 5, 0.005, 2, 0.99, 1024, 512, 256, 64, 8, 2, 0.2)
    gc.load_models()
    for a in agents:
        a.load_models()
    sd = torch.load(os.path.join(MODELS, "actor_3_ddpg"), map_location="cpu", weights_only=True)
    assert torch.equal(agents[3].actor.mu.weight.detach().cpu(), sd["mu.weight"])
    np.random.seed(0)
    act = agents[3].choose_action([0.4, 0.1, 0.2, 0.0, 0.05])
    assert act.shape == (2,) and np.all(np.isfinite(act)) and np.all(np.abs(act) < 2.0)
    agents[3].noise = 0.0
    a0 = agents[3].choose_action([0.4, 0.1, 0.2, 0.0, 0.05])
    assert np.all(np.abs(a0) <= 1.0)                       # tanh head
    q = gc.global_critic1(torch.zeros(1, 40, device=gc.device), torch.zeros(1, 16, device=gc.device))
    assert q.shape == (1, 1)
    for m in ("ddpg_torch", "global_critic"):
        sys.modules.pop(m, None)
