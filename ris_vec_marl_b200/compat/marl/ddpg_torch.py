"""`ddpg_torch.Agent` for the stale `Simulation-MARL-BCD/marl_test.py`.

The reference tree no longer ships this module; `marl_test.py:5,101-104,114,168` only needs an object
that is built with the 15 positional arguments below, loads the shipped per-agent actor checkpoint
`model2/3-BCD_RIS_marl_ddpg-8/actor_{i}_ddpg` and maps an observation to an action.  The network is
what the checkpoint's state dict describes: fc1 (obs -> 512), LayerNorm `bn1`, fc2 (512 -> 256),
LayerNorm `bn2`, `mu` (256 -> n_actions), tanh; exploration is Gaussian noise of scale `noise`
drawn from the global numpy stream.  Agent code is outside the hot path: plain torch, CPU or GPU.
"""
import os

import numpy as np
import torch
import torch.nn as nn

MODEL_SUBDIR = os.path.join("model2", "3-BCD_RIS_marl_ddpg-8")


def model_dir():
    """Checkpoint directory: $RISVEC_MARL_MODEL_DIR (set by compat/run_driver.py) or ./model2/..."""
    return os.environ.get("RISVEC_MARL_MODEL_DIR") or os.path.join(os.getcwd(), MODEL_SUBDIR)


class ActorNetwork(nn.Module):
    def __init__(self, input_dims, fc1_dims, fc2_dims, n_actions):
        super().__init__()
        self.fc1 = nn.Linear(input_dims, fc1_dims)
        self.fc2 = nn.Linear(fc1_dims, fc2_dims)
        self.bn1 = nn.LayerNorm(fc1_dims)
        self.bn2 = nn.LayerNorm(fc2_dims)
        self.mu = nn.Linear(fc2_dims, n_actions)

    def forward(self, state):
        x = torch.relu(self.bn1(self.fc1(state)))
        x = torch.relu(self.bn2(self.fc2(x)))
        return torch.tanh(self.mu(x))


class Agent:
    def __init__(self, alpha, beta, input_dims, tau, n_actions, gamma, C_fc1_dims, C_fc2_dims, C_fc3_dims,
                 A_fc1_dims, A_fc2_dims, batch_size, n_agents, agent_name, noise):
        self.n_actions, self.agent_name, self.noise = int(n_actions), agent_name, float(noise)
        self.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        self.actor = ActorNetwork(int(input_dims), int(A_fc1_dims), int(A_fc2_dims), self.n_actions).to(self.device)
        self.actor.eval()

    def choose_action(self, observation):
        state = torch.as_tensor(np.asarray(observation, dtype=np.float32), device=self.device).reshape(1, -1)
        with torch.no_grad():
            mu = self.actor(state)[0].cpu().numpy().astype(np.float64)
        return mu + np.random.normal(scale=self.noise, size=self.n_actions)

    def load_models(self):
        path = os.path.join(model_dir(), f"actor_{self.agent_name}_ddpg")
        # the checkpoints ship with the (untrusted) reference tree: tensors only, never unpickle code
        self.actor.load_state_dict(torch.load(path, map_location=self.device, weights_only=True))
        self.actor.eval()
