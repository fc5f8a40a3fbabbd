"""`global_critic.Global_Critic` for the stale `Simulation-MARL-BCD/marl_test.py`.

`marl_test.py:8,107-110` builds the object with 12 positional arguments and calls `load_models()`;
the test loop never evaluates it.  The twin critics have the layout of the shipped checkpoints
`global_critic{1,2}_ddpg`: fc1 (n_agents * obs -> 1024), fc2, fc3, LayerNorms `bn1..3`,
`action_value` (n_agents * n_actions -> 512) and `q` (256 -> 1).
"""
import os

import torch
import torch.nn as nn

from ddpg_torch import model_dir


class GlobalCriticNetwork(nn.Module):
    def __init__(self, state_dims, action_dims, fc1_dims, fc2_dims, fc3_dims):
        super().__init__()
        self.fc1 = nn.Linear(state_dims, fc1_dims)
        self.fc2 = nn.Linear(fc1_dims, fc2_dims)
        self.fc3 = nn.Linear(fc2_dims, fc3_dims)
        self.bn1, self.bn2, self.bn3 = nn.LayerNorm(fc1_dims), nn.LayerNorm(fc2_dims), nn.LayerNorm(fc3_dims)
        self.action_value = nn.Linear(action_dims, fc2_dims)
        self.q = nn.Linear(fc3_dims, 1)

    def forward(self, state, action):
        x = torch.relu(self.bn1(self.fc1(state)))
        x = self.bn2(self.fc2(x))
        x = torch.relu(x + self.action_value(action))
        x = torch.relu(self.bn3(self.fc3(x)))
        return self.q(x)


class Global_Critic:
    def __init__(self, beta, input_dims, tau, n_actions, gamma, C_fc1_dims, C_fc2_dims, C_fc3_dims, batch_size,
                 n_agents, update_actor_interval, noise):
        self.device = torch.device("cuda:0" if torch.cuda.is_available() else "cpu")
        mk = lambda: GlobalCriticNetwork(int(input_dims) * int(n_agents), int(n_actions) * int(n_agents),  # noqa: E731
                                         int(C_fc1_dims), int(C_fc2_dims), int(C_fc3_dims)).to(self.device)
        self.global_critic1, self.global_critic2 = mk(), mk()

    def load_models(self):
        for name, net in (("global_critic1_ddpg", self.global_critic1), ("global_critic2_ddpg", self.global_critic2)):
            # the checkpoints ship with the (untrusted) reference tree: tensors only, never unpickle code
            net.load_state_dict(torch.load(os.path.join(model_dir(), name), map_location=self.device,
                                           weights_only=True))
            net.eval()
