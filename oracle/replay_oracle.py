"""CPU oracle of the MARL driver's replay memory.  TEST INFRASTRUCTURE ONLY (see oracle/env_oracle.py).

Restates `ReplayBuffer` (`Simulation-MARL-BCD/buffer.py:3-39`) batched over E transitions, and the
driver's transition assembly (`marl_train_bcd.py:1390,1776-1799`).  Pinned by
`tests/test_replay_oracle.py` against the reference class loaded from the unmodified file.
"""
from __future__ import annotations

import numpy as np


class ReplayOracle:
    def __init__(self, max_size, input_shape, n_actions, n_agents):      # buffer.py:4-14
        self.mem_size, self.mem_cntr, self.n_agents = int(max_size), 0, int(n_agents)
        self.state_memory = np.zeros((self.mem_size, input_shape * n_agents), dtype=np.float32)
        self.action_memory = np.zeros((self.mem_size, n_actions * n_agents), dtype=np.float32)
        self.reward_global_memory = np.zeros(self.mem_size, dtype=np.float32)
        self.reward_local_memory = np.zeros((self.mem_size, n_agents), dtype=np.float32)
        self.new_state_memory = np.zeros((self.mem_size, input_shape * n_agents), dtype=np.float32)
        self.terminal_memory = np.zeros(self.mem_size, dtype=bool)
        self.mask_memory = np.zeros((self.mem_size, n_agents * n_agents), dtype=np.float32)

    def store_transitions(self, state, action, reward_g, reward_l, state_, done, mask_flat=None):
        """buffer.py:16-25 applied to rows e = 0..E-1 in order."""
        E = len(reward_g)
        done = np.broadcast_to(np.asarray(done, dtype=bool), (E,))
        for e in range(E):
            i = self.mem_cntr % self.mem_size
            self.state_memory[i] = state[e]
            self.action_memory[i] = action[e]
            self.reward_global_memory[i] = reward_g[e]
            self.reward_local_memory[i] = reward_l[e]
            self.new_state_memory[i] = state_[e]
            self.terminal_memory[i] = done[e]
            self.mask_memory[i] = 1.0 if mask_flat is None else mask_flat[e]
            self.mem_cntr += 1

    def sample(self, idx):                                               # buffer.py:30-38
        return (self.state_memory[idx], self.action_memory[idx], self.reward_global_memory[idx],
                self.reward_local_memory[idx], self.new_state_memory[idx], self.terminal_memory[idx],
                self.mask_memory[idx])


def assemble_marl_action(intent_probs, power_raw):
    """marl_train_bcd.py:1389-1390,1776-1783 for one env: [N,N] probs (diagonal zeroed), [N,2] power ->
    flat [N*(N+2)] float32 = per agent [probs[i] | power[i]]."""
    probs = np.array(intent_probs, dtype=np.float64)
    np.fill_diagonal(probs, 0)
    probs = probs.astype(np.float32)
    power = np.asarray(power_raw, dtype=np.float32)
    return np.concatenate([np.concatenate([probs[i], power[i]]) for i in range(probs.shape[0])]).astype(np.float32)


def assemble_marl_mask(mask_mat, n):
    """marl_train_bcd.py:1786-1789."""
    if mask_mat is None:
        return np.ones((n, n), dtype=np.float32).reshape(-1)
    return np.asarray(mask_mat).astype(np.float32).reshape(-1)
