// Device-resident replay memory of the MARL driver (Simulation-MARL-BCD/buffer.py:3-39) and the
// transition assembly the driver does before `store_transition` (marl_train_bcd.py:1776-1799),
// batched: E envs store E transitions per step into consecutive ring slots.  Pure copies, one block
// per transition row, every segment written coalesced.
#pragma once
#include "common.cuh"

namespace risvec {

struct ReplayMem {
    long long mem_size;
    int S, A, N;     // input_shape * n_agents, n_actions * n_agents, n_agents
    float *state, *action, *reward_g, *reward_l, *state_, *mask;
    unsigned char* terminal;
};

// generic rows: the arguments of buffer.py:16 for E transitions
__global__ void k_replay_store(ReplayMem m, long long first, int E, const float* __restrict__ state,
                               const float* __restrict__ action, const float* __restrict__ reward_g,
                               const float* __restrict__ reward_l, const float* __restrict__ state_,
                               const unsigned char* __restrict__ done, int done_all,
                               const float* __restrict__ mask) {
    const int e = blockIdx.x;
    if (e >= E) return;
    const long long slot = (first + e) % m.mem_size;
    const int NN = m.N * m.N;
    for (int c = threadIdx.x; c < m.S; c += blockDim.x) {
        m.state[slot * m.S + c] = state[(size_t)e * m.S + c];
        m.state_[slot * m.S + c] = state_[(size_t)e * m.S + c];
    }
    for (int c = threadIdx.x; c < m.A; c += blockDim.x) m.action[slot * m.A + c] = action[(size_t)e * m.A + c];
    for (int c = threadIdx.x; c < m.N; c += blockDim.x) m.reward_l[slot * m.N + c] = reward_l[(size_t)e * m.N + c];
    for (int c = threadIdx.x; c < NN; c += blockDim.x) m.mask[slot * NN + c] = mask ? mask[(size_t)e * NN + c] : 1.f;
    if (threadIdx.x == 0) {
        m.reward_g[slot] = reward_g[e];
        m.terminal[slot] = done ? (done[e] != 0) : (done_all != 0);
    }
}

// MARL driver assembly (marl_train_bcd.py:1776-1790): action row = per agent [intent probs (N,
// diagonal zeroed at :1390) | raw power (2)]; mask row = float32 of the u8 feasibility mask, all ones
// when the step built no mask (:1786-1789).
__global__ void k_replay_store_marl(ReplayMem m, long long first, int E, const float* __restrict__ state,
                                    const float* __restrict__ probs, const float* __restrict__ power,
                                    const float* __restrict__ reward_g, const float* __restrict__ reward_l,
                                    const float* __restrict__ state_, const unsigned char* __restrict__ done,
                                    int done_all, const unsigned char* __restrict__ mask) {
    const int e = blockIdx.x;
    if (e >= E) return;
    const long long slot = (first + e) % m.mem_size;
    const int N = m.N, NN = N * N, AW = N + 2;
    for (int c = threadIdx.x; c < m.S; c += blockDim.x) {
        m.state[slot * m.S + c] = state[(size_t)e * m.S + c];
        m.state_[slot * m.S + c] = state_[(size_t)e * m.S + c];
    }
    for (int c = threadIdx.x; c < m.A; c += blockDim.x) {
        const int i = c / AW, k = c - i * AW;
        float v;
        if (k < N) v = (k == i) ? 0.f : probs[((size_t)e * N + i) * N + k];
        else v = power[((size_t)e * N + i) * 2 + (k - N)];
        m.action[slot * m.A + c] = v;
    }
    for (int c = threadIdx.x; c < N; c += blockDim.x) m.reward_l[slot * N + c] = reward_l[(size_t)e * N + c];
    for (int c = threadIdx.x; c < NN; c += blockDim.x) m.mask[slot * NN + c] = mask ? (float)mask[(size_t)e * NN + c] : 1.f;
    if (threadIdx.x == 0) {
        m.reward_g[slot] = reward_g[e];
        m.terminal[slot] = done ? (done[e] != 0) : (done_all != 0);
    }
}

// sample_buffer (buffer.py:27-39) for caller-drawn indices: one block per sampled row
__global__ void k_replay_sample(ReplayMem m, int B, const long long* __restrict__ idx, float* __restrict__ states,
                                float* __restrict__ actions, float* __restrict__ rewards_g,
                                float* __restrict__ rewards_l, float* __restrict__ states_,
                                unsigned char* __restrict__ dones, float* __restrict__ masks) {
    const int b = blockIdx.x;
    if (b >= B) return;
    const long long slot = idx[b];
    const int NN = m.N * m.N;
    for (int c = threadIdx.x; c < m.S; c += blockDim.x) {
        states[(size_t)b * m.S + c] = m.state[slot * m.S + c];
        states_[(size_t)b * m.S + c] = m.state_[slot * m.S + c];
    }
    for (int c = threadIdx.x; c < m.A; c += blockDim.x) actions[(size_t)b * m.A + c] = m.action[slot * m.A + c];
    for (int c = threadIdx.x; c < m.N; c += blockDim.x) rewards_l[(size_t)b * m.N + c] = m.reward_l[slot * m.N + c];
    for (int c = threadIdx.x; c < NN; c += blockDim.x) masks[(size_t)b * NN + c] = m.mask[slot * NN + c];
    if (threadIdx.x == 0) {
        rewards_g[b] = m.reward_g[slot];
        dones[b] = m.terminal[slot];
    }
}

}  // namespace risvec
