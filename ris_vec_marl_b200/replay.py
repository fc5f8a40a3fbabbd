"""Device-resident `ReplayBuffer` with the reference's interface (Simulation-MARL-BCD/buffer.py),
batched over envs: `store_transitions` writes E transitions per call through `librisvec.so`."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import REPLAY_FIELDS, check
from .batched import _DevView


class ReplayBuffer:
    """`ReplayBuffer(max_size, input_shape, n_actions, n_agents)` (buffer.py:4).  The seven
    `*_memory` arrays are zero-copy torch views of library-owned device memory; `mem_cntr` counts
    stored transitions as in the reference."""

    def __init__(self, max_size, input_shape, n_actions, n_agents, device=0):
        self._lib = _lib.load_library()
        if not torch.cuda.is_available():
            raise _lib.RisvecLibraryError("no CUDA device: the replay memory has no CPU path")
        self.mem_size, self.n_agents = int(max_size), int(n_agents)
        self.input_shape, self.n_actions = int(input_shape), int(n_actions)
        self.device = torch.device("cuda", int(device))
        h = C.c_void_p()
        check(self._lib.risvec_replay_create(self.device.index, self.mem_size, self.input_shape, self.n_actions,
                                             self.n_agents, C.byref(h)))
        self._h = h
        for i, name in enumerate(REPLAY_FIELDS):
            ptr, rows, cols, eb = C.c_void_p(), C.c_int64(), C.c_int64(), C.c_int()
            check(self._lib.risvec_replay_field(self._h, i, C.byref(ptr), C.byref(rows), C.byref(cols), C.byref(eb)))
            shape = (rows.value,) if cols.value == 1 else (rows.value, cols.value)
            t = torch.as_tensor(_DevView(ptr.value, shape, "|u1" if eb.value == 1 else "<f4"), device=self.device)
            setattr(self, name, t.view(torch.bool) if eb.value == 1 else t)

    @property
    def mem_cntr(self):
        return int(self._lib.risvec_replay_count(self._h))

    def state_dict(self):
        return {**{n: getattr(self, n).clone() for n in REPLAY_FIELDS}, "_mem_cntr": self.mem_cntr}

    def load_state_dict(self, sd):
        for n in REPLAY_FIELDS:
            getattr(self, n).copy_(sd[n].to(self.device))
        check(self._lib.risvec_replay_set_count(self._h, int(sd["_mem_cntr"])))

    @property
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _f(self, x, shape, dtype=torch.float32):
        if x is None:
            return None
        t = torch.as_tensor(x)
        if t.dtype != dtype or t.device != self.device or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype).contiguous()
        if t.numel() != int(torch.Size(shape).numel()):
            raise ValueError(f"expected {tuple(shape)} elements, got {tuple(t.shape)}")
        return t

    @staticmethod
    def _p(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(None)

    def _done(self, done, E):
        if isinstance(done, (bool, int)):
            return None, int(bool(done))
        return self._f(done, (E,), torch.uint8 if torch.as_tensor(done).dtype != torch.bool else torch.bool), 0

    def store_transitions(self, state, action, reward_g, reward_l, state_, done, mask_flat=None):
        """`store_transition` (buffer.py:16-25) for E transitions: row e -> slot (mem_cntr + e) % mem_size."""
        E = int(torch.as_tensor(reward_g).numel())
        N, S, A = self.n_agents, self.input_shape * self.n_agents, self.n_actions * self.n_agents
        d, d_all = self._done(done, E)
        if d is not None and d.dtype == torch.bool:
            d = d.view(torch.uint8)
        args = [self._f(state, (E, S)), self._f(action, (E, A)), self._f(reward_g, (E,)), self._f(reward_l, (E, N)),
                self._f(state_, (E, S))]
        m = self._f(mask_flat, (E, N * N))
        check(self._lib.risvec_replay_store(self._h, E, *[self._p(t) for t in args], self._p(d), d_all, self._p(m),
                                            self._stream))

    def store_transition(self, state, action, reward_g, reward_l, state_, done, mask_flat):
        """Single transition, reference signature."""
        f = lambda x: torch.as_tensor(x, dtype=torch.float32).reshape(1, -1)  # noqa: E731
        self.store_transitions(f(state), f(action), torch.as_tensor([float(reward_g)]), f(reward_l), f(state_),
                               bool(done), f(mask_flat))

    def store_marl(self, state, intent_probs, power_raw, reward_g, reward_l, state_, done, mask_u8=None):
        """The driver's assembly (marl_train_bcd.py:1776-1799) fused with the store: `intent_probs`
        [E,N,N] (diagonal ignored), `power_raw` [E,N,2], `mask_u8` = `env.pair_mask` or None (ones)."""
        E = int(torch.as_tensor(reward_g).numel())
        N, S = self.n_agents, self.input_shape * self.n_agents
        d, d_all = self._done(done, E)
        if d is not None and d.dtype == torch.bool:
            d = d.view(torch.uint8)
        args = [self._f(state, (E, S)), self._f(intent_probs, (E, N, N)), self._f(power_raw, (E, N, 2)),
                self._f(reward_g, (E,)), self._f(reward_l, (E, N)), self._f(state_, (E, S))]
        m = self._f(mask_u8, (E, N * N), torch.uint8)
        check(self._lib.risvec_replay_store_marl(self._h, E, *[self._p(t) for t in args], self._p(d), d_all,
                                                 self._p(m), self._stream))

    def sample_buffer(self, batch_size, idx=None, generator=None):
        """`sample_buffer` (buffer.py:27-39); indices uniform with replacement over the filled part
        (`np.random.choice(max_mem, batch_size)`), drawn on the device unless `idx` is given."""
        B = int(batch_size)
        max_mem = min(self.mem_cntr, self.mem_size)
        if idx is None:
            idx = torch.randint(0, max_mem, (B,), device=self.device, generator=generator)
        idx = self._f(idx, (B,), torch.int64)
        N, S, A = self.n_agents, self.input_shape * self.n_agents, self.n_actions * self.n_agents
        mk = lambda *s, dt=torch.float32: torch.empty(*s, dtype=dt, device=self.device)  # noqa: E731
        out = [mk(B, S), mk(B, A), mk(B), mk(B, N), mk(B, S), mk(B, dt=torch.uint8), mk(B, N * N)]
        check(self._lib.risvec_replay_sample(self._h, B, self._p(idx), *[self._p(t) for t in out], self._stream))
        out[5] = out[5].view(torch.bool)
        return tuple(out)

    def close(self):
        if getattr(self, "_h", None):
            for name in REPLAY_FIELDS:
                if hasattr(self, name):
                    delattr(self, name)
            self._lib.risvec_replay_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
