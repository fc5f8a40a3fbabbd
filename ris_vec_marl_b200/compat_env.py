"""Drop-in `Environ` objects with the reference's duck-typed surface on top of
the CUDA library, so the reference drivers (marl_train_bcd.py, ddpg_train.py, marl_test.py)
can `import Environment` and run unchanged.

One `Environ` here is ONE env instance (E = 1) -- the view the reference scripts expect.  All
arithmetic runs in the sm_100a kernels; the host side only (i) draws randomness from the GLOBAL
numpy stream in the reference's call order and injects it (so a seeded run consumes exactly
the draws the reference would), and (ii) copies the small state arena back after each call.
For thousands of envs use `BatchedEnviron` directly.
"""
from __future__ import annotations

import random as _pyrandom

import numpy as np
import torch

from . import _lib
from .batched import _PARAM_NAMES, BatchedEnviron, encode_groups

_DIR_CHARS = "udlr"
_STAT_ATTR = {"last_" + n: i for i, n in enumerate(_lib.STAT_COLUMNS)}
# reference attribute -> kernel parameter (same name unless listed)
_ALIAS = {"L": "L", "rate": "rate"}


class Vehicle:
    """Read-only view with the reference's fields (MARL/Environment.py:45-53)."""

    def __init__(self, position, direction, velocity):
        self.position = position
        self.direction = direction
        self.velocity = velocity
        self.neighbors = []
        self.destinations = []


class _EnvironBase:
    _variant = "marl"

    def __init__(self, down_lane, up_lane, left_lane, right_lane, width, height, n_veh, M, control_bit):
        d = object.__getattribute__(self, "__dict__")
        d["_ready"] = False
        d["_dirty"] = False
        self.down_lanes, self.up_lanes = down_lane, up_lane
        self.left_lanes, self.right_lanes = left_lane, right_lane
        self.width, self.height = width, height
        self.n_veh, self.M, self.control_bit = n_veh, M, control_bit
        self._b = BatchedEnviron(self._variant, 1, n_veh, M, control_bit,
                                 lanes=(down_lane, up_lane, left_lane, right_lane), width=width, height=height)
        self._host = {}
        self.vehicles = []
        self.possible_angles = np.linspace(0, 2 * np.pi, 2 ** control_bit, endpoint=False)
        # python-only attributes the drivers read / write (MARL/Environment.py:70-143)
        p = self._b.params
        self.bandwidth_hz = p.bandwidth * 1e6
        self.N0_dBm_per_Hz = -174
        self.N0_W_per_Hz = 10 ** ((self.N0_dBm_per_Hz - 30) / 10)
        self.reward_norm_beta = 0.99
        self.sample_weights = True
        self.w_d_range, self.w_e_range = (0.2, 1.0), (2.0, 6.0)
        self.data_r = np.zeros(n_veh)
        d["_ready"] = True
        self._pull()

    # ---- attribute plumbing: parameter writes are forwarded lazily to the device struct
    def __setattr__(self, name, value):
        d = object.__getattribute__(self, "__dict__")
        if d.get("_ready") and name in _PARAM_NAMES and not name.endswith("_lanes") and name not in ("width", "height"):
            self._b._apply({name: value})
            d["_dirty"] = True
            return
        d[name] = value

    def __getattr__(self, name):
        d = object.__getattribute__(self, "__dict__")
        if name.startswith("__"):
            raise AttributeError(name)
        if name in _PARAM_NAMES and "_b" in d:
            v = getattr(d["_b"].params, name)
            if name == "channel_model":
                return {v: k for k, v in _lib.CHANNEL.items()}.get(v, "unknown")
            if name == "qos_enable":
                return bool(v)
            return v
        host = d.get("_host", {})
        if name in host:
            return host[name]
        if name in _STAT_ATTR and "stats" in host:
            return float(host["stats"][_STAT_ATTR[name]])
        raise AttributeError(name)

    def _flush(self):
        if self._dirty:
            self._b.set_params()
            object.__getattribute__(self, "__dict__")["_dirty"] = False

    def _pull(self):
        """One pass over the (tiny) state views -> host numpy, reference dtypes (float64)."""
        b, h = self._b, {}
        torch.cuda.current_stream(b.device).synchronize()
        f64 = lambda t: t.detach().cpu().numpy().astype(np.float64)
        for ref_name, field in (("DataBuf", "DataBuf"), ("data_t", "data_t"), ("data_p", "data_p"),
                                ("over_data", "over_data"), ("vehicle_rate", "vehicle_rate"),
                                ("channel_gains", "gains"), ("distances_R_i", "dist"), ("angles_R_i", "angle"),
                                ("elements_phase_shift_real", "phase_real"), ("_over_power", "over_power"),
                                ("_reward_user", "reward_user")):
            h[ref_name] = f64(b.state(field))[0]
        h["elements_phase_shift_complex"] = (f64(b.theta_re) + 1j * f64(b.theta_im))[0]
        h["mec_queue_cycles"] = float(f64(b.mec_queue_cycles)[0])
        h["last_mec_queue_cycles"] = h["mec_queue_cycles"]
        h["stats"] = f64(b.stats)[0]
        h["last_power_W"] = f64(b.last_power_W)[0]
        h["_reward"] = float(f64(b.reward)[0])
        px, py = f64(b.pos_x)[0], f64(b.pos_y)[0]
        dirs, vel = b.dir.cpu().numpy()[0], b.vel.cpu().numpy()[0]
        object.__getattribute__(self, "__dict__")["_host"] = h
        if self.vehicles:
            for i, v in enumerate(self.vehicles):
                v.position = [px[i], py[i]]
                v.direction = _DIR_CHARS[int(dirs[i])]
                v.velocity = int(vel[i])
        return px, py, dirs, vel

    # ---- reference methods ------------------------------------------------------------
    def make_new_game(self):
        """MARL/Environment.py:733-737: draws come from the global numpy stream in the
        reference's order and are injected into the reset kernel."""
        self._flush()
        V, ints, dirs = self.n_veh, [], []
        for _ in range(int(V / 4)):
            ints.append(np.random.randint(0, len(self.down_lanes)))
            for lo, hi in ((220, 230), (10, 15), (170, 180), (10, 15), (220, 230), (10, 15), (170, 180), (10, 15)):
                ints.append(np.random.randint(lo, hi))
        for _ in range(int(V % 4)):
            ints.append(np.random.randint(0, len(self.down_lanes)))
            dirs.append(_DIR_CHARS.index(_pyrandom.choice("dulr")))
            ints.append(np.random.randint(0, self.height))
            ints.append(np.random.randint(15, 20))
        self.V2I_Shadowing = np.random.normal(0, 8, V)  # (:409) read only by get_shadowing
        ints.append(np.random.randint(5, int(self._b.params.data_buf_size) - 1))
        self._b.make_new_game(np.asarray(ints, dtype=np.int32)[None],
                              np.asarray(dirs, dtype=np.int32)[None] if dirs else None)
        self._b.V2I_Shadowing.copy_(torch.as_tensor(self.V2I_Shadowing[None], device=self._b.device))
        self.vehicles = [Vehicle([0.0, 0.0], "d", 0) for _ in range(V)]
        self._pull()

    def renew_positions(self):
        """MARL/Environment.py:412-542.  The number of uniforms consumed is data dependent:
        offer 8 per vehicle, then rewind the numpy stream and advance it by the count used."""
        self._flush()
        state = np.random.get_state()
        n = 8 * self.n_veh
        u = np.random.uniform(0, 1, n)
        used = int(self._b.renew_positions(u[None]).cpu()[0])
        np.random.set_state(state)
        for _ in range(used):
            np.random.uniform(0, 1)
        self._pull()

    def compute_parms(self):
        self._flush()
        self._b.compute_parms()
        self._pull()

    def get_next_phase(self, action_phase):
        self._flush()
        self._b.get_next_phase(np.asarray(action_phase, dtype=np.float32)[None])
        self._pull()

    def optimize_phase_shift(self):
        self._flush()
        self._b.optimize_phase_shift()
        self._pull()

    def update_channel_gains(self):
        self._flush()
        b = self._b
        if b.params.channel_model == 0:
            b.update_channel_gains()
        else:  # per vehicle: rand, normal, then exponential or two normals (MARL:296-327)
            V = self.n_veh
            r, nrm, ex = np.zeros((1, V)), np.zeros((1, V, 3)), np.zeros((1, V))
            for i in range(V):
                r[0, i] = np.random.rand()
                nrm[0, i, 0] = np.random.normal(0.0, 1.0)
                if b.params.rician_K_dB <= 1e-6:
                    ex[0, i] = np.random.exponential(1.0)
                else:
                    nrm[0, i, 1] = np.random.normal(0.0, 1.0)
                    nrm[0, i, 2] = np.random.normal(0.0, 1.0)
            b.update_channel_gains(r, nrm, ex)
        self._pull()

    def get_channel_gains(self):
        return self._host["channel_gains"]

    def Random_phase(self):
        """MARL/Environment.py:203-206: M picks of python's `random.choice(possible_angles)`."""
        n = 2 ** self.control_bit
        idx = np.array([_pyrandom.choice(range(n)) for _ in range(self.M)], dtype=np.int32)
        self._flush()
        self._b.Random_phase(idx[None])
        self._pull()

    def get_path_loss(self, position_A):
        """MARL/Environment.py:192-196 (never called by the drivers).  The device computes it for the
        env's own vehicles; an arbitrary position is evaluated against them first."""
        pl = self._b.get_path_loss()[0].cpu().numpy()
        for i, veh in enumerate(self.vehicles):
            if veh.position[0] == position_A[0] and veh.position[1] == position_A[1]:
                return pl[i]
        raise ValueError("get_path_loss is only available for the positions of the env's vehicles")

    def get_shadowing(self, delta_distance, vehicle):
        """MARL/Environment.py:198-201 for `delta_distance = velocity * time_slow` (the only value the
        reference ever forms, :410); consumes one `np.random.normal(0, 8, 1)` like the reference."""
        nz = np.zeros((1, self.n_veh))
        nz[0, vehicle] = np.random.normal(0, 8, 1)[0]
        return self._b.get_shadowing(nz)[0, vehicle].cpu().numpy().reshape(1)

    def _arrivals(self):
        lam = float(self._b.params.rate)
        a = np.array([np.random.poisson(lam) for _ in range(self.n_veh)], dtype=np.int32)
        self.data_r = a.astype(np.float64)
        return a[None]


class MarlEnviron(_EnvironBase):
    """Simulation-MARL-BCD/Environment.py:56 `Environ`."""

    _variant = "marl"

    def step(self, action_power, noma_groups=None):
        """MARL/Environment.py:547-731.  `noma_groups=None` (the stale one-argument callers,
        marl_test.py:192) schedules every user as a singleton."""
        self._flush()
        V = self.n_veh
        if noma_groups is None:
            noma_groups = [[i] for i in range(V)]
        partner, ng = encode_groups(noma_groups, V)
        a = np.asarray(action_power, dtype=np.float32).reshape(1, 2, V)
        self._b.step_marl(a, partner[None], np.array([ng], dtype=np.int32), self._arrivals())
        self._pull()
        h = self._host
        return (h["_reward_user"], h["_reward"], h["DataBuf"], h["data_t"], h["data_p"], h["_over_power"],
                h["over_data"])


class SarlEnviron(_EnvironBase):
    """Simulation-SARL/Environment.py:36 `Environ`."""

    _variant = "sarl"

    def step(self, action_power, action_phase):
        """SARL/Environment.py:321-359."""
        self._flush()
        V = self.n_veh
        a = np.asarray(action_power, dtype=np.float32).reshape(1, 2, V)
        ph = np.asarray(action_phase, dtype=np.float32).reshape(1, self.M)
        self._b.step_sarl(a, ph, self._arrivals())
        self._pull()
        h = self._host
        self.Reward = h["_reward"]
        return (h["_reward"], h["DataBuf"], h["data_t"], h["data_p"], h["_over_power"], h["over_data"])
