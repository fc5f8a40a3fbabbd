"""Replay a golden scenario (tests/golden/*.npz) through a backend and collect the same
quantities the fixture holds.  Two backends share this driver: the CPU oracle
(`OracleBackend`, below) and the CUDA library (`tests/gpu_backend.py`)."""
from __future__ import annotations

import os

import numpy as np

from oracle.env_oracle import EnvOracle, InjectedDraws, OracleParams

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MARL_LAST = ("off_kbit_sum", "local_kbit_sum", "mec_queue_cycles", "delay_local_mean", "delay_edge_q_mean",
             "delay_edge_c_mean", "t_tx_mean", "backlog_kbit_mean", "mec_utilization", "local_util_mean",
             "qos_violation", "delay_mean", "energy_mean")


def load_golden(name):
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    meta = g.pop("meta")
    g["variant"] = str(meta[0])
    g["V"], g["M"], g["episodes"], g["T"], g["refresh_every"] = (int(x) for x in meta[1:6])
    g["params"] = str(meta[6])
    g["E"] = g["reset_ints"].shape[0]
    return g


def golden_names():
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("pairing_"))


def oracle_params(g):
    return OracleParams.marl_yaml() if (g["variant"] == "marl" and g["params"] == "yaml") else OracleParams()


class OracleBackend:
    """Adapter: golden inputs -> `EnvOracle` calls."""

    def __init__(self, g):
        self.draws = InjectedDraws(reset_ints=g["reset_ints"], reset_dirs=g["reset_dirs"])
        self.env = EnvOracle(g["variant"], g["V"], g["M"], 3, E=g["E"], params=oracle_params(g), draws=self.draws)

    def make_new_game(self):
        self.env.make_new_game()

    def vehicles(self):
        return self.env.pos.copy(), self.env.dir.copy(), self.env.vel.copy()

    def renew_positions(self, uniforms):
        self.draws.set_mobility_uniforms(uniforms)
        self.env.renew_positions()
        return self.draws.mob_draws_used.copy()

    def compute_parms(self):
        self.env.compute_parms()

    def geometry(self):
        return self.env.distances_R_i.copy(), self.env.angles_R_i.copy()

    def optimize_phase_shift(self):
        self.env.optimize_phase_shift()

    def update_channel_gains(self):
        self.env.update_channel_gains()

    def gains(self):
        return self.env.channel_gains.copy()

    def theta(self):
        return self.env.elements_phase_shift_complex.copy()

    def DataBuf(self):
        return self.env.DataBuf.copy()

    def mec_queue_cycles(self):
        return self.env.mec_queue_cycles.copy()

    def step_marl(self, actions, partner, ngroups, arrivals):
        self.draws.set_arrivals(arrivals[None])
        r_user, r_glob, over_p = self.env.step_marl(actions, partner, ngroups)
        e = self.env
        out = dict(reward_user=r_user, reward=r_glob, DataBuf=e.DataBuf, data_t=e.data_t, data_p=e.data_p,
                   over_power=over_p, rate=e.vehicle_rate, last_power_W=e.last["power_W"])
        for k in MARL_LAST:
            out["last_" + k] = e.last[k]
        # extras (prefix x_) feed the band-exclusion rules of tests/parity.py; not fixture keys
        out.update(x_delay=e.last["delay"], x_edge_in_sum=e.last["edge_in_sum"], x_q_before=e.last["q_before"])
        return out

    def step_sarl(self, actions, phases, arrivals):
        self.draws.set_arrivals(arrivals[None])
        r, over_p = self.env.step_sarl(actions, phases)
        e = self.env
        return dict(reward=r, DataBuf=e.DataBuf, data_t=e.data_t, data_p=e.data_p, over_power=over_p,
                    over_data=e.over_data, rate=e.vehicle_rate, x_buf_signed=e.last["buf_signed"])


def replay(g, backend):
    """Run the scenario; returns {key: array} with the fixture's `step_*` / `ep_*` keys."""
    variant, EP, T = g["variant"], g["episodes"], g["T"]
    out = {}

    def push(name, val):
        out.setdefault(name, []).append(np.array(val).copy())

    backend.make_new_game()
    pos, dirs, vel = backend.vehicles()
    out["reset_pos"], out["reset_dir"], out["reset_vel"] = pos, dirs, vel
    out["reset_DataBuf"] = backend.DataBuf()
    for ep in range(EP):
        if ep % g["refresh_every"] == 0:
            used = backend.renew_positions(g["ep_mob_uniforms"][ep])
            backend.compute_parms()
        else:
            used = np.zeros(g["E"], dtype=np.int64)
        push("ep_mob_used", used)
        pos, dirs, _ = backend.vehicles()
        push("ep_pos", pos); push("ep_dir", dirs)
        d, a = backend.geometry()
        push("ep_dist", d); push("ep_angle", a)
        if variant == "marl":
            backend.optimize_phase_shift()
            backend.update_channel_gains()
            push("ep_gains", backend.gains()); push("ep_theta", backend.theta())
        for t in range(T):
            i = ep * T + t
            if variant == "marl":
                res = backend.step_marl(g["actions"][i], g["ep_partner"][ep], g["ep_ngroups"][ep], g["arrivals"][i])
            else:
                res = backend.step_sarl(g["actions"][i], g["phases"][i], g["arrivals"][i])
            for k, v in res.items():
                push("step_" + k, v)
    res = {k: (np.stack(v, axis=0) if isinstance(v, list) else v) for k, v in out.items()}
    if variant == "marl":
        res["final_mec_queue_cycles"] = backend.mec_queue_cycles()
    return res
