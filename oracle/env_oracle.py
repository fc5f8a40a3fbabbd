"""TEST INFRASTRUCTURE ONLY -- the product path must never import this file.

CPU oracle: a float64 numpy restatement of the reference environment hot path
(`Environ` of Simulation-MARL-BCD/Environment.py and Simulation-SARL/Environment.py),
batched over E independent env instances (all arrays are `[E, V]` / `[E, M]`; E = 1
reproduces the reference object one to one).  Every method cites the reference lines it
follows (paths relative to the reference root; MARL = Simulation-MARL-BCD, SARL =
Simulation-SARL).

Pinning: `tests/test_oracle_vs_reference.py` checks this file against the unmodified
reference driven through `oracle/ref_harness.py` (container only) and
`tests/test_oracle_golden.py` checks it against the committed fixtures in
`tests/golden/` that were generated from the unmodified reference by
`tests/golden/make_golden.py` (runs anywhere).  The reference ships no tests or golden
vectors of its own, so those fixtures are the pin.

Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline / `--impl reference` legs
of `bench.py` may use it.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

# ----------------------------------------------------------------------------------------
# Module constants of the reference (MARL/Environment.py:29-42 == SARL/Environment.py:11-24)
# ----------------------------------------------------------------------------------------
RIS_XYZ = (220.0, 220.0, 25.0)
BS_XYZ = (0.0, 0.0, 25.0)
VEH_HEIGHT = 1.5
RO = 10 ** -2
LAMB = 1
ELEM_SPACING = 0.5
SIGMA = 10 ** (-7)
ALPHA1 = 2.2
ALPHA2 = 2.5

# direction codes shared with include/risvec.h (the reference uses the chars 'u','d','l','r')
DIR_UP, DIR_DOWN, DIR_LEFT, DIR_RIGHT = 0, 1, 2, 3
DIR_CHARS = "udlr"

# partner encoding shared with include/risvec.h (batched form of the ragged `noma_groups`)
PARTNER_SINGLE = -1
PARTNER_NONE = -2
PARTNER_SECOND = 1 << 16  # flag: this user is listed second in its pair


@dataclass
class OracleParams:
    """Scalar parameters.  Defaults are the reference CLASS defaults
    (MARL/Environment.py:70-143, 555; SARL/Environment.py:64-83)."""

    # shared
    time_slow: float = 0.1
    time_fast: float = 0.001
    bandwidth: float = 1.0
    k: float = 1e-28
    L: float = 500.0
    rate: float = 3.0
    data_buf_size: int = 10
    # MARL
    N0_W_per_Hz: float = 10 ** ((-174 - 30) / 10)
    noise_power: float = 10 ** ((-174 - 30) / 10) * 1.0e6
    P_max: float = 1.0
    power_scale: float = 0.7
    f_local_max: float = 1.0e9
    f_edge_max: float = 2.0e9
    cycles_per_bit: float = 500.0
    cpu_share_floor: float = 0.10
    w_d: float = 0.5
    w_e: float = 3.0
    qos_enable: bool = True
    R_min_bpsHz: float = 0.20
    D_max_s: float = 0.10
    qos_penalty: float = 5.0
    reward_clip: float = 50.0
    channel_model: str = "free"
    fc_GHz: float = 3.5
    shadow_std_los: float = 4.0
    shadow_std_nlos: float = 7.0
    rician_K_dB: float = 0.0
    veh_ant_gain: float = 3.0
    # SARL
    t_factor1: float = 1.0
    t_factor2: float = 0.6
    penalty1: float = 2.0
    penalty2: float = 2.0

    @staticmethod
    def marl_yaml() -> "OracleParams":
        """Values in force after marl_train_bcd.py:547-779 overlays config.yaml."""
        p = OracleParams(
            rate=1.0, f_local_max=3e9, f_edge_max=2e9, cycles_per_bit=300.0, k=1e-28, cpu_share_floor=0.10,
            P_max=2.0, bandwidth=5.0, power_scale=0.7, w_d=1.0, w_e=1.0, qos_enable=True, R_min_bpsHz=0.15,
            D_max_s=0.12, qos_penalty=1.5, reward_clip=50.0,
        )
        p.noise_power = p.N0_W_per_Hz * (p.bandwidth * 1e6)
        return p


@dataclass
class Lanes:
    """Lane constants passed to the constructor (MARL/Environment.py:57-64)."""

    down: list
    up: list
    left: list
    right: list
    width: float = 400
    height: float = 400

    @staticmethod
    def default() -> "Lanes":
        up = [i / 2.0 for i in [400 + 3.5 / 2, 400 + 3.5 + 3.5 / 2, 800 + 3.5 / 2, 800 + 3.5 + 3.5 / 2]]
        down = [i / 2.0 for i in [400 - 3.5 - 3.5 / 2, 400 - 3.5 / 2, 800 - 3.5 - 3.5 / 2, 800 - 3.5 / 2]]
        return Lanes(down=down, up=up, left=list(up), right=list(down))


# ----------------------------------------------------------------------------------------
# sources of randomness
# ----------------------------------------------------------------------------------------
class GlobalNumpyDraws:
    """Consumes the global legacy numpy stream in exactly the reference's call order
    (valid for E == 1 only); used for seeded known-answer tests."""

    def __init__(self):
        import random as _r

        self._r = _r

    def randint(self, e, lo, hi):
        return int(np.random.randint(lo, hi))

    def choice_dir(self, e):
        return DIR_CHARS.index(self._r.choice("dulr"))

    def dead_shadowing(self, n):
        np.random.normal(0, 8, n)  # MARL/Environment.py:409 -- consumed, never used

    def uniform(self, e):
        return float(np.random.uniform(0, 1))

    def new_mobility_call(self):
        pass

    def arrivals(self, lam, E, V):
        out = np.zeros((E, V), dtype=np.int64)
        for k in range(V):
            out[0, k] = np.random.poisson(lam)
        return out

    def rand(self, e):
        return float(np.random.rand())

    def std_normal(self, e):
        return float(np.random.normal(0.0, 1.0))

    def std_exponential(self, e):
        return float(np.random.exponential(1.0))


class InjectedDraws:
    """Serves pre-drawn values.  Layouts are the ones the CUDA library accepts:

    reset_ints   int   [E, n]      consumed left to right by make_new_game
    reset_dirs   int   [E, V % 4]  direction codes for the V % 4 extra vehicles
    mob_uniforms f64   [E, n]      one row per renew_positions call (cursor restarts at 0)
    arrivals     int   [T, E, V]   one slab per step
    chan_rand / chan_normal / chan_exp  f64 [E, n]  for the optional 3GPP channel branch
    """

    def __init__(self, reset_ints=None, reset_dirs=None, arrivals=None):
        self.reset_ints = None if reset_ints is None else np.asarray(reset_ints)
        self.reset_dirs = None if reset_dirs is None else np.asarray(reset_dirs)
        self.arrivals_buf = None if arrivals is None else np.asarray(arrivals)
        self._ri = None
        self._rd = None
        self._t = 0
        self.mob_uniforms = None
        self._mu = None
        self.chan = {}
        self._chan_pos = {}
        self.mob_draws_used = None

    # reset ---------------------------------------------------------------------------
    def begin_reset(self):
        E = self.reset_ints.shape[0]
        self._ri = np.zeros(E, dtype=np.int64)
        self._rd = np.zeros(E, dtype=np.int64)

    def randint(self, e, lo, hi):
        if self._ri is None:
            self.begin_reset()
        v = int(self.reset_ints[e, self._ri[e]])
        self._ri[e] += 1
        if not (lo <= v < hi):
            raise ValueError(f"injected reset int {v} outside [{lo},{hi})")
        return v

    def choice_dir(self, e):
        v = int(self.reset_dirs[e, self._rd[e]])
        self._rd[e] += 1
        return v

    def dead_shadowing(self, n):
        pass

    # mobility ------------------------------------------------------------------------
    def set_mobility_uniforms(self, u):
        self.mob_uniforms = np.asarray(u, dtype=np.float64)

    def new_mobility_call(self):
        self._mu = np.zeros(self.mob_uniforms.shape[0], dtype=np.int64)
        self.mob_draws_used = self._mu

    def uniform(self, e):
        v = float(self.mob_uniforms[e, self._mu[e]])
        self._mu[e] += 1
        return v

    # arrivals ------------------------------------------------------------------------
    def set_arrivals(self, a):
        self.arrivals_buf = np.asarray(a)
        self._t = 0

    def arrivals(self, lam, E, V):
        a = self.arrivals_buf[self._t]
        self._t += 1
        assert a.shape == (E, V)
        return a

    # optional 3GPP channel branch ----------------------------------------------------
    def set_channel_draws(self, rand, normal, exp):
        self.chan = {"rand": np.asarray(rand, float), "normal": np.asarray(normal, float),
                     "exp": np.asarray(exp, float)}
        E = self.chan["rand"].shape[0]
        self._chan_pos = {k: np.zeros(E, dtype=np.int64) for k in self.chan}

    def _chan_pop(self, kind, e):
        i = self._chan_pos[kind][e]
        self._chan_pos[kind][e] = i + 1
        return float(self.chan[kind][e, i])

    def rand(self, e):
        return self._chan_pop("rand", e)

    def std_normal(self, e):
        return self._chan_pop("normal", e)

    def std_exponential(self, e):
        return self._chan_pop("exp", e)


def encode_groups(noma_groups, V):
    """Ragged `noma_groups` (list of lists, one env) -> (partner[V] int32, ngroups).

    Mirrors how MARL/Environment.py:339-370 consumes the list: groups of length 1 are OMA
    singletons, length 2 are NOMA pairs (order matters only for the `gain1 > gain2` tie),
    any other length contributes to `len(noma_groups)` but yields rate 0, as do users that
    appear in no group.  Groups must be disjoint.
    """
    partner = np.full(V, PARTNER_NONE, dtype=np.int32)
    for g in noma_groups:
        if len(g) == 1:
            partner[g[0]] = PARTNER_SINGLE
        elif len(g) == 2:
            partner[g[0]] = g[1]
            partner[g[1]] = g[0] | PARTNER_SECOND
    return partner, len(noma_groups)


def _rowsum(a):
    """Sum over the vehicle axis with the summation order of a 1-D `np.sum` / `np.mean`
    (numpy's pairwise routine: sequential below 8 terms, 8 interleaved accumulators
    combined as a tree up to 128 terms), so that E = 1 and E > 1 agree bit for bit."""
    a = np.asarray(a)
    n = a.shape[-1]
    if n < 8:
        res = a[..., 0].copy()
        for i in range(1, n):
            res = res + a[..., i]
        return res
    if n > 128:
        return np.array([row.sum() for row in a.reshape(-1, n)]).reshape(a.shape[:-1])
    r = [a[..., i].copy() for i in range(8)]
    i = 8
    while i < n - (n % 8):
        for j in range(8):
            r[j] = r[j] + a[..., i + j]
        i += 8
    res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
    while i < n:
        res = res + a[..., i]
        i += 1
    return res


def _rowmean(a):
    return _rowsum(a) / a.shape[-1]


class EnvOracle:
    """Batched float64 restatement of the reference `Environ` (both variants)."""

    def __init__(self, variant, n_veh, M, control_bit=3, E=1, params=None, lanes=None, draws=None):
        assert variant in ("marl", "sarl")
        self.variant = variant
        self.V, self.M, self.E = int(n_veh), int(M), int(E)
        self.control_bit = int(control_bit)
        self.p = params if params is not None else OracleParams()
        self.lanes = lanes if lanes is not None else Lanes.default()
        self.draws = draws if draws is not None else GlobalNumpyDraws()
        E, V, M = self.E, self.V, self.M

        # per-vehicle queue / rate state (MARL/Environment.py:84,151-157; SARL :47,71-78)
        self.vehicle_rate = np.zeros((E, V))
        self.DataBuf = np.zeros((E, V))
        self.over_data = np.zeros((E, V))
        self.data_p = np.zeros((E, V))
        self.data_t = np.zeros((E, V))
        self.data_r = np.zeros((E, V))
        self.mec_queue_cycles = np.zeros(E)  # MARL/Environment.py:116

        # vehicles (MARL/Environment.py:45-53,99); n_active = vehicles created so far
        self.n_active = 0
        self.pos = np.zeros((E, V, 2))
        self.dir = np.zeros((E, V), dtype=np.int64)
        self.vel = np.zeros((E, V), dtype=np.int64)

        # RIS (MARL/Environment.py:162-182; SARL :86-104)
        self.phases_R_i = np.zeros((E, V, M), dtype=complex)
        self.distances_R_i = np.zeros((E, V))
        self.angles_R_i = np.zeros((E, V))
        self.possible_angles = np.linspace(0, 2 * math.pi, 2 ** self.control_bit, endpoint=False)
        self.elements_phase_shift_complex = np.zeros((E, M), dtype=complex)
        self.elements_phase_shift_real = np.zeros((E, M))
        self.distance_B_R = math.sqrt(
            (BS_XYZ[0] - RIS_XYZ[0]) ** 2 + (BS_XYZ[1] - RIS_XYZ[1]) ** 2 + (BS_XYZ[2] - RIS_XYZ[2]) ** 2)
        self.angle_B_R = (RIS_XYZ[0] - BS_XYZ[0]) / self.distance_B_R
        self.phase_R = np.zeros(M, dtype=complex)
        import cmath

        for m in range(M):  # MARL/Environment.py:178-179
            self.phase_R[m] = cmath.exp(2 * (math.pi / LAMB) * ELEM_SPACING * self.angle_B_R * m * 1j)
        self.channel_gains = np.zeros((E, V))  # MARL/Environment.py:183
        self.last = {}

    # ------------------------------------------------------------------ reset (row a2)
    def make_new_game(self):
        """MARL/Environment.py:733-737 + :381-410 (SARL :361-365 + :176-205)."""
        E, V, ln = self.E, self.V, self.lanes
        if hasattr(self.draws, "begin_reset"):
            self.draws.begin_reset()
        self.n_active = 0
        for e in range(E):
            slot = 0
            for _ in range(int(V / 4)):
                ind = self.draws.randint(e, 0, len(ln.down))
                x_down = ln.down[ind] if self.variant == "marl" else ln.down[0]  # SARL :181
                y = self.draws.randint(e, 220, 230)
                self._place(e, slot, x_down, y, DIR_DOWN, self.draws.randint(e, 10, 15)); slot += 1
                y = self.draws.randint(e, 170, 180)
                self._place(e, slot, ln.up[0], y, DIR_UP, self.draws.randint(e, 10, 15)); slot += 1
                x = self.draws.randint(e, 220, 230)
                self._place(e, slot, x, ln.left[0], DIR_LEFT, self.draws.randint(e, 10, 15)); slot += 1
                x = self.draws.randint(e, 170, 180)
                self._place(e, slot, x, ln.right[0], DIR_RIGHT, self.draws.randint(e, 10, 15)); slot += 1
            for _ in range(int(V % 4)):  # MARL/Environment.py:402-407
                ind = self.draws.randint(e, 0, len(ln.down))
                dcode = self.draws.choice_dir(e)
                y = self.draws.randint(e, 0, int(ln.height))
                self._place(e, slot, ln.down[ind], y, dcode, self.draws.randint(e, 15, 20)); slot += 1
            self.draws.dead_shadowing(slot)
            half = self.draws.randint(e, 5, self.p.data_buf_size - 1)
            self.DataBuf[e, :] = half / 2.0 * np.ones(V)  # MARL/Environment.py:737
        self.n_active = V

    def _place(self, e, slot, x, y, dcode, velocity):
        self.pos[e, slot, 0] = x
        self.pos[e, slot, 1] = y
        self.dir[e, slot] = dcode
        self.vel[e, slot] = velocity

    def set_vehicles(self, pos, dirs, vel):
        """Test helper: place vehicles directly (what SURVEY.md KAT-1 does by hand)."""
        self.pos[...] = np.asarray(pos, dtype=float).reshape(self.E, self.V, 2)
        self.dir[...] = np.asarray(dirs).reshape(self.E, self.V)
        self.vel[...] = np.asarray(vel).reshape(self.E, self.V)
        self.n_active = self.V

    # --------------------------------------------------------------- mobility (row a3)
    def renew_positions(self):
        """MARL/Environment.py:412-542 (identical to SARL :207-316)."""
        ln = self.lanes
        self.draws.new_mobility_call()
        # per heading: moving axis, sign of motion, first / second lane family to test,
        # heading after turning into the first / second family, sign applied to the
        # lateral shift for the first / second family and whether the second family ADDS
        # the gap (the reference writes `delta + gap` for u/d -> r, MARL :439-441,465-467)
        plan = {
            DIR_UP: (1, +1, ln.left, DIR_LEFT, ln.right, DIR_RIGHT),
            DIR_DOWN: (1, -1, ln.left, DIR_LEFT, ln.right, DIR_RIGHT),
            DIR_RIGHT: (0, +1, ln.up, DIR_UP, ln.down, DIR_DOWN),
            DIR_LEFT: (0, -1, ln.up, DIR_UP, ln.down, DIR_DOWN),
        }
        for e in range(self.E):
            for i in range(self.n_active):
                delta = int(self.vel[e, i]) * self.p.time_slow
                heading = int(self.dir[e, i])
                axis, sgn, fam1, head1, fam2, head2 = plan[heading]
                c = float(self.pos[e, i, axis])
                o = float(self.pos[e, i, 1 - axis])
                turned = False
                for which, fam, new_head in ((1, fam1, head1), (2, fam2, head2)):
                    for lane in fam:
                        if sgn > 0:
                            hit = (c <= lane) and ((c + delta) >= lane)
                            gap = lane - c
                        else:
                            hit = (c >= lane) and ((c - delta) <= lane)
                            gap = c - lane
                        if hit and self.draws.uniform(e) < 0.4:
                            if axis == 1:  # heading u/d: lateral coordinate is x
                                o = o - (delta - gap) if which == 1 else o + (delta + gap)
                            else:  # heading r/l: lateral coordinate is y
                                o = o + (delta - gap) if which == 1 else o - (delta - gap)
                            c = lane
                            heading = new_head
                            turned = True
                            break
                    if turned:
                        break
                if not turned:
                    c = c + delta if sgn > 0 else c - delta
                xy = [0.0, 0.0]
                xy[axis], xy[1 - axis] = c, o
                x, y = xy
                # leaving the map: MARL/Environment.py:522-540
                if (x < 0) or (y < 0) or (x > ln.width) or (y > ln.height):
                    if heading == DIR_UP:
                        heading, y = DIR_RIGHT, ln.right[-1]
                    elif heading == DIR_DOWN:
                        heading, y = DIR_LEFT, ln.left[0]
                    elif heading == DIR_LEFT:
                        heading, x = DIR_UP, ln.up[0]
                    elif heading == DIR_RIGHT:
                        heading, x = DIR_DOWN, ln.down[-1]
                self.pos[e, i, 0], self.pos[e, i, 1] = x, y
                self.dir[e, i] = heading

    # --------------------------------------------------------------- geometry (row a4)
    def compute_parms(self):
        """MARL/Environment.py:241-253 (== SARL :134-145)."""
        n = self.n_active
        if n == 0:
            return
        dx = self.pos[:, :n, 0] - RIS_XYZ[0]
        dy = self.pos[:, :n, 1] - RIS_XYZ[1]
        dz = VEH_HEIGHT - RIS_XYZ[2]
        d = np.sqrt(dx * dx + dy * dy + dz ** 2)
        self.distances_R_i[:, :n] = d
        self.angles_R_i[:, :n] = dx / d
        m = np.arange(self.M, dtype=float)
        # argument built in the reference's order: ((-2*(pi/lamb))*d*angle)*m
        arg = ((-2 * (math.pi / LAMB)) * ELEM_SPACING * self.angles_R_i[:, :n, None]) * m[None, None, :]
        self.phases_R_i[:, :n, :] = np.exp(arg * 1j)

    # ------------------------------------------------------------- RIS phases (row a5)
    def get_next_phase(self, action_phase):
        """MARL/Environment.py:233-239, SARL :125-131; `action_phase` is `[E, M]` radians."""
        ph = np.asarray(action_phase, dtype=float).reshape(self.E, self.M)
        self.elements_phase_shift_real = ph
        self.elements_phase_shift_complex = np.exp(ph * 1j)

    def random_phase(self, indices):
        """`Random_phase` (MARL/Environment.py:203-206) with the `random.choice` picks injected:
        `indices` [E, M] into `possible_angles = linspace(0, 2 pi, 2**control_bit, endpoint=False)` (:169)."""
        ph = self.possible_angles[np.asarray(indices, dtype=np.int64).reshape(self.E, self.M)]
        self.elements_phase_shift_real = ph
        self.elements_phase_shift_complex = np.exp(ph * 1j)

    # ------------------------------------------------- direct V2I link (row a9, dead code in the reference)
    def path_loss(self):
        """`get_path_loss(position)` (MARL/Environment.py:192-196) for every vehicle -> [E, V] dB."""
        d1 = np.abs(self.pos[..., 0] - BS_XYZ[0])
        d2 = np.abs(self.pos[..., 1] - BS_XYZ[1])
        dist = np.hypot(d1, d2)
        return 128.1 + 37.6 * np.log10(np.sqrt(dist ** 2 + (BS_XYZ[2] - 1.5) ** 2) / 1000)

    def shadowing(self, v2i_shadowing, normals8):
        """`get_shadowing(delta_distance, vehicle)` (MARL/Environment.py:198-201) for every vehicle with
        `delta_distance = velocity * time_slow` (:410); `normals8` are the N(0, 8) draws."""
        dd = self.vel.astype(float) * self.p.time_slow
        dec = 10  # Decorrelation_distance (:86)
        return (np.multiply(np.exp(-1 * (dd / dec)), np.asarray(v2i_shadowing, float))
                + np.sqrt(1 - np.exp(-2 * (dd / dec))) * np.asarray(normals8, float))

    # ------------------------------------------------------------------- BCD (row a6)
    def _objective(self):
        """MARL/Environment.py:222-231: note that `img` is the sum over ALL vehicles and
        elements, so it is the same number in every iteration of the vehicle loop."""
        prod = (self.elements_phase_shift_complex[:, None, :] * self.phases_R_i) * self.phase_R[None, None, :]
        E = self.E
        img = np.array([np.sum(prod[e]) for e in range(E)])
        total = np.zeros(E)
        for v in range(self.V):
            casc = (RO * img) / (np.sqrt(self.distances_R_i[:, v] ** ALPHA1) * math.sqrt(self.distance_B_R ** ALPHA2))
            total = total + (np.abs(casc) ** 2) / SIGMA ** 2
        return total

    def optimize_phase_shift(self):
        """MARL/Environment.py:208-220: coordinate search over the 2^control_bit angles,
        strictly-greater acceptance starting from best = 0 / best_phase = 0."""
        E = self.E
        import cmath

        cand = [cmath.exp(ph * 1j) for ph in self.possible_angles]  # :213
        for m in range(self.M):
            best = np.zeros(E)
            best_phase = np.zeros(E, dtype=complex)
            for c in cand:
                self.elements_phase_shift_complex[:, m] = c
                x = self._objective()
                better = best < x
                best = np.where(better, x, best)
                best_phase = np.where(better, c, best_phase)
            self.elements_phase_shift_complex[:, m] = best_phase

    # ------------------------------------------------------------ cascaded gain (a7/a8)
    def _cascade_sum(self):
        """sum_m theta_m * phases_R_i[v, m] * phase_R[m], accumulated left to right as the
        reference's python loop does (MARL :266-269, SARL :153-155)."""
        img = np.zeros((self.E, self.V), dtype=complex)
        th = self.elements_phase_shift_complex
        for m in range(self.M):
            img = img + (th[:, m, None] * self.phases_R_i[:, :, m]) * self.phase_R[m]
        return img

    def _cascaded_gain(self):
        img = self._cascade_sum()
        casc = (RO * img) / (np.sqrt(self.distances_R_i ** ALPHA1) * math.sqrt(self.distance_B_R ** ALPHA2))
        return np.abs(casc) ** 2

    def update_channel_gains(self):
        """MARL/Environment.py:255-327."""
        p = self.p
        if p.channel_model == "free":
            self.channel_gains = self._cascaded_gain()
            return
        fc = float(p.fc_GHz)
        for e in range(self.E):
            for i in range(self.V):
                dx = abs(self.pos[e, i, 0] - BS_XYZ[0])
                dy = abs(self.pos[e, i, 1] - BS_XYZ[1])
                dz = abs(BS_XYZ[2] - VEH_HEIGHT)
                d2d = math.hypot(dx, dy)
                d3d = math.sqrt(d2d * d2d + dz * dz)
                los = self.draws.rand(e) < 0.7 * np.exp(-d2d / 200.0)  # :296-299
                dd = max(d3d, 1.0)
                if p.channel_model == "3gpp_umi":  # :279-285
                    pl = (32.4 + 21.0 * np.log10(fc) + 20.0 * np.log10(dd)) if los else \
                        (36.7 + 22.7 * np.log10(fc) + 26.0 * np.log10(dd))
                elif p.channel_model == "3gpp_uma":  # :287-293
                    pl = (28.0 + 22.0 * np.log10(fc) + 20.0 * np.log10(dd)) if los else \
                        (13.54 + 39.08 * np.log10(dd) + 20.0 * np.log10(fc) - 0.6 * p.veh_ant_gain)
                else:
                    pl = 0.0  # :315-317
                large = 10 ** (-pl / 10.0)
                std = p.shadow_std_los if los else p.shadow_std_nlos
                shadow = 10 ** ((self.draws.std_normal(e) * std) / 10.0)  # :8-11
                if p.rician_K_dB <= 1e-6:  # :13-25
                    small = self.draws.std_exponential(e)
                else:
                    K = 10 ** (p.rician_K_dB / 10.0)
                    s = np.sqrt(K / (K + 1.0))
                    sg = 1.0 / np.sqrt(2.0 * (K + 1.0))
                    hr = s + sg * self.draws.std_normal(e)
                    hi = sg * self.draws.std_normal(e)
                    small = hr * hr + hi * hi
                self.channel_gains[e, i] = large * shadow * small

    def get_channel_gains(self):
        return self.channel_gains

    # ---------------------------------------------------------------- NOMA rates (a10)
    def compute_data_rate(self, power_W, partner, ngroups):
        """MARL/Environment.py:331-372 on the batched group encoding (`encode_groups`)."""
        E, V = self.E, self.V
        g = self.channel_gains
        partner = np.asarray(partner).reshape(E, V)
        frac = 1.0 / np.maximum(1, np.asarray(ngroups).reshape(E))  # :341-342
        rates = np.zeros((E, V))
        noise = self.p.noise_power
        for e in range(E):
            for u in range(V):
                code = int(partner[e, u])
                if code == PARTNER_NONE:
                    continue
                if code == PARTNER_SINGLE:  # :344-349
                    sinr = (power_W[e, 0, u] * g[e, u]) / noise
                    rates[e, u] = frac[e] * math.log2(1 + sinr)
                    continue
                second = bool(code & PARTNER_SECOND)
                other = code & (PARTNER_SECOND - 1)
                g1, g2 = (g[e, other], g[e, u]) if second else (g[e, u], g[e, other])
                first_is_near = g1 > g2  # :355 (ties make the SECOND listed user "near")
                i_am_near = (not first_is_near) if second else first_is_near
                if i_am_near:  # :367-369
                    sinr = (power_W[e, 0, u] * g[e, u]) / noise
                else:  # :362-365
                    sinr = (power_W[e, 0, u] * g[e, u]) / (power_W[e, 0, other] * g[e, u] + noise)
                rates[e, u] = frac[e] * math.log2(1 + sinr)
        return rates

    # ------------------------------------------------------------------ step (row a11)
    def step_marl(self, action_power, partner, ngroups):
        """MARL/Environment.py:547-731.  `action_power` is `[E, 2, V]`.
        Returns (per_user_reward[E,V], global_reward[E], over_power[E,V])."""
        p, E, V = self.p, self.E, self.V
        a = np.asarray(action_power, dtype=float).reshape(E, 2, V)
        proj = np.clip(a, 0.0, None) * p.power_scale  # :556
        s = proj[:, 0, :] + proj[:, 1, :]
        over = s > 1.0
        proj = np.where(over[:, None, :], proj / (s[:, None, :] + 1e-12), proj)  # :557-560
        power_W = proj * p.P_max

        self.vehicle_rate = self.compute_data_rate(power_W, partner, ngroups)
        self.data_t = self.vehicle_rate * p.time_fast * p.bandwidth * 1000.0  # :570

        cpu_share = np.clip(a[:, 1, :], 0.0, 1.0)  # :572
        floor = float(p.cpu_share_floor)
        if not np.isfinite(floor):
            floor = 0.10
        floor = max(0.0, min(floor, 0.95))
        cpu_share = np.maximum(cpu_share, floor)
        f_local = cpu_share * p.f_local_max
        Cpb = float(p.cycles_per_bit)

        backlog_kbit = self.DataBuf.copy()
        backlog_cyc = backlog_kbit * 1000.0 * Cpb
        cap = f_local * p.time_fast
        used = np.minimum(cap, backlog_cyc)
        local_done = used / (Cpb * 1000.0)
        self.data_p = local_done  # :592
        remaining = np.maximum(0.0, backlog_kbit - self.data_p)
        off = np.minimum(self.data_t, remaining)
        thr = self.vehicle_rate * p.bandwidth * 1000.0
        t_tx = off / (thr + 1e-12)

        edge_in = off * 1000.0 * Cpb
        q_before = self.mec_queue_cycles.copy()
        edge_in_sum = _rowsum(edge_in)
        q = self.mec_queue_cycles + edge_in_sum  # :606
        served = np.minimum(p.f_edge_max * p.time_fast, q)
        self.mec_queue_cycles = q - served

        self.DataBuf = np.maximum(0.0, self.DataBuf - (self.data_p + off))  # :617-618

        eps = 1e-12
        delay_local = np.maximum(0.0, backlog_cyc - edge_in) / (f_local + eps)
        share = edge_in / (edge_in_sum[:, None] + eps)
        delay_edge_q = share * (q_before / (p.f_edge_max + eps))[:, None]
        delay_edge_c = edge_in / (p.f_edge_max + eps)
        delay = delay_local + t_tx + delay_edge_q + delay_edge_c  # :633

        E_tx = power_W[:, 0, :] * t_tx
        E_loc = p.k * (f_local ** 2) * used
        energy = E_tx + E_loc

        pen = np.zeros((E, V))
        viol = np.zeros((E, V), dtype=bool)
        if p.qos_enable:  # :669-677
            viol = (self.vehicle_rate < float(p.R_min_bpsHz)) | (delay > float(p.D_max_s))
            pen = float(p.qos_penalty) * viol.astype(float)
        cost = float(p.w_d) * delay + float(p.w_e) * energy
        reward = np.clip(-cost - pen, -float(p.reward_clip), float(p.reward_clip))  # :696-703

        self.last = dict(  # :612-614, 636-656, 666, 677, 706-711
            off_kbit_sum=_rowsum(off), local_kbit_sum=_rowsum(local_done), mec_queue_cycles=self.mec_queue_cycles.copy(),
            delay_local_mean=_rowmean(delay_local), delay_edge_q_mean=_rowmean(delay_edge_q),
            delay_edge_c_mean=_rowmean(delay_edge_c), t_tx_mean=_rowmean(t_tx),
            backlog_kbit_mean=_rowmean(backlog_kbit),
            mec_utilization=served / (p.f_edge_max * p.time_fast + 1e-12),
            local_util_mean=_rowmean(used / (cap + 1e-12)),
            power_W=np.stack([E_tx / p.time_fast, E_loc / p.time_fast], axis=1),
            qos_violation=_rowmean(viol.astype(float)),
            delay_mean=_rowmean(delay), energy_mean=_rowmean(energy),
            # extras for tests (not reference attributes)
            delay=delay, energy=energy, edge_in_sum=edge_in_sum, q_before=q_before, off=off,
        )

        arr = self.draws.arrivals(p.rate, E, V)  # :717-719
        self.data_r = np.asarray(arr, dtype=float)
        self.DataBuf = self.DataBuf + self.data_r * p.time_fast * 1000
        global_reward = _rowmean(reward)  # :721
        over_power = np.maximum(0.0, (power_W[:, 0, :] + power_W[:, 1, :]) - p.P_max)  # :727-729
        return reward, global_reward, over_power

    # ------------------------------------------------------------------ step (row a12)
    def step_sarl(self, action_power, action_phase):
        """SARL/Environment.py:321-359 (+ :149-171, :318-319).
        Returns (reward[E], over_power[E,V])."""
        p, E, V = self.p, self.E, self.V
        a = np.asarray(action_power, dtype=float).reshape(E, 2, V)
        self.get_next_phase(action_phase)
        gain = self._cascaded_gain()
        self.vehicle_rate = np.log(1 + a[:, 0, :] * gain / SIGMA ** 2)  # :159
        self.data_t = self.vehicle_rate * p.time_fast * p.bandwidth * 1000
        self.data_p = np.power(a[:, 1, :] / p.k, 1.0 / 3.0) * p.time_fast / p.L / 1000  # :331

        buf = self.DataBuf - (self.data_t + self.data_p)
        buf_signed = buf.copy()  # pre-clamp value: the sign decides the reward penalties (:334-352)
        neg = buf < 0
        b_arg = np.fmax(0, buf + self.data_p)
        rev = np.power(b_arg * 1000 * p.L / p.time_fast, 3.0) * p.k  # :318-319
        over_power = np.where(neg, a[:, 1, :] - rev, 0.0)
        self.over_data = np.where(neg, -buf, 0.0)
        buf = np.where(neg, 0.0, buf)
        self.DataBuf = buf

        base = -(p.t_factor1 * (a[:, 0, :] + a[:, 1, :])) - (p.t_factor2 * buf)
        per_user = np.where(buf > 0, base - p.penalty1, np.where(self.over_data > 2, base - p.penalty2, base))

        arr = self.draws.arrivals(p.rate, E, V)
        self.data_r = np.asarray(arr, dtype=float)
        self.DataBuf = self.DataBuf + self.data_r * p.time_fast * 1000
        reward = _rowmean(per_user)
        self.last = dict(per_user_reward=per_user, gain=gain, buf_signed=buf_signed)
        return reward, over_power
