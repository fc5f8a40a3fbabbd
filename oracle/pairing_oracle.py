"""CPU oracle for the NOMA pairing stage that feeds `Environ.step(action_power, noma_groups)`.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the CPU-baseline legs of the
bench tools may import this module; the product (`ris_vec_marl_b200/`) never does.

This is a plain numpy / pure-python restatement (float64, one env at a time) of the pairing pipeline
that the reference driver runs at module level before every `env.step`
(`Simulation-MARL-BCD/marl_train_bcd.py:1315-1561`) and of the helpers it calls.  Each function cites
the reference lines it follows.

Pinning: `tests/test_pairing_oracle.py` runs the reference's own helper functions (extracted from the
unmodified driver source by AST, see `oracle/ref_harness.py:load_pairing_reference`) and the driver's
call-site statements on the same inputs and requires identical pairs / masks / thresholds; the
golden fixtures `tests/golden/pairing_*.npz` were generated the same way.

One deliberate, documented deviation: at exact ties the reference's `np.argsort` calls
(`:151`, `:271`) depend on which SIMD sort kernel numpy dispatches to on the host CPU (the AVX-512
kernel is not stable, the scalar insertion sort used for n <= 16 is).  The restatement (and the CUDA
kernel) use the stable order (lower index first); the reference is pinned with `argsort` forced to
`kind="stable"`.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np


# --------------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------------
@dataclass
class PairingConfig:
    """Pairing knobs.  Defaults = `Config.__init__` + the `getattr(config, name, default)` fall-backs at
    the call site (`marl_train_bcd.py:435-441,489-503,1404-1418,1481-1498`)."""

    mask_topk_start: int = 7            # n_veh - 1 (:499) -- set by `for_n_veh`
    mask_topk_end: int = 4              # max(4, n_veh // 2) (:500)
    mask_tau_q_start: float = 0.2       # :501
    mask_tau_q_end: float = 0.4         # :502
    mask_warmup_episodes: int = 200     # :503
    min_pair_target: int = 2            # max(1, n_veh // 4) (:489)
    mwm_accept_quantile: float = 0.10   # :439
    mwm_backoff_rounds: int = 5         # :440
    mwm_accept_q_step: float = 0.05     # :441
    completion_min_quantile: float = 0.30   # :282,300
    relax_q_step: float = 0.02          # :1481
    relax_topk_step: int = 1            # :1482
    relax_tau_factor_per_round: float = 0.95    # :1483
    tau_back_floor_db: float = 3.0      # :1498
    score_w_delta_db: float = 1.0       # :1416
    score_w_history: float = 0.3        # :1417
    abs_gain_min_db: float = -math.inf  # :1418
    qos_soft_penalty_dbscore: float = 6.0   # :1449
    qos_enable: bool = True             # :1427
    pair_hist_decay: float = 0.97       # :1404

    @staticmethod
    def for_n_veh(n_veh: int, **kw) -> "PairingConfig":
        c = PairingConfig(mask_topk_start=n_veh - 1, mask_topk_end=max(4, n_veh // 2),
                          min_pair_target=max(1, n_veh // 4))
        for k, v in kw.items():
            if not hasattr(c, k):
                raise AttributeError(k)
            setattr(c, k, v)
        return c

    @staticmethod
    def marl_yaml(n_veh: int = 8) -> "PairingConfig":
        """Effective values after the driver's YAML overlay (`marl_train_bcd.py:639-660,716-732`,
        `config.yaml`)."""
        return PairingConfig.for_n_veh(
            n_veh, mask_topk_start=7, mask_topk_end=7, mask_tau_q_start=0.10, mask_tau_q_end=0.25,
            min_pair_target=3, mwm_accept_quantile=0.10, mwm_backoff_rounds=3, mwm_accept_q_step=0.05,
            abs_gain_min_db=-120.0)


def anneal_topk(i_ep: int, n_agents: int, k_start: int, k_end: int, t_ep: int) -> int:
    """`_anneal_topk` (`marl_train_bcd.py:128-132`)."""
    i = max(0, min(i_ep, t_ep))
    k = round(k_end + (k_start - k_end) * (1.0 - i / max(1, t_ep)))
    return int(min(max(k, 1), n_agents - 1))


def mask_schedule(i_episode: int, n_veh: int, cfg: PairingConfig):
    """(K_now, q_now) of the mask curriculum (`marl_train_bcd.py:1323-1332`)."""
    prog = min(1.0, i_episode / max(1, cfg.mask_warmup_episodes))
    k_now = anneal_topk(i_episode, n_veh, cfg.mask_topk_start, cfg.mask_topk_end, cfg.mask_warmup_episodes)
    q_now = float(cfg.mask_tau_q_start + (cfg.mask_tau_q_end - cfg.mask_tau_q_start) * prog)
    return k_now, q_now


# --------------------------------------------------------------------------------------------
# numpy.quantile(method="linear") restated (numpy/lib/_function_base_impl.py: _quantile, _lerp)
# --------------------------------------------------------------------------------------------
def quantile_linear(vals, q: float) -> float:
    v = np.sort(np.asarray(vals, dtype=np.float64).ravel())
    n = v.size
    vi = (n - 1) * float(q)
    if vi >= n - 1:
        return float(v[-1])
    if vi < 0:
        return float(v[0])
    lo = math.floor(vi)
    a, b = float(v[int(lo)]), float(v[int(lo) + 1])
    t = vi - lo
    d = b - a
    r = a + d * t
    if t >= 0.5:
        r = b - d * (1 - t)
    return float(r)


def _stable_desc_order(x):
    """Indices that sort `x` descending, ties -> lower index first (== argsort(-x, kind='stable'))."""
    return sorted(range(len(x)), key=lambda j: (-x[j], j))


def _db(g, eps):
    return 10.0 * np.log10(np.maximum(np.asarray(g, dtype=np.float64), eps))


# --------------------------------------------------------------------------------------------
# helpers, one per reference function
# --------------------------------------------------------------------------------------------
def adaptive_threshold(gains, q: float) -> float:
    """`_adaptive_threshold_from_delta_g` (`marl_train_bcd.py:842-855`)."""
    g_db = _db(gains, 1e-15)
    n = g_db.size
    if n < 2:
        return 0.0
    order = np.argsort(g_db, kind="stable")
    weak, strong = order[: n // 2], order[n // 2:]
    if weak.size == 0 or strong.size == 0:
        return 0.0
    diffs = np.abs(g_db[strong][:, None] - g_db[weak][None, :]).ravel()
    return quantile_linear(diffs, q)


def build_feasible_mask(gains, tau: float, K: int) -> np.ndarray:
    """`_build_feasible_mask_from_delta_g` (`marl_train_bcd.py:134-156`) -> uint8 [N,N]."""
    g = _db(gains, 1e-15)
    N = g.size
    mask = np.ones((N, N), dtype=np.uint8)
    for i in range(N):
        mask[i, i] = 0
        for j in range(N):
            if i != j and abs(g[i] - g[j]) < tau:
                mask[i, j] = 0
    for i in range(N):
        cand = [j for j in range(N) if mask[i, j]]
        if len(cand) > K:
            diffs = [abs(g[i] - g[j]) for j in cand]
            keep = {cand[p] for p in _stable_desc_order(diffs)[:K]}
            for j in cand:
                if j not in keep:
                    mask[i, j] = 0
    return (mask * mask.T).astype(np.uint8)


def qos_pair_feasible(i, j, g, p01, noise_power, P_max, R_min) -> bool:
    """`_qos_pair_feasible` (`marl_train_bcd.py:858-881`)."""
    pi, pj = float(p01[i]) * float(P_max), float(p01[j]) * float(P_max)
    gi, gj = float(g[i]), float(g[j])
    if gi >= gj:
        g_near, g_far, p_near, p_far = gi, gj, pi, pj
    else:
        g_near, g_far, p_near, p_far = gj, gi, pj, pi
    sinr_far = (p_far * g_far) / (p_near * g_far + float(noise_power) + 1e-12)
    r_far = np.log2(1.0 + max(0.0, sinr_far))
    sinr_near = (p_near * g_near) / (float(noise_power) + 1e-12)
    r_near = np.log2(1.0 + max(0.0, sinr_near))
    return bool((r_far >= float(R_min)) and (r_near >= float(R_min)))


def qos_soft_mask(g, p01, noise_power, P_max, R_min) -> np.ndarray:
    """Call-site loop `marl_train_bcd.py:1428-1440` -> uint8 [N,N] (diagonal 0)."""
    N = len(g)
    m = np.zeros((N, N), dtype=np.uint8)
    for i in range(N):
        for j in range(N):
            if i != j:
                m[i, j] = 1 if qos_pair_feasible(i, j, g, p01, noise_power, P_max, R_min) else 0
    return m


def score_matrix(gains, feasible, hist_f32, w_delta_db, w_hist, abs_gain_min_db, qos_mask, qos_penalty):
    """`_score_matrix_from_gain_and_history` (`marl_train_bcd.py:164-194`).

    `hist_f32` is float32 (`:1288`), so `w_hist * hist` is a float32 product (numpy keeps the array's
    dtype against a python scalar) that is then widened to float64 by the sum."""
    g_db = _db(gains, 1e-12)
    delta = np.abs(g_db[:, None] - g_db[None, :])
    abs_ok = (g_db[:, None] >= abs_gain_min_db) | (g_db[None, :] >= abs_gain_min_db)
    hist_term = (np.float32(w_hist) * np.asarray(hist_f32, dtype=np.float32)).astype(np.float64)
    S = w_delta_db * delta + hist_term
    feas = np.asarray(feasible) > 0
    if not np.any(feas & abs_ok):
        abs_ok = np.ones_like(abs_ok)
    S = np.where(feas & abs_ok, S, -np.inf)
    if qos_mask is not None:
        S = np.where((np.asarray(qos_mask) <= 0) & np.isfinite(S), S - float(qos_penalty), S)
    np.fill_diagonal(S, -np.inf)
    return S


def relax_mask_once(mask, gains, tau_db: float, topk: int) -> np.ndarray:
    """`_relax_mask_once` (`marl_train_bcd.py:260-275`); row top-k by |delta dB| in stable order
    (the diagonal, delta = 0, competes like any other column)."""
    g_db = _db(gains, 1e-12)
    N = g_db.size
    delta = np.abs(g_db[:, None] - g_db[None, :])
    cand = delta >= tau_db
    np.fill_diagonal(cand, False)
    top = np.zeros((N, N), dtype=bool)
    if topk >= 1:
        k = min(topk, N - 1)
        for i in range(N):
            for j in _stable_desc_order(list(delta[i]))[:k]:
                top[i, j] = True
    return ((np.asarray(mask) > 0) | cand | top).astype(np.uint8)


def mwm_primary(S, feasible, accept_quantile: float):
    """`_mwm_primary` (`marl_train_bcd.py:326-398`) with `allow_singles=True`: exact maximum-weight
    matching by bitmask DP on the edges whose score is >= the (1-q) quantile.  The reference's
    memoised recursion `dp(mask)` is a pure function of `mask`, so the bottom-up table below holds the
    same float64 values; option order (single first, then j ascending, strict `>`) is preserved."""
    S = np.asarray(S, dtype=np.float64)
    N = S.shape[0]
    edge = (np.asarray(feasible) > 0) & np.isfinite(S)
    vals = S[edge]
    if vals.size == 0:
        return []
    q = min(max(float(accept_quantile), 0.0), 1.0)
    thr = quantile_linear(vals, 1.0 - q)
    W = np.where((S >= thr) & edge, S, -np.inf)
    full = (1 << N) - 1
    best = [0.0] * (full + 1)
    choice = [-2] * (full + 1)
    for mask in range(full - 1, -1, -1):
        i = 0
        while mask & (1 << i):
            i += 1
        bw, ch = -math.inf, -2
        w1 = best[mask | (1 << i)]
        if w1 > bw:
            bw, ch = w1, -1
        for j in range(i + 1, N):
            if mask & (1 << j):
                continue
            we = W[i, j]
            if not np.isfinite(we):
                continue
            w2 = best[mask | (1 << i) | (1 << j)]
            if np.isfinite(w2) and (we + w2 > bw):
                bw, ch = we + w2, j
        best[mask], choice[mask] = float(bw), ch
    pairs, mask = [], 0
    while mask != full:
        i = 0
        while mask & (1 << i):
            i += 1
        ch = choice[mask]
        if ch >= 0:
            pairs.append([i, ch])
            mask |= (1 << i) | (1 << ch)
        else:
            mask |= 1 << i
    pairs.sort(key=lambda x: (x[0], x[1]))
    return pairs


def mwm_completion(S, feasible, pairs_now, min_pairs: int, completion_min_quantile: float):
    """`_mwm_completion` (`marl_train_bcd.py:276-324`): greedy top-up on edges with score >= quantile."""
    S = np.asarray(S, dtype=np.float64)
    N = S.shape[0]
    used = {u for ab in pairs_now for u in ab}
    finite = np.isfinite(S) & (np.asarray(feasible) > 0)
    if not np.any(finite):
        return list(pairs_now)
    thr = quantile_linear(S[finite], completion_min_quantile)
    cand = [(S[i, j], i, j) for i in range(N) for j in range(i + 1, N) if finite[i, j] and S[i, j] >= thr]
    cand.sort(reverse=True)
    new_pairs, occupied = [], set(used)
    for _, i, j in cand:
        if i in occupied or j in occupied:
            continue
        new_pairs.append((i, j))
        occupied.update((i, j))
        if len(pairs_now) + len(new_pairs) >= int(min_pairs):
            break
    return list(pairs_now) + new_pairs


# --------------------------------------------------------------------------------------------
# the call site: one pairing step for one env
# --------------------------------------------------------------------------------------------
@dataclass
class PairingState:
    """What the driver keeps across steps of an episode (`marl_train_bcd.py:1282-1297`)."""

    n_veh: int
    hist: np.ndarray = None             # pair_affinity_hist float32 [N,N] (:1288)
    streak: np.ndarray = None           # unpaired_streak int32 [N] (:1290)
    mask: np.ndarray = None             # last_mask_mat uint8 [N,N] (:1282)
    tau: float = 0.0                    # last_tau_now
    K: int = 0                          # last_K_now
    q: float = 0.0                      # last_q_now
    groups: list = field(default_factory=list)      # episode_groups (:1297)

    def __post_init__(self):
        n = self.n_veh
        if self.hist is None:
            self.hist = np.zeros((n, n), dtype=np.float32)
        if self.streak is None:
            self.streak = np.zeros((n,), dtype=np.int32)


def pair_step(st: PairingState, gains, p01, cfg: PairingConfig, noise_power: float, P_max: float,
              R_min: float, K_now: int, q_now: float, recalc_mask: bool = True, reuse: bool = False,
              decay: bool = True, info: dict | None = None):
    """One driver step of the pairing stage (`marl_train_bcd.py:1315-1344,1404-1406,1413-1524,
    1542-1561`).  `reuse=True` is the frozen-groups path (`:1542-1547`): the solve is skipped (its
    results are discarded by the reference) and only the history / streak updates run.

    Returns `(pairs, groups)`: `pairs` in the reference's list order, `groups = pairs + singles`."""
    gains = np.asarray(gains, dtype=np.float64)
    p01 = np.asarray(p01, dtype=np.float64)
    N = st.n_veh
    if decay:                                               # :1406 (float32 array *= python float)
        st.hist *= np.float32(cfg.pair_hist_decay)
    mask_mat = None                                         # :1317 -- a per-step local in the reference
    if recalc_mask or st.mask is None:                      # :1319-1343
        st.tau = adaptive_threshold(gains, q_now)
        st.mask = mask_mat = build_feasible_mask(gains, st.tau, K_now)
        st.K, st.q = int(K_now), float(q_now)
    if reuse and st.groups:
        pairs = [(g[0], g[1]) for g in st.groups if len(g) == 2]
        groups = [list(g) for g in st.groups]
        rounds = 0
    else:
        min_pairs = max(1, cfg.min_pair_target)             # :1413
        # :1421-1424 -- quirk kept: on steps that do not recompute the mask the local `mask_mat` is None
        # and the solve runs on the all-ones (minus diagonal) mask; only tau / K are carried over.
        feasible = (mask_mat.astype(np.uint8) if mask_mat is not None
                    else (np.ones((N, N), dtype=np.uint8) - np.eye(N, dtype=np.uint8)))
        qos = qos_soft_mask(gains, p01, noise_power, P_max, R_min) if cfg.qos_enable else None   # :1426-1440
        S = score_matrix(gains, feasible, st.hist, cfg.score_w_delta_db, cfg.score_w_history,
                         cfg.abs_gain_min_db, qos, cfg.qos_soft_penalty_dbscore)                 # :1441-1450
        accept_q = float(cfg.mwm_accept_quantile)
        pairs = mwm_primary(S, feasible, accept_q)          # :1456-1461
        if len(pairs) < min_pairs:                          # :1464-1465
            pairs = mwm_completion(S, feasible, pairs, min_pairs, cfg.completion_min_quantile)
        rounds, K_back, tau_back = 0, int(st.K), float(st.tau)      # :1480-1491
        while len(pairs) < min_pairs and rounds < int(cfg.mwm_backoff_rounds):      # :1493-1524
            rounds += 1
            K_back = min(N - 1, K_back + int(cfg.relax_topk_step))
            tau_back = max(float(cfg.tau_back_floor_db), tau_back * float(cfg.relax_tau_factor_per_round))
            feasible = relax_mask_once(feasible, gains, tau_back, K_back)
            S = score_matrix(gains, feasible, st.hist, cfg.score_w_delta_db, cfg.score_w_history,
                             cfg.abs_gain_min_db, qos, cfg.qos_soft_penalty_dbscore)
            accept_q = max(0.05, accept_q - float(cfg.mwm_accept_q_step))
            pairs = mwm_primary(S, feasible, accept_q)
            if len(pairs) < min_pairs:
                pairs = mwm_completion(S, feasible, pairs, min_pairs, cfg.completion_min_quantile)
        pairs = [(int(a), int(b)) for a, b in pairs]
        used = {u for ab in pairs for u in ab}              # :1550-1553
        groups = [[i, j] for (i, j) in pairs] + [[k] for k in range(N) if k not in used]
        st.groups = [list(g) for g in groups]
    used = {u for ab in pairs for u in ab}
    for (i, j) in pairs:                                    # :1556-1558
        st.hist[i, j] += np.float32(1.0)
        st.hist[j, i] += np.float32(1.0)
    for u in range(N):                                      # :1560-1561
        st.streak[u] = 0 if u in used else st.streak[u] + 1
    if info is not None:
        info.update(rounds=rounds, tau=st.tau)
    return [tuple(p) for p in pairs], groups
