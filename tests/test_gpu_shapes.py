"""Shape coverage of the generic kernels: random (V, M, E, T) against the float64 oracle,
including V = 1, M = 1, odd element slices, V % 4 != 0 (extra vehicles) and the M = 1024 cap."""
import numpy as np
import pytest
import torch

from oracle.env_oracle import EnvOracle, InjectedDraws, OracleParams, encode_groups
from tests.parity import RTOL, sarl_rate_atol, sarl_reward_band

pytestmark = pytest.mark.gpu

SHAPES = [(1, 1), (3, 2), (2, 17), (8, 16), (8, 24), (16, 63), (9, 129), (17, 100), (31, 255), (32, 512), (5, 1024),
          (4, 100), (12, 200), (20, 66), (32, 256), (8, 130),   # k_sarl_mma_big
          (4, 1024), (2, 512), (1, 256),                        # few vehicles x many elements (8-lane mapping) (V % 4 == 0, M even, M <= 256)
          (32, 1024)]     # the last one is the largest shape the library accepts


def reset_inputs(rng, E, V):
    pat = [(0, 4), (220, 230), (10, 15), (170, 180), (10, 15), (220, 230), (10, 15), (170, 180), (10, 15)] * (V // 4)
    pat += [(0, 4), (0, 400), (15, 20)] * (V % 4) + [(5, 9)]
    ints = np.stack([rng.integers(lo, hi, E) for lo, hi in pat], axis=1).astype(np.int32)
    dirs = rng.integers(0, 4, (E, V % 4)).astype(np.int32)
    return ints, dirs


def close(got, want, atol, what):
    got, want = np.asarray(got, float), np.asarray(want, float)
    bad = np.abs(got - want) > atol + RTOL * np.abs(want)
    assert not bad.any(), f"{what}: {bad.sum()}/{bad.size} out of tolerance, max abs err {np.abs(got - want)[bad].max():.3g}"


@pytest.mark.parametrize("V,M", SHAPES)
def test_sarl_random_shapes(V, M):
    from ris_vec_marl_b200 import BatchedEnviron

    E, T = 7, 9
    rng = np.random.default_rng(1000 * V + M)
    ints, dirs = reset_inputs(rng, E, V)
    mob = rng.random((E, 8 * V))
    acts = rng.random((T, E, 2, V)).astype(np.float32)
    phs = (rng.random((T, E, M)) * 2 * np.pi).astype(np.float32)
    arr = rng.poisson(3.0, (T, E, V)).astype(np.int32)
    env = BatchedEnviron("sarl", E, V, M)
    env.make_new_game(ints, dirs if V % 4 else None)
    env.renew_positions(mob); env.compute_parms()
    got = {k: v.cpu().numpy() for k, v in env.rollout_sarl(acts, phs, arr).items()}
    d = InjectedDraws(reset_ints=ints, reset_dirs=dirs, arrivals=arr)
    o = EnvOracle("sarl", V, M, 3, E=E, draws=d)
    o.make_new_game(); d.set_mobility_uniforms(mob); o.renew_positions(); o.compute_parms()
    assert np.array_equal(env.pos_x.cpu().numpy(), o.pos[..., 0]) and np.array_equal(env.dir.cpu().numpy(), o.dir)
    ra = sarl_rate_atol(M)
    n_band = 0
    for t in range(T):
        rew, over_p = o.step_sarl(acts[t], phs[t])
        band = sarl_reward_band(o.last["buf_signed"], o.over_data, M).any(axis=1)
        n_band += int(band.sum())
        close(got["rate"][t], o.vehicle_rate, ra, f"rate t={t}")
        close(got["DataBuf"][t], o.DataBuf, 4 * ra, f"DataBuf t={t}")
        close(got["over_power"][t], over_p, 4e-6 + 4 * ra, f"over_power t={t}")
        close(got["reward"][t][~band], rew[~band], 4e-6 + ra, f"reward t={t}")
    assert n_band <= 0.05 * E * T + 1, f"{n_band} of {E * T} env-steps band-excluded from the reward check"


@pytest.mark.parametrize("V,M", [(1, 3), (2, 5), (7, 12), (13, 40), (24, 9), (32, 64)])
def test_marl_random_shapes(V, M):
    from ris_vec_marl_b200 import BatchedEnviron, marl_yaml_overrides

    E, T = 6, 8
    rng = np.random.default_rng(77 * V + M)
    ints, dirs = reset_inputs(rng, E, V)
    mob = rng.random((E, 8 * V))
    acts = rng.random((T, E, 2, V)).astype(np.float32)
    acts[:, :, 1, :] = np.maximum(acts[:, :, 1, :], np.float32(0.1))
    arr = rng.poisson(1.0, (T, E, V)).astype(np.int32)
    groups = [[i, i + 1] for i in range(0, V - 1, 3)] + [[i + 2] for i in range(0, V - 2, 3)]
    part, ng = encode_groups(groups, V)
    partner, ngroups = np.tile(part, (E, 1)), np.full(E, ng, dtype=np.int32)
    env = BatchedEnviron("marl", E, V, M, **marl_yaml_overrides())
    env.make_new_game(ints, dirs if V % 4 else None)
    env.renew_positions(mob); env.compute_parms(); env.optimize_phase_shift(); env.update_channel_gains()
    got = {k: v.cpu().numpy() for k, v in env.rollout_marl(acts, partner, ngroups, arr).items()}
    d = InjectedDraws(reset_ints=ints, reset_dirs=dirs, arrivals=arr)
    o = EnvOracle("marl", V, M, 3, E=E, params=OracleParams.marl_yaml(), draws=d)
    o.make_new_game(); d.set_mobility_uniforms(mob); o.renew_positions(); o.compute_parms()
    o.optimize_phase_shift(); o.update_channel_gains()
    np.testing.assert_allclose(env.gains.cpu().numpy(), o.channel_gains, rtol=1e-9)
    p = o.p
    for t in range(T):
        r_user, r_glob, over_p = o.step_marl(acts[t], partner, ngroups)
        L = o.last
        band = (np.abs(o.vehicle_rate - p.R_min_bpsHz) < 1e-5 * p.R_min_bpsHz) | (np.abs(L["delay"] - p.D_max_s) < 1e-5 * p.D_max_s)
        band |= ((L["edge_in_sum"] < 1e-3) & (L["q_before"] > 0) & (L["edge_in_sum"] > 0))[:, None]
        close(got["rate"][t], o.vehicle_rate, 1e-6, f"rate t={t}")
        close(got["DataBuf"][t], o.DataBuf, 1e-5, f"DataBuf t={t}")
        close(np.where(band, 0, got["reward_user"][t]), np.where(band, 0, r_user), 2e-6, f"reward_user t={t}")
        close(got["stats"][t][:, 0], L["delay_mean"], 1e-9 + 1e-3 * band.any(axis=1), f"delay_mean t={t}")


@pytest.mark.parametrize("variant,V,M,E,T", [("sarl", 8, 40, 4098, 7), ("sarl", 5, 33, 9, 4), ("sarl", 32, 256, 3, 5),
                                             ("sarl", 17, 100, 5, 3), ("marl", 8, 40, 4097, 6), ("marl", 6, 7, 10, 5),
                                             ("marl", 32, 64, 3, 4)])
def test_traces_stay_inside_their_buffers(variant, V, M, E, T):
    """compute-sanitizer is not available on this pool, so bounds are checked with canaries:
    every trace is carved out of a sentinel-filled arena with guard bands on both sides; after a
    rollout each trace must be fully written (no sentinel left) and every guard band intact."""
    from ris_vec_marl_b200 import BatchedEnviron, MARL_TRACES, SARL_TRACES, marl_yaml_overrides

    over = marl_yaml_overrides() if variant == "marl" else {}
    env = BatchedEnviron(variant, E, V, M, **over)
    env.make_new_game(); env.renew_positions(); env.compute_parms()
    names = SARL_TRACES if variant == "sarl" else MARL_TRACES
    shapes = {k: tuple(v.shape) for k, v in env._alloc_traces(names, T, names).items()}
    guard, sentinel = 1024, -7.0e30
    total = sum(int(np.prod(sh)) + guard for sh in shapes.values()) + guard
    arena = torch.full((total,), sentinel, dtype=torch.float32, device="cuda")
    out, off = {}, guard
    for k, sh in shapes.items():
        n = int(np.prod(sh))
        out[k] = arena[off:off + n].view(sh)
        off += n + guard
    gen = torch.Generator(device="cuda").manual_seed(3)
    acts = torch.rand(T, E, 2, V, device="cuda", generator=gen)
    arr = torch.poisson(torch.full((T, E, V), 2.0, device="cuda"), generator=gen).to(torch.int32)
    if variant == "sarl":
        env.rollout_sarl(acts, torch.rand(T, E, M, device="cuda", generator=gen) * 6.28, arr, out=out)
    else:
        env.optimize_phase_shift(); env.update_channel_gains()
        part = torch.full((E, V), -1, dtype=torch.int32, device="cuda")
        env.rollout_marl(acts, part, torch.full((E,), V, dtype=torch.int32, device="cuda"), arr, out=out)
    torch.cuda.synchronize()
    mask = torch.ones(total, dtype=torch.bool, device="cuda")
    off = guard
    for k, sh in shapes.items():
        n = int(np.prod(sh))
        mask[off:off + n] = False
        assert not (out[k] == sentinel).any(), f"{k} not fully written"
        assert torch.isfinite(out[k]).all(), k
        off += n + guard
    assert (arena[mask] == sentinel).all(), "a kernel wrote outside its trace buffers"
