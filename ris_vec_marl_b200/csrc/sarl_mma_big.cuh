// SARL rollout for MANY vehicles / RIS elements (BASELINE config 4: V = 32, M = 256): the cascaded
// reduction of one env is the real GEMM  [2V x 2M] . [2M x T]  (= [64 x 512] per step column) on the
// tensor cores, one thread BLOCK per env.  Reference: Simulation-SARL/Environment.py:125-131, 149-171,
// 318-359.  Same operand split (two binary16 pieces, 3 mma per product) and the same max-plus treatment
// of the DataBuf recursion as k_sarl_mma_tma (sarl_mma.cuh); what differs is the decomposition:
//
//   16 warps = 4 row tiles r (8 vehicles: rows Re S_v, Im S_v) x 4 K-quarters q (M / 4 elements each).
//   Warp (r, q) keeps ITS slice of the geometry phasors in registers as mma A fragments (8 k-tiles x
//   8 registers) for the whole rollout.  theta = exp(j*phase) of a 16-step stage is evaluated ONCE per
//   block (every warp produces 4 of the 64 B-fragment sets: one packed sin/cos per lane and set),
//   split, and parked in shared memory in fragment order (double buffered: stage k + 1 is produced
//   while stage k multiplies); the warps then stream the fragments of their K-quarter with LDS.128.
//   The four K-quarter partial sums meet in shared memory, laid out so that warp (r, q) picks up
//   steps 4 q .. 4 q + 3 of its 8 vehicles: one (vehicle, step) per lane for the per-step part.  The
//   recursion is scanned over the 4 lanes of a vehicle by shuffles and over the 4 warps through
//   16-byte composites in shared memory (a 16-lane shuffle scan per vehicle measured 20 % slower).
//   Traces leave through an out tile [6][16][V] and one TMA tensor store per trace and stage.
#pragma once
#include "sarl_mma.cuh"

namespace risvec {

struct SarlBigOutMaps {
    CUtensorMap trace[6];  // DataBuf, data_t, data_p, over_power, over_data, rate: [T, E*V] f32, box {V, 16}
};

constexpr int kBigThreads = 512;
constexpr int kBigPartStride = 40;  // floats per (warp, g) row of the partial-sum area (32 used; pad breaks bank conflicts)
constexpr int kBigPartFloats = 16 * 8 * kBigPartStride;  // one stage of partial sums (double buffered)
__host__ __device__ constexpr int sarl_big_smem_bytes(int KQ, int V) {
    return 2 * (4 * KQ) * 2 * 32 * 16            // B fragments, two stages
           + 2 * kBigPartFloats * 4              // K-quarter partial sums, two stages
           + 4 * 4 * 8 * 16                      // max-plus composites [r][q][g]
           + 4 * 16 * 4                          // reward partial sums [r][step]
           + 6 * 16 * V * 4                      // out tile
           + 256;                                // alignment slack
}

template <int KQ>
__global__ void __launch_bounds__(kBigThreads, 1)
    k_sarl_mma_big(Dims d, State s, const SarlConsts c, SarlArgs a, const __grid_constant__ SarlBigOutMaps tm_out) {
    constexpr int KT = 4 * KQ, R = 16;
    extern __shared__ unsigned char big_smem_raw[];
    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0);
    const int r = warp >> 2, q = warp & 3;
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int E = d.E, V = d.V, M = d.M, T = a.T;
    const int e = blockIdx.x;
    const int v = 8 * r + g;
    const bool vact = v < V;
    const int vc = min(v, V - 1);
    const size_t ev = (size_t)e * V + vc;

    // ---- shared memory carve-up (128 B aligned: the out tile is a TMA source)
    const uint32_t base_s = (smem_u32(big_smem_raw) + 127u) & ~127u;
    unsigned char* base_g = big_smem_raw + (base_s - smem_u32(big_smem_raw));
    constexpr int BF_STAGE = KT * 2 * 32 * 16;
    uint4* const bf = reinterpret_cast<uint4*>(base_g);                                  // [2][KT][2][32]
    float* const part = reinterpret_cast<float*>(base_g + 2 * BF_STAGE);                 // [2][16][8][kBigPartStride]
    double2* const comps = reinterpret_cast<double2*>(part + 2 * kBigPartFloats);        // [4][4][8]
    float* const rsum = reinterpret_cast<float*>(comps + 4 * 4 * 8);                     // [4][16]
    float* const out_g = rsum + 4 * 16;                                                  // [6][16][V]
    const uint32_t out_s = base_s + (uint32_t)((unsigned char*)out_g - base_g);
    const int TRACE_WORDS = R * V;

    // ---- A operand: vehicle v, elements 8 j + 2 tig + {0, 1} of the k-tiles j = KQ q .. KQ q + KQ - 1
    uint32_t Ah[KQ][4], Al[KQ][4];
    {
        const double2 z = unit_phasor64(d.angle_BR - s.angle[ev]);  // w(v, m) = z^m, float64 (SARL:134-145)
        const double2 z2 = cmul64(z, z), z4 = cmul64(z2, z2), z8 = cmul64(z4, z4);
        double2 w = cpow64(z, 2u * (unsigned)tig + 8u * (unsigned)(KQ * q));
#pragma unroll
        for (int jj = 0; jj < KQ; ++jj) {
            const int ma = 8 * (KQ * q + jj) + 2 * tig;
            double2 wa = w, wb = cmul64(w, z);
            if (!(vact && ma < M)) wa = make_double2(0.0, 0.0);
            if (!(vact && ma + 1 < M)) wb = make_double2(0.0, 0.0);
            split_h2(wa.x, -wa.y, Ah[jj][0], Al[jj][0]);  // row g     (Re S_v): ( Re w, -Im w)
            split_h2(wa.y, wa.x, Ah[jj][1], Al[jj][1]);   // row g + 8 (Im S_v): ( Im w,  Re w)
            split_h2(wb.x, -wb.y, Ah[jj][2], Al[jj][2]);
            split_h2(wb.y, wb.x, Ah[jj][3], Al[jj][3]);
            w = cmul64(w, z8);
        }
    }
    // per-step part: warp (r, q) serves steps 4 q .. 4 q + 3 of the row tile's 8 vehicles, one (vehicle, step) per lane
    const int sstep = 4 * q + tig;
    const int ve = v;
    const bool eact = vact;
    const int vec = vc;
    const size_t eve = ev;
    double buf = s.databuf[eve];  // replicated over the lanes / warps that serve the vehicle
    const float coef = eact ? (float)(s.amp[eve] / (kSigma * kSigma)) : 0.f;  // SARL:157-159
    const long long step0 = s.step_ctr[e];
    const int NS = (T + R - 1) / R;

    // ---- producer side: this warp makes the B fragments of k-tiles (warp, warp + 16, ...) x both n-tiles.
    // B column n = g of tile A is step 4 (g >> 1) + (g & 1) of the stage, of tile B the step two later, so
    // that an accumulator lane (g, tig') ends up with the four consecutive steps 4 tig' .. 4 tig' + 3.
    // (set = warp + 16 u: k-tile (warp >> 1) + 8 u, n-tile warp & 1 -- the same n-tile and therefore the same
    //  stage row for all of a warp's sets, element offsets 64 apart)
    constexpr int NSETS = (KT * 2) / 16;        // fragment sets per warp and stage
    static_assert(KT * 2 == NSETS * 16, "the 2 KT fragment sets divide evenly over the 16 warps");
    const int row_w = 4 * (g >> 1) + (g & 1) + 2 * (warp & 1);
    const unsigned sM = (unsigned)E * M, s2V = (unsigned)E * 2 * V, sVv = (unsigned)E * V;  // 32-bit indices (host-checked)
    const int m_w = 8 * (warp >> 1) + 2 * tig;                       // element of set u: m_w + 64 u
    const float* const ph_w = a.phase + (unsigned)e * M + m_w;       // + t * sM + 64 u
    uint4* const bf_w = bf + warp * 32 + lane;                       // + buffer * (KT * 64) + u * 512
    float2 phn[NSETS];                          // phases of the NEXT stage to produce (register prefetch)
    auto load_phases = [&](int k) {
        const float* q0 = ph_w + (unsigned)min(k * R + row_w, T - 1) * sM;
#pragma unroll
        for (int u = 0; u < NSETS; ++u)
            phn[u] = (m_w + 64 * u < M) ? __ldg(reinterpret_cast<const float2*>(q0 + 64 * u)) : make_float2(0.f, 0.f);
    };
    auto produce = [&](int k) {  // phn (stage k) -> bf[k & 1]
        uint4* dst = bf_w + (k & 1) * (KT * 64);
#pragma unroll
        for (int u = 0; u < NSETS; ++u) {
            float2 sn, cs;
            sincos_pi2(phn[u], &sn, &cs);  // theta = exp(j*phase) (SARL:125-131)
            uint4 f;  // (b0 hi, b1 hi, b0 lo, b1 lo): every mma B operand is an adjacent register pair
            split_h2(cs.x, sn.x, f.x, f.z);  // b0: element 8 j + 2 tig
            split_h2(cs.y, sn.y, f.y, f.w);  // b1: element 8 j + 2 tig + 1
            dst[u * 512] = f;
        }
    };
    // ---- per-step inputs of my (vehicle, step) item, one stage ahead in registers
    const float* const ac_w = a.action + (unsigned)e * 2 * V + vec;
    const int* const ar_w = a.arrivals != nullptr ? a.arrivals + (unsigned)e * V + vec : nullptr;
    float na0 = 0.f, na1 = 0.f;
    int narr = 0;
    auto load_scalars = [&](int k) {
        const unsigned t = (unsigned)min(k * R + sstep, T - 1);
        na0 = __ldg(ac_w + t * s2V);
        na1 = __ldg(ac_w + t * s2V + V);
        narr = ar_w != nullptr ? __ldg(ar_w + t * sVv) : 0;
    };
    const bool is_t0 = threadIdx.x == 0;
    const int tid16 = threadIdx.x < 16 ? (int)threadIdx.x : -1;
    // step s of vehicle 8 r + g sits in accumulator lane (g, tig' = s >> 2), slot s & 3, of the warps (r, 0..3)
    const float* const part_r = part + (r * 32 + g) * kBigPartStride + 2 * sstep;          // + qq * 8 * stride
    double2* const comps_r = comps + r * 32 + g;                                            // + qq * 8
    float4* const part_w = reinterpret_cast<float4*>(part + (warp * 8 + g) * kBigPartStride + tig * 8);
    float* const out_w = out_g + sstep * V + ve;

    // my K-quarter of stage k's GEMM: fragments bf[k & 1] -> partial sums part[k & 1]
    auto mma_stage = [&](int k) {
        float mA[4] = {0.f, 0.f, 0.f, 0.f}, xA[4] = {0.f, 0.f, 0.f, 0.f};
        float mB[4] = {0.f, 0.f, 0.f, 0.f}, xB[4] = {0.f, 0.f, 0.f, 0.f};
        const uint4* bk = bf + ((k & 1) * KT * 2 + 2 * (KQ * q)) * 32 + lane;  // sets 2 j, 2 j + 1 of my k-tiles
#pragma unroll
        for (int jj = 0; jj < KQ; ++jj) {
            const uint4 fa = bk[(2 * jj) * 32], fb = bk[(2 * jj + 1) * 32];
            mma_16816(mA, Ah[jj], fa.x, fa.y);
            mma_16816(xA, Ah[jj], fa.z, fa.w);
            mma_16816(xA, Al[jj], fa.x, fa.y);
            mma_16816(mB, Ah[jj], fb.x, fb.y);
            mma_16816(xB, Ah[jj], fb.z, fb.w);
            mma_16816(xB, Al[jj], fb.x, fb.y);
        }
        // partial S of my lane's four steps 4 tig + i: (Re, Im) pairs, i = 0, 1 from tile A, 2, 3 from tile B
        float4* pw = part_w + (k & 1) * (kBigPartFloats / 4);
        pw[0] = make_float4(mA[0] + xA[0], mA[2] + xA[2], mA[1] + xA[1], mA[3] + xA[3]);
        pw[1] = make_float4(mB[0] + xB[0], mB[2] + xB[2], mB[1] + xB[1], mB[3] + xB[3]);
    };

    // ---- software pipeline over the 16-step stages.  Between two block barriers every warp holds two
    // INDEPENDENT instruction streams, so the float64 / shuffle chains of the per-step part overlap with
    // sin/cos and tensor-core work of later stages:
    //   segment 1:  produce theta fragments of stage k + 2   ||  per-step part of stage k, first half
    //   segment 2:  GEMM of stage k + 1                      ||  per-step part of stage k, second half
    load_phases(0);
    load_scalars(0);
    produce(0);
    load_phases(1 < NS ? 1 : 0);
    __syncthreads();
    if (1 < NS) {
        produce(1);
        load_phases(2 < NS ? 2 : 1);
    }
    mma_stage(0);
    __syncthreads();

    // values of step T - 1 (they become the env's state)
    float f_rate = 0.f, f_dt = 0.f, f_dp = 0.f, f_overp = 0.f, f_overd = 0.f;
    int f_arr = 0;
    bool f_mine = false;

    for (int k = 0; k < NS; ++k) {
        // ---- segment 1
        if (k + 2 < NS) {  // theta of stage k + 2 -> the fragment buffer stage k used; its phases were requested earlier
            produce(k + 2);
            load_phases(k + 3 < NS ? k + 3 : k + 2);
        }
        const float a0 = na0, a1 = na1;
        int arr = narr;
        if (k + 1 < NS) load_scalars(k + 1);
        // per-step part (SARL:327-358) of my item: vehicle ve, step t = 16 k + sstep
        const int t = k * R + sstep;
        const bool ok = eact && t < T;
        float re = 0.f, im = 0.f;
        const float* pr = part_r + (k & 1) * kBigPartFloats;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const float2 pq = *reinterpret_cast<const float2*>(pr + qq * 8 * kBigPartStride);
            re += pq.x;
            im += pq.y;
        }
        const float g2 = __fmaf_rn(re, re, __fmul_rn(im, im));
        if (a.arrivals == nullptr) arr = ok ? draw_arrival(d, e, vec, step0 + t, c.lam) : 0;
        const float rate = log1p_sfu(__fmul_rn(a0, __fmul_rn(coef, g2)));  // natural log, SARL:159
        const float dt = __fmul_rn(rate, c.c_dt);
        const float dp = __fmul_rn(cbrt_sfu(a1), c.c_dp);                   // SARL:331
        const double dd = ok ? __dadd_rn((double)dt, (double)dp) : 0.0;     // identity step when not ok
        const double inc = ok ? __dmul_rn(__dmul_rn((double)arr, c.tf), 1000.0) : 0.0;
        // scan over the 4 lanes of the vehicle (steps 4 q .. 4 q + 3), then over the 4 warps q via shared memory
        MaxPlus f{inc - dd, inc};
        {
            MaxPlus pq{__shfl_up_sync(kFull, f.a, 1, 4), __shfl_up_sync(kFull, f.b, 1, 4)};
            const MaxPlus f1 = mp_then(pq, f);
            if (tig >= 1) f = f1;
            pq = MaxPlus{__shfl_up_sync(kFull, f.a, 2, 4), __shfl_up_sync(kFull, f.b, 2, 4)};
            const MaxPlus f2m = mp_then(pq, f);
            if (tig >= 2) f = f2m;
        }
        const MaxPlus ex{__shfl_up_sync(kFull, f.a, 1, 4), __shfl_up_sync(kFull, f.b, 1, 4)};
        if (tig == 3) comps_r[q * 8] = make_double2(f.a, f.b);
        if (is_t0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // out tile free again
        __syncthreads();  // S2
        // ---- segment 2
        if (k + 1 < NS) mma_stage(k + 1);
        MaxPlus before{0.0, -1.0e300}, whole{0.0, -1.0e300};  // identity maps (x -> max(x, -huge))
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const double2 cq = comps_r[qq * 8];
            const MaxPlus m{cq.x, cq.y};
            whole = mp_then(whole, m);
            if (qq < q) before = whole;
        }
        const double xq = mp_apply(before, buf);                 // DataBuf at step 4 q of the stage
        const double xin = tig == 0 ? xq : mp_apply(ex, xq);     // ... at my step
        buf = mp_apply(whole, buf);
        // my step in the reference's order (SARL:333-358)
        const double raw = __dsub_rn(xin, dd);  // SARL:334
        const bool neg = raw < 0.0;
        const float b = __fmul_rn(fmaxf(0.f, (float)(raw + (double)dp)), c.c_rev);
        const float overp = neg ? __fsub_rn(a1, __fmul_rn(__fmul_rn(b, b), b)) : 0.f;  // SARL:336-339
        const float overd = neg ? (float)(-raw) : 0.f;
        const double nb = neg ? 0.0 : raw;
        const float basev = __fmaf_rn((float)nb, c.nt2, __fmul_rn(__fadd_rn(a0, a1), c.nt1));
        const float pen = (nb > 0.0) ? c.pen1 : ((overd > 2.0f) ? c.pen2 : 0.f);  // SARL:343-352
        const double cur = __dadd_rn(nb, inc);  // SARL:354-356
        if (eact) {
            float* o = out_w;
            o[0 * TRACE_WORDS] = (float)cur;
            o[1 * TRACE_WORDS] = dt;
            o[2 * TRACE_WORDS] = dp;
            o[3 * TRACE_WORDS] = overp;
            o[4 * TRACE_WORDS] = overd;
            o[5 * TRACE_WORDS] = rate;
        }
        if (ok && t == T - 1) {
            f_rate = rate; f_dt = dt; f_dp = dp; f_overp = overp; f_overd = overd; f_arr = arr; f_mine = true;
        }
        float ru = ok ? __fsub_rn(basev, pen) : 0.f;  // reward: mean over all vehicles (8 here, 4 row tiles via smem)
        ru += __shfl_xor_sync(kFull, ru, 4);
        ru += __shfl_xor_sync(kFull, ru, 8);
        ru += __shfl_xor_sync(kFull, ru, 16);
        if (g == 0) rsum[r * 16 + sstep] = ru;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();  // S3: out tile, reward sums, partial sums of stage k + 1 and fragments of stage k + 2 complete
        if (is_t0) {
#pragma unroll
            for (int n = 0; n < 6; ++n) tma_store_2d(&tm_out.trace[n], out_s + n * (TRACE_WORDS * 4), e * V, k * R);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        if (tid16 >= 0 && k * R + tid16 < T) {
            const int st = tid16;
            const float rew = __fmul_rn((rsum[st] + rsum[16 + st]) + (rsum[32 + st] + rsum[48 + st]), 1.0f / (float)V);
            if (a.out.reward != nullptr) a.out.reward[(size_t)(k * R + st) * E + e] = rew;
            if (k * R + st == T - 1) s.reward[e] = rew;
        }
    }

    // ---- registers -> state
    for (int m = threadIdx.x; m < M; m += kBigThreads)  // elements_phase_shift_real = the last action_phase (SARL:128)
        s.phase_real[(size_t)e * M + m] = __ldg(a.phase + ((size_t)(T - 1) * E + e) * M + m);
    if (f_mine) {
        s.rate[eve] = f_rate;
        s.data_t[eve] = f_dt;
        s.data_p[eve] = f_dp;
        s.over_power[eve] = f_overp;
        s.over_data[eve] = f_overd;
        s.data_r[eve] = f_arr;
    }
    if (q == 0 && tig == 0 && eact) s.databuf[eve] = buf;
    if (is_t0) {
        s.step_ctr[e] = step0 + T;
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before exit
    }
}

}  // namespace risvec
