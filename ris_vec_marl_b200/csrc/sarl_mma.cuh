// SARL rollout with the cascaded RIS reduction on the tensor cores (row a12 of SURVEY.md 8a).
// Reference: Simulation-SARL/Environment.py:125-131 (get_next_phase), :149-171 (compute_data_rate),
//            :318-359 (localProcRev, step).
//
// With T steps fused, the per-step reduction  S_v(t) = sum_m exp(j*phase_m(t)) * w(v, m)  of one env is
// a real GEMM  [2V x 2M] . [2M x T]:  rows = (Re S_v, Im S_v), K = (cos, sin) of every element,
// columns = time steps.  One WARP owns one env and walks the rollout in tiles of 8 steps:
//
//   A (16 x 16 per k-tile, row major) = geometry phasors of the env, constant over the rollout, kept in
//       registers as mma fragments:  A[v][(m, cos)] = Re w, A[v][(m, sin)] = -Im w,
//                                    A[8 + v][(m, cos)] = Im w, A[8 + v][(m, sin)] = Re w;
//   B (16 x 8 per k-tile, column major) = theta = (cos, sin)(phase[t0 + n][m]) -- lane (g, tig) evaluates
//       exactly the two elements its B fragment holds for step t0 + g (one packed sin/cos), so theta
//       goes from the sin/cos polynomial straight into the mma operand registers (no shared memory);
//   D (16 x 8, float32) -> lane (g, tig) ends with Re/Im S_g of steps t0 + 2 tig and t0 + 2 tig + 1.
//
// Precision: every operand is split into two binary16 pieces x = hi + lo (22 significant bits) and
// the product is formed as hi*hi + hi*lo + lo*hi with float32 accumulation (hi*hi in its own
// accumulator); the dropped lo*lo term is below 2^-24.  Measured against the float64 oracle this is
// as accurate as the float32 FFMA chain of k_sarl_v8 (tests/parity.py, sarl_rate_atol).
//
// The per-vehicle queue recursion (SARL:333-358) stays the reference's sequential float64 chain:
// the four lanes that hold one vehicle's eight steps pass DataBuf along by shuffles, so results do
// not depend on where a rollout is cut into launches or tiles.
#pragma once
#include <cuda_fp16.h>

#include "step.cuh"

namespace risvec {

// D += A * B, m16n8k16, binary16 operands, float32 accumulate (legacy tensor path: HMMA in SASS)
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// (x0, x1) -> packed binary16 pairs hi and lo with x = hi + lo up to 2^-22 |x| (and 2^-25 absolute
// in the binary16 subnormal range); x0 goes to the low half (= the lower k / column index)
__device__ __forceinline__ void split_h2(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn(__fsub_rn(x0, back.x), __fsub_rn(x1, back.y));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void split_h2(double x0, double x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn((float)x0, (float)x1);
    const float2 back = __half22float2(h);
    const __half2 l = __floats2half2_rn((float)(x0 - (double)back.x), (float)(x1 - (double)back.y));
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ double shfl_f64(double x, int src) { return __shfl_sync(kFull, x, src); }

struct SarlMmaIn {     // inputs of one 8-step tile, as one lane needs them
    float a0[2], a1[2];  // action rows of vehicle g at steps t0 + 2 tig + {0, 1}
    int arr[2];
};

// KT = k-tiles of 8 RIS elements (M <= 8 KT, M even); FULL = V == 8, M == 8 KT, every trace and the
// arrivals supplied (no per-access predicates).  Block = 4 warps = 4 adjacent envs.
template <int KT, bool FULL>
__global__ void __launch_bounds__(128, 4) k_sarl_mma(Dims d, State s, risvec_params_t p, SarlArgs a) {
    const int lane = threadIdx.x & 31, g = lane >> 2, tig = lane & 3;
    const int E = d.E, V = d.V, M = d.M, T = a.T;
    const int e = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (e >= E) return;  // warps are independent: no block-level synchronisation below
    const bool vact = FULL ? true : (g < V);
    const int vc = FULL ? g : min(g, V - 1);
    const size_t ev = (size_t)e * V + vc;

    // ---- A operand: phasors of vehicle g at the elements m = 8 j + 2 tig + {0, 1} of every k-tile
    // (fragment columns 2 tig, 2 tig + 1 hold element 8 j + 2 tig, columns 2 tig + 8, + 9 element
    //  8 j + 2 tig + 1: the same element order the B fragments below use)
    uint32_t Ah[KT][4], Al[KT][4];
    {
        const double2 z = unit_phasor64(d.angle_BR - s.angle[ev]);  // w(v, m) = z^m, float64 (SARL:134-145)
        const double2 z2 = cmul64(z, z), z4 = cmul64(z2, z2), z8 = cmul64(z4, z4);
        double2 w = cpow64(z, 2u * (unsigned)tig);
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            const int ma = 8 * j + 2 * tig;
            double2 wa = w, wb = cmul64(w, z);
            if (!(vact && (FULL || ma < M))) wa = make_double2(0.0, 0.0);
            if (!(vact && (FULL || ma + 1 < M))) wb = make_double2(0.0, 0.0);
            split_h2(wa.x, -wa.y, Ah[j][0], Al[j][0]);  // row g     (Re S_g): ( Re w, -Im w)
            split_h2(wa.y, wa.x, Ah[j][1], Al[j][1]);   // row g + 8 (Im S_g): ( Im w,  Re w)
            split_h2(wb.x, -wb.y, Ah[j][2], Al[j][2]);
            split_h2(wb.y, wb.x, Ah[j][3], Al[j][3]);
            w = cmul64(w, z8);
        }
    }
    double buf = s.databuf[ev];  // replicated over the 4 lanes of vehicle g
    const float coef = (float)(s.amp[ev] / (kSigma * kSigma));  // SARL:157-159
    const long long step0 = s.step_ctr[e];
    const float c_dt = (float)(p.time_fast * p.bandwidth * 1000.0);
    const float c_dp = (float)(cbrt(1.0 / p.k) * p.time_fast / p.L / 1000.0);  // SARL:331
    const float c_rev = (float)(1000.0 * p.L / p.time_fast * cbrt(p.k));        // SARL:318-319
    const float t1 = (float)p.t_factor1, t2 = (float)p.t_factor2, pen1 = (float)p.penalty1, pen2 = (float)p.penalty2;
    const float invV = 1.0f / (float)V;
    const float lam = (float)p.rate;
    const double tf = p.time_fast;

    const size_t sM = (size_t)E * M, s2V = (size_t)E * 2 * V, sV = (size_t)E * V;
    const float* const ph_b = a.phase + (size_t)e * M + 2 * tig;
    const float* const ac_b = a.action + (size_t)e * 2 * V + vc;
    const int* const ar_b = (FULL || a.arrivals != nullptr) ? a.arrivals + ev : nullptr;
    const int Tm1 = T - 1;

    auto load_phases = [&](float2 (&ph)[KT], int t0) {
        const float* q = ph_b + (size_t)min(t0 + g, Tm1) * sM;  // B fragment column g = step t0 + g
#pragma unroll
        for (int j = 0; j < KT; ++j)
            ph[j] = (FULL || 8 * j + 2 * tig < M) ? __ldg(reinterpret_cast<const float2*>(q + 8 * j))
                                                  : make_float2(0.f, 0.f);
    };
    auto load_scalars = [&](SarlMmaIn& in, int t0) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const size_t t = (size_t)min(t0 + 2 * tig + i, Tm1);
            in.a0[i] = __ldg(ac_b + t * s2V);
            in.a1[i] = __ldg(ac_b + t * s2V + V);
            in.arr[i] = (FULL || ar_b != nullptr) ? __ldg(ar_b + t * sV) : 0;
        }
    };
    // L2 prefetch of the env's rows of a later tile: lane (g, tig) covers sectors of row t0 + g
    auto prefetch_tile = [&](int t0) {
        const size_t t = (size_t)min(t0 + g, Tm1);
        const float* q = a.phase + t * sM + (size_t)e * M + tig * 8;  // M <= 64: at most two sectors per lane
        if (FULL || tig * 8 < M) asm volatile("prefetch.global.L2 [%0];" ::"l"(q));
        if (tig * 8 + 32 < (FULL ? 8 * KT : M)) asm volatile("prefetch.global.L2 [%0];" ::"l"(q + 32));
        if (tig * 8 < 2 * V) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.action + t * s2V + (size_t)e * 2 * V + tig * 8));
        if (tig == 3 && (FULL || ar_b != nullptr)) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.arrivals + t * sV + (size_t)e * V));
    };

    // values of the two steps of the latest tile (the owner of step T - 1 writes them to the state)
    float o_rate[2] = {0.f, 0.f}, o_dt[2] = {0.f, 0.f}, o_dp[2] = {0.f, 0.f}, o_overp[2] = {0.f, 0.f};
    float o_overd[2] = {0.f, 0.f}, o_rew[2] = {0.f, 0.f};
    int o_arr[2] = {0, 0};

    auto tile = [&](const float2 (&ph)[KT], const SarlMmaIn& in, int t0) {
        // ---- cascaded reduction of 8 steps on the tensor cores
        float accM[4] = {0.f, 0.f, 0.f, 0.f}, accX[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < KT; ++j) {
            float2 sn, cs;
            sincos_fast2(ph[j], &sn, &cs);  // theta = exp(j*phase) of elements 8 j + 2 tig + {0, 1} (SARL:125-131)
            uint32_t b0h, b0l, b1h, b1l;
            split_h2(cs.x, sn.x, b0h, b0l);
            split_h2(cs.y, sn.y, b1h, b1l);
            mma_16816(accM, Ah[j], b0h, b1h);
            mma_16816(accX, Ah[j], b0l, b1l);
            mma_16816(accX, Al[j], b0h, b1h);
        }
        // lane (g, tig): S_g of steps t0 + 2 tig (acc[0] + j acc[2]) and t0 + 2 tig + 1 (acc[1] + j acc[3])
        double dd[2], inc[2];
        float dp[2], rate[2], dt[2];
        int arr[2];
        bool ok[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int t = t0 + 2 * tig + i;
            ok[i] = t < T;
            const float re = __fadd_rn(accM[i], accX[i]), im = __fadd_rn(accM[2 + i], accX[2 + i]);
            const float g2 = __fmaf_rn(re, re, __fmul_rn(im, im));
            rate[i] = log1p_sfu(__fmul_rn(in.a0[i], __fmul_rn(coef, g2)));  // natural log, SARL:159
            dt[i] = __fmul_rn(rate[i], c_dt);
            dp[i] = __fmul_rn(cbrt_sfu(in.a1[i]), c_dp);                    // SARL:331
            arr[i] = in.arr[i];
            if (!FULL && ar_b == nullptr) arr[i] = (vact && ok[i]) ? draw_arrival(d, e, vc, step0 + t, lam) : 0;
            // steps past the end of the rollout are the identity of the recursion (DataBuf >= 0)
            dd[i] = ok[i] ? __dadd_rn((double)dt[i], (double)dp[i]) : 0.0;
            inc[i] = ok[i] ? __dmul_rn(__dmul_rn((double)arr[i], tf), 1000.0) : 0.0;
        }
        // ---- DataBuf recursion (SARL:333-356), sequential float64 in the reference's order: the lanes
        // tig = 0..3 of vehicle g hold steps (0,1), (2,3), (4,5), (6,7); the value is passed along.
        double x = buf, xin = buf;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double r0 = __dsub_rn(x, dd[0]);
            const double y = __dadd_rn(r0 < 0.0 ? 0.0 : r0, inc[0]);
            const double r1 = __dsub_rn(y, dd[1]);
            const double x2 = __dadd_rn(r1 < 0.0 ? 0.0 : r1, inc[1]);
            x = shfl_f64(x2, (lane & ~3) | q);  // the lane that really holds steps (2q, 2q + 1)
            if (q < 3 && tig == q + 1) xin = x;
        }
        buf = x;
        // ---- my two steps again from their true start value: overflow terms, reward, traces
        double cur = xin;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int t = t0 + 2 * tig + i;
            const double raw = __dsub_rn(cur, dd[i]);  // SARL:334
            const bool neg = raw < 0.0;
            const float b = __fmul_rn((float)fmax(0.0, raw + (double)dp[i]), c_rev);
            const float overp = neg ? __fsub_rn(in.a1[i], __fmul_rn(__fmul_rn(b, b), b)) : 0.f;  // SARL:336-339
            const float overd = neg ? (float)(-raw) : 0.f;
            const double nb = neg ? 0.0 : raw;
            const float base = __fsub_rn(-__fmul_rn(t1, __fadd_rn(in.a0[i], in.a1[i])), __fmul_rn(t2, (float)nb));
            const float pen = (nb > 0.0) ? pen1 : ((overd > 2.0f) ? pen2 : 0.f);  // SARL:343-352
            float ru = vact ? __fsub_rn(base, pen) : 0.f;
            ru += __shfl_xor_sync(kFull, ru, 4);   // mean over the vehicles: same tree as seg_sum<8>
            ru += __shfl_xor_sync(kFull, ru, 8);
            ru += __shfl_xor_sync(kFull, ru, 16);
            const float rew = __fmul_rn(ru, invV);
            cur = __dadd_rn(nb, inc[i]);  // SARL:354-356
            if (vact && ok[i]) {
                const size_t o = (size_t)t * sV + ev;
                if (FULL || a.out.DataBuf) a.out.DataBuf[o] = (float)cur;
                if (FULL || a.out.data_t) a.out.data_t[o] = dt[i];
                if (FULL || a.out.data_p) a.out.data_p[o] = dp[i];
                if (FULL || a.out.over_power) a.out.over_power[o] = overp;
                if (FULL || a.out.over_data) a.out.over_data[o] = overd;
                if (FULL || a.out.rate) a.out.rate[o] = rate[i];
                if (g == 0 && (FULL || a.out.reward)) a.out.reward[(size_t)t * E + e] = rew;
            }
            o_rate[i] = rate[i]; o_dt[i] = dt[i]; o_dp[i] = dp[i]; o_overp[i] = overp; o_overd[i] = overd;
            o_rew[i] = rew; o_arr[i] = arr[i];
        }
    };

    // ---- software pipeline over tiles: the inputs of tile k + 1 are in flight (registers) and tile
    // k + 3 is being pulled into L2 while tile k computes; two register sets alternate roles
    const int NT = (T + 7) >> 3;
    float2 phX[KT], phY[KT];
    SarlMmaIn inX, inY;
    load_phases(phX, 0);
    load_scalars(inX, 0);
    prefetch_tile(8);
    prefetch_tile(16);
    int k = 0;
    for (; k + 2 <= NT; k += 2) {
        prefetch_tile(8 * (k + 3));
        load_phases(phY, 8 * (k + 1));
        load_scalars(inY, 8 * (k + 1));
        tile(phX, inX, 8 * k);
        prefetch_tile(8 * (k + 4));
        load_phases(phX, 8 * (k + 2));  // clamped past the end: unused
        load_scalars(inX, 8 * (k + 2));
        tile(phY, inY, 8 * (k + 1));
    }
    if (k < NT) tile(phX, inX, 8 * k);

    // ---- registers -> state (what the reference object holds after the last step)
    for (int m = lane; m < M; m += 32)  // elements_phase_shift_real = the last action_phase (SARL:128)
        s.phase_real[(size_t)e * M + m] = __ldg(a.phase + (size_t)Tm1 * sM + (size_t)e * M + m);
    const int last = Tm1 & 7;  // position of step T - 1 in its tile: lane tig = last / 2, slot last % 2
    if (tig == (last >> 1) && vact) {
        const int i = last & 1;
        s.databuf[ev] = buf;
        s.rate[ev] = i ? o_rate[1] : o_rate[0];
        s.data_t[ev] = i ? o_dt[1] : o_dt[0];
        s.data_p[ev] = i ? o_dp[1] : o_dp[0];
        s.over_power[ev] = i ? o_overp[1] : o_overp[0];
        s.over_data[ev] = i ? o_overd[1] : o_overd[0];
        s.data_r[ev] = i ? o_arr[1] : o_arr[0];
        if (g == 0) {
            s.reward[e] = i ? o_rew[1] : o_rew[0];
            s.step_ctr[e] = step0 + T;
        }
    }
}

}  // namespace risvec
