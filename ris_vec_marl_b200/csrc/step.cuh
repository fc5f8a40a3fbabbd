// Fused T-step rollouts of Environ.step for E env instances (T = 1 is a plain step).
// Reference: Simulation-MARL-BCD/Environment.py:331-372,547-731 and
//            Simulation-SARL/Environment.py:149-171,318-359.
//
// Thread mapping: an env owns VP = pow2ceil(V) adjacent lanes of one warp (lane v = vehicle v),
// so a warp carries 32 / VP envs and every cross-vehicle term (NOMA partner lookup, sum of edge
// cycles, means) is a segmented warp shuffle.  State (DataBuf, MEC queue, gains, phasor table)
// lives in registers for the whole rollout; per step only actions / phases / arrivals stream
// in from HBM and the requested traces stream out.
#pragma once
#include "common.cuh"

namespace risvec {

struct MarlArgs {
    int T;
    const float* action;   // [T,E,2,V]
    const int* partner;    // [E,V]
    const int* ngroups;    // [E]
    const int* arrivals;   // [T,E,V] or null
    risvec_marl_out_t out;
};

struct SarlArgs {
    int T;
    const float* action;  // [T,E,2,V]
    const float* phase;   // [T,E,M]
    const int* arrivals;  // [T,E,V] or null
    risvec_sarl_out_t out;
};

__device__ inline int draw_arrival(const Dims& d, int e, int v, long long step, float lam) {
    const uint4 r = rng_draw(d, e, (unsigned long long)step, (unsigned)v, kRngArrival);
    return poisson_inv(lam, u01f(r.x));
}

// ---------------------------------------------------------------------------------------
// MARL step (row a11 + a10 of SURVEY.md 8a)
// ---------------------------------------------------------------------------------------
template <int VP>
__global__ void __launch_bounds__(128) k_marl_rollout(Dims d, State s, risvec_params_t p, MarlArgs a) {
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int e = gtid / VP, v = gtid % VP;
    const int E = d.E, V = d.V;
    const bool env_ok = e < E;
    const bool act = env_ok && v < V;
    const size_t ev = (size_t)e * V + v;

    // ---- per-rollout state -> registers
    double buf = act ? s.databuf[ev] : 0.0;
    const double g = act ? s.gains[ev] : 0.0;
    const int code = act ? a.partner[ev] : RISVEC_PARTNER_NONE;
    const int ng = env_ok ? a.ngroups[e] : 1;
    double Q = env_ok ? s.mecq[e] : 0.0;
    const long long step0 = env_ok ? s.step_ctr[e] : 0;

    const bool paired = code >= 0;
    const bool second = paired && (code & RISVEC_PARTNER_SECOND);
    int other = paired ? (code & (RISVEC_PARTNER_SECOND - 1)) : v;
    other = min(max(other, 0), VP - 1);
    const int src = (lane & ~(VP - 1)) + other;
    const double g_o = __shfl_sync(kFull, g, src);
    // MARL:355: the first listed user is "near" only if its gain is strictly larger
    const bool first_near = second ? (g_o > g) : (g > g_o);
    const bool near = paired ? (second ? !first_near : first_near) : true;
    const float gn = (float)(g / p.noise_power);
    const float frac = (float)(1.0 / (double)max(1, ng));  // MARL:341-342
    const bool scheduled = code != RISVEC_PARTNER_NONE;

    // ---- constants
    const float ps = (float)p.power_scale, Pmax = (float)p.P_max;
    const float c_dt = (float)(p.time_fast * p.bandwidth * 1000.0);
    const float c_thr = (float)(p.bandwidth * 1000.0);
    double floor_d = p.cpu_share_floor;
    if (!isfinite(floor_d)) floor_d = 0.10;
    floor_d = fmax(0.0, fmin(floor_d, 0.95));
    const float floor_f = (float)floor_d;
    const double Cpb = p.cycles_per_bit;
    const double kbit2cyc_den = Cpb * 1000.0;
    const double edge_cap = p.f_edge_max * p.time_fast;
    const float inv_fedge = (float)(1.0 / (p.f_edge_max + 1e-12));
    const float inv_tf = (float)(1.0 / p.time_fast);
    const float wd = (float)p.w_d, we = (float)p.w_e, pen = (float)p.qos_penalty, clipv = (float)p.reward_clip;
    const float Rmin = (float)p.R_min_bpsHz, Dmax = (float)p.D_max_s;
    const float invV = 1.0f / (float)V;
    const float lam = (float)p.rate;

    float o_rate = 0.f, o_dt = 0.f, o_dp = 0.f, o_rew = 0.f, o_glob = 0.f, o_overp = 0.f;
    int o_arr = 0;

    for (int t = 0; t < a.T; ++t) {
        const size_t tev = ((size_t)t * E + e) * V + v;
        const size_t ta = ((size_t)t * E + e) * 2 * V + v;
        const float a0 = act ? a.action[ta] : 0.f;
        const float a1 = act ? a.action[ta + V] : 0.f;
        int arr = 0;
        if (act) arr = a.arrivals != nullptr ? a.arrivals[tev] : draw_arrival(d, e, v, step0 + t, lam);

        // power projection (MARL:555-561)
        float p0 = fmaxf(a0, 0.f) * ps, p1 = fmaxf(a1, 0.f) * ps;
        const float sm = p0 + p1;
        if (sm > 1.0f) {
            const float den = sm + 1e-12f;
            p0 = p0 / den;
            p1 = p1 / den;
        }
        const float P0 = p0 * Pmax, P1 = p1 * Pmax;
        const float P0_o = __shfl_sync(kFull, P0, src);

        // NOMA / OMA rate (MARL:339-370); log2(1+x) as log1p(x)/ln2 keeps tiny far-user SINRs
        const float sig = P0 * gn;
        const float sinr = near ? sig : sig / (P0_o * gn + 1.0f);
        const float rate = scheduled ? frac * (log1pf(sinr) * 1.44269504088896341f) : 0.f;
        const float data_t = rate * c_dt;

        // local CPU (MARL:572-592) -- float64 with the reference's operation order so that the
        // rounding residues of DataBuf - data_p (which min/max select) are reproduced
        const float share = fmaxf(fminf(fmaxf(a1, 0.f), 1.f), floor_f);
        const double f_local = __dmul_rn((double)share, p.f_local_max);
        const double backlog_kbit = buf;
        const double backlog_cyc = __dmul_rn(__dmul_rn(backlog_kbit, 1000.0), Cpb);
        const double cap = __dmul_rn(f_local, p.time_fast);
        const double used = fmin(cap, backlog_cyc);
        const double local_done = __ddiv_rn(used, kbit2cyc_den);
        const double remaining = fmax(0.0, __dsub_rn(backlog_kbit, local_done));
        const double off = fmin((double)data_t, remaining);  // MARL:595-596
        const float thr = rate * c_thr;
        const float off_f = (float)off;
        const float t_tx = off_f / (thr + 1e-12f);  // MARL:599-601

        // MEC FCFS queue, one per env (MARL:604-610)
        const double edge_in = __dmul_rn(__dmul_rn(off, 1000.0), Cpb);
        const double edge_sum = seg_sum<VP>(edge_in);
        const double q_before = Q;
        Q = Q + edge_sum;
        const double served = fmin(edge_cap, Q);
        Q = Q - served;

        buf = fmax(0.0, __dsub_rn(buf, __dadd_rn(local_done, off)));  // MARL:617-618

        // delays (MARL:622-633), energy (MARL:659-661)
        const float f_local_f = (float)f_local;
        const float d_local = (float)fmax(0.0, backlog_cyc - edge_in) / (f_local_f + 1e-12f);
        const float edge_in_f = (float)edge_in;
        const float sh = edge_in_f / (float)(edge_sum + 1e-12);
        const float d_eq = sh * ((float)q_before * inv_fedge);
        const float d_ec = edge_in_f * inv_fedge;
        const float delay = ((d_local + t_tx) + d_eq) + d_ec;
        const float E_tx = P0 * t_tx;
        const float E_loc = (float)(p.k * f_local * f_local) * (float)used;
        const float energy = E_tx + E_loc;

        // QoS penalty and reward (MARL:669-703)
        const bool viol = p.qos_enable && ((rate < Rmin) || (delay > Dmax));
        float rew = -(wd * delay + we * energy) - (viol ? pen : 0.f);
        rew = fminf(fmaxf(rew, -clipv), clipv);
        const float glob = seg_sum<VP>(act ? rew : 0.f) * invV;  // MARL:721
        const float overp = fmaxf(0.f, (P0 + P1) - Pmax);        // MARL:727-729

        // arrivals (MARL:717-719): DataBuf += (data_r * time_fast) * 1000
        buf = __dadd_rn(buf, __dmul_rn(__dmul_rn((double)arr, p.time_fast), 1000.0));

        const bool last = (t == a.T - 1);
        if (a.out.stats != nullptr || a.out.last_power != nullptr || last) {
            // last_* scalars (MARL:612-614,636-656,677,706-711)
            const float m_delay = seg_sum<VP>(act ? delay : 0.f) * invV;
            const float m_energy = seg_sum<VP>(act ? energy : 0.f) * invV;
            const float m_dl = seg_sum<VP>(act ? d_local : 0.f) * invV;
            const float m_dq = seg_sum<VP>(act ? d_eq : 0.f) * invV;
            const float m_dc = seg_sum<VP>(act ? d_ec : 0.f) * invV;
            const float m_ttx = seg_sum<VP>(act ? t_tx : 0.f) * invV;
            const float m_back = seg_sum<VP>(act ? (float)backlog_kbit : 0.f) * invV;
            const float util = (float)(used / (cap + 1e-12));
            const float m_util = seg_sum<VP>(act ? util : 0.f) * invV;
            const float m_viol = seg_sum<VP>((act && viol) ? 1.f : 0.f) * invV;
            const float s_off = seg_sum<VP>(act ? off_f : 0.f);
            const float s_loc = seg_sum<VP>(act ? (float)local_done : 0.f);
            const float mec_util = (float)(served / (edge_cap + 1e-12));
            const float q_f = (float)Q;
            auto stat_of = [&](int col) -> float {
                switch (col) {
                    case RISVEC_S_DELAY_MEAN: return m_delay;
                    case RISVEC_S_ENERGY_MEAN: return m_energy;
                    case RISVEC_S_DELAY_LOCAL_MEAN: return m_dl;
                    case RISVEC_S_DELAY_EDGE_Q_MEAN: return m_dq;
                    case RISVEC_S_DELAY_EDGE_C_MEAN: return m_dc;
                    case RISVEC_S_T_TX_MEAN: return m_ttx;
                    case RISVEC_S_BACKLOG_KBIT_MEAN: return m_back;
                    case RISVEC_S_MEC_UTILIZATION: return mec_util;
                    case RISVEC_S_LOCAL_UTIL_MEAN: return m_util;
                    case RISVEC_S_QOS_VIOLATION: return m_viol;
                    case RISVEC_S_OFF_KBIT_SUM: return s_off;
                    case RISVEC_S_LOCAL_KBIT_SUM: return s_loc;
                    case RISVEC_S_MEC_QUEUE_CYCLES: return q_f;
                    default: return 0.f;
                }
            };
            // lane v of the env writes stat columns v, v + VP, ... (coalesced rows of NSTAT floats)
            if (env_ok) {
                for (int col = v; col < RISVEC_NSTAT; col += VP) {
                    const float val = stat_of(col);
                    if (a.out.stats != nullptr) a.out.stats[((size_t)t * E + e) * RISVEC_NSTAT + col] = val;
                    if (last) s.stats[(size_t)e * RISVEC_NSTAT + col] = val;
                }
            }
            if (act) {  // last_power_W = [E_tx, E_loc] / time_fast (MARL:664-666)
                const float ptx = E_tx * inv_tf, ploc = E_loc * inv_tf;
                if (a.out.last_power != nullptr) {
                    a.out.last_power[ta] = ptx;
                    a.out.last_power[ta + V] = ploc;
                }
                if (last) {
                    s.last_power[(size_t)e * 2 * V + v] = ptx;
                    s.last_power[(size_t)e * 2 * V + V + v] = ploc;
                }
            }
        }
        if (act) {
            if (a.out.reward_user != nullptr) a.out.reward_user[tev] = rew;
            if (a.out.DataBuf != nullptr) a.out.DataBuf[tev] = (float)buf;
            if (a.out.data_t != nullptr) a.out.data_t[tev] = data_t;
            if (a.out.data_p != nullptr) a.out.data_p[tev] = (float)local_done;
            if (a.out.rate != nullptr) a.out.rate[tev] = rate;
            if (a.out.over_power != nullptr) a.out.over_power[tev] = overp;
            if (v == 0 && a.out.reward != nullptr) a.out.reward[(size_t)t * E + e] = glob;
        }
        o_rate = rate; o_dt = data_t; o_dp = (float)local_done; o_rew = rew; o_glob = glob; o_overp = overp;
        o_arr = arr;
    }

    // ---- registers -> state
    if (act && a.T > 0) {
        s.databuf[ev] = buf;
        s.rate[ev] = o_rate;
        s.data_t[ev] = o_dt;
        s.data_p[ev] = o_dp;
        s.reward_user[ev] = o_rew;
        s.over_power[ev] = o_overp;
        s.data_r[ev] = o_arr;
        if (v == 0) {
            s.mecq[e] = Q;
            s.reward[e] = o_glob;
            s.step_ctr[e] = step0 + a.T;
        }
    }
}

// ---------------------------------------------------------------------------------------
// SARL step (row a12): per step the V x M cascaded RIS reduction
//   S_v = sum_m exp(j*phase_m) * phasor(v, m),  rate_v = ln(1 + P0_v * amp_v * |S_v|^2 / sigma^2)
// The geometry phasor table (depends only on positions) is built once per launch in float64
// and kept in registers: lane (env, v) of warp w holds elements [w*slice, (w+1)*slice).
// WPE warps share one env group and combine their partial sums through shared memory.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ void phasor_f32(double x, float* re, float* im) {
    double s64, c64;
    sincospi(x, &s64, &c64);
    *re = (float)c64;
    *im = (float)s64;
}

template <int VP, int MPL, int WPE>
__global__ void __launch_bounds__(32 * WPE) k_sarl_rollout(Dims d, State s, risvec_params_t p, SarlArgs a) {
    constexpr int EPW = 32 / VP;  // envs per warp (= per block)
    extern __shared__ float2 sarl_smem[];
    const int E = d.E, V = d.V, M = d.M;
    const int MS = M + 2;  // padded env stride of the theta table (bank spread)
    float2* th = sarl_smem;                // [EPW][MS]
    float2* part = sarl_smem + EPW * MS;   // [WPE][32] (WPE > 1 only)
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int el = lane / VP, v = lane % VP;
    const int e0 = blockIdx.x * EPW;
    const int e = e0 + el;
    const bool env_ok = e < E;
    const bool act = env_ok && v < V;
    const size_t ev = (size_t)e * V + v;
    const int slice = (M + WPE - 1) / WPE;
    const int m0 = w * slice;
    const int m1 = min(M, m0 + slice);

    // ---- geometry phasor table -> registers (float64 argument reduction, float32 storage)
    float wr[MPL], wi[MPL];
    {
        const double delta = act ? d.angle_BR - s.angle[ev] : 0.0;
#pragma unroll
        for (int i = 0; i < MPL; ++i) {
            const int m = m0 + i;
            float re = 0.f, im = 0.f;
            if (act && m < m1) phasor_f32((double)m * delta, &re, &im);
            wr[i] = re;
            wi[i] = im;
        }
    }
    const bool cphase = (w == 0);
    double buf = (act && cphase) ? s.databuf[ev] : 0.0;
    const float coef = act ? (float)(s.amp[ev] / (kSigma * kSigma)) : 0.f;  // SARL:157-159
    const long long step0 = env_ok ? s.step_ctr[e] : 0;

    const float c_dt = (float)(p.time_fast * p.bandwidth * 1000.0);
    const float c_dp = (float)(cbrt(1.0 / p.k) * p.time_fast / p.L / 1000.0);       // SARL:331
    const float c_rev = (float)(1000.0 * p.L / p.time_fast * cbrt(p.k));             // SARL:318-319
    const float t1 = (float)p.t_factor1, t2 = (float)p.t_factor2, pen1 = (float)p.penalty1, pen2 = (float)p.penalty2;
    const float invV = 1.0f / (float)V;
    const float lam = (float)p.rate;
    const int n_env_here = min(EPW, E - e0);

    float o_rate = 0.f, o_dt = 0.f, o_dp = 0.f, o_overp = 0.f, o_overd = 0.f, o_rew = 0.f;
    int o_arr = 0;

    for (int t = 0; t < a.T; ++t) {
        // (1) theta_m = exp(j*phase_m) for the block's envs (SARL:125-131); the EPW rows of
        //     phase[t] are contiguous in HBM
        const float* ph_t = a.phase + ((size_t)t * E + e0) * M;
        const bool last = (t == a.T - 1);
        for (int idx = threadIdx.x; idx < n_env_here * M; idx += 32 * WPE) {
            const float ph = ph_t[idx];
            float sn, cs;
            sincosf(ph, &sn, &cs);
            const int el2 = idx / M, m = idx - el2 * M;
            th[el2 * MS + m] = make_float2(cs, sn);
            if (last) s.phase_real[(size_t)e0 * M + idx] = ph;
        }
        if (WPE > 1) __syncthreads(); else __syncwarp();

        // (2) cascaded reduction over this warp's element slice
        float sr = 0.f, si = 0.f;
        const float2* th_e = th + el * MS + m0;
#pragma unroll
        for (int i = 0; i < MPL; ++i) {
            if (m0 + i < m1) {
                const float2 tq = th_e[i];
                sr = fmaf(tq.x, wr[i], sr);
                sr = fmaf(-tq.y, wi[i], sr);
                si = fmaf(tq.x, wi[i], si);
                si = fmaf(tq.y, wr[i], si);
            }
        }
        if (WPE > 1) {
            part[w * 32 + lane] = make_float2(sr, si);
            __syncthreads();
            if (cphase) {
                sr = 0.f; si = 0.f;
#pragma unroll
                for (int k = 0; k < WPE; ++k) {
                    const float2 q = part[k * 32 + lane];
                    sr += q.x; si += q.y;
                }
            }
        } else {
            __syncwarp();
        }

        // (3) per-vehicle queue update and reward (SARL:327-358), first warp of the block
        if (cphase) {
            const size_t tev = ((size_t)t * E + e) * V + v;
            const size_t ta = ((size_t)t * E + e) * 2 * V + v;
            const float a0 = act ? a.action[ta] : 0.f;
            const float a1 = act ? a.action[ta + V] : 0.f;
            int arr = 0;
            if (act) arr = a.arrivals != nullptr ? a.arrivals[tev] : draw_arrival(d, e, v, step0 + t, lam);

            const float g2 = sr * sr + si * si;
            const float rate = log1pf(a0 * (coef * g2));  // natural log, SARL:159
            const float data_t = rate * c_dt;
            const float data_p = cbrtf(a1) * c_dp;
            double nb = buf - ((double)data_t + (double)data_p);  // SARL:334
            float overp = 0.f, overd = 0.f;
            if (nb < 0.0) {  // SARL:336-339
                const float b = (float)fmax(0.0, nb + (double)data_p) * c_rev;
                overp = a1 - b * b * b;
                overd = (float)(-nb);
                nb = 0.0;
            }
            const float nbf = (float)nb;
            const float base = -(t1 * (a0 + a1)) - (t2 * nbf);
            const float ru = (nb > 0.0) ? base - pen1 : ((overd > 2.0f) ? base - pen2 : base);  // SARL:343-352
            const float rew = seg_sum<VP>(act ? ru : 0.f) * invV;
            buf = __dadd_rn(nb, __dmul_rn(__dmul_rn((double)arr, p.time_fast), 1000.0));  // SARL:354-356

            if (act) {
                if (a.out.DataBuf != nullptr) a.out.DataBuf[tev] = (float)buf;
                if (a.out.data_t != nullptr) a.out.data_t[tev] = data_t;
                if (a.out.data_p != nullptr) a.out.data_p[tev] = data_p;
                if (a.out.over_power != nullptr) a.out.over_power[tev] = overp;
                if (a.out.over_data != nullptr) a.out.over_data[tev] = overd;
                if (a.out.rate != nullptr) a.out.rate[tev] = rate;
                if (v == 0 && a.out.reward != nullptr) a.out.reward[(size_t)t * E + e] = rew;
            }
            o_rate = rate; o_dt = data_t; o_dp = data_p; o_overp = overp; o_overd = overd; o_rew = rew; o_arr = arr;
        }
    }

    if (cphase && act && a.T > 0) {
        s.databuf[ev] = buf;
        s.rate[ev] = o_rate;
        s.data_t[ev] = o_dt;
        s.data_p[ev] = o_dp;
        s.over_power[ev] = o_overp;
        s.over_data[ev] = o_overd;
        s.data_r[ev] = o_arr;
        if (v == 0) {
            s.reward[e] = o_rew;
            s.step_ctr[e] = step0 + a.T;
        }
    }
}

// =========================================================================================
// SARL fast path for V <= 8, M <= 8 * MPI (BASELINE configs 1-2: V = 8, M = 40, MPI = 5)
// =========================================================================================
// One warp = 4 envs; lane = (env el, part).  For the cascaded reduction the 8 lanes of an env
// split the ELEMENT axis: lane `part` owns elements m = part + 8 i (i < MPI), evaluates
// exp(j*phase_m) for them once, and multiplies them into the partial sums of ALL 8 vehicles
// (geometry phasors of its elements x 8 vehicles sit in registers as packed float2 pairs so
// the MACs issue as FFMA2).  A 3-stage butterfly (reduce-scatter over lanes xor 4, 2, 1) then
// leaves the total S_v on lane part = v, which runs the per-vehicle queue update.
// No shared memory, no block barrier.  U consecutive steps are processed together: their
// state-independent parts (phasors, reduction, rate, data_t, data_p) are independent
// instruction streams the scheduler interleaves; only the DataBuf recursion is sequential.
// Inputs for the next U steps are prefetched into registers while the current ones compute.

// sin/cos of a float32 angle in radians; |x| < ~1e4 (RIS phases live in [0, 2*pi]).
// Cody-Waite reduction to [-pi/4, pi/4] + Cephes minimax polynomials (<= 1 ulp there).
__device__ __forceinline__ void sincos_fast(float x, float* sn, float* cs) {
    const float kf = rintf(x * 0.63661977236758134f);
    float r = fmaf(kf, -1.5707963705062866f, x);
    r = fmaf(kf, 4.3711390001862426e-8f, r);
    const int q = (int)kf;
    const float r2 = r * r;
    float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(ps, r2, -1.6666654611e-1f);
    const float s0 = fmaf(ps * r2, r, r);
    float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(pc, r2, 4.166664568298827e-2f);
    const float c0 = fmaf(pc * r2, r2, fmaf(r2, -0.5f, 1.0f));
    const bool swap = q & 1;
    const float s1 = swap ? c0 : s0, c1 = swap ? s0 : c0;
    *sn = (q & 2) ? -s1 : s1;
    *cs = ((q + 1) & 2) ? -c1 : c1;
}

__device__ __forceinline__ float2 shfl_xor2(float2 x, int o) {
    return make_float2(__shfl_xor_sync(kFull, x.x, o), __shfl_xor_sync(kFull, x.y, o));
}
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }

template <int MPI>
struct SarlStepIn {
    float ph[MPI];
    float a0, a1;
    int arr;
};

template <int MPI, int U>
__global__ void __launch_bounds__(32) k_sarl_v8(Dims d, State s, risvec_params_t p, SarlArgs a) {
    const int warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, el = lane >> 3, part = lane & 7;
    const int E = d.E, V = d.V, M = d.M, T = a.T;
    if (warp * 4 >= E) return;
    const int e = warp * 4 + el, v = part;
    const bool env_ok = e < E;
    const bool act = env_ok && v < V;
    const size_t ev = (size_t)e * V + v;

    // ---- geometry phasors of my elements for all 8 vehicles -> registers
    float2 WX[MPI][4], WY[MPI][4];
    {
        const double my_delta = act ? d.angle_BR - s.angle[ev] : 0.0;
#pragma unroll
        for (int v2 = 0; v2 < 8; ++v2) {
            const double dv = __shfl_sync(kFull, my_delta, (lane & ~7) + v2);
            const bool ok2 = env_ok && v2 < V;
#pragma unroll
            for (int i = 0; i < MPI; ++i) {
                const int m = part + 8 * i;
                float re = 0.f, im = 0.f;
                if (ok2 && m < M) phasor_f32((double)m * dv, &re, &im);
                if (v2 & 1) { WX[i][v2 >> 1].y = re; WY[i][v2 >> 1].y = im; }
                else        { WX[i][v2 >> 1].x = re; WY[i][v2 >> 1].x = im; }
            }
        }
    }
    double buf = act ? s.databuf[ev] : 0.0;
    const float coef = act ? (float)(s.amp[ev] / (kSigma * kSigma)) : 0.f;  // SARL:157-159
    const long long step0 = env_ok ? s.step_ctr[e] : 0;
    const float c_dt = (float)(p.time_fast * p.bandwidth * 1000.0);
    const float c_dp = (float)(cbrt(1.0 / p.k) * p.time_fast / p.L / 1000.0);  // SARL:331
    const float c_rev = (float)(1000.0 * p.L / p.time_fast * cbrt(p.k));        // SARL:318-319
    const float t1 = (float)p.t_factor1, t2 = (float)p.t_factor2, pen1 = (float)p.penalty1, pen2 = (float)p.penalty2;
    const float invV = 1.0f / (float)V;
    const float lam = (float)p.rate;
    const bool b2 = part & 4, b1 = part & 2, b0 = part & 1;

    auto load_in = [&](SarlStepIn<MPI>& in, int t) {
        const float* ph_t = a.phase + ((size_t)t * E + e) * M;
#pragma unroll
        for (int i = 0; i < MPI; ++i) {
            const int m = part + 8 * i;
            in.ph[i] = (env_ok && m < M) ? __ldg(ph_t + m) : 0.f;
        }
        const size_t ta = ((size_t)t * E + e) * 2 * V + v;
        in.a0 = act ? __ldg(a.action + ta) : 0.f;
        in.a1 = act ? __ldg(a.action + ta + V) : 0.f;
        in.arr = (act && a.arrivals != nullptr) ? __ldg(a.arrivals + ((size_t)t * E + e) * V + v) : 0;
    };

    SarlStepIn<MPI> cur[U], nxt[U];
#pragma unroll
    for (int u = 0; u < U; ++u)
        if (u < T) load_in(cur[u], u);

    float o_rate = 0.f, o_dt = 0.f, o_dp = 0.f, o_overp = 0.f, o_overd = 0.f, o_rew = 0.f;
    int o_arr = 0;

    for (int t0 = 0; t0 < T; t0 += U) {
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (t0 + U + u < T) load_in(nxt[u], t0 + U + u);

        // ---- state-independent part of the U steps
        float rate[U], data_t[U], data_p[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float2 RE[4], IM[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) RE[k] = IM[k] = make_float2(0.f, 0.f);
#pragma unroll
            for (int i = 0; i < MPI; ++i) {
                float sn, cs;
                sincos_fast(cur[u].ph[i], &sn, &cs);  // theta_m = exp(j*phase_m), SARL:125-131
                const float2 TX = make_float2(cs, cs), TY = make_float2(sn, sn), NTY = make_float2(-sn, -sn);
#pragma unroll
                for (int k = 0; k < 4; ++k) {  // theta_m * w_vm for the vehicle pair (2k, 2k+1)
                    RE[k] = __ffma2_rn(TX, WX[i][k], RE[k]);
                    RE[k] = __ffma2_rn(NTY, WY[i][k], RE[k]);
                    IM[k] = __ffma2_rn(TX, WY[i][k], IM[k]);
                    IM[k] = __ffma2_rn(TY, WX[i][k], IM[k]);
                }
            }
            // reduce-scatter over the env's 8 lanes: lane `part` ends with vehicle v = part
            float2 r0 = b2 ? RE[2] : RE[0], r1 = b2 ? RE[3] : RE[1];
            float2 i0 = b2 ? IM[2] : IM[0], i1 = b2 ? IM[3] : IM[1];
            r0 = add2(r0, shfl_xor2(b2 ? RE[0] : RE[2], 4)); r1 = add2(r1, shfl_xor2(b2 ? RE[1] : RE[3], 4));
            i0 = add2(i0, shfl_xor2(b2 ? IM[0] : IM[2], 4)); i1 = add2(i1, shfl_xor2(b2 ? IM[1] : IM[3], 4));
            float2 r = add2(b1 ? r1 : r0, shfl_xor2(b1 ? r0 : r1, 2));
            float2 im = add2(b1 ? i1 : i0, shfl_xor2(b1 ? i0 : i1, 2));
            const float sr = (b0 ? r.y : r.x) + __shfl_xor_sync(kFull, b0 ? r.x : r.y, 1);
            const float si = (b0 ? im.y : im.x) + __shfl_xor_sync(kFull, b0 ? im.x : im.y, 1);

            const float g2 = sr * sr + si * si;
            rate[u] = log1pf(cur[u].a0 * (coef * g2));  // natural log, SARL:159
            data_t[u] = rate[u] * c_dt;
            data_p[u] = cbrtf(cur[u].a1) * c_dp;
        }

        // ---- DataBuf recursion and reward, sequential over the U steps (SARL:333-358)
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int t = t0 + u;
            if (t < T) {
                const float a0 = cur[u].a0, a1 = cur[u].a1;
                int arr = cur[u].arr;
                if (act && a.arrivals == nullptr) arr = draw_arrival(d, e, v, step0 + t, lam);
                double nb = buf - ((double)data_t[u] + (double)data_p[u]);
                float overp = 0.f, overd = 0.f;
                if (nb < 0.0) {
                    const float b = (float)fmax(0.0, nb + (double)data_p[u]) * c_rev;
                    overp = a1 - b * b * b;
                    overd = (float)(-nb);
                    nb = 0.0;
                }
                const float nbf = (float)nb;
                const float base = -(t1 * (a0 + a1)) - (t2 * nbf);
                const float ru = (nb > 0.0) ? base - pen1 : ((overd > 2.0f) ? base - pen2 : base);
                const float rew = seg_sum<8>(act ? ru : 0.f) * invV;
                buf = __dadd_rn(nb, __dmul_rn(__dmul_rn((double)arr, p.time_fast), 1000.0));
                if (act) {
                    const size_t tev = ((size_t)t * E + e) * V + v;
                    if (a.out.DataBuf != nullptr) a.out.DataBuf[tev] = (float)buf;
                    if (a.out.data_t != nullptr) a.out.data_t[tev] = data_t[u];
                    if (a.out.data_p != nullptr) a.out.data_p[tev] = data_p[u];
                    if (a.out.over_power != nullptr) a.out.over_power[tev] = overp;
                    if (a.out.over_data != nullptr) a.out.over_data[tev] = overd;
                    if (a.out.rate != nullptr) a.out.rate[tev] = rate[u];
                    if (v == 0 && a.out.reward != nullptr) a.out.reward[(size_t)t * E + e] = rew;
                }
                if (t == T - 1) {
                    o_rate = rate[u]; o_dt = data_t[u]; o_dp = data_p[u]; o_overp = overp; o_overd = overd;
                    o_rew = rew; o_arr = arr;
#pragma unroll
                    for (int i = 0; i < MPI; ++i) {  // elements_phase_shift_real = action_phase
                        const int m = part + 8 * i;
                        if (env_ok && m < M) s.phase_real[(size_t)e * M + m] = cur[u].ph[i];
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = nxt[u];
    }

    if (act && T > 0) {
        s.databuf[ev] = buf;
        s.rate[ev] = o_rate;
        s.data_t[ev] = o_dt;
        s.data_p[ev] = o_dp;
        s.over_power[ev] = o_overp;
        s.over_data[ev] = o_overd;
        s.data_r[ev] = o_arr;
        if (v == 0) {
            s.reward[e] = o_rew;
            s.step_ctr[e] = step0 + T;
        }
    }
}

}  // namespace risvec
