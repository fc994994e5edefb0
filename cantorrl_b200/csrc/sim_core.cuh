// Device-side core of the path simulator, shared by K1 (path_sim.cu) and the episode-fused rollout (rollout.cu):
// Philox -> Box-Muller normals, the log-Euler GBM / Heston step and the float32 ATM Black-Scholes price.
// Both kernels call the SAME functions on the SAME counters, so an on-the-fly rollout reproduces, bit for bit,
// a replay over a book that K1 wrote.
//
// Reference semantics: src/sim/rbergomi_sim.py:454-464 (step), :418 (ATM strike), :19 (tenor);
// src/sim/option_price_assignment.py:10-21 (Black-Scholes).
#pragma once
#include "bs_math.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace cantor {

constexpr unsigned kStreamPaths = 0x50415448u;   // "PATH": 4th counter word of the path-simulation stream

struct SimConsts {
    float s0, v0, r, dt, sqrt_dt, kappa, theta, sigma_v, rho, rho_c;
    float drift_gbm, vol_gbm;            // (r - v0/2) dt, sqrt(v0 dt)
    float drift_gbm_l2, vol_gbm_l2;      // the same two times log2(e): the step's exponential is one ex2
    float tenor, sqrt_tenor, disc;       // option tenor, its sqrt, exp(-r tenor)
    float inv_sqrt_tenor, inv_disc;
    GreekConsts g;                       // the simulator's own (r, tenor) in the greeks' terms (atm_quote_f32 fallback)
    unsigned seed_lo, seed_hi;
    long long path_offset;
    int model, reprice, T;
};

// Box-Muller on two Philox words (oracle/sim_oracle.py: philox_normals), entirely on the SFU: lg2, sqrt, sin, cos are
// one MUFU instruction each (absolute error ~2^-21 on sin / cos over [0, 2 pi], ~2^-22 on lg2), i.e. ~1e-6 absolute on
// a normal draw -- two orders below the float32 parity tolerance of the paths, invisible to their distribution.
__device__ __forceinline__ float mufu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void box_muller(unsigned x0, unsigned x1, float& n0, float& n1) {
    const float u1 = ((float)x0 + 1.0f) * 2.3283064365386963e-10f;       // (0, 1]
    const float u2 = (float)x1 * 2.3283064365386963e-10f;                 // [0, 1]
    const float rad = mufu_sqrt(-2.0f * kLn2f * mufu_lg2(u1));            // sqrt(-2 ln u1)
    const float ang = 6.283185307179586f * u2;
    n0 = rad * __cosf(ang);
    n1 = rad * __sinf(ang);
}

// ATM call / put, float32: K = rint(S), sigma = sqrt(max(v, 0)) floored at 1e-8 (option_price_assignment.py:10-21).
__device__ __forceinline__ void atm_call_put_f32(float S, float v, const SimConsts& k, float& call, float& put) {
    const float K = rintf(S);
    const float sigma = fmaxf(sqrtf(fmaxf(v, 0.0f)), 1e-8f);
    const float sst = sigma * k.sqrt_tenor;
    const float d1 = (logf(S / K) + (k.r + 0.5f * sigma * sigma) * k.tenor) / sst;
    const float d2 = d1 - sst;
    float c1, c1m, c2, c2m;
    normal_pdf_cdf(d1, &c1, &c1m);
    normal_pdf_cdf(d2, &c2, &c2m);
    const float kd = K * k.disc;
    call = S * c1 - kd * c2;
    put = S * c1m - kd * c2m;            // K disc Phi(-d2) - S Phi(-d1), with Phi(-x) = -(Phi(x) - 1)
}

// ATM call / put AND the ATM greeks of the same state from one evaluation of d1: the price shares phi(d1), Phi(d1) with
// the greeks (bs_math.cuh: atm_core_f32) and gets Phi(d2) without a second exponential through
// S phi(d1) = K e^{-rT} phi(d2).  Used by K1 (prices only; the greeks are dead code there) and by the on-the-fly
// rollout (both), so a replayed K1 book and an on-the-fly rollout see bit-identical marks.
// A variance below the greeks' floor (v < 1e-8: sigma floors differ, hedging_env_v2.py:84 vs option_price_assignment.py:12)
// or a degenerate price takes the two separate evaluations.
struct AtmQuote {
    float call, put;
    Greeks g;
};
__device__ __forceinline__ AtmQuote atm_quote_f32(float S, float v, const SimConsts& k) {
    AtmQuote q;
    const float K = rintf(S);
    if (v >= 1e-8f && S > 1e-6f && k.tenor > 1e-6f) {
        const AtmCore c = atm_core_f32(S, K, v, k.r, k.tenor, k.inv_sqrt_tenor);
        const float sst = (v * c.rs) * k.sqrt_tenor;                       // sigma sqrt(T), sigma = v / sqrt(v)
        const float d2 = c.d1 - sst;
        const float kd = fmaxf(K, 1e-6f) * k.disc;
        const float pdf2 = c.pdf * (S * mufu_rcp(kd));
        const float t2 = mufu_rcp(fmaf(0.2316419f, fabsf(d2), 1.0f));
        float poly = fmaf(t2, 1.330274429f, -1.821255978f);
        poly = fmaf(t2, poly, 1.781477937f);
        poly = fmaf(t2, poly, -0.356563782f);
        poly = fmaf(t2, poly, 0.319381530f);
        const float q2 = (c.pdf > 0.f) ? pdf2 * (poly * t2) : 0.f;
        const bool pos2 = d2 >= 0.f;
        const float c2 = pos2 ? 1.0f - q2 : q2;
        const float c2m = pos2 ? -q2 : q2 - 1.0f;
        q.call = fmaf(S, c.cdf, -kd * c2);
        q.put = fmaf(S, c.cdf_m1, -kd * c2m);
        q.g.call_delta = c.cdf;
        q.g.put_delta = c.cdf_m1;
        q.g.gamma = (S < 1e-9f * c.inv_sst) ? 0.f : c.pdf * c.inv_sst * mufu_rcp(S);
    } else {
        atm_call_put_f32(S, v, k, q.call, q.put);
        q.g = atm_greeks_f32(S, K, v, k.g);
    }
    return q;
}

// The four normals of Philox call `call` of global path `gp` (two Box-Muller pairs).
__device__ __forceinline__ void path_normals(const SimConsts& k, unsigned long long gp, unsigned call, float (&z)[4]) {
    const uint4 x = philox4x32_10(make_uint4((unsigned)gp, (unsigned)(gp >> 32), call, kStreamPaths),
                                  make_uint2(k.seed_lo, k.seed_hi));
    box_muller(x.x, x.y, z[0], z[1]);
    box_muller(x.z, x.w, z[2], z[3]);
}

// One log-Euler day (rbergomi_sim.py:454-464).  MODEL 0: GBM, one normal; MODEL 1: Heston full truncation, two.
template <int MODEL>
__device__ __forceinline__ void sim_advance(const SimConsts& k, float& S, float& v, float z1, float z2) {
    // the growth factor is one FFMA + one MUFU ex2 (2 ulp): expf's range reduction and fix-ups were 16 of the 130
    // instructions of a path-step; 252 such factors drift < 1e-5 relative from the correctly rounded product
    if (MODEL == 0) {
        S = fmaxf(S * mufu_ex2(fmaf(k.vol_gbm_l2, z1, k.drift_gbm_l2)), 1e-8f);
    } else {
        const float vp = fmaxf(v, 0.0f);
        const float sq = mufu_sqrt(vp * k.dt);
        const float zv = k.rho * z1 + k.rho_c * z2;                      // :457
        S = fmaxf(S * mufu_ex2(fmaf(sq, z1, (k.r - 0.5f * vp) * k.dt) * kLog2ef), 1e-8f);  // :460-464
        v = v + k.kappa * (k.theta - vp) * k.dt + k.sigma_v * sq * zv;   // full-truncation Euler
    }
}

// GreekConsts of the simulator's own (r, tenor): what atm_quote_f32 needs for its rare separate-evaluation branch.
inline GreekConsts sim_greek_consts(const cantor_sim_params* p) {
    GreekConsts g;
    g.r_f = (float)p->r; g.T_f = (float)p->tenor; g.T_d = p->tenor; g.sqrtT_d = sqrt(p->tenor);
    g.inv_sqrtT_f = p->tenor > 0 ? (float)(1.0 / sqrt(p->tenor)) : 0.f;
    g.record_metrics = 1;
    return g;
}

inline void fill_sim_consts(const cantor_sim_params* p, int T, SimConsts* k) {
    k->s0 = (float)p->s0; k->v0 = (float)p->v0; k->r = (float)p->r; k->dt = (float)p->dt;
    k->sqrt_dt = (float)sqrt(p->dt);
    k->kappa = (float)p->kappa; k->theta = (float)p->theta; k->sigma_v = (float)p->sigma_v; k->rho = (float)p->rho;
    k->rho_c = (float)sqrt(fmax(0.0, 1.0 - p->rho * p->rho));
    k->drift_gbm = (float)((p->r - 0.5 * p->v0) * p->dt);
    k->vol_gbm = (float)sqrt(p->v0 * p->dt);
    k->drift_gbm_l2 = (float)((p->r - 0.5 * p->v0) * p->dt * 1.4426950408889634);
    k->vol_gbm_l2 = (float)(sqrt(p->v0 * p->dt) * 1.4426950408889634);
    k->tenor = (float)p->tenor; k->sqrt_tenor = (float)sqrt(p->tenor); k->disc = (float)exp(-p->r * p->tenor);
    k->inv_sqrt_tenor = (float)(1.0 / sqrt(p->tenor)); k->inv_disc = (float)exp(p->r * p->tenor);
    k->seed_lo = (unsigned)(p->seed & 0xffffffffull); k->seed_hi = (unsigned)(p->seed >> 32);
    k->path_offset = p->path_offset;
    k->model = p->model; k->reprice = p->reprice; k->T = T;
    k->g = sim_greek_consts(p);
}


}  // namespace cantor
