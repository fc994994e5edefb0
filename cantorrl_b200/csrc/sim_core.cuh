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
    float tenor, sqrt_tenor, disc;       // option tenor, its sqrt, exp(-r tenor)
    unsigned seed_lo, seed_hi;
    long long path_offset;
    int model, reprice, T;
};

// Box-Muller on two Philox words (oracle/sim_oracle.py: philox_normals).
__device__ __forceinline__ void box_muller(unsigned x0, unsigned x1, float& n0, float& n1) {
    const float u1 = ((float)x0 + 1.0f) * 2.3283064365386963e-10f;       // (0, 1]
    const float u2 = (float)x1 * 2.3283064365386963e-10f;                 // [0, 1]
    const float rad = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(6.283185307179586f * u2, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
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
    if (MODEL == 0) {
        S = fmaxf(S * expf(k.drift_gbm + k.vol_gbm * z1), 1e-8f);
    } else {
        const float vp = fmaxf(v, 0.0f);
        const float sq = sqrtf(vp * k.dt);
        const float zv = k.rho * z1 + k.rho_c * z2;                      // :457
        S = fmaxf(S * expf((k.r - 0.5f * vp) * k.dt + sq * z1), 1e-8f);  // :460-464
        v = v + k.kappa * (k.theta - vp) * k.dt + k.sigma_v * sq * zv;   // full-truncation Euler
    }
}

inline void fill_sim_consts(const cantor_sim_params* p, int T, SimConsts* k) {
    k->s0 = (float)p->s0; k->v0 = (float)p->v0; k->r = (float)p->r; k->dt = (float)p->dt;
    k->sqrt_dt = (float)sqrt(p->dt);
    k->kappa = (float)p->kappa; k->theta = (float)p->theta; k->sigma_v = (float)p->sigma_v; k->rho = (float)p->rho;
    k->rho_c = (float)sqrt(fmax(0.0, 1.0 - p->rho * p->rho));
    k->drift_gbm = (float)((p->r - 0.5 * p->v0) * p->dt);
    k->vol_gbm = (float)sqrt(p->v0 * p->dt);
    k->tenor = (float)p->tenor; k->sqrt_tenor = (float)sqrt(p->tenor); k->disc = (float)exp(-p->r * p->tenor);
    k->seed_lo = (unsigned)(p->seed & 0xffffffffull); k->seed_hi = (unsigned)(p->seed >> 32);
    k->path_offset = p->path_offset;
    k->model = p->model; k->reprice = p->reprice; k->T = T;
}


}  // namespace cantor
