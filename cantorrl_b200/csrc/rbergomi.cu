// Rough-Bergomi path generator and nested-Monte-Carlo ATM pricer -- the reference's own training-data generator
// (src/sim/rbergomi_sim.py), which is the one place the reference itself used a GPU (CuPy: cuRAND + cuFFT + ~10
// elementwise kernels per inner step, 7.6e12 inner path-steps for the shipped configuration).
//
// Reference semantics
//   per-path parameters     :363-367  S0, xi, H, eta, rho = base * (1 + std * N(0,1)) with the clips of :35-40
//   Brownian increments     :377-382  dW1, dW2 = Re / Im ifft(Z) sqrt(M): iid N(0,1) (an orthogonal transform of iid Z)
//   "fractional" driver     :206-229  X = sqrt(2H) eta Re ifft(fft(lambda) * Z),  lambda_k = t_k^(2H) / 2
//   variance                :231-243  v_k = xi exp(X_k - eta^2 t_k^(2H) / 2)
//   log-Euler step          :454-464 / :285-295
//   nested MC ATM price     :246-306  5000 inner paths x 30 steps from (S_t, K = round(S_t), xi := v_t, H, eta, rho)
//
// Because lambda is real and Z = fft(dW1 + i dW2) / sqrt(M), the FFT pipeline collapses EXACTLY to a circular FIR
// filter of the first increment stream (oracle/rbergomi_oracle.py: fgn_conv, checked against the reference to 1e-16):
//       X_k = sqrt(2H) eta / sqrt(M) * sum_n lambda_n dW1[(k - n) mod M]
// so no FFT, no complex arithmetic and no (B, 5000, 32) complex128 intermediates are needed: an inner path is 62
// normals, a 30-tap filter held in registers (900 FFMA), 30 exp2 / sqrt and one log-space accumulation.
//
//   rbergomi_paths_kernel   one CTA per outer path, float64 like the reference (the outer problem is tiny)
//   rbergomi_price_kernel   one CTA per (path, day, call|put); thread q walks inner paths q, q + 128, ...; float32;
//                           block reduction of the payoff sum; writes C or P of the packed book
// Counter-based Philox4x32-10 everywhere: counter = (global path, day | kind << 24, inner path, call# | "RBMC"), so
// the book does not depend on launch geometry, on the day range of a launch, or on the sharding over GPUs.
#include "bs_math.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace cantor {

constexpr unsigned kStreamRbParams = 0x52425000u;   // "RBP\0"
constexpr unsigned kStreamRbMain = 0x52424D00u;     // "RBM\0"
constexpr unsigned kStreamRbInner = 0x52424900u;    // "RBI\0"
constexpr int kInnerSteps = 30;                     // int(T_OPTION_TENOR / DT)  (:250)
constexpr int kInnerM = 32;                         // next_power_of_two(31)     (:262)
constexpr int kPriceThreads = 128;
constexpr int kOuterThreads = 256;
constexpr int kOuterMaxM = 1024;

struct RbConsts {
    double s0, xi, H, eta, rho;
    double p_s0, p_xi, p_H, p_eta, p_rho;
    double min_xi, min_eta, H_lo, H_hi, rho_lo, rho_hi;
    double r, dt;
    float r_f, dt_f, sqrt_dt_f, disc_f;
    int n_mc, shared_draws;
    unsigned seed_lo, seed_hi;
    long long path_offset;
};

__device__ __forceinline__ void box_muller_f64(unsigned x0, unsigned x1, double& n0, double& n1) {
    const double u1 = ((double)x0 + 1.0) * 2.3283064365386963e-10;
    const double u2 = (double)x1 * 2.3283064365386963e-10;
    const double rad = sqrt(-2.0 * log(u1));
    double s, c;
    sincospi(2.0 * u2, &s, &c);
    n0 = rad * c;
    n1 = rad * s;
}
__device__ __forceinline__ float mufu_sqrt_rb(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void box_muller_f32(unsigned x0, unsigned x1, float& n0, float& n1) {
    const float u1 = ((float)x0 + 1.0f) * 2.3283064365386963e-10f;
    const float u2 = (float)x1 * 2.3283064365386963e-10f;
    const float rad = mufu_sqrt_rb(-2.0f * kLn2f * mufu_lg2(u1));
    const float ang = 6.283185307179586f * u2;
    n0 = rad * __cosf(ang);
    n1 = rad * __sinf(ang);
}

// ---- outer generator --------------------------------------------------------------------------------------------
// params_in [5, n_paths] (S0, xi, H, eta, rho) and dW1_in / dW2_in [n_paths, M] replace the Philox draws when given
// (parity runs on the reference's exported values).  Outputs: packed book S, v (C, P = 0), per-path parameters
// path_params [5, n_paths] (float64), optional float64 copies paths64 / v64 [n_paths, T + 1].
__global__ void __launch_bounds__(kOuterThreads)
rbergomi_paths_kernel(const RbConsts k, int n_paths, int T, int M, const double* __restrict__ params_in,
                      const double* __restrict__ dW1_in, const double* __restrict__ dW2_in, float4* __restrict__ rec,
                      long long ld, double* __restrict__ path_params, double* __restrict__ paths64, double* __restrict__ v64) {
    extern __shared__ double sm[];
    double* w1 = sm;                 // [M]
    double* w2 = w1 + M;             // [M]
    double* lamv = w2 + M;           // [T + 1]   lambda_n
    double* vv = lamv + (T + 1);     // [T + 1]   variance
    double* ss = vv + (T + 1);       // [T + 1]   prices
    __shared__ double prm[5];
    const int p = blockIdx.x;
    const unsigned long long gp = (unsigned long long)(k.path_offset + p);
    const uint2 key = make_uint2(k.seed_lo, k.seed_hi);
    if (threadIdx.x == 0) {
        double z[6];
        if (params_in != nullptr) {
            for (int j = 0; j < 5; ++j) prm[j] = params_in[(long long)j * n_paths + p];
        } else {
            const uint4 a = philox4x32_10(make_uint4((unsigned)gp, (unsigned)(gp >> 32), 0u, kStreamRbParams), key);
            const uint4 b = philox4x32_10(make_uint4((unsigned)gp, (unsigned)(gp >> 32), 1u, kStreamRbParams), key);
            box_muller_f64(a.x, a.y, z[0], z[1]);
            box_muller_f64(a.z, a.w, z[2], z[3]);
            box_muller_f64(b.x, b.y, z[4], z[5]);
            prm[0] = k.s0 * (1.0 + k.p_s0 * z[0]);                                             // :363
            prm[1] = k.xi * fmax(k.min_xi, 1.0 + k.p_xi * z[1]);                               // :364
            prm[2] = fmin(fmax(k.H * (1.0 + k.p_H * z[2]), k.H_lo), k.H_hi);                   // :365
            prm[3] = k.eta * fmax(k.min_eta, 1.0 + k.p_eta * z[3]);                            // :366
            prm[4] = fmin(fmax(k.rho * (1.0 + k.p_rho * z[4]), k.rho_lo), k.rho_hi);           // :367
        }
        if (path_params != nullptr)
            for (int j = 0; j < 5; ++j) path_params[(long long)j * n_paths + p] = prm[j];
    }
    for (int j = threadIdx.x; j < M; j += kOuterThreads) {                                      // :377-382
        if (dW1_in != nullptr) {
            w1[j] = dW1_in[(long long)p * M + j];
            w2[j] = dW2_in[(long long)p * M + j];
        } else {
            const uint4 x = philox4x32_10(make_uint4((unsigned)gp, (unsigned)(gp >> 32), (unsigned)(j >> 1), kStreamRbMain), key);
            double a, b, c, d;
            box_muller_f64(x.x, x.y, a, b);            // call j/2 serves increments j (even) and j + 1 (odd)
            box_muller_f64(x.z, x.w, c, d);
            w1[j] = (j & 1) ? b : a;
            w2[j] = (j & 1) ? d : c;
        }
    }
    __syncthreads();
    const double S0 = prm[0], xi = prm[1], H = prm[2], eta = prm[3], rho = prm[4];
    for (int n = threadIdx.x; n <= T; n += kOuterThreads)
        lamv[n] = (n == 0) ? 0.0 : 0.5 * exp(2.0 * H * log((double)n * k.dt));                  // :206-207
    __syncthreads();
    const double c = sqrt(2.0 * H) * eta / sqrt((double)M);
    for (int kk = threadIdx.x; kk <= T; kk += kOuterThreads) {
        double acc = 0.0;
        for (int n = 1; n <= T; ++n) acc = fma(lamv[n], w1[(kk - n) & (M - 1)], acc);           // the FIR form of :217-229
        vv[kk] = xi * exp(c * acc - eta * eta * lamv[kk]);                                      // :231-243
    }
    __syncthreads();
    if (threadIdx.x == 0) {                                                                     // :454-464
        const double sq = sqrt(k.dt), rc = sqrt(fmax(0.0, 1.0 - rho * rho));
        double S = S0;
        ss[0] = S;
        for (int j = 1; j <= T; ++j) {
            const double dW = rho * (sq * w1[j - 1]) + rc * (sq * w2[j - 1]);
            const double vt = vv[j - 1];
            S = fmax(S * exp((k.r - 0.5 * vt) * k.dt + sqrt(fmax(0.0, vt)) * dW), 1e-8);
            ss[j] = S;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t <= T; t += kOuterThreads) {
        if (rec != nullptr) rec[(long long)t * ld + p] = make_float4((float)ss[t], (float)vv[t], 0.f, 0.f);
        if (paths64 != nullptr) paths64[(long long)p * (T + 1) + t] = ss[t];
        if (v64 != nullptr) v64[(long long)p * (T + 1) + t] = vv[t];
    }
}

// ---- one inner path -----------------------------------------------------------------------------------------------
// w1[32], w2[30]: unscaled increments; L[n] = sqrt(2H) eta / sqrt(32) * lambda_n (n = 1..30); dk[k] = -eta^2 lambda_k.
// Returns log S_T.  The per-step floor S >= 1e-8 (:295) is applied in log space (monotone, hence identical).
struct InnerConsts {
    float L[kInnerSteps + 1];        // L[0] = 0 (t_0 = 0)
    float log2_xi_plus_dk[kInnerSteps];   // log2(xi) + dk * log2(e): v_k = exp2(X_k log2e + this)
    float rho, rho_c, r_dt, half_dt, sqrt_dt, log_s0;
};
// `L` = the filter taps in REGISTERS (copied once per thread); `c` may live in shared memory (one read per step).
__device__ __forceinline__ float inner_log_terminal(const float (&w1)[kInnerM], const float (&w2)[kInnerSteps],
                                                    const float (&L)[kInnerSteps + 1], const InnerConsts& c) {
    float logS = c.log_s0;
#pragma unroll
    for (int kk = 0; kk < kInnerSteps; ++kk) {
        float acc = 0.f;
#pragma unroll
        for (int n = 1; n <= kInnerSteps; ++n) acc = fmaf(L[n], w1[(kk - n) & (kInnerM - 1)], acc);
        const float v = mufu_ex2(fmaf(acc, kLog2ef, c.log2_xi_plus_dk[kk]));                    // xi exp(X_k - eta^2 lambda_k)
        const float dW = fmaf(c.rho, w1[kk], c.rho_c * w2[kk]);
        logS += fmaf(mufu_sqrt_rb(v) * c.sqrt_dt, dW, fmaf(-c.half_dt, v, c.r_dt));             // (r - v/2) dt + sqrt(v) sqrt(dt) dW
        logS = fmaxf(logS, -18.420680743952367f);                                               // ln(1e-8)
    }
    return logS;
}

__device__ __forceinline__ void fill_inner_consts(InnerConsts& c, float S, float xi, double H, double eta, double rho,
                                                  const RbConsts& k) {
    const double cc = sqrt(2.0 * H) * eta / sqrt((double)kInnerM);
    c.L[0] = 0.f;
    const double l2xi = log2((double)fmaxf(xi, 1e-30f));
#pragma unroll
    for (int n = 0; n <= kInnerSteps; ++n) {
        const double lam = (n == 0) ? 0.0 : 0.5 * exp(2.0 * H * log((double)n * k.dt));
        if (n >= 1) c.L[n] = (float)(cc * lam);
        if (n < kInnerSteps) c.log2_xi_plus_dk[n] = (float)(l2xi - eta * eta * lam * 1.4426950408889634);
    }
    c.rho = (float)rho;
    c.rho_c = (float)sqrt(fmax(0.0, 1.0 - rho * rho));
    c.r_dt = (float)(k.r * k.dt);
    c.half_dt = (float)(0.5 * k.dt);
    c.sqrt_dt = k.sqrt_dt_f;
    c.log_s0 = logf(S);
}

template <int NW>
__device__ __forceinline__ float block_sum(float x, float* smem /* [NW] */) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_down_sync(0xffffffffu, x, off);
    if ((threadIdx.x & 31) == 0) smem[threadIdx.x >> 5] = x;
    __syncthreads();
    float t = 0.f;
    if (threadIdx.x == 0) {
#pragma unroll
        for (int w = 0; w < NW; ++w) t += smem[w];
    }
    return t;
}

// grid = (n_paths, n_days, 2): CTA (p, d, kind) prices the ATM call (kind 0) or put (1) of path p at day t_begin + d.
__global__ void __launch_bounds__(kPriceThreads)
rbergomi_price_kernel(const RbConsts k, float4* __restrict__ rec, long long ld, int n_paths, int T, int t_begin,
                      const double* __restrict__ path_params) {
    __shared__ InnerConsts sc;
    __shared__ float red[kPriceThreads / 32];
    const int p = blockIdx.x, t = t_begin + blockIdx.y, kind = blockIdx.z;
    const float4 st = rec[(long long)t * ld + p];
    const float S = st.x, K = rintf(S);                                                         // :418
    if (threadIdx.x == 0)
        fill_inner_consts(sc, S, st.y, path_params[2LL * n_paths + p], path_params[3LL * n_paths + p],
                          path_params[4LL * n_paths + p], k);                                   // xi := v_t (:439)
    __syncthreads();
    float L[kInnerSteps + 1];
#pragma unroll
    for (int n = 0; n <= kInnerSteps; ++n) L[n] = sc.L[n];                                      // taps in registers
    const unsigned long long gp = (unsigned long long)(k.path_offset + p);
    const uint2 key = make_uint2(k.seed_lo, k.seed_hi);
    const unsigned c1 = (unsigned)t | ((k.shared_draws ? 0u : (unsigned)kind) << 24);           // independent draws per kind (:437-446)
    float pay = 0.f;
    for (int q = threadIdx.x; q < k.n_mc; q += kPriceThreads) {
        float w1[kInnerM], w2[kInnerM];
#pragma unroll
        for (int j = 0; j < kInnerM / 2; ++j) {                                                 // 16 Philox calls -> 64 normals
            const uint4 x = philox4x32_10(make_uint4((unsigned)gp, c1, (unsigned)q, kStreamRbInner | (unsigned)j), key);
            box_muller_f32(x.x, x.y, w1[2 * j], w1[2 * j + 1]);
            box_muller_f32(x.z, x.w, w2[2 * j], w2[2 * j + 1]);
        }
        float w2s[kInnerSteps];
#pragma unroll
        for (int j = 0; j < kInnerSteps; ++j) w2s[j] = w2[j];
        const float ST = __expf(inner_log_terminal(w1, w2s, L, sc));
        pay += kind == 0 ? fmaxf(ST - K, 0.f) : fmaxf(K - ST, 0.f);                             // :299-302
    }
    const float tot = block_sum<kPriceThreads / 32>(pay, red);
    if (threadIdx.x == 0) {
        const float price = tot / (float)k.n_mc * k.disc_f;                                     // :304
        float* out = reinterpret_cast<float*>(rec + (long long)t * ld + p) + (kind == 0 ? 2 : 3);
        *out = price;
        if (t == T - 1) reinterpret_cast<float*>(rec + (long long)T * ld + p)[kind == 0 ? 2 : 3] = price;   // stale marks of row T
    }
}

// Parity entry: the same inner-path function on exported increments.  One CTA per batch element.
__global__ void __launch_bounds__(kPriceThreads)
rbergomi_price_from_increments_kernel(const RbConsts k, const double* __restrict__ S0, const double* __restrict__ Kk,
                                      const double* __restrict__ xi, const double* __restrict__ H, const double* __restrict__ eta,
                                      const double* __restrict__ rho, const double* __restrict__ dW1, const double* __restrict__ dW2,
                                      int n_mc, int is_put, double* __restrict__ price) {
    __shared__ InnerConsts sc;
    __shared__ float red[kPriceThreads / 32];
    const int b = blockIdx.x;
    if (threadIdx.x == 0) fill_inner_consts(sc, (float)S0[b], (float)xi[b], H[b], eta[b], rho[b], k);
    __syncthreads();
    float L[kInnerSteps + 1];
#pragma unroll
    for (int n = 0; n <= kInnerSteps; ++n) L[n] = sc.L[n];
    const float K = (float)Kk[b];
    float pay = 0.f;
    for (int q = threadIdx.x; q < n_mc; q += kPriceThreads) {
        const double* a = dW1 + ((long long)b * n_mc + q) * kInnerM;
        const double* d = dW2 + ((long long)b * n_mc + q) * kInnerM;
        float w1[kInnerM], w2[kInnerSteps];
#pragma unroll
        for (int j = 0; j < kInnerM; ++j) w1[j] = (float)a[j];
#pragma unroll
        for (int j = 0; j < kInnerSteps; ++j) w2[j] = (float)d[j];
        const float ST = __expf(inner_log_terminal(w1, w2, L, sc));
        pay += is_put ? fmaxf(K - ST, 0.f) : fmaxf(ST - K, 0.f);
    }
    const float tot = block_sum<kPriceThreads / 32>(pay, red);
    if (threadIdx.x == 0) price[b] = (double)(tot / (float)n_mc * k.disc_f);
}

static int make_rb_consts(const cantor_rbergomi_params* p, RbConsts* k) {
    CANTOR_REQUIRE(p != nullptr, "params is NULL");
    CANTOR_REQUIRE(p->dt > 0 && p->tenor > 0, "dt and tenor must be positive");
    CANTOR_REQUIRE(p->H > 0 && p->eta >= 0 && p->xi > 0, "H, xi must be positive, eta non-negative");
    k->s0 = p->s0; k->xi = p->xi; k->H = p->H; k->eta = p->eta; k->rho = p->rho;
    k->p_s0 = p->perturb_s0; k->p_xi = p->perturb_xi; k->p_H = p->perturb_H; k->p_eta = p->perturb_eta; k->p_rho = p->perturb_rho;
    k->min_xi = p->min_xi_factor; k->min_eta = p->min_eta_factor;
    k->H_lo = p->clip_H_min; k->H_hi = p->clip_H_max; k->rho_lo = p->clip_rho_min; k->rho_hi = p->clip_rho_max;
    k->r = p->r; k->dt = p->dt;
    k->r_f = (float)p->r; k->dt_f = (float)p->dt; k->sqrt_dt_f = (float)sqrt(p->dt); k->disc_f = (float)exp(-p->r * p->tenor);
    k->n_mc = p->n_mc; k->shared_draws = p->shared_draws;
    k->seed_lo = (unsigned)(p->seed & 0xffffffffull); k->seed_hi = (unsigned)(p->seed >> 32);
    k->path_offset = p->path_offset;
    return CANTOR_OK;
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_rbergomi_paths(const cantor_rbergomi_params* params, int32_t n_paths, int32_t episode_length,
                                     const double* params_in, const double* dW1_in, const double* dW2_in, int32_t M_in,
                                     float* svcp, int64_t ld, double* path_params, double* paths64, double* v64, void* stream) {
    RbConsts k;
    int rc = make_rb_consts(params, &k);
    if (rc) return rc;
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0, "bad shape");
    CANTOR_REQUIRE(svcp == nullptr || (ld >= n_paths && aligned16(svcp)), "svcp must be 16-byte aligned with ld >= n_paths");
    CANTOR_REQUIRE((dW1_in == nullptr) == (dW2_in == nullptr), "dW1_in and dW2_in go together");
    int M = 1;
    while (M < episode_length + 1) M <<= 1;                                                   // next_power_of_two (:200-204)
    CANTOR_REQUIRE(M <= kOuterMaxM, "episode_length too large (M = next_power_of_two(T + 1) must be <= 1024)");
    CANTOR_REQUIRE(dW1_in == nullptr || M_in == M, "exported increments must have M = next_power_of_two(T + 1) columns");
    const size_t smem = (2 * (size_t)M + 3 * (size_t)(episode_length + 1)) * sizeof(double);
    rbergomi_paths_kernel<<<(unsigned)n_paths, kOuterThreads, smem, (cudaStream_t)stream>>>(
        k, n_paths, episode_length, M, params_in, dW1_in, dW2_in, (float4*)svcp, ld, path_params, paths64, v64);
    return check_launch("rbergomi_paths_kernel");
}

extern "C" int cantor_rbergomi_price_atm(const cantor_rbergomi_params* params, float* svcp, int64_t ld, int32_t n_paths,
                                         int32_t episode_length, const double* path_params, int32_t t_begin, int32_t t_end,
                                         void* stream) {
    RbConsts k;
    int rc = make_rb_consts(params, &k);
    if (rc) return rc;
    CANTOR_REQUIRE(svcp != nullptr && path_params != nullptr && aligned16(svcp), "svcp / path_params");
    CANTOR_REQUIRE(n_paths > 0 && ld >= n_paths && episode_length > 0, "bad shape");
    CANTOR_REQUIRE(0 <= t_begin && t_begin <= t_end && t_end <= episode_length, "day range must lie in [0, T]");
    CANTOR_REQUIRE(params->n_mc > 0, "n_mc must be positive");
    CANTOR_REQUIRE((int)(params->tenor / params->dt) == kInnerSteps, "the nested-MC pricer is built for int(tenor / dt) == 30 inner steps");
    CANTOR_REQUIRE(t_end - t_begin <= 65535, "at most 65535 days per launch");
    if (t_end == t_begin) return CANTOR_OK;
    const dim3 grid((unsigned)n_paths, (unsigned)(t_end - t_begin), 2u);
    rbergomi_price_kernel<<<grid, kPriceThreads, 0, (cudaStream_t)stream>>>(k, (float4*)svcp, ld, n_paths, episode_length,
                                                                            t_begin, path_params);
    return check_launch("rbergomi_price_kernel");
}

extern "C" int cantor_rbergomi_price_from_increments(const cantor_rbergomi_params* params, const double* S0, const double* K,
                                                     const double* xi, const double* H, const double* eta, const double* rho,
                                                     const double* dW1, const double* dW2, int32_t batch, int32_t n_mc,
                                                     int32_t M, int32_t is_put, double* price, void* stream) {
    RbConsts k;
    int rc = make_rb_consts(params, &k);
    if (rc) return rc;
    CANTOR_REQUIRE(S0 && K && xi && H && eta && rho && dW1 && dW2 && price, "array is NULL");
    CANTOR_REQUIRE(batch > 0 && n_mc > 0 && M == kInnerM, "increments must be [batch, n_mc, 32]");
    CANTOR_REQUIRE((int)(params->tenor / params->dt) == kInnerSteps, "the nested-MC pricer is built for int(tenor / dt) == 30 inner steps");
    rbergomi_price_from_increments_kernel<<<(unsigned)batch, kPriceThreads, 0, (cudaStream_t)stream>>>(
        k, S0, K, xi, H, eta, rho, dW1, dW2, n_mc, is_put, price);
    return check_launch("rbergomi_price_from_increments_kernel");
}
