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
#ifndef CANTOR_RB_PRICE_BLOCKS
#define CANTOR_RB_PRICE_BLOCKS 5          // resident pricing CTAs per SM the tensor-core pricer is compiled for (register cap 65536 / (128 x this));
                                          // measured 3 / 4 / 5: 11.39 / 11.20 / 11.07 ms on the 512 x 32-day bench, no spills at 96 registers
#endif
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
// Uniforms by the mantissa trick (bits -> [1, 2) -> subtract): ALU-pipe instructions instead of int-to-float conversions,
// which share the quarter-rate XU pipe with the MUFU transcendentals this kernel is bound by.  23-bit resolution.
__device__ __forceinline__ void box_muller_f32(unsigned x0, unsigned x1, float& n0, float& n1) {
    const float u1 = 2.0f - __uint_as_float(0x3F800000u | (x0 >> 9));         // (0, 1]
    const float u2 = __uint_as_float(0x3F800000u | (x1 >> 9)) - 1.0f;         // [0, 1)
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
// Called by threads 0 .. kInnerSteps of the CTA: thread n computes tap n (one float64 log + exp each -- done by a single
// thread the 31 taps were 27 % of the kernel's instructions and a serial prologue in front of every pricing), thread 0 the scalars.
__device__ __forceinline__ void fill_inner_consts(InnerConsts& c, int n, float S, float xi, double H, double eta, double rho,
                                                  const RbConsts& k) {
    const double cc = sqrt(2.0 * H) * eta / sqrt((double)kInnerM);
    const double l2xi = log2((double)fmaxf(xi, 1e-30f));
    const double lam = (n == 0) ? 0.0 : 0.5 * exp(2.0 * H * log((double)n * k.dt));
    c.L[n] = (n == 0) ? 0.f : (float)(cc * lam);
    if (n < kInnerSteps) c.log2_xi_plus_dk[n] = (float)(l2xi - eta * eta * lam * 1.4426950408889634);
    if (n == 0) {
        c.rho = (float)rho;
        c.rho_c = (float)sqrt(fmax(0.0, 1.0 - rho * rho));
        c.r_dt = (float)(k.r * k.dt);
        c.half_dt = (float)(0.5 * k.dt);
        c.sqrt_dt = k.sqrt_dt_f;
        c.log_s0 = logf(S);
    }
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

// ---- tensor-core FIR ------------------------------------------------------------------------------------------------
// X[128 inner paths x 32] = W1[128 x 32] * Lambda^T, Lambda[k][j] = L[(k - j) & 31] (the circulant of the filter taps):
// GEMM-shaped, so it runs on tcgen05 (kind::tf32, M = 128 = the CTA's inner paths, N = 32, K = 32 in four K = 8 steps,
// float32 accumulators in 32 TMEM columns).  TF32 keeps 10 mantissa bits, so both operands are split hi + lo
// (hi = the upper 19 bits, lo = x - hi, exact) and X = A_hi B_hi + A_hi B_lo + A_lo B_hi: float32-level accuracy from
// 12 small MMAs.  Operand tiles use the canonical K-major no-swizzle core-matrix layout (8 rows x 16 bytes).
namespace rbtc {
constexpr int kLbo = 128, kSbo = (kInnerM / 4) * 128;              // fp32: 4 elements per 16-byte core-matrix row; K = 32 -> 8 chunks
constexpr int kABytes = kPriceThreads * kInnerM * 4;              // 16 KB per A tile
constexpr int kBBytes = kInnerM * kInnerM * 4;                    // 4 KB per B tile
constexpr int kSmemBytes = 2 * kABytes + 2 * kBBytes + 16;
constexpr int kTmemCols = 32;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(kLbo >> 4) << 16) | ((uint64_t)(kSbo >> 4) << 32) | (1ull << 46);
}
// kind::tf32: D = F32 (1 << 4), A = B = TF32 (2 << 7, 2 << 10), K-major, M = 128, N = 32
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kInnerM >> 3) << 17) | ((uint32_t)(kPriceThreads >> 4) << 24);
__device__ __forceinline__ void umma_tf32(uint32_t d, uint64_t a, uint64_t b, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d), "l"(a), "l"(b), "r"(kIdesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&r)[16]) {
    uint32_t u[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
                   "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) r[j] = __uint_as_float(u[j]);
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
}  // namespace rbtc

// Where an inner path's increments come from: Philox (production) or the caller's exported arrays (parity).
// Philox layout: call j < 8 of (path, day | kind, inner path q) gives w1[4j .. 4j + 3]; call 8 + c gives w2[4c .. 4c + 3].
struct InnerDraws {
    const double* dW1;      // exported [*, n_mc, 32] or NULL
    const double* dW2;
    unsigned c0, c1;
    uint2 key;
};
template <bool EX>
__device__ __forceinline__ void draw_w1(const InnerDraws& d, long long row, unsigned q, float (&w1)[kInnerM]) {
    if (EX) {
        const double* a = d.dW1 + (row + q) * kInnerM;
#pragma unroll
        for (int j = 0; j < kInnerM; ++j) w1[j] = (float)a[j];
    } else {
#pragma unroll
        for (int j = 0; j < kInnerM / 4; ++j) {
            const uint4 x = philox4x32_10(make_uint4(d.c0, d.c1, q, kStreamRbInner | (unsigned)j), d.key);
            box_muller_f32(x.x, x.y, w1[4 * j], w1[4 * j + 1]);
            box_muller_f32(x.z, x.w, w1[4 * j + 2], w1[4 * j + 3]);
        }
    }
}
template <bool EX>
__device__ __forceinline__ void draw_w2(const InnerDraws& d, long long row, unsigned q, int chunk, float (&w2)[4]) {
    if (EX) {
        const double* a = d.dW2 + (row + q) * kInnerM + 4 * chunk;
#pragma unroll
        for (int j = 0; j < 4; ++j) w2[j] = (float)a[j];
    } else {
        const uint4 x = philox4x32_10(make_uint4(d.c0, d.c1, q, kStreamRbInner | (unsigned)(8 + chunk)), d.key);
        box_muller_f32(x.x, x.y, w2[0], w2[1]);
        box_muller_f32(x.z, x.w, w2[2], w2[3]);
    }
}

// One step of the inner log-Euler loop given X_k (:285-295), in log space.
__device__ __forceinline__ float inner_step(float logS, float Xk, float w1k, float w2k, int kk, const InnerConsts& c) {
    // sqrt(v) = exp2(arg / 2), v = sqrt(v)^2: one MUFU per step instead of two.  v = xi exp(X_k - eta^2 lambda_k)
    const float sv = mufu_ex2(0.5f * fmaf(Xk, kLog2ef, c.log2_xi_plus_dk[kk]));
    const float v = sv * sv;
    const float dW = fmaf(c.rho, w1k, c.rho_c * w2k);
    logS += fmaf(sv * c.sqrt_dt, dW, fmaf(-c.half_dt, v, c.r_dt));
    return fmaxf(logS, -18.420680743952367f);                                                    // S >= 1e-8 (:295)
}

// grid = (n_paths, n_days, 2): CTA (p, d, kind) prices the ATM call (kind 0) or put (1) of path p at day t_begin + d;
// with exported draws (EX, the parity form): grid = (batch), inputs from the arrays, price written to price_out.
template <bool TC, bool EX>
__global__ void __launch_bounds__(kPriceThreads, TC ? CANTOR_RB_PRICE_BLOCKS : 1)
rbergomi_price_kernel(const RbConsts k, float4* __restrict__ rec, long long ld, int n_paths, int T, int t_begin,
                      const double* __restrict__ path_params, const double* __restrict__ ex_S0, const double* __restrict__ ex_K,
                      const double* __restrict__ ex_xi, const double* __restrict__ ex_dW1, const double* __restrict__ ex_dW2,
                      int ex_is_put, double* __restrict__ price_out) {
    __shared__ InnerConsts sc;
    __shared__ float red[kPriceThreads / 32];
    extern __shared__ __align__(128) unsigned char dyn[];                                       // TC: A_hi, A_lo, B_hi, B_lo, mbarrier, TMEM slot
    constexpr bool exported = EX;
    const int p = blockIdx.x, t = t_begin + blockIdx.y;
    const int kind = exported ? ex_is_put : blockIdx.z;
    float S, K, xi0;
    if (exported) {
        S = (float)ex_S0[p];
        K = (float)ex_K[p];
        xi0 = (float)ex_xi[p];
    } else {
        const float4 st = rec[(long long)t * ld + p];
        S = st.x;
        K = rintf(S);                                                                           // :418
        xi0 = st.y;                                                                             // xi := v_t (:439)
    }
    const int n_mc = k.n_mc;
    if (threadIdx.x <= kInnerSteps)
        fill_inner_consts(sc, (int)threadIdx.x, S, xi0, path_params[2LL * n_paths + p], path_params[3LL * n_paths + p], path_params[4LL * n_paths + p], k);
    const unsigned long long gp = (unsigned long long)(k.path_offset + p);
    InnerDraws dr{ex_dW1, ex_dW2, (unsigned)gp, (unsigned)t | ((k.shared_draws ? 0u : (unsigned)kind) << 24),   // independent draws per kind (:437-446)
                  make_uint2(k.seed_lo, k.seed_hi)};
    const long long row = (long long)p * n_mc;
    float pay = 0.f;

    if (!TC) {
        __syncthreads();
        float L[kInnerSteps + 1];
#pragma unroll
        for (int n = 0; n <= kInnerSteps; ++n) L[n] = sc.L[n];                                  // filter taps in registers
        for (int q = threadIdx.x; q < n_mc; q += kPriceThreads) {
            float w1[kInnerM];
            draw_w1<EX>(dr, row, (unsigned)q, w1);
            float logS = sc.log_s0;
            float w2[4] = {0.f, 0.f, 0.f, 0.f};                                                 // one draw call serves four consecutive steps
#pragma unroll
            for (int kk = 0; kk < kInnerSteps; ++kk) {
                if ((kk & 3) == 0) draw_w2<EX>(dr, row, (unsigned)q, kk >> 2, w2);
                float acc = 0.f;
#pragma unroll
                for (int n = 1; n <= kInnerSteps; ++n) acc = fmaf(L[n], w1[(kk - n) & (kInnerM - 1)], acc);
                logS = inner_step(logS, acc, w1[kk], w2[kk & 3], kk, sc);
            }
            const float ST = __expf(logS);
            pay += kind == 0 ? fmaxf(ST - K, 0.f) : fmaxf(K - ST, 0.f);                         // :299-302
        }
    } else {
        unsigned char* a_hi = dyn;
        unsigned char* a_lo = a_hi + rbtc::kABytes;
        unsigned char* b_hi = a_lo + rbtc::kABytes;
        unsigned char* b_lo = b_hi + rbtc::kBBytes;
        uint64_t* bar = reinterpret_cast<uint64_t*>(b_lo + rbtc::kBBytes);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
        const uint32_t mbar = rbtc::smem_u32(bar);
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (threadIdx.x < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(rbtc::smem_u32(tmem_slot)), "r"((uint32_t)rbtc::kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        __syncthreads();                                                                        // sc.L is ready
        // B[n = k][j] = L[(k - j) & 31] split hi / lo, core-matrix layout: (n / 8) SBO + (j / 4) LBO + (n % 8) 16 + (j % 4) 4
        for (int e = threadIdx.x; e < kInnerM * kInnerM; e += kPriceThreads) {
            const int n = e >> 5, j = e & 31;
            const int tap = (n - j) & (kInnerM - 1);
            const float v = tap <= kInnerSteps ? sc.L[tap] : 0.f;                               // L[0] = 0, tap 31 does not exist
            const float hi = rbtc::tf32_hi(v);
            const int off = (n >> 3) * rbtc::kSbo + (j >> 2) * rbtc::kLbo + (n & 7) * 16 + (j & 3) * 4;
            *reinterpret_cast<float*>(b_hi + off) = hi;
            *reinterpret_cast<float*>(b_lo + off) = v - hi;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        fence_proxy_async_smem();
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem = *tmem_slot;
        const uint32_t lane_addr = tmem + ((uint32_t)(threadIdx.x & ~31) << 16);
        unsigned char* row_hi = a_hi + (threadIdx.x >> 3) * rbtc::kSbo + (threadIdx.x & 7) * 16;
        unsigned char* row_lo = a_lo + (threadIdx.x >> 3) * rbtc::kSbo + (threadIdx.x & 7) * 16;
        uint32_t phase = 0;
        bool timed_out = false;
        const int iters = (n_mc + kPriceThreads - 1) / kPriceThreads;
        for (int it = 0; it < iters; ++it) {
            const int q = it * kPriceThreads + threadIdx.x;
            const bool live = q < n_mc;
            float w1[kInnerM];
            draw_w1<EX>(dr, row, (unsigned)(live ? q : 0), w1);
#pragma unroll
            for (int cch = 0; cch < kInnerM / 4; ++cch) {                                       // this thread's A row, hi and lo
                float4 h, l;
                h.x = rbtc::tf32_hi(w1[4 * cch]);     l.x = w1[4 * cch] - h.x;
                h.y = rbtc::tf32_hi(w1[4 * cch + 1]); l.y = w1[4 * cch + 1] - h.y;
                h.z = rbtc::tf32_hi(w1[4 * cch + 2]); l.z = w1[4 * cch + 2] - h.z;
                h.w = rbtc::tf32_hi(w1[4 * cch + 3]); l.w = w1[4 * cch + 3] - h.w;
                *reinterpret_cast<float4*>(row_hi + cch * rbtc::kLbo) = h;
                *reinterpret_cast<float4*>(row_lo + cch * rbtc::kLbo) = l;
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");                    // earlier tcgen05.ld of the accumulator
            fence_proxy_async_smem();
            __syncthreads();
            if (threadIdx.x == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = rbtc::smem_u32(a_hi), al = rbtc::smem_u32(a_lo), bh = rbtc::smem_u32(b_hi), bl = rbtc::smem_u32(b_lo);
#pragma unroll
                for (int ks = 0; ks < kInnerM / 8; ++ks) {                                      // K = 8 per instruction = 2 chunks
                    const uint32_t o = ks * 2 * rbtc::kLbo;
                    rbtc::umma_tf32(tmem, rbtc::smem_desc(ah + o), rbtc::smem_desc(bh + o), ks > 0 ? 1u : 0u);
                    rbtc::umma_tf32(tmem, rbtc::smem_desc(ah + o), rbtc::smem_desc(bl + o), 1u);
                    rbtc::umma_tf32(tmem, rbtc::smem_desc(al + o), rbtc::smem_desc(bh + o), 1u);
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
            }
            {
                uint32_t done = 0;
                unsigned spins = 0;
                while (!done && !timed_out) {
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                                 : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
                    if (!done && ++spins > (1u << 24)) timed_out = true;
                }
                phase ^= 1;
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            float logS = sc.log_s0;
            float w2[4] = {0.f, 0.f, 0.f, 0.f};                                                 // persists over the four steps of a draw call
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                float X[16];
                rbtc::tmem_ld16(lane_addr + half * 16, X);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int kk = half * 16 + j;
                    if (kk < kInnerSteps) {
                        if ((kk & 3) == 0) draw_w2<EX>(dr, row, (unsigned)(live ? q : 0), kk >> 2, w2);
                        logS = inner_step(logS, X[j], w1[kk], w2[kk & 3], kk, sc);
                    }
                }
            }
            const float ST = __expf(logS);
            if (live) pay += kind == 0 ? fmaxf(ST - K, 0.f) : fmaxf(K - ST, 0.f);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)rbtc::kTmemCols) : "memory");
        if (timed_out) pay = __int_as_float(0x7fc00000);                                        // NaN price: an MMA never completed
    }

    const float tot = block_sum<kPriceThreads / 32>(pay, red);
    if (threadIdx.x == 0) {
        const float price = tot / (float)n_mc * k.disc_f;                                       // :304
        if (exported) {
            price_out[p] = (double)price;
        } else {
            reinterpret_cast<float*>(rec + (long long)t * ld + p)[kind == 0 ? 2 : 3] = price;
            if (t == T - 1) reinterpret_cast<float*>(rec + (long long)T * ld + p)[kind == 0 ? 2 : 3] = price;   // stale marks of row T
        }
    }
}

// The tensor-core pricer needs 41 KB of shared memory per CTA; without a stated preference the driver picks a carve-out that
// holds only three of them per SM.  Ask for the largest one (per device, once) so that CANTOR_RB_PRICE_BLOCKS CTAs fit.
static void prefer_shared_memory_for_pricer() {
    static bool done[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || done[dev]) return;
    cudaFuncSetAttribute(rbergomi_price_kernel<true, false>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(rbergomi_price_kernel<true, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    done[dev] = true;
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
    cudaStream_t s = (cudaStream_t)stream;
    if (params->tensor_cores) prefer_shared_memory_for_pricer();
    if (params->tensor_cores)
        rbergomi_price_kernel<true, false><<<grid, kPriceThreads, rbtc::kSmemBytes, s>>>(k, (float4*)svcp, ld, n_paths, episode_length, t_begin,
                                                                                  path_params, nullptr, nullptr, nullptr, nullptr,
                                                                                  nullptr, 0, nullptr);
    else
        rbergomi_price_kernel<false, false><<<grid, kPriceThreads, 0, s>>>(k, (float4*)svcp, ld, n_paths, episode_length, t_begin, path_params,
                                                                    nullptr, nullptr, nullptr, nullptr, nullptr, 0, nullptr);
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
    // path_params layout [5, batch] = {S0, xi, H, eta, rho}: the kernel reads rows 2..4; S0 / K / xi come from their own arrays
    CANTOR_REQUIRE(H + batch == eta && eta + batch == rho, "H, eta, rho must be consecutive rows of one [3, batch] array");
    k.n_mc = n_mc;
    const double* pp = H - 2LL * batch;
    cudaStream_t s = (cudaStream_t)stream;
    if (params->tensor_cores) prefer_shared_memory_for_pricer();
    if (params->tensor_cores)
        rbergomi_price_kernel<true, true><<<(unsigned)batch, kPriceThreads, rbtc::kSmemBytes, s>>>(k, nullptr, 0, batch, 0, 0, pp, S0, K, xi, dW1,
                                                                                             dW2, is_put, price);
    else
        rbergomi_price_kernel<false, true><<<(unsigned)batch, kPriceThreads, 0, s>>>(k, nullptr, 0, batch, 0, 0, pp, S0, K, xi, dW1, dW2, is_put,
                                                                               price);
    return check_launch("rbergomi_price_kernel (exported increments)");
}
