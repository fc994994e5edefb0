// Black-Scholes building blocks: normal CDF / PDF and the ATM greeks of the observation.
//
// Reference: HedgingEnv._calculate_greeks (src/env/hedging_env_v2.py:79-107),
//            black_scholes_vectorized (src/sim/option_price_assignment.py:10-21).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

// 1: norm.cdf / norm.pdf of the F64 ledger's observation greeks in float64.  0: d1 and the exponent's argument in float64, erfc / exp
// in float32 with a low-half correction.  Measured: the float32 specials move the F64 step only from 279 to 266 us per launch at
// 2^23 envs (0.72 -> 0.75 of the HBM roofline; 34.2 -> 30.9 us at 2^20) and cost ~1e-7 of absolute accuracy on the deltas, which
// VecNormalize amplifies by 1 / std of those columns beyond the 1e-5 the wrapper's parity test allows -- parity first: float64 stays.
#ifndef CANTOR_F64_GREEKS_IN_FP64
#define CANTOR_F64_GREEKS_IN_FP64 1
#endif

namespace cantor {

constexpr double kInvSqrt2Pi = 0.3989422804014326779;   // 1/sqrt(2*pi)
constexpr double kSqrtHalf = 0.7071067811865475244;
constexpr float kLn2f = 0.6931471805599453f;
constexpr float kLog2ef = 1.4426950408889634f;

// ---- SFU (MUFU) primitives: one instruction each, ~1-2 ulp ------------------------------------------
__device__ __forceinline__ float mufu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_rsqrt(float x) { float y; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float mufu_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// Standard normal pdf and cdf in float32 on the SFU + FP32 pipes.
// cdf: Abramowitz & Stegun 26.2.17 (|error| < 7.5e-8), sharing the exponential with the pdf:
//   Q(|x|) = phi(x) * t (b1 + t (b2 + t (b3 + t (b4 + t b5)))),  t = 1 / (1 + p |x|).
// Returns phi(x); *cdf = Phi(x); *cdf_m1 = Phi(x) - 1 without cancellation for x >= 0.
__device__ __forceinline__ float normal_pdf_cdf(float x, float* cdf, float* cdf_m1) {
    const float ax = fabsf(x);
    const float t = mufu_rcp(fmaf(0.2316419f, ax, 1.0f));
    const float pdf = mufu_ex2(-0.5f * kLog2ef * x * x) * (float)kInvSqrt2Pi;
    float poly = fmaf(t, 1.330274429f, -1.821255978f);
    poly = fmaf(t, poly, 1.781477937f);
    poly = fmaf(t, poly, -0.356563782f);
    poly = fmaf(t, poly, 0.319381530f);
    const float q = pdf * (poly * t);                 // upper tail of |x|
    const bool pos = x >= 0.f;
    *cdf = pos ? 1.0f - q : q;
    *cdf_m1 = pos ? -q : q - 1.0f;
    return pdf;
}

struct Greeks {
    float call_delta, gamma, put_delta;
};

// Constants of one env configuration that the greeks need (prepared on the host once per launch).
struct GreekConsts {
    float r_f;          // (float) risk_free_rate           -- python float is weak next to float32
    float T_f;          // (float) option_tenor_years
    float inv_sqrtT_f;  // 1 / sqrt(T)
    double sqrtT_d;     // sqrt(T) in float64 (np.sqrt of a python float)
    double T_d;
    int record_metrics;
};

// hedging_env_v2.py:79-107, float64 ledger: float32 numerator, float64 denominator, float64 Phi / phi.
__device__ __forceinline__ Greeks atm_greeks_f64(float S, float K, float v_spot, const GreekConsts& g) {
    Greeks out{0.f, 0.f, 0.f};
    if (!g.record_metrics) return out;                                         // :80-81
    const float sigma = __fsqrt_rn(fmaxf(v_spot, 1e-8f));                      // :84 (>= 1e-4, so :90's sigma test never fires)
    if (S <= 1e-6f) {                                                          // :87-89
        out.call_delta = (K == 0.f) ? 0.5f : (K > 0.f ? 0.f : 1.f);
        out.put_delta = (K == 0.f) ? -0.5f : (K < 0.f ? 0.f : -1.f);
        return out;
    }
    if (g.T_d <= 1e-6 || sigma <= 1e-6f) {                                     // :90-92
        out.call_delta = (S > K) ? 1.f : (S == K ? 0.5f : 0.f);
        out.put_delta = (S < K) ? -1.f : (S == K ? -0.5f : 0.f);
        return out;
    }
    const float Kc = fmaxf(K, 1e-6f);                                          // :94
    // float32 numerator exactly as NumPy evaluates it (no FMA contraction)
    const float drift = __fmul_rn(__fadd_rn(g.r_f, __fmul_rn(0.5f, __fmul_rn(sigma, sigma))), g.T_f);
    const float num = __fadd_rn(logf(__fdiv_rn(S, Kc)), drift);
    const double sst = __dmul_rn((double)sigma, g.sqrtT_d);                    // :95 float32 * float64
#if CANTOR_F64_GREEKS_IN_FP64
    double d1;
    if (sst < 1e-9) d1 = (num > 0.f ? 10.0 : (num < 0.f ? -10.0 : 0.0));        // :96-97
    else d1 = (double)num / sst;                                               // :99
    const double cdf = 0.5 * erfc(-d1 * kSqrtHalf);                            // norm.cdf
    out.call_delta = (float)cdf;
    out.put_delta = (float)(cdf - 1.0);                                        // :101
    const double den = __dmul_rn((double)S, sst);                              // :102
    out.gamma = (fabs(den) < 1e-9) ? 0.f : (float)(exp(-0.5 * d1 * d1) * kInvSqrt2Pi / den);   // :103-106
#else
    // The three results are float32 observations (tolerance 1e-6 relative / 2e-7 absolute): only d1 and the exponent's argument
    // need float64; Phi and the exponential run in float32 (CUDA erfcf <= 4 ulp, expf <= 2 ulp), with the argument's low half
    // applied as a first-order correction.  The all-float64 form (erfc + exp + two divisions: ~45 % of the kernel's instructions,
    // FP64 pipe 32 % busy, profiles/r02_hedge_step_f64_2p20_ncu_summary.txt) held the F64 step at 0.72 of the HBM roofline.
    double d1;
    if (sst < 1e-9) {
        d1 = (num > 0.f ? 10.0 : (num < 0.f ? -10.0 : 0.0));                    // :96-97
    } else {                                                                   // :99  num / sst to ~1e-14: float quotient + one float64 residual step
        const float sst_f = (float)sst;
        const float q0 = __fdiv_rn(num, sst_f);
        const double r = fma(-(double)q0, sst, (double)num);
        d1 = (double)q0 + r * (double)mufu_rcp(sst_f);
    }
    const float z = (float)(d1 * kSqrtHalf);
    const float q = 0.5f * erfcf(fabsf(z));                                    // upper tail of |d1|, relative accuracy
    const bool pos = z >= 0.f;
    out.call_delta = pos ? 1.0f - q : q;                                       // norm.cdf(d1)
    out.put_delta = pos ? -q : q - 1.0f;                                       // :101, without cancellation for d1 >= 0
    const double den = __dmul_rn((double)S, sst);                              // :102
    if (fabs(den) < 1e-9) {
        out.gamma = 0.f;                                                       // :103-106
    } else {
        const double xa = -0.5 * d1 * d1;
        const float hi = (float)xa;
        const float lo = (float)(xa - (double)hi);
        const float e = expf(hi);
        out.gamma = __fdiv_rn(fmaf(e, lo, e) * (float)kInvSqrt2Pi, (float)den);
    }
#endif
    return out;
}

// Shared core of the float32 ATM greeks and the float32 ATM price: d1, phi(d1), Phi(d1), Phi(d1) - 1 and
// 1 / (sigma sqrt(T)) for S > 1e-6, K = rint(S), vv = max(v, 1e-8), T > 1e-6.  6 MUFU + ~30 FP32 instructions.
struct AtmCore {
    float d1, pdf, cdf, cdf_m1, inv_sst, rs;     // rs = 1 / sigma
};
__device__ __forceinline__ AtmCore atm_core_f32(float S, float K, float vv, float r, float T, float inv_sqrtT) {
    AtmCore c;
    const float Kc = fmaxf(K, 1e-6f);
    c.rs = mufu_rsqrt(vv);                                                     // 1 / sigma
    c.inv_sst = c.rs * inv_sqrtT;                                              // 1 / (sigma sqrt(T))
    // log(S / K) with K = rint(S), through u = S / K - 1.  The reference rounds the quotient to float32 before the log
    // (:99); that rounding (<= 6e-8 in u) is amplified by 1 / (sigma sqrt(T)) in d1 -- 3e4 with a floored sigma -- so
    // there it is reproduced bit for bit with an IEEE division.  At ordinary volatilities (amplification <= 64: a d1
    // difference < 5e-6, far inside the 1e-4 tolerance) u = (S - K) / K, with the subtraction exact (|S - K| <= 1/2,
    // Sterbenz) and one MUFU reciprocal, is both cheaper (3 instructions instead of ~15) and closer to the exact value.
    float u;
    if (c.inv_sst <= 64.0f) u = (S - Kc) * mufu_rcp(Kc);
    else u = __fdiv_rn(S, Kc) - 1.0f;
    // log(1 + u) through the log1p series (|u| <= 1/16: error < u^7 / 7) because lg2.approx only bounds the ABSOLUTE
    // error near 1.
    float lg;
    if (fabsf(u) <= 0.0625f) {
        float s = fmaf(u, -1.0f / 6.0f, 0.2f);
        s = fmaf(u, s, -0.25f);
        s = fmaf(u, s, 1.0f / 3.0f);
        s = fmaf(u, s, -0.5f);
        lg = fmaf(u * u, s, u);
    } else {
        lg = mufu_lg2(1.0f + u) * kLn2f;
    }
    const float num = fmaf(fmaf(0.5f, vv, r), T, lg);
    c.d1 = num * c.inv_sst;
    c.pdf = normal_pdf_cdf(c.d1, &c.cdf, &c.cdf_m1);
    return c;
}

// hedging_env_v2.py:79-107 for the float32 throughput path: no divisions beyond the one the reference rounds.
__device__ __forceinline__ Greeks atm_greeks_f32(float S, float K, float v_spot, const GreekConsts& g) {
    Greeks out{0.f, 0.f, 0.f};
    if (!g.record_metrics) return out;
    const float vv = fmaxf(v_spot, 1e-8f);
    if (S <= 1e-6f) {
        out.call_delta = (K == 0.f) ? 0.5f : (K > 0.f ? 0.f : 1.f);
        out.put_delta = (K == 0.f) ? -0.5f : (K < 0.f ? 0.f : -1.f);
        return out;
    }
    if (g.T_d <= 1e-6) {
        out.call_delta = (S > K) ? 1.f : (S == K ? 0.5f : 0.f);
        out.put_delta = (S < K) ? -1.f : (S == K ? -0.5f : 0.f);
        return out;
    }
    const AtmCore c = atm_core_f32(S, K, vv, g.r_f, g.T_f, g.inv_sqrtT_f);
    out.call_delta = c.cdf;
    out.put_delta = c.cdf_m1;
    // :102-106  gamma = phi / (S sigma sqrt(T)), 0 when that denominator is < 1e-9  (S < 1e-9 / (sigma sqrt(T)))
    out.gamma = (S < 1e-9f * c.inv_sst) ? 0.f : c.pdf * c.inv_sst * mufu_rcp(S);
    return out;
}

}  // namespace cantor
