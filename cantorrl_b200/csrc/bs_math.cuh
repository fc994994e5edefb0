// Black-Scholes building blocks: normal CDF / PDF and the ATM greeks of the observation.
//
// Reference: HedgingEnv._calculate_greeks (src/env/hedging_env_v2.py:79-107),
//            black_scholes_vectorized (src/sim/option_price_assignment.py:10-21).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace cantor {

constexpr double kInvSqrt2Pi = 0.3989422804014326779;   // 1/sqrt(2*pi)
constexpr double kSqrtHalf = 0.7071067811865475244;

struct Greeks {
    float call_delta, gamma, put_delta;
};

// Constants of one env configuration that the greeks need (prepared on the host once per launch).
struct GreekConsts {
    float r_f;          // (float) risk_free_rate           -- python float is weak next to float32
    float T_f;          // (float) option_tenor_years
    float sqrtT_f;      // (float) sqrt(T)
    double sqrtT_d;     // sqrt(T) in float64 (np.sqrt of a python float)
    double T_d;
    int record_metrics;
};

// hedging_env_v2.py:79-107.  S, K = rint(S), v_spot float32.
// F64 = true follows the reference's ledger: float32 numerator, float64 denominator, float64 Phi / phi.
// F64 = false is the throughput path: float32 everywhere, erfcf + one MUFU.EX2 for phi.
template <bool F64>
__device__ __forceinline__ Greeks atm_greeks(float S, float K, float v_spot, const GreekConsts& g) {
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
    if (F64) {
        // float32 numerator exactly as NumPy evaluates it (no FMA contraction)
        const float drift = __fmul_rn(__fadd_rn(g.r_f, __fmul_rn(0.5f, __fmul_rn(sigma, sigma))), g.T_f);
        const float num = __fadd_rn(logf(__fdiv_rn(S, Kc)), drift);
        const double sst = __dmul_rn((double)sigma, g.sqrtT_d);               // :95 float32 * float64
        double d1;
        if (sst < 1e-9) d1 = (num > 0.f ? 10.0 : (num < 0.f ? -10.0 : 0.0));   // :96-97
        else d1 = (double)num / sst;                                           // :99
        const double cdf = 0.5 * erfc(-d1 * kSqrtHalf);                        // norm.cdf
        out.call_delta = (float)cdf;
        out.put_delta = (float)(cdf - 1.0);                                    // :101
        const double den = __dmul_rn((double)S, sst);                          // :102
        out.gamma = (fabs(den) < 1e-9) ? 0.f : (float)(exp(-0.5 * d1 * d1) * kInvSqrt2Pi / den);   // :103-106
    } else {
        const float num = logf(S / Kc) + (g.r_f + 0.5f * sigma * sigma) * g.T_f;
        const float sst = sigma * g.sqrtT_f;
        const float inv_sst = __frcp_rn(sst);
        const float d1 = num * inv_sst;
        const float cdf = 0.5f * erfcf(-d1 * (float)kSqrtHalf);
        out.call_delta = cdf;
        out.put_delta = cdf - 1.f;
        const float pdf = __expf(-0.5f * d1 * d1) * (float)kInvSqrt2Pi;
        out.gamma = pdf * inv_sst / S;
    }
    return out;
}

}  // namespace cantor
