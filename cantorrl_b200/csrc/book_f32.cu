// K2, throughput form -- float32 Black-Scholes price / delta / gamma of a ladder of M strikes along every path
// of a packed book, the "multi-strike option book in paths_options.npz layout" of BASELINE.json configs[2].
//
// Reference semantics: process_price_paths (src/sim/option_price_assignment.py:33-52) = strike K = round(S_0)
// (:36), time to expiry T_t = clip(1 - t/252, 0) running to the episode end (:38), realised volatility of the
// path prefix (:23-31; column 0 -> sigma floor, column 1 -> NaN), black_scholes_vectorized (:10-21); delta and
// gamma in the closed form of HedgingEnv._calculate_greeks (src/env/hedging_env_v2.py:94-106).  Generalised to
// K_m = round(S_0) * mult[m] and, optionally, to the book's own instantaneous variance (sigma = sqrt(max(v_t, 0)),
// the Heston case) and to a fixed tenor.  M = 1, mult = 1, realised volatility, maturity to the episode end IS the
// reference pipeline and is checked against its shipped known-answer pair (data/paths.npy -> paths_options.npz).
//
// One thread owns one path and walks time (the realised volatility is a running statistic); the 32 paths of a
// warp read one 512-byte line of float4 records per step and write one 128-byte line per strike per output.
// Per (path, t): one IEEE division + one logf for ln(S / K_0), one sqrt; per strike: 3 MUFU (2 rcp + 1 ex2) and
// ~35 FP32 instructions -- Phi(d2) reuses the exponential of Phi(d1) through S phi(d1) = K e^{-rT} phi(d2).
// The strike ladder {mult, ln mult, 1/mult} and the maturity grid {T_t, sqrt T_t, 1/sqrt T_t, e^{-rT_t}, e^{rT_t}}
// are staged in shared memory once per CTA.
#include "bs_math.cuh"
#include "common.cuh"

namespace cantor {

constexpr int kBookThreads = 128;
constexpr int kBookMaxStrikes = 32;
constexpr int kBookMaxGrid = 4096;

struct BookOut {
    float* calls;    // [M, T+1, ld]
    float* puts;
    float* deltas;   // call delta (put delta = call delta - 1), or NULL
    float* gammas;   // or NULL
};

// Welford recurrence over log-returns (option_price_assignment.py:23-31: std with ddof = 1, times sqrt(252)).  The
// accumulators are float64 -- five FP64 operations per (path, t), nothing per strike -- so sigma keeps float32 accuracy
// even after hundreds of steps; the log-return itself is float32.
struct RunningVolF32 {
    int n = 0;
    double mean = 0.0, m2 = 0.0;
    __device__ __forceinline__ void push(float xf) {
        const double x = (double)xf;
        ++n;
        const double d = x - mean;
        mean += d / (double)n;
        m2 += d * (x - mean);
    }
    // n == 1 -> 0 / 0 = NaN, like np.std(ddof=1) of one sample
    __device__ __forceinline__ float sigma_annual() const { return (float)(sqrt(m2 / (double)(n - 1)) * 15.874507866387544); }
};

template <bool GREEKS>
__global__ void __launch_bounds__(kBookThreads)
book_f32_kernel(const float4* __restrict__ rec, long long ld, long long out_ld, int n_paths, int T, float r,
                const float* __restrict__ strike_mult, int M, int sigma_from_book, float fixed_tenor, const BookOut out) {
    __shared__ float s_mult[kBookMaxStrikes], s_lnm[kBookMaxStrikes], s_invm[kBookMaxStrikes];
    extern __shared__ float s_grid[];                 // [5][T+1]: T_t, sqrt, 1/sqrt, disc, 1/disc
    float* g_T = s_grid;
    float* g_sq = g_T + (T + 1);
    float* g_isq = g_sq + (T + 1);
    float* g_disc = g_isq + (T + 1);
    float* g_idisc = g_disc + (T + 1);
    for (int j = threadIdx.x; j < M; j += kBookThreads) {
        const double m = (double)strike_mult[j];
        s_mult[j] = (float)m;
        s_lnm[j] = (float)log(m);
        s_invm[j] = (float)(1.0 / m);
    }
    for (int t = threadIdx.x; t <= T; t += kBookThreads) {
        // :38 (maturity to the episode end); T <= 0 takes the intrinsic-value branch (:17-20)
        const double Tt = fixed_tenor > 0.f ? (double)fixed_tenor : fmax(1.0 - (double)t / 252.0, 0.0);
        const double Ts = Tt <= 0.0 ? 1e-8 : Tt;
        g_T[t] = (float)Tt;
        g_sq[t] = (float)sqrt(Ts);
        g_isq[t] = (float)(1.0 / sqrt(Ts));
        g_disc[t] = (float)exp(-(double)r * Tt);
        g_idisc[t] = (float)exp((double)r * Tt);
    }
    __syncthreads();
    const int p = blockIdx.x * kBookThreads + threadIdx.x;
    if (p >= n_paths) return;

    const float S0 = __ldg(&rec[p].x);
    const float K0 = rintf(S0);                                               // np.round: half to even (:36)
    RunningVolF32 rv;
    float prev = S0;
    const long long plane = (long long)(T + 1) * out_ld;
    for (int t = 0; t <= T; ++t) {
        const float4 rc = __ldcs(rec + (long long)t * ld + p);
        const float S = rc.x;
        float sigma;
        if (sigma_from_book) {
            sigma = sqrtf(fmaxf(rc.y, 0.f));
        } else {
            sigma = 0.f;                                                      // column 0 (:25)
            if (t > 0) {
                rv.push(logf(__fdiv_rn(S, prev)));
                sigma = rv.sigma_annual();
            }
            prev = S;
        }
        sigma = (sigma < 1e-8f) ? 1e-8f : sigma;                              // NaN < eps is false: NaN propagates (:12)
        const float Tt = g_T[t], sq = g_sq[t], disc = g_disc[t];
        const float sst = sigma * sq;
        const float inv_sst = mufu_rcp(sigma) * g_isq[t];
        const float l0 = logf(__fdiv_rn(S, K0));                              // ln(S / K_0), |.| small: ~1e-7 absolute
        const float base = (l0 + fmaf(0.5f * sigma, sigma, r) * Tt) * inv_sst;    // d1 of the strike K_0 itself
        const float k0d = K0 * disc;                                          // K_0 e^{-rT}
        const float s_over_k0d = S * mufu_rcp(K0) * g_idisc[t];               // S / (K_0 e^{-rT})
        const float gamma_scale = GREEKS ? inv_sst * mufu_rcp(S) : 0.f;
        float* pc = out.calls + (long long)t * out_ld + p;                        // strike m lives `m * plane` further on
        float* pp = out.puts + (long long)t * out_ld + p;
        float* pd = (GREEKS && out.deltas != nullptr) ? out.deltas + (long long)t * out_ld + p : nullptr;
        float* pg = (GREEKS && out.gammas != nullptr) ? out.gammas + (long long)t * out_ld + p : nullptr;
        if (Tt <= 0.f) {                                                      // :17-20 intrinsic value at expiry (uniform in t)
            for (int m = 0; m < M; ++m) {
                const float K = K0 * s_mult[m], kd = k0d * s_mult[m];
                __stcs(pc, fmaxf(S - kd, 0.f));
                __stcs(pp, fmaxf(kd - S, 0.f));
                if (pd != nullptr) { __stcs(pd, (S > K) ? 1.f : (S == K ? 0.5f : 0.f)); pd += plane; }   // hedging_env_v2.py:90-92
                if (pg != nullptr) { __stcs(pg, 0.f); pg += plane; }
                pc += plane;
                pp += plane;
            }
            continue;
        }
#pragma unroll 4
        for (int m = 0; m < M; ++m) {
            const float kd = k0d * s_mult[m];
            const float d1 = fmaf(-s_lnm[m], inv_sst, base);                  // (ln(S / K_m) + (r + sigma^2 / 2) T) / (sigma sqrt T)
            const float d2 = d1 - sst;
            // Phi(d1) with its own exponential; phi(d2) = phi(d1) * S / (K e^{-rT})
            float c1, c1m;
            const float pdf1 = normal_pdf_cdf(d1, &c1, &c1m);
            const float pdf2 = pdf1 * (s_over_k0d * s_invm[m]);
            const float t2 = mufu_rcp(fmaf(0.2316419f, fabsf(d2), 1.0f));
            float poly = fmaf(t2, 1.330274429f, -1.821255978f);
            poly = fmaf(t2, poly, 1.781477937f);
            poly = fmaf(t2, poly, -0.356563782f);
            poly = fmaf(t2, poly, 0.319381530f);
            // far in the tail pdf1 underflows while the ratio overflows (0 * inf): the tail mass is 0 there
            const float q2 = (pdf1 > 0.f) ? pdf2 * (poly * t2) : 0.f;
            const bool pos2 = d2 >= 0.f;
            const float c2 = pos2 ? 1.0f - q2 : q2;
            const float c2m = pos2 ? -q2 : q2 - 1.0f;
            float call = fmaf(S, c1, -kd * c2);                               // S Phi(d1) - K e^{-rT} Phi(d2)
            float put = fmaf(S, c1m, -kd * c2m);                              // K e^{-rT} Phi(-d2) - S Phi(-d1)
            call = (call < 0.f) ? 0.f : call;                                 // cancellation noise of a price that is >= 0
            put = (put < 0.f) ? 0.f : put;                                    // (a NaN stays a NaN)
            __stcs(pc, call);
            __stcs(pp, put);
            pc += plane;
            pp += plane;
            if (GREEKS) {
                if (pd != nullptr) { __stcs(pd, c1); pd += plane; }           // call delta = Phi(d1)
                if (pg != nullptr) { __stcs(pg, pdf1 * gamma_scale); pg += plane; }   // phi(d1) / (S sigma sqrt(T))
            }
        }
    }
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_reprice_book_strided(const float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length, double r,
                                           const float* strike_mult, int32_t n_strikes, int32_t sigma_source, double fixed_tenor,
                                           int64_t out_ld, float* calls, float* puts, float* deltas, float* gammas, void* stream) {
    CANTOR_REQUIRE(svcp && strike_mult && calls && puts, "array is NULL");
    CANTOR_REQUIRE(aligned16(svcp), "svcp must be 16-byte aligned");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths && out_ld >= n_paths, "bad shape");
    CANTOR_REQUIRE(n_strikes >= 1 && n_strikes <= kBookMaxStrikes, "n_strikes must be in [1, 32]");
    CANTOR_REQUIRE(episode_length + 1 <= kBookMaxGrid, "episode_length too large for the shared-memory maturity grid");
    CANTOR_REQUIRE(sigma_source == CANTOR_SIGMA_REALISED || sigma_source == CANTOR_SIGMA_BOOK_VARIANCE, "sigma_source");
    CANTOR_REQUIRE(fixed_tenor >= 0, "fixed_tenor < 0");
    const BookOut out{calls, puts, deltas, gammas};
    const unsigned grid = (unsigned)((n_paths + kBookThreads - 1) / kBookThreads);
    const size_t smem = 5 * (size_t)(episode_length + 1) * sizeof(float);
    cudaStream_t s = (cudaStream_t)stream;
    if (deltas != nullptr || gammas != nullptr)
        book_f32_kernel<true><<<grid, kBookThreads, smem, s>>>((const float4*)svcp, ld, out_ld, n_paths, episode_length, (float)r,
                                                               strike_mult, n_strikes, sigma_source, (float)fixed_tenor, out);
    else
        book_f32_kernel<false><<<grid, kBookThreads, smem, s>>>((const float4*)svcp, ld, out_ld, n_paths, episode_length, (float)r,
                                                                strike_mult, n_strikes, sigma_source, (float)fixed_tenor, out);
    return check_launch("book_f32_kernel");
}

extern "C" int cantor_reprice_book(const float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length, double r,
                                   const float* strike_mult, int32_t n_strikes, int32_t sigma_source, double fixed_tenor,
                                   float* calls, float* puts, float* deltas, float* gammas, void* stream) {
    return cantor_reprice_book_strided(svcp, ld, n_paths, episode_length, r, strike_mult, n_strikes, sigma_source, fixed_tenor, ld,
                                       calls, puts, deltas, gammas, stream);
}
