// K2 -- Black-Scholes repricing along paths in float64 (the reference's option_price_assignment pipeline)
// and the single-call Black-Scholes delta hedge of bs_delta.py.
//
// Reference semantics:
//   black_scholes_vectorized          src/sim/option_price_assignment.py:10-21   -> bs_price_f64 / cantor_bs_price
//   calculate_annualized_vol_matrix   src/sim/option_price_assignment.py:23-31   -> realised-vol recurrence (Welford)
//   process_price_paths               src/sim/option_price_assignment.py:33-52   -> cantor_schema_b_book
//   bs_delta_hedge                    src/tools/bs_delta.py:11-55                -> cantor_bs_delta_hedge
//
// All arrays are time-major [T+1, ld]: one thread owns one path and walks time (the realised volatility is a
// running statistic of the path prefix), so the 32 paths of a warp read / write one contiguous line per step.
// The strike ladder and the maturity grid (T_t, sqrt(T_t), exp(-r T_t)) of the book are staged in shared memory.
#include "bs_math.cuh"
#include "common.cuh"

namespace cantor {

__device__ __forceinline__ double norm_cdf_f64(double x) { return 0.5 * erfc(-x * kSqrtHalf); }

// option_price_assignment.py:10-21 for one element.
__device__ __forceinline__ void bs_price_f64(double S, double K, double T, double r, double sigma, double eps,
                                             double& call, double& put) {
    if (T <= 0.0) {                                                 // :17-20 intrinsic value at expiry
        const double kd = K * exp(-r * T);
        call = fmax(S - kd, 0.0);
        put = fmax(kd - S, 0.0);
        return;
    }
    const double sig = (sigma < eps) ? eps : sigma;                 // NaN < eps is false: NaN propagates (:12)
    const double sq = sqrt(T);
    const double d1 = (log(S / K) + (r + 0.5 * sig * sig) * T) / (sig * sq);
    const double d2 = d1 - sig * sq;
    const double disc = exp(-r * T);
    call = S * norm_cdf_f64(d1) - K * disc * norm_cdf_f64(d2);
    put = K * disc * norm_cdf_f64(-d2) - S * norm_cdf_f64(-d1);
}

// Elementwise black_scholes_vectorized with NumPy-style broadcasting through strides (0 = scalar).
__global__ void __launch_bounds__(256)
bs_price_kernel(const double* __restrict__ S, const double* __restrict__ K, const double* __restrict__ T,
                const double* __restrict__ sigma, long long n, int sS, int sK, int sT, int sSig, double r, double eps,
                double* __restrict__ call, double* __restrict__ put) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double c, p;
    bs_price_f64(S[i * sS], K[i * sK], T[i * sT], r, sigma[i * sSig], eps, c, p);
    call[i] = c;
    put[i] = p;
}

// Running sample standard deviation of the log-returns of the path prefix (ddof = 1), annualised:
// after n returns sigma = sqrt(M2 / (n - 1)) * sqrt(252); n = 1 gives 0/0 = NaN like np.std(ddof=1) (:29).
struct RunningVol {
    double mean = 0.0, m2 = 0.0;
    int n = 0;
    __device__ __forceinline__ void push(double x) {
        ++n;
        const double d = x - mean;
        mean += d / (double)n;
        m2 += d * (x - mean);
    }
    __device__ __forceinline__ double sigma_annual() const { return sqrt(m2 / (double)(n - 1)) * 15.874507866387544; }
};

constexpr int kMaxStrikes = 32;
constexpr int kMaxGrid = 4096;

// process_price_paths generalised to a ladder of M strikes K_m = round(S_0) * mult[m] (M = 1, mult = 1 is the
// reference).  Maturity runs to the episode end: T_t = clip(1 - t/252, 0) (:38).  Outputs [M, T+1, ld].
__global__ void __launch_bounds__(128)
schema_b_book_kernel(const double* __restrict__ paths, int n_paths, int T, long long ld, double r,
                     const double* __restrict__ strike_mult, int M, double* __restrict__ vols,
                     double* __restrict__ calls, double* __restrict__ puts) {
    extern __shared__ double smem[];
    double* s_mult = smem;                   // [M]
    double* s_T = smem + kMaxStrikes;        // [T+1] time to expiry
    for (int j = threadIdx.x; j < M; j += blockDim.x) s_mult[j] = strike_mult[j];
    for (int t = threadIdx.x; t <= T; t += blockDim.x) s_T[t] = fmax(1.0 - (double)t / 252.0, 0.0);
    __syncthreads();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const double S0 = paths[p];
    const double K0 = rint(S0);                                     // np.round: half to even (:36)
    RunningVol rv;
    double prev = S0;
    for (int t = 0; t <= T; ++t) {
        const double S = paths[(long long)t * ld + p];
        double sigma = 0.0;                                         // column 0 (:25)
        if (t > 0) {
            rv.push(log(S / prev));
            sigma = rv.sigma_annual();
        }
        prev = S;
        if (vols != nullptr) vols[(long long)t * ld + p] = sigma;
        for (int m = 0; m < M; ++m) {
            double c, q;
            bs_price_f64(S, K0 * s_mult[m], s_T[t], r, sigma, 1e-8, c, q);
            calls[((long long)m * (T + 1) + t) * ld + p] = c;
            puts[((long long)m * (T + 1) + t) * ld + p] = q;
        }
    }
}

// bs_delta.py:36-55: K = S_0 (unrounded), T_total = (T+1) dt, sigma = realised vol of the prefix (0 with fewer
// than two returns), delta = Phi(d1) or 1{S > K}, no premium, no transaction costs.  pnl [T+1, ld].
__global__ void __launch_bounds__(128)
bs_delta_hedge_kernel(const double* __restrict__ paths, int n_paths, int T, long long ld, double r, double dt,
                      double* __restrict__ pnl) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const double K = paths[p];
    const double T_total = (double)(T + 1) * dt;
    RunningVol rv;
    double prev = K, cash = 0.0, prev_delta = 0.0;
    for (int t = 0; t <= T; ++t) {
        const double S = paths[(long long)t * ld + p];
        if (t > 0) rv.push(log(S / prev));
        prev = S;
        const double sigma = (rv.n < 2) ? 0.0 : rv.sigma_annual();  // bs_delta.py:26-34
        const double T_rem = fmax(T_total - (double)t * dt, 0.0);
        double delta, price;
        if (sigma < 1e-8 || T_rem <= 0.0) {                         // :12-14, :21-22
            price = fmax(S - K * exp(-r * T_rem), 0.0);
            delta = (S > K) ? 1.0 : 0.0;
        } else {
            const double sq = sigma * sqrt(T_rem);
            const double d1 = (log(S / K) + (r + 0.5 * sigma * sigma) * T_rem) / sq;
            const double d2 = d1 - sq;
            delta = norm_cdf_f64(d1);
            price = S * delta - K * exp(-r * T_rem) * norm_cdf_f64(d2);
        }
        cash -= (delta - prev_delta) * S;                           // :50-51
        prev_delta = delta;
        pnl[(long long)t * ld + p] = cash + prev_delta * S - price; // :54
    }
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_bs_price(const double* S, const double* K, const double* T, const double* sigma, int64_t n,
                               int32_t stride_S, int32_t stride_K, int32_t stride_T, int32_t stride_sigma, double r,
                               double epsilon, double* call, double* put, void* stream) {
    CANTOR_REQUIRE(S && K && T && sigma && call && put, "array is NULL");
    CANTOR_REQUIRE(n >= 0, "n < 0");
    CANTOR_REQUIRE((stride_S | stride_K | stride_T | stride_sigma) >= 0 && stride_S <= 1 && stride_K <= 1 &&
                   stride_T <= 1 && stride_sigma <= 1, "strides must be 0 (scalar) or 1");
    if (n == 0) return CANTOR_OK;
    bs_price_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(S, K, T, sigma, n, stride_S, stride_K,
                                                                                   stride_T, stride_sigma, r, epsilon,
                                                                                   call, put);
    return check_launch("bs_price_kernel");
}

extern "C" int cantor_schema_b_book(const double* paths, int32_t n_paths, int32_t episode_length, int64_t ld, double r,
                                    const double* strike_mult, int32_t n_strikes, double* vols, double* calls,
                                    double* puts, void* stream) {
    CANTOR_REQUIRE(paths && strike_mult && calls && puts, "array is NULL");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths, "bad shape");
    CANTOR_REQUIRE(n_strikes >= 1 && n_strikes <= kMaxStrikes, "n_strikes must be in [1, 32]");
    CANTOR_REQUIRE(episode_length + 1 <= kMaxGrid, "episode_length too large for the shared-memory maturity grid");
    const size_t smem = (kMaxStrikes + episode_length + 1) * sizeof(double);
    schema_b_book_kernel<<<(unsigned)((n_paths + 127) / 128), 128, smem, (cudaStream_t)stream>>>(
        paths, n_paths, episode_length, ld, r, strike_mult, n_strikes, vols, calls, puts);
    return check_launch("schema_b_book_kernel");
}

extern "C" int cantor_bs_delta_hedge(const double* paths, int32_t n_paths, int32_t episode_length, int64_t ld, double r,
                                     double dt, double* pnl, void* stream) {
    CANTOR_REQUIRE(paths && pnl, "array is NULL");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths, "bad shape");
    bs_delta_hedge_kernel<<<(unsigned)((n_paths + 127) / 128), 128, 0, (cudaStream_t)stream>>>(paths, n_paths,
                                                                                               episode_length, ld, r, dt, pnl);
    return check_launch("bs_delta_hedge_kernel");
}
