// K1 -- path simulation (GBM / Heston log-Euler from counter-based Philox normals), optionally fused with
// K2's ATM Black-Scholes repricing, writing the packed time-major book {S, v, C, P} directly.
//
// One thread owns one path and walks time; the 32 paths of a warp store 512 contiguous bytes (one float4
// record each) per time step.  No reads: the kernel is bound by the FP32/SFU pipes and the HBM write stream.
//
// Reference semantics: the outer log-Euler step of generate_paths_and_options,
//   S_j = max(S_{j-1} * exp((r - v/2) dt + sqrt(max(v, 0)) * sqrt(dt) * dW), 1e-8)
// (src/sim/rbergomi_sim.py:454-464), the ATM strike K = round(S) (:418) and tenor 30/252 (:19) of the option
// columns, and the output schema (:528).  The reference's nested-MC pricer (:246-306) is replaced by the
// closed form, as BASELINE.json's north_star prescribes.  cantor_euler_from_normals replays the reference's
// own float64 step on exported normals for parity.
#include "sim_core.cuh"

namespace cantor {

template <int MODEL>   // 0 GBM, 1 Heston
__global__ void __launch_bounds__(128)
sim_paths_kernel(const SimConsts k, int n_paths, float4* __restrict__ rec, long long ld) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const unsigned long long gp = (unsigned long long)(k.path_offset + p);
    const uint2 key = make_uint2(k.seed_lo, k.seed_hi);
    float S = k.s0, v = k.v0;
    float4 out = make_float4(S, fmaxf(v, 0.0f), 0.f, 0.f);
    if (k.reprice) { const AtmQuote q = atm_quote_f32(out.x, out.y, k); out.z = q.call; out.w = q.put; }
    float4* dst = rec + p;                                                   // stepped by one row per day (no 64-bit index arithmetic in the loop)
    __stcs(dst, out);                                                        // row 0
    constexpr int NPS = MODEL == 0 ? 1 : 2;                                  // normals per step
    int t = 0;
    for (unsigned call = 0; t < k.T; ++call) {                               // one Philox call = 4 normals
        float z[4];
        path_normals(k, gp, call, z);
#pragma unroll
        for (int j = 0; j < 4 / NPS; ++j) {
            if (t < k.T) {
                sim_advance<MODEL>(k, S, v, z[NPS * j], z[NPS * j + NPS - 1]);
                ++t;
                out.x = S;
                out.y = fmaxf(v, 0.0f);
                // row T has no option columns in the reference schema: it repeats the marks of row T-1
                if (k.reprice && t < k.T) { const AtmQuote q = atm_quote_f32(out.x, out.y, k); out.z = q.call; out.w = q.put; }
                dst += ld;
                __stcs(dst, out);
            }
        }
    }
}

// rbergomi_sim.py:454-464 in float64 on exported normals.  Time-major inputs: v [T+1, ld], dW1/dW2 [T, ld].
__global__ void __launch_bounds__(128)
euler_from_normals_kernel(const double* __restrict__ S0, const double* __restrict__ v, const double* __restrict__ dW1,
                          const double* __restrict__ dW2, const double* __restrict__ rho, int n_paths, int T, long long ld,
                          double r, double dt, double* __restrict__ paths) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_paths) return;
    const double sqrt_dt = sqrt(dt);
    const double rh = rho[p];
    const double rc = sqrt(fmax(0.0, 1.0 - rh * rh));
    double S = S0[p];
    paths[p] = S;
    for (int j = 1; j <= T; ++j) {
        const double dw1 = sqrt_dt * dW1[(long long)(j - 1) * ld + p];
        const double dw2 = sqrt_dt * dW2[(long long)(j - 1) * ld + p];
        const double dW = rh * dw1 + rc * dw2;
        const double vt = v[(long long)(j - 1) * ld + p];
        const double drift = (r - 0.5 * vt) * dt;
        const double diff = sqrt(fmax(0.0, vt)) * dW;
        S = fmax(S * exp(drift + diff), 1e-8);
        paths[(long long)j * ld + p] = S;
    }
}

// In-place ATM repricing of a packed book that already holds S and v (e.g. loaded paths): fills C, P.
__global__ void __launch_bounds__(256)
reprice_atm_kernel(const SimConsts k, int n_paths, float4* __restrict__ rec, long long ld) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (p >= n_paths) return;
    const int ts = min(t, k.T - 1);                                          // row T repeats row T-1's marks
    float4 me = rec[(long long)t * ld + p];
    float S = me.x, v = me.y;
    if (ts != t) {
        const float4 src = rec[(long long)ts * ld + p];
        S = src.x;
        v = src.y;
    }
    const AtmQuote q = atm_quote_f32(S, v, k);
    me.z = q.call;
    me.w = q.put;
    rec[(long long)t * ld + p] = me;
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_sim_paths(const cantor_sim_params* params, int32_t n_paths, int32_t episode_length, float* svcp,
                                int64_t ld, void* stream) {
    CANTOR_REQUIRE(params != nullptr && svcp != nullptr, "params/svcp is NULL");
    CANTOR_REQUIRE(params->model == CANTOR_MODEL_GBM || params->model == CANTOR_MODEL_HESTON, "model");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths, "bad shape");
    CANTOR_REQUIRE(aligned16(svcp), "svcp must be 16-byte aligned");
    CANTOR_REQUIRE(params->dt > 0 && params->tenor > 0, "dt and tenor must be positive");
    SimConsts k;
    fill_sim_consts(params, episode_length, &k);
    const unsigned grid = (unsigned)((n_paths + 127) / 128);
    cudaStream_t s = (cudaStream_t)stream;
    if (params->model == CANTOR_MODEL_GBM) sim_paths_kernel<0><<<grid, 128, 0, s>>>(k, n_paths, (float4*)svcp, ld);
    else sim_paths_kernel<1><<<grid, 128, 0, s>>>(k, n_paths, (float4*)svcp, ld);
    return check_launch("sim_paths_kernel");
}

extern "C" int cantor_reprice_atm(float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length, double r,
                                  double tenor, void* stream) {
    CANTOR_REQUIRE(svcp != nullptr && aligned16(svcp), "svcp is NULL or misaligned");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths && tenor > 0, "bad shape");
    cantor_sim_params p = {};
    p.r = r; p.tenor = tenor; p.dt = 1.0 / 252;
    SimConsts k;
    fill_sim_consts(&p, episode_length, &k);
    const dim3 grid((n_paths + 255) / 256, episode_length + 1);
    reprice_atm_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(k, n_paths, (float4*)svcp, ld);
    return check_launch("reprice_atm_kernel");
}

extern "C" int cantor_euler_from_normals(const double* S0, const double* v, const double* dW1, const double* dW2,
                                         const double* rho, int32_t n_paths, int32_t episode_length, int64_t ld,
                                         double r, double dt, double* paths, void* stream) {
    CANTOR_REQUIRE(S0 && v && dW1 && dW2 && rho && paths, "array is NULL");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths, "bad shape");
    const unsigned grid = (unsigned)((n_paths + 127) / 128);
    euler_from_normals_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(S0, v, dW1, dW2, rho, n_paths, episode_length, ld,
                                                                      r, dt, paths);
    return check_launch("euler_from_normals_kernel");
}
