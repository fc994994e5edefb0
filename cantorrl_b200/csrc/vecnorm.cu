// VecNormalize on the device: running mean / variance of observations and of discounted returns, and the
// normalisation + clipping of a step's observations and rewards, without the batch ever leaving HBM.
//
// Reference semantics: the reference wraps its env in Stable-Baselines3's VecNormalize
// (src/agents/train_ppo_v2.py:204-208, 305-309: norm_obs = True, norm_reward = True, clip_obs = 10, gamma) and ships
// the resulting statistics (quantconnect/model_wrapper.py:131: (obs - mean) / sqrt(var + 1e-8)).  SB3 itself is an
// un-vendored dependency (2.6.0); what is restated here is its documented algorithm:
//   step_wait():  obs_rms.update(obs);  obs <- clip((obs - mean) / sqrt(var + eps), +-clip_obs)
//                 returns <- returns * gamma + reward;  ret_rms.update(returns)
//                 reward <- clip(reward / sqrt(ret_var + eps), +-clip_reward);  returns[done] <- 0
//   RunningMeanStd.update(x): batch mean / population variance / count folded in with the parallel (Chan) formula,
//                 initial mean 0, var 1, count 1e-4.
// Two kernels chained with programmatic dependent launch: batch moments (16-byte vector loads of 128-env chunks, per-CTA
// partial sums, no atomics on the data) whose LAST CTA (atomic ticket) folds the partials in a fixed order -- bitwise
// reproducible statistics whichever CTA that is -- and does the Chan update; then the in-place apply, which derives
// 1 / sqrt(var + eps) itself and normalises in float32 against a two-float mean (no float64 conversions per element).
// Memory-bound: 52 B read for the moments, 104 + ~18 B for the apply, per env-step (L2 hits right behind the step kernel).
#include "common.cuh"

#ifndef CANTOR_VN_APPLY_REVERSE
#define CANTOR_VN_APPLY_REVERSE 1
#endif

namespace cantor {

constexpr int kVnCols = CANTOR_OBS_DIM;            // 13
constexpr int kVnThreads = kVnCols * 32;           // 416: a thread's flat index keeps (index % 13) fixed across the stride
// rms[] layout (doubles)
constexpr int kObsMean = 0, kObsVar = 13, kObsCount = 26, kRetMean = 27, kRetVar = 28, kRetCount = 29;
constexpr int kDerived = 32;                       // [32..46): 1 / sqrt(var + eps) [13], return 1 / std (informational; the apply kernel derives its own)
constexpr int kTicket = 56;                        // [56]: arrival counter of the moments kernel's CTAs (unsigned, zero between launches)
constexpr int kPartial = 64;                       // [64, 64 + 28 * 592): per-CTA partial sums of the moments kernel, [statistic][CTA]

constexpr int kVnRows = 128;                       // envs per chunk: 128 x 13 floats = 416 float4 = one float4 per thread
constexpr int kVnSums = 2 * kVnCols + 2;           // per-CTA partial sums: obs sum[13], obs sumsq[13], ret sum, ret sumsq
constexpr int kVnMaxGrid = 148 * 4;

// The chunk mapping shared by the moments and the apply kernel: a CTA walks chunks of 128 envs; thread i owns the
// float4 number i of a chunk, i.e. elements 4i .. 4i + 3, whose (row within the chunk, column) never change from chunk
// to chunk because 128 * 13 is a multiple of 4 and of 13.  Full chunks move as aligned 16-byte vectors.
struct ChunkMap {
    int col[4], row[4];
    __device__ __forceinline__ ChunkMap() {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int e = 4 * (int)threadIdx.x + j;
            col[j] = e % kVnCols;
            row[j] = e / kVnCols;
        }
    }
};

// RunningMeanStd.update_from_moments for the 13 observation columns and the returns, from the batch sums in cols[]
// ({obs sum [13], obs sumsq [13], return sum, return sumsq}); called by every thread of ONE CTA (>= 14 threads); resets the ticket.
__device__ __forceinline__ void running_update(double* __restrict__ rms, const double* cols, long long n, int norm_obs,
                                               int norm_reward, double epsilon, unsigned* ticket) {
    const int j = threadIdx.x;
    const double bc = (double)n;
    const double count = rms[kObsCount];
    if (j < kVnCols) {                                                               // per column
        if (norm_obs) {
            const double bmean = cols[j] / bc;
            const double bvar = fmax(cols[kVnCols + j] / bc - bmean * bmean, 0.0);
            const double mean = rms[kObsMean + j], var = rms[kObsVar + j];
            const double t = count + bc, delta = bmean - mean;
            rms[kObsMean + j] = mean + delta * bc / t;
            rms[kObsVar + j] = (var * count + bvar * bc + delta * delta * count * bc / t) / t;
        }
        rms[kDerived + j] = 1.0 / sqrt(rms[kObsVar + j] + epsilon);
    } else if (j == kVnCols) {
        if (norm_reward) {
            const double bmean = cols[2 * kVnCols] / bc;
            const double bvar = fmax(cols[2 * kVnCols + 1] / bc - bmean * bmean, 0.0);
            const double rc = rms[kRetCount], mean = rms[kRetMean], var = rms[kRetVar];
            const double t = rc + bc, delta = bmean - mean;
            rms[kRetMean] = mean + delta * bc / t;
            rms[kRetVar] = (var * rc + bvar * bc + delta * delta * rc * bc / t) / t;
            rms[kRetCount] = t;
        }
        rms[kDerived + kVnCols] = 1.0 / sqrt(rms[kRetVar] + epsilon);
    }
    __syncthreads();
    if (j == 0) {
        if (norm_obs) rms[kObsCount] = count + bc;                                   // after every column used the old count
        *ticket = 0u;                                                                // ready for the next launch
    }
}

__global__ void __launch_bounds__(kVnThreads, 2)
vecnorm_moments_kernel(double* __restrict__ rms, double* __restrict__ returns, long long n,
                       const float* __restrict__ obs, const void* __restrict__ reward, int reward_f64, double gamma,
                       int norm_obs, int norm_reward, int vec_ok, double epsilon) {
    double* __restrict__ partial = rms + kPartial;                                   // [kVnSums, gridDim.x]
    __shared__ double part[8][kVnThreads];
    __shared__ double cols[kVnSums];
    const ChunkMap cm;
    pdl_wait_prior_grid();
    double sum[4] = {0.0, 0.0, 0.0, 0.0}, sq[4] = {0.0, 0.0, 0.0, 0.0};
    double rsum = 0.0, rsq = 0.0;
    const long long n_chunks = (n + kVnRows - 1) / kVnRows;
    const long long n_full = vec_ok ? n / kVnRows : 0;                               // chunks that move as whole float4 vectors
    constexpr int U = 8;                                                             // independent 16-byte loads in flight per thread
    long long c = blockIdx.x;
    if (norm_obs) {
        for (; c + (U - 1) * (long long)gridDim.x < n_full; c += U * (long long)gridDim.x) {   // full batches of U chunks:
            float4 v[U];                                                             // every load of a batch is in flight before the first add
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = __ldg(reinterpret_cast<const float4*>(obs) + (c + u * (long long)gridDim.x) * kVnThreads + threadIdx.x);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double d0 = v[u].x, d1 = v[u].y, d2 = v[u].z, d3 = v[u].w;
                sum[0] += d0; sq[0] = fma(d0, d0, sq[0]);
                sum[1] += d1; sq[1] = fma(d1, d1, sq[1]);
                sum[2] += d2; sq[2] = fma(d2, d2, sq[2]);
                sum[3] += d3; sq[3] = fma(d3, d3, sq[3]);
            }
        }
        constexpr int UL = 4;
        for (; c < n_full; c += UL * (long long)gridDim.x) {                        // the leftovers, in partly empty batches of UL
            float4 v[UL];
#pragma unroll
            for (int u = 0; u < UL; ++u) {
                const long long cu = c + u * (long long)gridDim.x;
                v[u] = cu < n_full ? __ldg(reinterpret_cast<const float4*>(obs) + cu * kVnThreads + threadIdx.x) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < UL; ++u) {
                const double d0 = v[u].x, d1 = v[u].y, d2 = v[u].z, d3 = v[u].w;
                sum[0] += d0; sq[0] = fma(d0, d0, sq[0]);
                sum[1] += d1; sq[1] = fma(d1, d1, sq[1]);
                sum[2] += d2; sq[2] = fma(d2, d2, sq[2]);
                sum[3] += d3; sq[3] = fma(d3, d3, sq[3]);
            }
        }
        // the chunks that cannot move as whole vectors: the ragged last one, or all of them when obs is not 16-byte aligned
        for (c = n_full + blockIdx.x; c < n_chunks; c += gridDim.x) {
            const long long base = c * kVnRows;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (base + cm.row[j] < n) {
                    const double d = (double)__ldg(obs + base * kVnCols + 4 * threadIdx.x + j);
                    sum[j] += d;
                    sq[j] = fma(d, d, sq[j]);
                }
        }
    }
    if (norm_reward) {                                                               // discounted returns: envs strided over the whole grid,
        const long long stride = (long long)gridDim.x * kVnThreads;                 // R independent load pairs in flight per thread
        long long i = (long long)blockIdx.x * kVnThreads + threadIdx.x;
        constexpr int R = 4;
        auto load_reward = [&](long long e) {
            return reward_f64 ? __ldg(reinterpret_cast<const double*>(reward) + e) : (double)__ldg(reinterpret_cast<const float*>(reward) + e);
        };
        for (; i < n; i += R * stride) {
            double ret[R], r[R];
#pragma unroll
            for (int u = 0; u < R; ++u)
                if (i + u * stride < n) {
                    ret[u] = returns[i + u * stride];
                    r[u] = load_reward(i + u * stride);
                }
#pragma unroll
            for (int u = 0; u < R; ++u)
                if (i + u * stride < n) {
                    ret[u] = fma(ret[u], gamma, r[u]);
                    returns[i + u * stride] = ret[u];
                    rsum += ret[u];
                    rsq = fma(ret[u], ret[u], rsq);
                }
        }
    }
    pdl_launch_dependents();
    // deterministic CTA reduction: threads with equal (i % 13) own the same four columns
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        part[j][threadIdx.x] = sum[j];
        part[4 + j][threadIdx.x] = sq[j];
    }
    __shared__ double p104[8 * kVnCols];
    __syncthreads();
    if (threadIdx.x < 8 * kVnCols) {                                                 // (value v, group g): sum the 32 members g, g + 13, ...
        const int v = threadIdx.x / kVnCols, g = threadIdx.x % kVnCols;
        double a = 0.0;
#pragma unroll 8
        for (int m = 0; m < kVnThreads / kVnCols; ++m) a += part[v][g + kVnCols * m];
        p104[threadIdx.x] = a;
    }
    __syncthreads();
    if (threadIdx.x < 2 * kVnCols) {                                                 // column c of {sum, sumsq}: its four (group, slot) owners
        const int kind = threadIdx.x / kVnCols, c = threadIdx.x % kVnCols;
        double a = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int g = ((c - j + kVnCols) * 10) % kVnCols;                        // 4 g + j = c (mod 13); 4^-1 = 10 (mod 13)
            a += p104[(4 * kind + j) * kVnCols + g];
        }
        cols[threadIdx.x] = a;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        rsum += __shfl_down_sync(0xffffffffu, rsum, off);
        rsq += __shfl_down_sync(0xffffffffu, rsq, off);
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) part[0][threadIdx.x >> 5] = rsum, part[1][threadIdx.x >> 5] = rsq;
    __syncthreads();
    if (threadIdx.x < 2) {                                                           // fixed order over the 13 warps
        double a = 0.0;
#pragma unroll
        for (int q = 0; q < kVnThreads / 32; ++q) a += part[threadIdx.x][q];
        cols[2 * kVnCols + threadIdx.x] = a;
    }
    __syncthreads();
    if (threadIdx.x < kVnSums) partial[(long long)threadIdx.x * gridDim.x + blockIdx.x] = cols[threadIdx.x];
    // ---- the last CTA to get here folds everybody's partials (fixed order) and updates the running statistics ----------
    __shared__ int is_last;
    __threadfence();
    __syncthreads();
    unsigned* ticket = reinterpret_cast<unsigned*>(rms + kTicket);
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_partial = (int)gridDim.x;
    for (int s = w; s < kVnSums; s += kVnThreads / 32) {                             // 13 warps, 28 statistics
        double a = 0.0;
        for (int j = lane; j < n_partial; j += 32) a += __ldcg(partial + (long long)s * n_partial + j);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) a += __shfl_down_sync(0xffffffffu, a, off);
        if (lane == 0) cols[s] = a;
    }
    __syncthreads();
    running_update(rms, cols, n, norm_obs, norm_reward, epsilon, ticket);
}

// Fused form (cantor_vecnorm_step_fused): the step kernel already wrote the per-CTA partial sums partial[CTA][28]
// (hedge_step.cu: vecnorm_partials).  Two levels, both with coalesced reads and a fixed summation order: fold CTA f sums the
// records of source CTAs [128 f, 128 f + 128) -- 252 threads = 9 parts x 28 statistics, each thread's <= 15 loads in flight at
// once, then the 9 parts in order -- into level1[f][28]; the last fold CTA (atomic ticket) sums level1 the same way and runs the
// running-statistics update.  (A one-CTA-per-statistic fold over a statistic-major layout took 8 us: the step kernel's 8-byte
// scattered writes left partially valid sectors that every read had to fill from DRAM.)
constexpr int kFoldThreads = 256;
constexpr int kFoldParts = 9, kFoldRows = 128, kFoldDepth = (kFoldRows + kFoldParts - 1) / kFoldParts;   // 15

// sums rows [0, n_rows) of a [n_rows][28] array (n_rows <= 128 per call) into out28 (shared), deterministic
__device__ __forceinline__ void fold_rows(const double* __restrict__ rows, int n_rows, double* sh /* [9 * 28] */, double* out28) {
    const int t = threadIdx.x;
    if (t < kFoldParts * kVnSums) {
        const int part = t / kVnSums;
        double v[kFoldDepth];
#pragma unroll
        for (int k = 0; k < kFoldDepth; ++k) {
            const int r = k * kFoldParts + part;
            v[k] = r < n_rows ? __ldcg(rows + (long long)r * kVnSums + (t - part * kVnSums)) : 0.0;
        }
        double a = 0.0;
#pragma unroll
        for (int k = 0; k < kFoldDepth; ++k) a += v[k];
        sh[t] = a;
    }
    __syncthreads();
    if (t < kVnSums) {
        double a = 0.0;
#pragma unroll
        for (int p = 0; p < kFoldParts; ++p) a += sh[p * kVnSums + t];
        out28[t] = a;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kFoldThreads)
vecnorm_fold_kernel(double* __restrict__ rms, const double* __restrict__ partial, double* __restrict__ level1, int n_cta,
                    long long n, int norm_obs, int norm_reward, double epsilon) {
    __shared__ double sh[kFoldParts * kVnSums];
    __shared__ double cols[kVnSums];
    __shared__ double acc28[kVnSums];
    __shared__ int is_last;
    pdl_wait_prior_grid();
    const int first = (int)blockIdx.x * kFoldRows;
    fold_rows(partial + (long long)first * kVnSums, min(kFoldRows, n_cta - first), sh, cols);
    if (threadIdx.x < kVnSums) level1[(long long)blockIdx.x * kVnSums + threadIdx.x] = cols[threadIdx.x];
    pdl_launch_dependents();
    unsigned* ticket = reinterpret_cast<unsigned*>(rms + kTicket);
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < kVnSums) acc28[threadIdx.x] = 0.0;
    __syncthreads();
    for (int base = 0; base < (int)gridDim.x; base += kFoldRows) {                   // level 2, chunks of 128 level-1 records in order
        fold_rows(level1 + (long long)base * kVnSums, min(kFoldRows, (int)gridDim.x - base), sh, cols);
        if (threadIdx.x < kVnSums) acc28[threadIdx.x] += cols[threadIdx.x];
        __syncthreads();
    }
    running_update(rms, acc28, n, norm_obs, norm_reward, epsilon, ticket);
}

__global__ void __launch_bounds__(kVnThreads, 2)
vecnorm_apply_kernel(const double* __restrict__ rms, double* __restrict__ returns, long long n, float* __restrict__ obs,
                     void* __restrict__ reward, int reward_f64, const unsigned char* __restrict__ done,
                     float* __restrict__ terminal_obs, double clip_obs, double clip_reward, int norm_obs, int norm_reward,
                     int vec_ok, double epsilon) {
    const ChunkMap cm;
    pdl_wait_prior_grid();
    float mh[4], ml[4], inv[4];                                                      // mean = mh + ml to ~2^-48: (x - mh) - ml loses nothing
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double m = rms[kObsMean + cm.col[j]];
        mh[j] = (float)m;
        ml[j] = (float)(m - (double)mh[j]);
        inv[j] = (float)(1.0 / sqrt(rms[kObsVar + cm.col[j]] + epsilon));
    }
    const float lo = (float)-clip_obs, hi = (float)clip_obs;
    const double rinv = 1.0 / sqrt(rms[kRetVar] + epsilon);
    // the subtraction against a two-float mean has no cancellation error; scale and clip in float32 -- the result is a float32 anyway
    auto norm1 = [&](float x, int j) { return fminf(fmaxf(((x - mh[j]) - ml[j]) * inv[j], lo), hi); };
    const long long n_chunks = (n + kVnRows - 1) / kVnRows;
    const long long n_full = vec_ok ? n / kVnRows : 0;
    constexpr int U = 8;
    // full chunks are walked LAST-WRITTEN FIRST: the step kernel that produced the batch wrote chunk 0 first and chunk n - 1 last, so what
    // is still in L2 is the tail of the batch; walking forward would start with the lines most likely to have left
#if CANTOR_VN_APPLY_REVERSE
    auto rev = [n_full](long long c) { return n_full - 1 - c; };
#else
    auto rev = [](long long c) { return c; };
#endif
    long long c = blockIdx.x;
    if (norm_obs) {
        for (; c + (U - 1) * (long long)gridDim.x < n_full; c += U * (long long)gridDim.x) {   // full batches of U chunks
            float4* p[U];
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                p[u] = reinterpret_cast<float4*>(obs) + rev(c + u * (long long)gridDim.x) * kVnThreads + threadIdx.x;
                v[u] = *p[u];
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                v[u].x = norm1(v[u].x, 0); v[u].y = norm1(v[u].y, 1); v[u].z = norm1(v[u].z, 2); v[u].w = norm1(v[u].w, 3);
                *p[u] = v[u];
            }
        }
        constexpr int UL = 4;
        for (; c < n_full; c += UL * (long long)gridDim.x) {                        // the leftovers, in partly empty batches of UL
            float4 v[UL];
#pragma unroll
            for (int u = 0; u < UL; ++u) {
                const long long cu = c + u * (long long)gridDim.x;
                if (cu < n_full) v[u] = *(reinterpret_cast<float4*>(obs) + rev(cu) * kVnThreads + threadIdx.x);
            }
#pragma unroll
            for (int u = 0; u < UL; ++u) {
                const long long cu = c + u * (long long)gridDim.x;
                if (cu < n_full) {
                    v[u].x = norm1(v[u].x, 0); v[u].y = norm1(v[u].y, 1); v[u].z = norm1(v[u].z, 2); v[u].w = norm1(v[u].w, 3);
                    *(reinterpret_cast<float4*>(obs) + rev(cu) * kVnThreads + threadIdx.x) = v[u];
                }
            }
        }
        for (c = n_full + blockIdx.x; c < n_chunks; c += gridDim.x) {               // ragged last chunk / unaligned obs
            const long long base = c * kVnRows;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (base + cm.row[j] < n) {
                    float* p = obs + base * kVnCols + 4 * threadIdx.x + j;
                    *p = norm1(*p, j);
                }
        }
    }
    // rewards, returns and the finished envs' terminal observations: envs strided over the whole grid, R loads in flight
    {
        const long long stride = (long long)gridDim.x * kVnThreads;
        constexpr int R = 4;
        auto finish = [&](long long i) {
            returns[i] = 0.0;
            if (norm_obs && terminal_obs != nullptr) {                               // the finished env's pre-reset observation (rare: whole row here)
                float* row = terminal_obs + i * kVnCols;
#pragma unroll
                for (int j = 0; j < kVnCols; ++j)
                    row[j] = fminf(fmaxf((float)((double)row[j] - rms[kObsMean + j]) * (float)(1.0 / sqrt(rms[kObsVar + j] + epsilon)), lo), hi);
            }
        };
        long long i = (long long)blockIdx.x * kVnThreads + threadIdx.x;
        for (; i < n; i += R * stride) {
            unsigned char d[R];
            double r[R];
#pragma unroll
            for (int u = 0; u < R; ++u) {
                d[u] = 0;
                r[u] = 0.0;
                if (i + u * stride < n) {
                    d[u] = __ldg(done + i + u * stride);
                    if (norm_reward) r[u] = reward_f64 ? reinterpret_cast<double*>(reward)[i + u * stride] : (double)reinterpret_cast<float*>(reward)[i + u * stride];
                }
            }
#pragma unroll
            for (int u = 0; u < R; ++u) {
                if (i + u * stride >= n) continue;
                if (norm_reward) {
                    const double y = fmin(fmax(r[u] * rinv, -clip_reward), clip_reward);
                    if (reward_f64) reinterpret_cast<double*>(reward)[i + u * stride] = y;
                    else reinterpret_cast<float*>(reward)[i + u * stride] = (float)y;
                }
                if (d[u]) finish(i + u * stride);
            }
        }
    }
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_vecnorm_init(double* rms, double* returns, int64_t n_envs, void* stream) {
    CANTOR_REQUIRE(rms != nullptr && returns != nullptr && n_envs >= 0, "bad arguments");
    double h[32] = {0};
    for (int j = 0; j < kVnCols; ++j) h[kObsVar + j] = 1.0;
    h[kObsCount] = 1e-4;
    h[kRetVar] = 1.0;
    h[kRetCount] = 1e-4;
    cudaStream_t s = (cudaStream_t)stream;
    CANTOR_CUDA(cudaMemsetAsync(rms, 0, sizeof(double) * CANTOR_VECNORM_DOUBLES, s));
    CANTOR_CUDA(cudaMemcpyAsync(rms, h, sizeof(h), cudaMemcpyHostToDevice, s));
    CANTOR_CUDA(cudaStreamSynchronize(s));                                   // h is a stack buffer
    CANTOR_CUDA(cudaMemsetAsync(returns, 0, sizeof(double) * (size_t)n_envs, s));
    return CANTOR_OK;
}

static int launch_apply(double* rms, double* returns, long long n, float* obs, void* reward, int f64, const uint8_t* done,
                        float* terminal_obs, double clip_obs, double clip_reward, double epsilon, int norm_obs, int norm_reward,
                        cudaStream_t s);

extern "C" int cantor_vecnorm_step_fused(double* rms, const cantor_vecnorm_fuse* fuse, int64_t n_envs, float* obs, void* reward,
                                         int32_t reward_precision, const uint8_t* done, float* terminal_obs, double clip_obs,
                                         double clip_reward, double epsilon, void* stream) {
    CANTOR_REQUIRE(rms && fuse && obs && reward && done, "array is NULL");
    CANTOR_REQUIRE(fuse->partial != nullptr && fuse->returns != nullptr, "fuse.partial / fuse.returns is NULL");
    CANTOR_REQUIRE(reward_precision == CANTOR_F32 || reward_precision == CANTOR_F64, "reward_precision must be 32 or 64");
    CANTOR_REQUIRE(n_envs >= 0, "n_envs < 0");
    if (n_envs == 0) return CANTOR_OK;
    const long long n_cta_ll = (n_envs + 127) / 128;                              // the step kernels run 128 envs per CTA
    CANTOR_REQUIRE(fuse->n_partial_ctas >= n_cta_ll && n_cta_ll <= 0x7fffffffLL, "fuse.partial is too small for n_envs");
    cudaStream_t s = (cudaStream_t)stream;
    long long n = n_envs;
    int n_cta = (int)n_cta_ll;
    const double* partial = fuse->partial;
    int norm_obs = fuse->norm_obs, norm_reward = fuse->norm_reward;
    const int n_fold = (n_cta + kFoldRows - 1) / kFoldRows;
    // level-1 records live behind the step kernel's partials: fuse->partial holds 28 * (n_partial_ctas + ceil(n_partial_ctas / 128)) doubles
    double* level1 = fuse->partial + (long long)fuse->n_partial_ctas * kVnSums;
    void* a1[] = {&rms, &partial, &level1, &n_cta, &n, &norm_obs, &norm_reward, &epsilon};
    int rc = launch_pdl((const void*)vecnorm_fold_kernel, dim3(n_fold), dim3(kFoldThreads), s, a1);
    if (rc) return rc;
    return launch_apply(rms, fuse->returns, n, obs, reward, reward_precision == CANTOR_F64, done, terminal_obs, clip_obs, clip_reward,
                        epsilon, norm_obs, norm_reward, s);
}

// occupancy-sized grids (cached per device: the query costs microseconds and this runs every env-step)
static int vn_occupancy(int* occ_moments, int* occ_apply, int* n_sm) {
    static int occ_cache[64][3];                                             // [device] = {CTAs/SM moments, CTAs/SM apply, SM count}
    int dev = 0;
    CANTOR_CUDA(cudaGetDevice(&dev));
    CANTOR_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
    if (occ_cache[dev][2] == 0) {
        int om = 0, oa = 0, sm = 0;
        CANTOR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&om, vecnorm_moments_kernel, kVnThreads, 0));
        CANTOR_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&oa, vecnorm_apply_kernel, kVnThreads, 0));
        CANTOR_CUDA(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev));
        occ_cache[dev][0] = om;
        occ_cache[dev][1] = oa;
        occ_cache[dev][2] = sm;                                              // written last: a racing thread recomputes the same values
    }
    *occ_moments = occ_cache[dev][0];
    *occ_apply = occ_cache[dev][1];
    *n_sm = occ_cache[dev][2];
    return CANTOR_OK;
}

static int launch_apply(double* rms, double* returns, long long n, float* obs, void* reward, int f64, const uint8_t* done,
                        float* terminal_obs, double clip_obs, double clip_reward, double epsilon, int norm_obs, int norm_reward,
                        cudaStream_t s) {
    int om, oa, n_sm;
    int rc = vn_occupancy(&om, &oa, &n_sm);
    if (rc) return rc;
    const long long n_chunks = (n + kVnRows - 1) / kVnRows;
    const long long cap_a = (long long)n_sm * (oa > 0 ? oa : 1);
    const unsigned grid_apply = (unsigned)(n_chunks < cap_a ? n_chunks : cap_a);
    int vec_ok = aligned16(obs) && (terminal_obs == nullptr || aligned16(terminal_obs)) ? 1 : 0;
    const double* rms_c = rms;
    void* a3[] = {&rms_c, &returns, &n, &obs, &reward, (void*)&f64, &done, &terminal_obs, &clip_obs, &clip_reward, &norm_obs,
                  &norm_reward, &vec_ok, &epsilon};
    return launch_pdl((const void*)vecnorm_apply_kernel, dim3(grid_apply), dim3(kVnThreads), s, a3);
}

extern "C" int cantor_vecnorm_step(double* rms, double* returns, int64_t n_envs, float* obs, void* reward,
                                   int32_t reward_precision, const uint8_t* done, float* terminal_obs, double gamma,
                                   double clip_obs, double clip_reward, double epsilon, int32_t training, int32_t norm_obs,
                                   int32_t norm_reward, void* stream) {
    CANTOR_REQUIRE(rms && returns && obs && reward && done, "array is NULL");
    CANTOR_REQUIRE(reward_precision == CANTOR_F32 || reward_precision == CANTOR_F64, "reward_precision must be 32 or 64");
    CANTOR_REQUIRE(n_envs >= 0, "n_envs < 0");
    if (n_envs == 0) return CANTOR_OK;
    cudaStream_t s = (cudaStream_t)stream;
    long long n = n_envs;
    const int f64 = reward_precision == CANTOR_F64;
    const long long n_chunks = (n + kVnRows - 1) / kVnRows;
    // grid-stride over chunks of 128 envs with exactly one resident wave: SMs x (CTAs that fit per SM), so no partial tail wave
    int om, oa, n_sm;
    int rc = vn_occupancy(&om, &oa, &n_sm);
    if (rc) return rc;
    const long long cap_m = (long long)n_sm * (om > 0 ? om : 1);
    const unsigned grid = (unsigned)(n_chunks < cap_m ? n_chunks : (cap_m < kVnMaxGrid ? cap_m : kVnMaxGrid));
    int vec_ok = aligned16(obs) && (terminal_obs == nullptr || aligned16(terminal_obs)) ? 1 : 0;
    const float* obs_c = obs;
    const void* rew_c = reward;
    if (training) {
        void* a1[] = {&rms, &returns, &n, &obs_c, &rew_c, (void*)&f64, &gamma, &norm_obs, &norm_reward, &vec_ok, &epsilon};
        rc = launch_pdl((const void*)vecnorm_moments_kernel, dim3(grid), dim3(kVnThreads), s, a1);
        if (rc) return rc;
    }
    return launch_apply(rms, returns, n, obs, reward, f64, done, terminal_obs, clip_obs, clip_reward, epsilon, norm_obs, norm_reward, s);
}
