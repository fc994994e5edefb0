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
// Three small kernels chained with programmatic dependent launch: batch moments (coalesced, column-aligned grid-stride
// loop + one atomic per column per CTA), the fold (one warp), the in-place apply.  HBM-bound: 52 B read for the
// moments, 104 + ~18 B for the apply, per env-step.
#include "common.cuh"

namespace cantor {

constexpr int kVnCols = CANTOR_OBS_DIM;            // 13
constexpr int kVnThreads = kVnCols * 32;           // 416: a thread's flat index keeps (index % 13) fixed across the stride
// rms[] layout (doubles)
constexpr int kObsMean = 0, kObsVar = 13, kObsCount = 26, kRetMean = 27, kRetVar = 28, kRetCount = 29;
constexpr int kScratch = 32;                       // [32..58): batch sums: obs sum[13], obs sumsq[13], ret sum, ret sumsq
constexpr int kDerived = 64;                       // [64..78): float-free derived values for the apply kernel: inv_std[13], ret_inv_std

__global__ void __launch_bounds__(kVnThreads)
vecnorm_moments_kernel(double* __restrict__ rms, double* __restrict__ returns, long long n, const float* __restrict__ obs,
                       const void* __restrict__ reward, int reward_f64, double gamma, int norm_obs, int norm_reward) {
    __shared__ double s_sum[kVnThreads], s_sq[kVnThreads];
    pdl_wait_prior_grid();
    const long long total = n * kVnCols;
    const long long stride = (long long)gridDim.x * kVnThreads;              // multiple of 13: the column of a thread is fixed
    double sum = 0.0, sq = 0.0;
    if (norm_obs) {
        for (long long e = (long long)blockIdx.x * kVnThreads + threadIdx.x; e < total; e += stride) {
            const double x = (double)__ldg(obs + e);
            sum += x;
            sq += x * x;
        }
    }
    s_sum[threadIdx.x] = sum;
    s_sq[threadIdx.x] = sq;
    // discounted returns (one thread per env), their batch sums in registers
    double rsum = 0.0, rsq = 0.0;
    if (norm_reward) {
        for (long long i = (long long)blockIdx.x * kVnThreads + threadIdx.x; i < n; i += stride) {
            const double r = reward_f64 ? reinterpret_cast<const double*>(reward)[i] : (double)reinterpret_cast<const float*>(reward)[i];
            const double ret = returns[i] * gamma + r;
            returns[i] = ret;
            rsum += ret;
            rsq += ret * ret;
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        rsum += __shfl_down_sync(0xffffffffu, rsum, off);
        rsq += __shfl_down_sync(0xffffffffu, rsq, off);
    }
    __syncthreads();
    pdl_launch_dependents();
    if (norm_obs && threadIdx.x < kVnCols) {                                 // column c: threads c, c + 13, c + 26, ...
        double a = 0.0, q = 0.0;
        for (int j = threadIdx.x; j < kVnThreads; j += kVnCols) {
            a += s_sum[j];
            q += s_sq[j];
        }
        atomicAdd(rms + kScratch + threadIdx.x, a);
        atomicAdd(rms + kScratch + kVnCols + threadIdx.x, q);
    }
    if (norm_reward && (threadIdx.x & 31) == 0) {
        atomicAdd(rms + kScratch + 2 * kVnCols, rsum);
        atomicAdd(rms + kScratch + 2 * kVnCols + 1, rsq);
    }
}

// RunningMeanStd.update_from_moments for each observation column and for the returns; clears the scratch sums.
__global__ void vecnorm_fold_kernel(double* __restrict__ rms, long long n, double epsilon, int training, int norm_obs,
                                    int norm_reward) {
    pdl_wait_prior_grid();
    const int j = threadIdx.x;
    const double bc = (double)n;
    if (j < kVnCols) {
        if (training && norm_obs) {
            const double bmean = rms[kScratch + j] / bc;
            const double bvar = fmax(rms[kScratch + kVnCols + j] / bc - bmean * bmean, 0.0);
            const double count = rms[kObsCount], mean = rms[kObsMean + j], var = rms[kObsVar + j];
            const double tot = count + bc, delta = bmean - mean;
            rms[kObsMean + j] = mean + delta * bc / tot;
            rms[kObsVar + j] = (var * count + bvar * bc + delta * delta * count * bc / tot) / tot;
        }
        rms[kDerived + j] = 1.0 / sqrt(rms[kObsVar + j] + epsilon);
    } else if (j == kVnCols) {
        if (training && norm_reward) {
            const double bmean = rms[kScratch + 2 * kVnCols] / bc;
            const double bvar = fmax(rms[kScratch + 2 * kVnCols + 1] / bc - bmean * bmean, 0.0);
            const double count = rms[kRetCount], mean = rms[kRetMean], var = rms[kRetVar];
            const double tot = count + bc, delta = bmean - mean;
            rms[kRetMean] = mean + delta * bc / tot;
            rms[kRetVar] = (var * count + bvar * bc + delta * delta * count * bc / tot) / tot;
            rms[kRetCount] = tot;
        }
        rms[kDerived + kVnCols] = 1.0 / sqrt(rms[kRetVar] + epsilon);
    }
    __syncwarp();
    if (j == 0 && training && norm_obs) rms[kObsCount] += bc;                  // after every column used the old count
    if (j < 2 * kVnCols + 2) rms[kScratch + j] = 0.0;
    pdl_launch_dependents();
}

__global__ void __launch_bounds__(kVnThreads)
vecnorm_apply_kernel(const double* __restrict__ rms, double* __restrict__ returns, long long n, float* __restrict__ obs,
                     void* __restrict__ reward, int reward_f64, const unsigned char* __restrict__ done,
                     float* __restrict__ terminal_obs, double clip_obs, double clip_reward, int norm_obs, int norm_reward) {
    pdl_wait_prior_grid();
    const long long total = n * kVnCols;
    const long long stride = (long long)gridDim.x * kVnThreads;
    const int col = threadIdx.x % kVnCols;                                   // fixed across the stride
    if (norm_obs) {
        const double mean = rms[kObsMean + col], inv = rms[kDerived + col];
        for (long long e = (long long)blockIdx.x * kVnThreads + threadIdx.x; e < total; e += stride) {
            obs[e] = (float)fmin(fmax(((double)obs[e] - mean) * inv, -clip_obs), clip_obs);
            if (terminal_obs != nullptr && done[e / kVnCols])
                terminal_obs[e] = (float)fmin(fmax(((double)terminal_obs[e] - mean) * inv, -clip_obs), clip_obs);
        }
    }
    const double rinv = rms[kDerived + kVnCols];
    for (long long i = (long long)blockIdx.x * kVnThreads + threadIdx.x; i < n; i += stride) {
        if (norm_reward) {
            if (reward_f64) {
                double* r = reinterpret_cast<double*>(reward) + i;
                *r = fmin(fmax(*r * rinv, -clip_reward), clip_reward);
            } else {
                float* r = reinterpret_cast<float*>(reward) + i;
                *r = (float)fmin(fmax((double)*r * rinv, -clip_reward), clip_reward);
            }
        }
        if (done[i]) returns[i] = 0.0;
    }
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_vecnorm_init(double* rms, double* returns, int64_t n_envs, void* stream) {
    CANTOR_REQUIRE(rms != nullptr && returns != nullptr && n_envs >= 0, "bad arguments");
    double h[CANTOR_VECNORM_DOUBLES] = {0};
    for (int j = 0; j < kVnCols; ++j) h[kObsVar + j] = 1.0;
    h[kObsCount] = 1e-4;
    h[kRetVar] = 1.0;
    h[kRetCount] = 1e-4;
    cudaStream_t s = (cudaStream_t)stream;
    CANTOR_CUDA(cudaMemcpyAsync(rms, h, sizeof(h), cudaMemcpyHostToDevice, s));
    CANTOR_CUDA(cudaStreamSynchronize(s));                                   // h is a stack buffer
    CANTOR_CUDA(cudaMemsetAsync(returns, 0, sizeof(double) * (size_t)n_envs, s));
    return CANTOR_OK;
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
    const long long want = (n * kVnCols + kVnThreads - 1) / kVnThreads;
    const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);       // grid-stride: at most 8 CTAs per SM
    int rc;
    const float* obs_c = obs;
    const void* rew_c = reward;
    if (training) {
        void* a1[] = {&rms, &returns, &n, &obs_c, &rew_c, (void*)&f64, &gamma, &norm_obs, &norm_reward};
        rc = launch_pdl((const void*)vecnorm_moments_kernel, dim3(grid), dim3(kVnThreads), s, a1);
        if (rc) return rc;
    }
    void* a2[] = {&rms, &n, &epsilon, &training, &norm_obs, &norm_reward};
    rc = launch_pdl((const void*)vecnorm_fold_kernel, dim3(1), dim3(32), s, a2);
    if (rc) return rc;
    const double* rms_c = rms;
    void* a3[] = {&rms_c, &returns, &n, &obs, &reward, (void*)&f64, &done, &terminal_obs, &clip_obs, &clip_reward, &norm_obs,
                  &norm_reward};
    return launch_pdl((const void*)vecnorm_apply_kernel, dim3(grid), dim3(kVnThreads), s, a3);
}
