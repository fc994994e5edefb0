// Episode-fused rollout: path step (replayed or simulated on the fly) -> ATM repricing -> observation ->
// policy -> fused hedge step -> episode statistics, for many consecutive steps in ONE kernel with the whole
// env state in registers.  Nothing is written per step unless rollout storage is requested, so this mode is
// bound by the FP32 / SFU pipes, not by HBM.  Episode statistics are reduced warp -> block -> one atomic per
// block, giving the small vector (and histogram) that the multi-GPU all-reduce sums.
//
// Reference semantics: the evaluation loops around HedgingEnv --
//   evaluate_baseline_policy      src/agents/baselines.py:32-72       (per-episode mean |dPnL|, mean cost)
//   run_evaluation statistics     src/agents/train_ppo_v2.py:482-530  (|sum dPnL| / T, cost / T, CVaR95)
//   policy_delta_every_step       src/agents/baselines.py:77-103
//   delta_hedging_action_selector src/benchmark/delta_and_nothing.py:122-163
//   random policy                 src/agents/test_inf.py:29 (action_space.sample())
//   MLP actor 13-64-64-2          quantconnect/model_wrapper.py:131,177-185 (normalise, ReLU MLP, clip)
// and the env step itself (hedge_core.cuh).
#include "hedge_core.cuh"
#include "lstm_tc.cuh"
#include "mlp_tc.cuh"
#include "sim_core.cuh"

namespace cantor {

constexpr int kRollThreads = 128;
#ifndef CANTOR_MLP_TC_BLOCKS
#define CANTOR_MLP_TC_BLOCKS 6       // resident CTAs per SM of the tensor-core MLP actor (80 registers, 36 KB of shared memory each)
#endif
// How many copies of the step body the inner loop (one Philox call = up to 4 steps) is unrolled into.  The body is ~1 000
// SASS instructions; four copies (64 KB) showed 23 % of the stall samples in `no_instruction`.  Measured on the B200
// (2^20 envs x 252 steps, GBM on the fly, delta policy): 1 copy 2.09 ms, 2 copies 2.04 ms, 4 copies 2.12 ms; the recurrent
// actor, with one warp per scheduler, wants a single copy (84.9 vs 88.0 ms).
constexpr int kRolloutMaxUnroll = 2;
constexpr unsigned kStreamActions = 0x4143544Eu;   // "ACTN": 4th counter word of the random-policy stream
constexpr int kMlpIn = 13, kMlpHidden = 64, kMlpOut = 2;
constexpr int kMlpFloats = kMlpIn * kMlpHidden + kMlpHidden + kMlpHidden * kMlpHidden + kMlpHidden +
                           kMlpHidden * kMlpOut + kMlpOut + 2 * kMlpIn;          // 5212

struct PolicyConsts {
    int kind, put_disabled, squash;
    float obs_clip;             // network policies: clip of the normalised observation (+inf = none)
    const float* mlp;
    const float2* actions;      // CANTOR_POLICY_ACTIONS: [n_steps, n_envs] open-loop actions
    unsigned seed_lo, seed_hi;
};

struct RolloutOut {
    float* obs;                 // [n_steps, n_envs, 13] observation the policy saw
    float2* actions;            // [n_steps, n_envs]
    float* reward;              // [n_steps, n_envs]
    unsigned char* done;        // [n_steps, n_envs]
};

// ---- policies ------------------------------------------------------------------------------------------
// baselines.py:77-103 evaluated in float32 like NumPy does on a float32 observation (python ints are weak).
// Returns CONTRACT COUNTS clipped to +-max_trade, which the env then multiplies by max_trade again and clips
// (the reference's behaviour).
__device__ __forceinline__ float2 policy_delta_baselines(const float* o, const StepConsts& k) {
    const float cd = o[7], pd = o[9];
    const float mc = (float)k.max_contracts;
    const float cpos = __fmul_rn(o[3], mc), ppos = __fmul_rn(o[4], mc);
    const float opt_delta = __fmul_rn(__fadd_rn(__fmul_rn(cpos, cd), __fmul_rn(ppos, pd)), k.mult_f);
    const float target = -__fadd_rn(k.shares_f, opt_delta);
    float tc = 0.f, tp = 0.f;
    const float cdm = __fmul_rn(cd, k.mult_f), pdm = __fmul_rn(pd, k.mult_f);
    if (fabsf(cdm) > 0.1f) tc = __fdiv_rn(target, cdm);
    else if (fabsf(pdm) > 0.1f) tp = __fdiv_rn(target, pdm);
    return make_float2(fminf(fmaxf(tc, -k.max_trade_f), k.max_trade_f), fminf(fmaxf(tp, -k.max_trade_f), k.max_trade_f));
}

// delta_and_nothing.py:122-163 (float32 here; the reference mixes float32 / float64).
__device__ __forceinline__ float2 policy_delta_benchmark(const float* o, const StepConsts& k, int pos_c, int pos_p) {
    const float cd = o[7], pd = o[9];
    const float cur = ((float)pos_c * cd + (float)pos_p * pd) * k.mult_f;
    const float change = -k.shares_f - cur;
    if (fabsf(change) < 0.5f * fabsf(cd) * k.mult_f) return make_float2(0.f, 0.f);
    float rc = 0.f, rp = 0.f;
    if (change > 0.f) {
        if (fabsf(cd) > 1e-6f) rc = fminf(fmaxf(change / (cd * k.mult_f), -k.max_trade_f), k.max_trade_f);
    } else if (change < 0.f) {
        if (fabsf(pd) > 1e-6f) rp = fminf(fmaxf(change / (pd * k.mult_f), -k.max_trade_f), k.max_trade_f);
    }
    return make_float2(rc, rp);
}

// uniform(-1, 1) pair from Philox(seed; global env, global step, "ACTN")
__device__ __forceinline__ float2 policy_random(const PolicyConsts& pc, unsigned long long genv, unsigned step) {
    const uint4 x = philox4x32_10(make_uint4((unsigned)genv, (unsigned)(genv >> 32), step, kStreamActions),
                                  make_uint2(pc.seed_lo, pc.seed_hi));
    return make_float2(fmaf((float)(x.x >> 8), 1.1920928955078125e-07f, -1.0f),      // [-1, 1)
                       fmaf((float)(x.y >> 8), 1.1920928955078125e-07f, -1.0f));
}

// ReLU MLP 13 -> 64 -> 64 -> 2 on the normalised observation; weights broadcast from shared memory.
// noinline: inlined into the 4x-unrolled step loop the 64-wide accumulator arrays no longer fit the register file
__device__ __noinline__ float2 policy_mlp(const float* o, const float* __restrict__ w, float obs_clip) {
    const float* W1 = w;
    const float* b1 = W1 + kMlpIn * kMlpHidden;
    const float* W2 = b1 + kMlpHidden;
    const float* b2 = W2 + kMlpHidden * kMlpHidden;
    const float* W3 = b2 + kMlpHidden;
    const float* b3 = W3 + kMlpHidden * kMlpOut;
    const float* mean = b3 + kMlpOut;
    const float* inv_std = mean + kMlpIn;
    float x[kMlpIn];
#pragma unroll
    for (int i = 0; i < kMlpIn; ++i) x[i] = fminf(fmaxf((o[i] - mean[i]) * inv_std[i], -obs_clip), obs_clip);   // VecNormalize clip
    float h1[kMlpHidden];
#pragma unroll
    for (int j = 0; j < kMlpHidden; ++j) h1[j] = b1[j];
#pragma unroll
    for (int i = 0; i < kMlpIn; ++i) {
#pragma unroll
        for (int j = 0; j < kMlpHidden; j += 4) {
            const float4 wv = *reinterpret_cast<const float4*>(W1 + i * kMlpHidden + j);
            h1[j] = fmaf(x[i], wv.x, h1[j]); h1[j + 1] = fmaf(x[i], wv.y, h1[j + 1]);
            h1[j + 2] = fmaf(x[i], wv.z, h1[j + 2]); h1[j + 3] = fmaf(x[i], wv.w, h1[j + 3]);
        }
    }
    float out0 = b3[0], out1 = b3[1];
    // layer 2 is produced 4 outputs at a time and consumed immediately by layer 3 (h2 never materialised)
#pragma unroll 1
    for (int j = 0; j < kMlpHidden; j += 4) {
        float a0 = b2[j], a1 = b2[j + 1], a2 = b2[j + 2], a3 = b2[j + 3];
#pragma unroll
        for (int i = 0; i < kMlpHidden; ++i) {
            const float hi = fmaxf(h1[i], 0.f);
            const float4 wv = *reinterpret_cast<const float4*>(W2 + i * kMlpHidden + j);
            a0 = fmaf(hi, wv.x, a0); a1 = fmaf(hi, wv.y, a1); a2 = fmaf(hi, wv.z, a2); a3 = fmaf(hi, wv.w, a3);
        }
        a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f);
        out0 = fmaf(a0, W3[(j + 0) * 2], out0); out1 = fmaf(a0, W3[(j + 0) * 2 + 1], out1);
        out0 = fmaf(a1, W3[(j + 1) * 2], out0); out1 = fmaf(a1, W3[(j + 1) * 2 + 1], out1);
        out0 = fmaf(a2, W3[(j + 2) * 2], out0); out1 = fmaf(a2, W3[(j + 2) * 2 + 1], out1);
        out0 = fmaf(a3, W3[(j + 3) * 2], out0); out1 = fmaf(a3, W3[(j + 3) * 2 + 1], out1);
    }
    return make_float2(out0, out1);                     // action means; the caller squashes them
}

// SRC: 0 replay (packed book), 1 GBM on the fly, 2 Heston on the fly.
// MLP: 0 = one of the closed-form / tabulated policies, 1 = the MLP actor in float32 FFMAs (parity form),
//      2 = the MLP actor on the tensor cores (bf16 tcgen05.mma, mlp_tc.cuh), 3 = the recurrent LSTM + MLP actor on the
//      tensor cores (lstm_tc.cuh); compile-time, so the other policies do not pay for its registers and shared memory.
//      The recurrent form runs 256 envs per CTA (two groups of 128 that ping-pong) plus a ninth warp that only issues MMAs /
//      TMA copies (lstmtc::Actor::issuer_loop).
template <int SRC, int MLP, bool WRITE>
__global__ void __launch_bounds__(MLP == 3 ? lstmtc::kThreads : kRollThreads, MLP == 2 ? CANTOR_MLP_TC_BLOCKS : 0)   // 0 = no occupancy hint (a hint of 1 let ptxas take 104 registers for the plain policies: 15.8 -> 17.9 ms)
rollout_kernel(const StepConsts k, const Book b, const SimConsts sk, const PolicyConsts pc, long long n_envs,
               long long env_offset, long long total_envs, long long first_episode, int n_steps, const StatsOut st,
               const RolloutOut out, int obs_tma_ok, int share_quote) {
    extern __shared__ __align__(128) float smem_f[];
    float* w_mlp = smem_f;                                                     // [kMlpFloats] float32 weights (MLP == 1)
    constexpr int mlp_floats = MLP == 1 ? (kMlpFloats + 3) / 4 * 4 : (MLP == 2 ? mlptc::kSmemBytes / 4 : (MLP == 3 ? lstmtc::kSmemBytes / 4 : 0));
    constexpr int kEnv = MLP == 3 ? lstmtc::kEnvs : kRollThreads;             // envs (= env threads) per CTA
    constexpr int kThreads = MLP == 3 ? lstmtc::kThreads : kRollThreads;
    mlptc::Actor actor;
    lstmtc::Actor lstm;
    if (MLP == 3) lstm.setup(reinterpret_cast<unsigned char*>(smem_f), reinterpret_cast<const unsigned char*>(pc.mlp), pc.obs_clip);
    // [kEnv * 13] observation staging tile when WRITE; the recurrent actor lends its A2 tiles for it (no shared memory left):
    // one [128 x 13] piece per group
    float* tile = MLP == 3 ? lstm.obs_staging(0) : smem_f + mlp_floats;
    auto tile_row = [&](int r) { return MLP == 3 ? lstm.obs_staging(r >> 7) + (r & 127) * CANTOR_OBS_DIM : tile + r * CANTOR_OBS_DIM; };
    double* red = reinterpret_cast<double*>(smem_f + mlp_floats + ((WRITE && MLP != 3) ? kEnv * CANTOR_OBS_DIM : 0));
    if (MLP == 1) {
        for (int j = threadIdx.x; j < kMlpFloats; j += kRollThreads) w_mlp[j] = pc.mlp[j];
    }
    if (MLP == 2) actor.setup(reinterpret_cast<unsigned char*>(smem_f), pc.mlp, pc.obs_clip);
    __syncthreads();

    const bool env_thread = threadIdx.x < kEnv;                                // the others: the recurrent actor's issuer warp
    auto env_sync = [&]() {                                                    // barrier of the env threads only
        if (MLP == 3) lstmtc::env_sync();
        else __syncthreads();
    };
    const long long first_env = (long long)blockIdx.x * kEnv;
    const long long i = first_env + threadIdx.x;
    const bool live = env_thread && i < n_envs;
    const int rows = (int)min((long long)kEnv, n_envs - first_env);
    const unsigned long long genv = (unsigned long long)(env_offset + (live ? i : 0));
    constexpr int MODEL = SRC == 2 ? 1 : 0;
    constexpr int NPS = SRC == 2 ? 2 : 1;
    constexpr int STEPS_PER_CALL = SRC == 0 ? 1 : 4 / NPS;

    // env state, all in registers
    int pos_c = 0, pos_p = 0, t = 0;
    long long episode = first_episode;
    const unsigned first_step = (unsigned)((unsigned long long)first_episode * (unsigned long long)k.T);   // random-action stream continues too
    unsigned long long gp = genv;                                              // global path of the current episode
    float4 cur = make_float4(0.f, 0.f, 0.f, 0.f), prev;
    Greeks gk{0.f, 0.f, 0.f};                                                  // ATM greeks of `cur` (on-the-fly, shared with its price)
    float S = 0.f, v = 0.f, s0 = 1.f, inv_s0 = 1.f;
    float acc_pps = 0.f, acc_abs = 0.f, acc_cost = 0.f, acc_reward = 0.f;
    double stat[11];
#pragma unroll
    for (int s = 0; s < 11; ++s) stat[s] = 0.0;

    auto begin_episode = [&]() {
        gp = (unsigned long long)episode * (unsigned long long)total_envs + genv;
        if (SRC == 0) {
            cur = b.rec[(long long)(gp % (unsigned long long)b.n_paths)];
        } else {
            S = sk.s0;
            v = sk.v0;
            cur.x = S;
            cur.y = fmaxf(v, 0.f);
            const AtmQuote q = atm_quote_f32(cur.x, cur.y, sk);
            cur.z = q.call;
            cur.w = q.put;
            gk = q.g;
        }
        prev = cur;
        s0 = (cur.x < 1e-6f) ? 1.0f : cur.x;                                   // hedging_env_v2.py:157
        inv_s0 = mufu_rcp(fmaxf(s0, 25.0f));
        pos_c = pos_p = 0;
        t = 0;
        acc_pps = acc_abs = acc_cost = acc_reward = 0.f;
    };
    if (env_thread) begin_episode();
    if (MLP == 3 && !env_thread) lstm.issuer_loop(n_steps, k.T);

    int g = env_thread ? 0 : n_steps;                                          // global step of the rollout
    while (g < n_steps) {
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        if (SRC != 0) path_normals(sk, gp, (unsigned)((t * NPS) >> 2), z);
#pragma unroll(MLP == 3 ? 1 : kRolloutMaxUnroll)          // see kRolloutMaxUnroll (a factor >= the trip count unrolls fully)
        for (int j = 0; j < STEPS_PER_CALL; ++j) {
            if (g >= n_steps) break;
            // ---- observation of the current state and the policy's action --------------------------------
            float o[CANTOR_OBS_DIM];
            // on the fly, the greeks of this state came with its price one step ago (same r, tenor: share_quote)
            if (SRC != 0 && share_quote) make_observation_f32(o, k, cur.x, cur.y, cur.z, cur.w, inv_s0, pos_c, pos_p, t, prev.x, prev.y, gk);
            else make_observation_f32(o, k, cur.x, cur.y, cur.z, cur.w, inv_s0, pos_c, pos_p, t, prev.x, prev.y);
            // ---- next path record: does not depend on the action.  The recurrent actor computes it between its gate phase and its head
            // (under the head's first product, where the thread would only wait); the other policies after the action, as before.
            float4 nxt;
            auto next_record = [&]() {
                if (SRC == 0) {
                    nxt = live ? __ldcs(b.rec + ((long long)(t + 1) * b.ld + (long long)(gp % (unsigned long long)b.n_paths))) : cur;
                } else {
                    sim_advance<MODEL>(sk, S, v, z[NPS * j], z[NPS * j + NPS - 1]);
                    nxt.x = S;
                    nxt.y = fmaxf(v, 0.f);
                    nxt.z = cur.z;
                    nxt.w = cur.w;                                                 // stale marks at the terminal step (:226-231)
                    if (t + 1 < k.T) {
                        const AtmQuote q = atm_quote_f32(nxt.x, nxt.y, sk);
                        nxt.z = q.call;
                        nxt.w = q.put;
                        gk = q.g;
                    }
                }
            };
            float2 a;
            if (MLP == 3) {
                lstm.forward_gates(o);                                         // CTA-collective; carries h, c across steps
                next_record();
                a = lstm.forward_head();
            } else if (MLP == 2) {
                a = actor.forward(o);                                          // CTA-collective: all 128 threads, every step
            } else if (MLP == 1) {
                a = policy_mlp(o, w_mlp, pc.obs_clip);
            } else {
                switch (pc.kind) {
                    case CANTOR_POLICY_RANDOM: a = policy_random(pc, genv, first_step + (unsigned)g); break;
                    case CANTOR_POLICY_DELTA_BASELINES: a = policy_delta_baselines(o, k); break;
                    case CANTOR_POLICY_DELTA_BENCHMARK: a = policy_delta_benchmark(o, k, pos_c, pos_p); break;
                    case CANTOR_POLICY_ACTIONS: a = live ? __ldcs(pc.actions + (long long)g * n_envs + i) : make_float2(0.f, 0.f); break;
                    default: a = make_float2(0.f, 0.f);
                }
            }
            if (MLP != 0) {                                                    // network policies return action means
                if (pc.squash == CANTOR_SQUASH_TANH) a = make_float2(tanhf(a.x), tanhf(a.y));         // quantconnect/model_wrapper.py:202
                else a = make_float2(fminf(fmaxf(a.x, -1.f), 1.f), fminf(fmaxf(a.y, -1.f), 1.f));   // SB3: clip to the Box
            }
            if (pc.put_disabled) a.y = 0.f;
            if (MLP != 3) next_record();
            // ---- fused hedge step ---------------------------------------------------------------------------
            const LedgerF32 L = ledger_f32(k, a.x, a.y, pos_c, pos_p, t, inv_s0, cur, nxt, false);
            const bool terminated = t + 1 >= k.T;
            if (WRITE) {
                float* orow = tile_row(threadIdx.x);
#pragma unroll
                for (int q = 0; q < CANTOR_OBS_DIM; ++q) orow[q] = o[q];
                if (MLP == 3) {
                    // the recurrent actor's groups are half a step apart: each group of 128 stores its own [128 x 13] piece behind its
                    // own named barrier (ids 2 / 3, 128 threads), never waiting for the other group
                    const int grp = threadIdx.x >> 7, m = threadIdx.x & 127;
                    const int grows = max(0, min(rows - 128 * grp, 128));
                    float* dst = out.obs + ((long long)g * n_envs + first_env + 128 * grp) * CANTOR_OBS_DIM;
                    const bool tma = obs_tma_ok && (grows % 4 == 0) && ((((long long)g * n_envs) & 3) == 0);
                    if (tma) fence_proxy_async_smem();
                    lstmtc::group_sync(grp);
                    if (tma) {
                        if (m == 0 && grows > 0) {
                            tma_store_1d_evict_first(dst, lstm.obs_staging(grp), (uint32_t)(grows * CANTOR_OBS_DIM * sizeof(float)));
                            tma_store_commit();
                            tma_store_wait_read();
                        }
                    } else {
                        const float* src = lstm.obs_staging(grp);
                        for (int q = m; q < grows * CANTOR_OBS_DIM; q += 128) dst[q] = src[q];
                    }
                    lstmtc::group_sync(grp);
                } else {
                float* dst = out.obs + ((long long)g * n_envs + first_env) * CANTOR_OBS_DIM;
                const bool tma = obs_tma_ok && (rows % 4 == 0) && ((((long long)g * n_envs) & 3) == 0);
                if (tma) {
                    fence_proxy_async_smem();
                    env_sync();
                    if (threadIdx.x == 0) {
                        tma_store_1d_evict_first(dst, tile, (uint32_t)(rows * CANTOR_OBS_DIM * sizeof(float)));
                        tma_store_commit();
                        tma_store_wait_read();
                    }
                    env_sync();
                } else {
                    env_sync();
                    for (int q = threadIdx.x; q < rows * CANTOR_OBS_DIM; q += kEnv) dst[q] = tile[q];
                    env_sync();
                }
                }
                if (live) {
                    const long long at = (long long)g * n_envs + i;
                    __stcs(out.actions + at, a);
                    __stcs(out.reward + at, L.reward);
                    out.done[at] = terminated ? 1 : 0;
                }
            }
            acc_pps += L.pps;
            acc_abs += fabsf(L.pps);
            acc_cost += L.costs;
            acc_reward += L.reward;
            pos_c = L.new_c;
            pos_p = L.new_p;
            prev = cur;
            cur = nxt;
            ++t;
            ++g;
            if (terminated) {
                if (live) {
                    const float eb = episode_statistics(stat, acc_reward, acc_pps, acc_abs, acc_cost, k.inv_T_f, st);
                    if (st.episode_b != nullptr && episode - first_episode < st.episode_slots)
                        st.episode_b[(episode - first_episode) * n_envs + i] = eb;
                }
                ++episode;
                begin_episode();
                if (MLP == 3) lstm.reset_state();                              // SB3 zeroes the LSTM state at an episode start
                break;                                                         // the new path starts a new Philox call
            }
        }
    }
    // ---- one reduction at the end: warp shuffle -> shared -> one atomic per statistic per block ---------------
    if (st.sums != nullptr) block_accumulate<11, kEnv>(stat, st.sums, red);          // warps beyond the env warps only join its barriers
    if (st.sums != nullptr && threadIdx.x == 0 && blockIdx.x == 0) atomicAdd(st.sums + 11, (double)n_envs * (double)n_steps);
    if (MLP == 2) {
        if (actor.timed_out && st.sums != nullptr) atomicAdd(st.sums + 15, 1.0);   // an MMA never completed: results are invalid
        actor.teardown();
    }
    if (MLP == 3) {
        if (lstm.timed_out && st.sums != nullptr) atomicAdd(st.sums + 15, 1.0);
        lstm.teardown();
    }
    // fused all-reduce: the last CTA adds this launch's statistics into every rank's global block (NVLS multimem.red / peer atomics)
    push_statistics_to_all_ranks<kThreads>(st);
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_rollout(const cantor_env_params* params, const cantor_replay_book* book,
                              const cantor_sim_params* sim, int32_t episode_length, const cantor_policy* policy,
                              int64_t n_envs, int64_t env_offset, int64_t total_envs, int64_t first_episode, int32_t n_steps,
                              const cantor_stats_out* stats, const cantor_rollout_out* out, void* stream) {
    CANTOR_REQUIRE(params != nullptr && policy != nullptr, "params/policy is NULL");
    CANTOR_REQUIRE((book != nullptr) != (sim != nullptr), "exactly one of book / sim must be given");
    CANTOR_REQUIRE(n_envs > 0 && n_steps >= 0 && total_envs >= n_envs && env_offset >= 0 && first_episode >= 0, "bad sizes");
    CANTOR_REQUIRE(policy->kind >= CANTOR_POLICY_NO_HEDGE && policy->kind <= CANTOR_POLICY_LSTM, "policy kind");
    CANTOR_REQUIRE((policy->kind != CANTOR_POLICY_MLP && policy->kind != CANTOR_POLICY_LSTM) || policy->mlp != nullptr, "policy.mlp is NULL");
    CANTOR_REQUIRE(policy->kind != CANTOR_POLICY_LSTM || aligned16(policy->mlp), "the LSTM weight image must be 16-byte aligned");
    CANTOR_REQUIRE(policy->kind != CANTOR_POLICY_ACTIONS || policy->actions != nullptr, "policy.actions is NULL");
    StepConsts k;
    Book b{nullptr, 0, 1};
    SimConsts sk = {};
    int T = episode_length;
    int src = 0;
    if (book != nullptr) {
        int rc = make_book(book, &b);
        if (rc) return rc;
        T = book->episode_length;
    } else {
        CANTOR_REQUIRE(sim->model == CANTOR_MODEL_GBM || sim->model == CANTOR_MODEL_HESTON, "sim.model");
        CANTOR_REQUIRE(sim->dt > 0 && sim->tenor > 0, "dt and tenor must be positive");
        fill_sim_consts(sim, T, &sk);
        src = sim->model == CANTOR_MODEL_GBM ? 1 : 2;
    }
    int rc = make_step_consts(params, T, &k);
    if (rc) return rc;
    CANTOR_REQUIRE(policy->action_squash == CANTOR_SQUASH_CLIP || policy->action_squash == CANTOR_SQUASH_TANH, "policy.action_squash");
    const float obs_clip = (policy->obs_clip > 0.f) ? policy->obs_clip : __builtin_inff();      // <= 0 / NaN: no clip
    PolicyConsts pc{policy->kind, policy->put_leg_disabled, policy->action_squash, obs_clip, policy->mlp, (const float2*)policy->actions,
                    (unsigned)(policy->seed & 0xffffffffull), (unsigned)(policy->seed >> 32)};
    StatsOut so;
    rc = make_stats_out(stats, &so);
    if (rc) return rc;
    RolloutOut ro{nullptr, nullptr, nullptr, nullptr};
    const bool write = out != nullptr;
    if (write) {
        CANTOR_REQUIRE(out->obs && out->actions && out->reward && out->done, "rollout output array is NULL");
        ro = RolloutOut{out->obs, (float2*)out->actions, out->reward, out->done};
    }
    if (n_steps == 0) return CANTOR_OK;
    const int mlp_mode = policy->kind == CANTOR_POLICY_LSTM ? 3 : (policy->kind != CANTOR_POLICY_MLP ? 0 : (policy->mlp_tensor_cores ? 2 : 1));
    const int env_per_cta = mlp_mode == 3 ? lstmtc::kEnvs : kRollThreads;
    const unsigned grid = (unsigned)((n_envs + env_per_cta - 1) / env_per_cta);
    const size_t smem = (mlp_mode == 1 ? (kMlpFloats + 3) / 4 * 4 * sizeof(float)
                         : (mlp_mode == 2 ? (size_t)mlptc::kSmemBytes : (mlp_mode == 3 ? (size_t)lstmtc::kSmemBytes : 0))) +
                        ((write && mlp_mode != 3) ? kRollThreads * CANTOR_OBS_DIM * sizeof(float) : 0) + 11 * (env_per_cta / 32) * sizeof(double) + 16;
    const int tma_ok = write && aligned16(out->obs) ? 1 : 0;
    cudaStream_t s = (cudaStream_t)stream;
    // the observation's greeks can ride on the price evaluation when the env and the simulator agree on (r, tenor)
    const int share = (src != 0 && params->record_metrics && k.g.r_f == sk.r && k.g.T_f == sk.tenor && sk.tenor > 1e-6f) ? 1 : 0;
#define LAUNCH(SRC, MLP, WRITE)                                                                                                         \
    do {                                                                                                                             \
        if (smem > 48 * 1024) {                                                                                                      \
            cudaError_t e_ = cudaFuncSetAttribute(rollout_kernel<SRC, MLP, WRITE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e_ != cudaSuccess) return cuda_fail(e_, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");                         \
        }                                                                                                                            \
        rollout_kernel<SRC, MLP, WRITE><<<grid, MLP == 3 ? lstmtc::kThreads : kRollThreads, smem, s>>>(k, b, sk, pc, n_envs, env_offset, total_envs, first_episode, n_steps, so, ro, tma_ok, share); \
    } while (0)
#define LAUNCH_SRC(SRC)                                                      \
    do {                                                                     \
        if (mlp_mode == 3) { if (write) LAUNCH(SRC, 3, true); else LAUNCH(SRC, 3, false); }      \
        else if (mlp_mode == 2) { if (write) LAUNCH(SRC, 2, true); else LAUNCH(SRC, 2, false); } \
        else if (mlp_mode == 1) { if (write) LAUNCH(SRC, 1, true); else LAUNCH(SRC, 1, false); } \
        else { if (write) LAUNCH(SRC, 0, true); else LAUNCH(SRC, 0, false); }                    \
    } while (0)
    if (src == 0) LAUNCH_SRC(0);
    else if (src == 1) LAUNCH_SRC(1);
    else LAUNCH_SRC(2);
#undef LAUNCH_SRC
#undef LAUNCH
    return check_launch("rollout_kernel");
}

// ---------------------------------------------------------------------------------------------------
// cantor_umma_probe: what ONE tcgen05.mma (cta_group::1, kind::f16, bf16 operands, M = 128, K = 16) costs on this part as a function
// of N and of where the A operand lives -- the number the two tensor-core actors are designed around.  One CTA per SM,
// one thread issues `reps` back-to-back K-loops of `k_steps` MMAs into one accumulator and waits for the commit.
namespace cantor {
__global__ void __launch_bounds__(128) umma_probe_kernel(int n, int k_steps, int reps, int a_in_tmem, long long* out) {
    extern __shared__ __align__(1024) unsigned char probe_smem[];
    __shared__ __align__(8) unsigned long long bar_mem;
    __shared__ unsigned tmem_slot;
    const int a_bytes = 128 * 16 * k_steps * 2, b_bytes = n * 16 * k_steps * 2;
    for (int j = threadIdx.x; j < (a_bytes + b_bytes) / 16; j += 128) reinterpret_cast<uint4*>(probe_smem)[j] = make_uint4(0u, 0u, 0u, 0u);
    const uint32_t bar = mlptc::smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(mlptc::smem_u32(&tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_proxy_async_smem();
    mlptc::fence_before_sync();
    __syncthreads();
    mlptc::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    {   // tensor-memory columns [256, 256 + 8 k_steps): the A operand of the TS form (bf16 pairs; zeros)
        uint32_t z = 0u;
        for (int c = 0; c < 8 * k_steps; ++c)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(tmem + ((uint32_t)(threadIdx.x & 96) << 16) + 256 + c), "r"(z) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    mlptc::fence_before_sync();
    __syncthreads();
    mlptc::fence_after_sync();
    if (threadIdx.x == 0) {
        const int sbo = 2 * k_steps * 128;                                    // bytes between 8-row groups: K / 8 core matrices of 128 B
        const uint64_t da0 = mlptc::smem_desc(mlptc::smem_u32(probe_smem), mlptc::kLbo, sbo);
        const uint64_t db0 = mlptc::smem_desc(mlptc::smem_u32(probe_smem + a_bytes), mlptc::kLbo, sbo);
        const uint32_t idesc = mlptc::instr_desc(128, n);
        const long long t0 = clock64();
        // straight-line batches (mlptc::umma_batch: descriptors stepped on the uniform datapath), as the actors issue them -- a rolled
        // loop that rebuilds the descriptors per instruction costs ~120 clk per MMA in the ISSUING thread and hides the pipe's own rate
#define PROBE_TS_STEP(ACC) "tcgen05.mma.cta_group::1.kind::f16 [%0], [ta], db, %3, " ACC ";\n\tadd.u32 ta, ta, 8;\n\tadd.s64 db, db, 16;\n\t"
#define PROBE_TS_HEAD "{\n\t.reg .pred pt, pf;\n\t.reg .b32 ta;\n\t.reg .b64 db;\n\tsetp.eq.u32 pt, 0, 0;\n\tsetp.ne.u32 pf, 0, 0;\n\tmov.b32 ta, %1;\n\tmov.b64 db, %2;\n\t"
        for (int r = 0; r < reps; ++r) {
            if (a_in_tmem) {
                if (k_steps == 1)
                    asm volatile(PROBE_TS_HEAD PROBE_TS_STEP("pf") "}" :: "r"(tmem), "r"(tmem + 256), "l"(db0), "r"(idesc) : "memory");
                else if (k_steps == 5)
                    asm volatile(PROBE_TS_HEAD PROBE_TS_STEP("pf") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") "}"
                                 :: "r"(tmem), "r"(tmem + 256), "l"(db0), "r"(idesc) : "memory");
                else
                    asm volatile(PROBE_TS_HEAD PROBE_TS_STEP("pf") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt")
                                 PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") PROBE_TS_STEP("pt") "}"
                                 :: "r"(tmem), "r"(tmem + 256), "l"(db0), "r"(idesc) : "memory");
            } else {
                if (k_steps == 1) mlptc::umma_batch<1>(da0, db0, idesc, tmem);
                else if (k_steps == 5) mlptc::umma_batch<5>(da0, db0, idesc, tmem);
                else mlptc::umma_batch<9>(da0, db0, idesc, tmem);
            }
        }
#undef PROBE_TS_STEP
#undef PROBE_TS_HEAD
        const long long t1 = clock64();
        mlptc::umma_commit(bar);
        uint32_t done = 0;
        unsigned spins = 0;
        while (!done && ++spins < (1u << 26))
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
        const long long t2 = clock64();
        out[2 * blockIdx.x] = t1 - t0;                                        // issue
        out[2 * blockIdx.x + 1] = done ? t2 - t0 : -1;                        // issue + completion
    }
    mlptc::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
}
}  // namespace cantor

extern "C" int cantor_umma_probe(int32_t n, int32_t k_steps, int32_t reps, int32_t a_in_tmem, int32_t n_ctas,
                                 double* issue_clk_per_mma, double* total_clk_per_mma) {
    CANTOR_REQUIRE(n >= 8 && n <= 256 && n % 16 == 0, "n must be a multiple of 16 in [16, 256]");
    CANTOR_REQUIRE(k_steps == 1 || k_steps == 5 || k_steps == 9, "k_steps must be 1, 5 or 9 (the batch shapes the actors use)");
    CANTOR_REQUIRE(reps >= 1 && n_ctas >= 1 && n_ctas <= 1024, "reps / n_ctas");
    CANTOR_REQUIRE(issue_clk_per_mma && total_clk_per_mma, "output pointer is NULL");
    const size_t smem = (size_t)(128 + n) * 16 * k_steps * 2;
    cudaError_t e_ = cudaFuncSetAttribute(cantor::umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e_ != cudaSuccess) return cantor::cuda_fail(e_, "cudaFuncSetAttribute(MaxDynamicSharedMemorySize)");
    long long* out = nullptr;
    CANTOR_CUDA(cudaMalloc(&out, sizeof(long long) * 2 * n_ctas));
    // 120 KB of dynamic shared memory per CTA whatever the tile: one CTA per SM, like the actors that use all 512 TMEM columns
    cantor::umma_probe_kernel<<<n_ctas, 128, smem > 120 * 1024 ? smem : 120 * 1024>>>(n, k_steps, reps, a_in_tmem, out);
    long long* host = (long long*)malloc(sizeof(long long) * 2 * n_ctas);
    cudaError_t e2 = cudaMemcpy(host, out, sizeof(long long) * 2 * n_ctas, cudaMemcpyDeviceToHost);
    cudaFree(out);
    if (e2 != cudaSuccess) { free(host); return cantor::cuda_fail(e2, "umma_probe_kernel"); }
    double issue = 0.0, total = 0.0;
    bool ok = true;
    for (int c = 0; c < n_ctas; ++c) { issue += (double)host[2 * c]; total += (double)host[2 * c + 1]; ok = ok && host[2 * c + 1] >= 0; }
    free(host);
    if (!ok) return cantor::fail(CANTOR_ERR_CUDA, "cantor_umma_probe: an MMA batch never completed");
    *issue_clk_per_mma = issue / n_ctas / ((double)reps * k_steps);
    *total_clk_per_mma = total / n_ctas / ((double)reps * k_steps);
    return CANTOR_OK;
}

#ifdef LSTM_TRACE
// debugging aid of variant builds: copies the recurrent actor's timestamp trace to the host (3 actors x 256 x {tag, clock})
extern "C" int cantor_debug_lstm_trace(long long* out, int* counts) {
    CANTOR_CUDA(cudaDeviceSynchronize());
    CANTOR_CUDA(cudaMemcpyFromSymbol(out, cantor::lstmtc::g_lstm_trace, sizeof(long long) * 3 * 256 * 2));
    CANTOR_CUDA(cudaMemcpyFromSymbol(counts, cantor::lstmtc::g_lstm_trace_n, sizeof(int) * 3));
    return CANTOR_OK;
}
#endif
