// K3 -- the fused hedge step in replay mode, and the env reset.
//
// One thread owns one env.  A step reads, once, the env's 16-byte state record, its cash, its action and
// the two 16-byte path records {S, v, C, P} at (t, path) and (t+1, path), and writes, once, the 13-float
// observation, the reward, the done flag and the new state: 137 algorithmic bytes (F32) / 157 (F64) per
// env-step, no re-reads.  The observation tile of a CTA (contiguous in the caller's [n_envs, 13] array) is
// staged in shared memory and leaves the SM as one 1-D TMA bulk store (UBLKCP).  Launches are chained with
// programmatic dependent launch so the next step's CTAs are resident before the previous grid drains.
//
// Reference semantics: HedgingEnv.step / _get_observation / _calculate_greeks / reset,
// src/env/hedging_env_v2.py:175-294 / :109-143 / :79-107 / :145-173 (v1: src/env/hedging_env.py).
#include "hedge_core.cuh"

#ifndef CANTOR_STEP_THREADS
#define CANTOR_STEP_THREADS 128
#endif
#ifndef CANTOR_STEP_MIN_BLOCKS
#define CANTOR_STEP_MIN_BLOCKS 16   // 32 registers, no spills; round-1 sweep: 12 -> 17.97 us, 14/16 -> 17.5 us per 2^20-env launch
#endif
#ifndef CANTOR_STEP_PREFETCH          // prefetch.global.L2 of the path record the NEXT step needs.  Measured (round 1): it HURTS
#define CANTOR_STEP_PREFETCH 0        // (21.25 -> 23.63 us per launch; 18.05 -> 19.84 with evict-first stores), so it is off.
#endif
#ifndef CANTOR_OBS_EVICT_FIRST        // observation tiles leave through L2 with an evict-first policy: they are never re-read by
#define CANTOR_OBS_EVICT_FIRST 1      // the env, and must not displace the state / path rows that the next launch re-reads
#endif                                // (21.25 -> 18.05 us per launch).

namespace cantor {

constexpr int kStepThreads = CANTOR_STEP_THREADS;

struct ResetRule {
    int mode;
    const int* __restrict__ next_path;
    unsigned long long seed;
    long long env_offset;
    long long episode_counter;
};

struct InfoOut {
    double* f64;
    int* i32;
};

// hedging_env_v2.py:150-170 for one env; returns the reset state and fills the reset observation.
template <bool F64>
__device__ __forceinline__ void reset_one(const StepConsts& k, const Book& b, int path, float* __restrict__ o,
                                          int4& core, double& cash, double& pv_prev) {
    const float4 r0 = b.rec[path];                                            // row 0: S0, v0, C0, P0
    const float s0 = (r0.x < 1e-6f) ? 1.0f : r0.x;                            // :157
    core.x = 0;                                                               // no contracts
    core.y = 0;                                                               // current_step
    core.z = path;
    core.w = __float_as_int(s0);
    cash = k.initial_cash;                                                    // :165
    // :167-168 evaluated in float32: (shares * S) + 0 + cash
    pv_prev = (double)__fadd_rn(__fmul_rn(k.shares_f, r0.x), k.initial_cash_f);
    if (F64) make_observation_f64(o, k, r0.x, r0.y, r0.z, r0.w, s0, 0, 0, 0, r0.x, r0.y);
    else make_observation_f32(o, k, r0.x, r0.y, r0.z, r0.w, mufu_rcp(fmaxf(s0, 25.0f)), 0, 0, 0, r0.x, r0.y);
}

__device__ __forceinline__ int next_episode_path(const ResetRule& rr, const Book& b, long long i, int current) {
    if (rr.mode == CANTOR_RESET_FROM_ARRAY) return rr.next_path[i];
    if (rr.mode == CANTOR_RESET_PHILOX) {
        const unsigned long long gid = (unsigned long long)(rr.env_offset + i);
        const uint4 r = philox4x32_10(make_uint4((unsigned)gid, (unsigned)(gid >> 32), (unsigned)rr.episode_counter,
                                                 (unsigned)((unsigned long long)rr.episode_counter >> 32) ^ 0x52455345u),
                                      make_uint2((unsigned)rr.seed, (unsigned)(rr.seed >> 32)));
        return (int)__umulhi(r.x, (unsigned)b.n_paths);                       // uniform in [0, n_paths)
    }
    return current;
}

// Stage a CTA's observation rows in shared memory, then store the contiguous tile.
__device__ __forceinline__ void store_obs_tile(float* __restrict__ obs, const float* tile, long long first_env,
                                               int rows, bool use_tma, bool keep_in_l2) {
    if (use_tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
#if CANTOR_OBS_EVICT_FIRST
            if (!keep_in_l2) tma_store_1d_evict_first(obs + first_env * CANTOR_OBS_DIM, tile, (uint32_t)(rows * CANTOR_OBS_DIM * sizeof(float)));
            else
#endif
            tma_store_1d(obs + first_env * CANTOR_OBS_DIM, tile, (uint32_t)(rows * CANTOR_OBS_DIM * sizeof(float)));
            tma_store_commit();
            tma_store_wait_read();
        }
    } else {
        __syncthreads();
        float* dst = obs + first_env * CANTOR_OBS_DIM;
        for (int j = threadIdx.x; j < rows * CANTOR_OBS_DIM; j += kStepThreads) dst[j] = tile[j];
    }
}

// ---------------------------------------------------------------------------------------------------
// Optional Monitor state / outputs (cantor_env_state.episode_*): per-env running sums of the current episode.
struct Monitor {
    void* acc;                 // [n * 4] float / double: {reward, pps, |pps|, cost}
    void* episode_return;      // [n] float / double, written at episode end
    int* episode_length;       // [n]
    StatsOut stats;            // stats.sums == NULL: no reduction
};

template <bool F64, bool INFO, bool MON>
__global__ void __launch_bounds__(kStepThreads, (MON || INFO || F64) ? 8 : CANTOR_STEP_MIN_BLOCKS)
hedge_step_kernel(const StepConsts k, const Book b, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                  double* __restrict__ pv_arr, long long n_envs, const float2* __restrict__ actions,
                  float* __restrict__ obs, void* __restrict__ reward_arr, unsigned char* __restrict__ done_arr,
                  float* __restrict__ terminal_obs, int auto_reset, const ResetRule rr, const InfoOut info,
                  int obs_tma_ok, const Monitor mon) {
    __shared__ __align__(128) float tile[kStepThreads * CANTOR_OBS_DIM];
    __shared__ double red[MON ? 11 * (kStepThreads / 32) : 1];
    double stat[11];
    bool finished_episode = false;                                            // MON: this thread's env just ended an episode
    if (MON) {
#pragma unroll
        for (int s = 0; s < 11; ++s) stat[s] = 0.0;
    }
    const long long first_env = (long long)blockIdx.x * kStepThreads;
    const long long i = first_env + threadIdx.x;
    const int rows = (int)min((long long)kStepThreads, n_envs - first_env);
    float* o = tile + threadIdx.x * CANTOR_OBS_DIM;                            // stride 13 words: conflict-free

    pdl_wait_prior_grid();          // state / actions may come from the kernel launched just before this one

    if (i < n_envs) {
        // ---- independent loads first -----------------------------------------------------------------
        int4 core = core_arr[i];
        const float2 a = __ldcs(actions + i);
        double cash = 0.0, pv_prev = 0.0;
        float cash_f = 0.f;
        if (F64) {
            cash = reinterpret_cast<const double*>(cash_arr)[i];
            pv_prev = pv_arr[i];
        } else {
            cash_f = reinterpret_cast<const float*>(cash_arr)[i];
        }
        int pos_c = unpack_lo(core.x), pos_p = unpack_hi(core.x);
        int step = core.y;
        const int path = core.z;
        const float s0 = __int_as_float(core.w);
        const bool already_done = step >= k.T;                                // only reachable with auto_reset = 0

        // ---- the two time records of this env's path ----------------------------------------------------
        const int t_prev = already_done ? k.T - 1 : step;
        const float4* rp = b.rec + ((long long)t_prev * b.ld + path);
        const float4 prev = __ldcs(rp);                                       // last use of slab t
        const float4 cur = __ldg(rp + b.ld);                                  // slab t+1 is read again next step
        pdl_launch_dependents();
        const int t_new = t_prev + 1;
        const bool terminated = t_new >= k.T;                                 // :220
        const float S_prev = prev.x, v_prev = prev.y, C_prev = prev.z, P_prev = prev.w;
        // :226-231: row T of the packed book repeats the option marks of row T-1 (stale marks at the end)
        const float S_new = cur.x, v_new = cur.y, C_new = cur.z, P_new = cur.w;

        // ---- (i) action -> trade  :178-200 --------------------------------------------------------------
        const float cf_c = __fmul_rn(a.x, k.max_trade_f);
        const float cf_p = __fmul_rn(a.y, k.max_trade_f);
        const int req_c = requested_trade(cf_c, k.max_trade_f);
        const int req_p = requested_trade(cf_p, k.max_trade_f);
        const int new_c = already_done ? pos_c : max(-k.max_contracts, min(k.max_contracts, pos_c + req_c));
        const int new_p = already_done ? pos_p : max(-k.max_contracts, min(k.max_contracts, pos_p + req_p));
        const int tc = new_c - pos_c, tp = new_p - pos_p;
        const int atc = abs(tc), atp = abs(tp);

        float s0_floor = fmaxf(s0, 25.0f);
        float inv_s0 = 0.f;
        if (F64) {
            // ---- (ii) commission, slippage on the PRE-advance option prices, cash  :203-213 -------------
            const double commission = __dmul_rn((double)(atc + atp), k.cost_per_contract);
            const double slip_c = __dmul_rn(__dmul_rn(__dmul_rn((double)atc, (double)C_prev), k.mult_d), k.bps_frac);
            const double slip_p = __dmul_rn(__dmul_rn(__dmul_rn((double)atp, (double)P_prev), k.mult_d), k.bps_frac);
            const double slippage = __dadd_rn(slip_c, slip_p);
            const double costs = __dadd_rn(commission, slippage);
            const double cash_new = __dsub_rn(cash, costs);
            // ---- (iv) mark to market  :233-238 ----------------------------------------------------------
            const float stock_new = __fmul_rn(k.shares_f, S_new);             // float32 stock leg (:235)
            const double opt_new = __dadd_rn(__dmul_rn(__dmul_rn((double)new_c, (double)C_new), k.mult_d),
                                             __dmul_rn(__dmul_rn((double)new_p, (double)P_new), k.mult_d));
            const double pv = __dadd_rn(__dadd_rn((double)stock_new, opt_new), cash_new);
            const double step_pnl = __dsub_rn(pv, pv_prev);                   // :237
            const double pps = k.shares != 0 ? __ddiv_rn(step_pnl, k.shares_d) : step_pnl;   // :238
            // ---- (v) reward  :243-262 -------------------------------------------------------------------
            double term;
            if (k.loss_mse) term = __ddiv_rn(__dmul_rn(pps, pps), (double)__fadd_rn(__fmul_rn(s0_floor, s0_floor), 1e-9f));
            else term = __ddiv_rn(fabs(pps), (double)__fadd_rn(s0_floor, 1e-9f));
            const double rpc = __dmul_rn(k.neg_w, term);
            const double tcp = __dmul_rn(k.lambda_cost, costs);
            const double theta_pen = __dmul_rn(k.theta_weight, __ddiv_rn((double)(k.T - t_new), 252.0));
            const double reward = already_done ? 0.0 : __dsub_rn(__dsub_rn(rpc, tcp), theta_pen);
            if (INFO && !already_done) {
                double* f = info.f64 + i;
                const long long n = n_envs;
                f[0 * n] = step_pnl;   f[1 * n] = pps;        f[2 * n] = fabs(pps);  f[3 * n] = costs;
                f[4 * n] = commission; f[5 * n] = slippage;   f[6 * n] = rpc;        f[7 * n] = tcp;
                f[8 * n] = theta_pen;  f[9 * n] = reward;     f[10 * n] = pv;        f[11 * n] = cash_new;
            }
            if (!already_done) {
                cash = cash_new;
                pv_prev = pv;
            }
            __stcs(reinterpret_cast<double*>(reward_arr) + i, reward);
            if (MON && !already_done) {
                double2* ap = reinterpret_cast<double2*>(mon.acc) + 2 * i;
                double2 a0 = ap[0], a1 = ap[1];                               // {reward, pps}, {|pps|, cost}
                a0.x += reward; a0.y += pps; a1.x += fabs(pps); a1.y += costs;
                if (terminated) {
                    if (mon.episode_return != nullptr) reinterpret_cast<double*>(mon.episode_return)[i] = a0.x;
                    if (mon.episode_length != nullptr) mon.episode_length[i] = t_new;
                    if (mon.stats.sums != nullptr) {
                        episode_statistics(stat, (float)a0.x, (float)a0.y, (float)a1.x, (float)a1.y, k.inv_T_f, mon.stats);
                        finished_episode = true;
                    }
                    a0 = make_double2(0.0, 0.0);
                    a1 = a0;
                }
                ap[0] = a0;
                ap[1] = a1;
            }
        } else {
            // float32 ledger (hedge_core.cuh): P&L as a sum of small differences, no portfolio value formed
            inv_s0 = mufu_rcp(s0_floor);
            const LedgerF32 L = ledger_f32(k, a.x, a.y, pos_c, pos_p, step, inv_s0, prev, cur, already_done);
            const float commission = L.commission, slippage = L.slippage, costs = L.costs, step_pnl = L.step_pnl;
            const float pps = L.pps, rpc = L.rpc, tcp = L.tcp, theta_pen = L.theta_pen, reward = L.reward;
            const float opt_new = L.opt_new;
            const float cash_new = cash_f - costs;
            if (INFO && !already_done) {
                double* f = info.f64 + i;
                const long long n = n_envs;
                f[0 * n] = step_pnl;   f[1 * n] = pps;        f[2 * n] = fabsf(pps); f[3 * n] = costs;
                f[4 * n] = commission; f[5 * n] = slippage;   f[6 * n] = rpc;        f[7 * n] = tcp;
                f[8 * n] = theta_pen;  f[9 * n] = reward;     f[11 * n] = cash_new;
                f[10 * n] = (double)__fmul_rn(k.shares_f, S_new) + (double)(opt_new * k.mult_f) + (double)cash_new;
            }
            if (!already_done) cash_f = cash_new;
            __stcs(reinterpret_cast<float*>(reward_arr) + i, reward);
            if (MON && !already_done) {
                float4* ap = reinterpret_cast<float4*>(mon.acc) + i;
                float4 a = *ap;                                               // {reward, pps, |pps|, cost}
                a.x += reward; a.y += pps; a.z += fabsf(pps); a.w += costs;
                if (terminated) {
                    if (mon.episode_return != nullptr) reinterpret_cast<float*>(mon.episode_return)[i] = a.x;
                    if (mon.episode_length != nullptr) mon.episode_length[i] = t_new;
                    if (mon.stats.sums != nullptr) {
                        episode_statistics(stat, a.x, a.y, a.z, a.w, k.inv_T_f, mon.stats);
                        finished_episode = true;
                    }
                    a = make_float4(0.f, 0.f, 0.f, 0.f);
                }
                *ap = a;
            }
        }
        if (INFO && !already_done) {
            double* f = info.f64 + i;
            const long long n = n_envs;
            f[12 * n] = a.x; f[13 * n] = a.y; f[14 * n] = cf_c; f[15 * n] = cf_p; f[16 * n] = s0;
            int* q = info.i32 + i;
            q[0 * n] = new_c; q[1 * n] = new_p; q[2 * n] = req_c; q[3 * n] = req_p; q[4 * n] = tc; q[5 * n] = tp;
        }
        pos_c = new_c;
        pos_p = new_p;
        step = t_new;

        // ---- observation of the advanced state  :266 ------------------------------------------------------
        if (F64) make_observation_f64(o, k, S_new, v_new, C_new, P_new, s0, pos_c, pos_p, step, S_prev, v_prev);
        else make_observation_f32(o, k, S_new, v_new, C_new, P_new, inv_s0, pos_c, pos_p, step, S_prev, v_prev);
        core.x = pack_pos(pos_c, pos_p);
        core.y = step;
#if CANTOR_STEP_PREFETCH
        // the next step of this env reads rows `step` (just read: L2-resident) and `step + 1` (new): start that DRAM read now
        if (!terminated) prefetch_l2(rp + 2 * b.ld);
#endif

        if (terminated) {
            if (terminal_obs != nullptr) {
                float* to = terminal_obs + i * CANTOR_OBS_DIM;
#pragma unroll
                for (int j = 0; j < CANTOR_OBS_DIM; ++j) to[j] = o[j];
            }
            if (auto_reset) {                                                 // VecEnv convention: next episode starts now
                const int next = next_episode_path(rr, b, i, path);
                double cash_d = 0.0;
                reset_one<F64>(k, b, next, o, core, cash_d, pv_prev);
                cash = cash_d;
                cash_f = (float)cash_d;
            }
        }

        // ---- stores -------------------------------------------------------------------------------------
        core_arr[i] = core;
        if (F64) {
            reinterpret_cast<double*>(cash_arr)[i] = cash;
            pv_arr[i] = pv_prev;
        } else {
            reinterpret_cast<float*>(cash_arr)[i] = cash_f;
        }
        done_arr[i] = terminated ? 1 : 0;
    } else {
        pdl_launch_dependents();
    }
    store_obs_tile(obs, tile, first_env, rows, (obs_tma_ok & 1) && (rows % 4 == 0), (obs_tma_ok & 2) != 0);
    if (MON) {
        // finished episodes -> statistics vector: warp shuffle -> shared -> one atomic per statistic per CTA, only on the
        // steps where some env of this CTA finished (block-uniform vote, so the barrier inside is safe)
        if (mon.stats.sums != nullptr && __syncthreads_or(finished_episode)) {
            block_accumulate<11, kStepThreads>(stat, mon.stats.sums, red);
        }
        if (mon.stats.sums != nullptr && blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(mon.stats.sums + 11, (double)n_envs);
    }
}

// ---------------------------------------------------------------------------------------------------
template <bool F64>
__global__ void __launch_bounds__(kStepThreads)
env_reset_kernel(const StepConsts k, const Book b, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                 double* __restrict__ pv_arr, long long n_envs, const unsigned char* __restrict__ mask,
                 const int* __restrict__ path_idx, float* __restrict__ obs, void* __restrict__ episode_acc) {
    const long long i = (long long)blockIdx.x * kStepThreads + threadIdx.x;
    if (i >= n_envs) return;
    if (mask != nullptr && mask[i] == 0) return;
    if (episode_acc != nullptr) {
        if (F64) {
            reinterpret_cast<double2*>(episode_acc)[2 * i] = make_double2(0.0, 0.0);
            reinterpret_cast<double2*>(episode_acc)[2 * i + 1] = make_double2(0.0, 0.0);
        } else {
            reinterpret_cast<float4*>(episode_acc)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    float o[CANTOR_OBS_DIM];
    int4 core;
    double cash, pv_prev;
    reset_one<F64>(k, b, path_idx[i], o, core, cash, pv_prev);
    core_arr[i] = core;
    if (F64) {
        reinterpret_cast<double*>(cash_arr)[i] = cash;
        pv_arr[i] = pv_prev;
    } else {
        reinterpret_cast<float*>(cash_arr)[i] = (float)cash;
    }
    if (obs != nullptr) {
#pragma unroll
        for (int j = 0; j < CANTOR_OBS_DIM; ++j) obs[i * CANTOR_OBS_DIM + j] = o[j];
    }
}

// ---------------------------------------------------------------------------------------------------
static int make_consts(const cantor_env_params* p, const cantor_replay_book* book, StepConsts* k, Book* b) {
    CANTOR_REQUIRE(p != nullptr && book != nullptr, "params/book is NULL");
    int rc = make_book(book, b);
    if (rc) return rc;
    return make_step_consts(p, book->episode_length, k);
}

static int check_state(const cantor_env_state* st, int precision) {
    CANTOR_REQUIRE(st != nullptr && st->core != nullptr && st->cash != nullptr, "state array is NULL");
    CANTOR_REQUIRE(aligned16(st->core), "state.core must be 16-byte aligned");
    CANTOR_REQUIRE(precision == CANTOR_F32 || precision == CANTOR_F64, "precision must be 32 or 64");
    CANTOR_REQUIRE(precision == CANTOR_F32 || st->pv_prev != nullptr, "state.pv_prev is required in F64 mode");
    CANTOR_REQUIRE(st->episode_acc == nullptr || aligned16(st->episode_acc), "state.episode_acc must be 16-byte aligned");
    CANTOR_REQUIRE(st->episode_acc != nullptr || (st->episode_return == nullptr && st->episode_length == nullptr && st->stats == nullptr),
                   "episode_return / episode_length / stats need state.episode_acc");
    return CANTOR_OK;
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_env_reset(const cantor_env_params* params, const cantor_replay_book* book,
                                const cantor_env_state* state, int64_t n_envs, int32_t precision,
                                const uint8_t* mask, const int32_t* path_idx, float* obs, void* stream) {
    StepConsts k;
    Book b;
    int rc = make_consts(params, book, &k, &b);
    if (rc) return rc;
    rc = check_state(state, precision);
    if (rc) return rc;
    CANTOR_REQUIRE(n_envs >= 0, "n_envs < 0");
    CANTOR_REQUIRE(path_idx != nullptr, "path_idx is NULL");
    if (n_envs == 0) return CANTOR_OK;
    const unsigned grid = (unsigned)((n_envs + kStepThreads - 1) / kStepThreads);
    cudaStream_t s = (cudaStream_t)stream;
    if (precision == CANTOR_F64)
        env_reset_kernel<true><<<grid, kStepThreads, 0, s>>>(k, b, (int4*)state->core, state->cash, state->pv_prev,
                                                             n_envs, mask, path_idx, obs, state->episode_acc);
    else
        env_reset_kernel<false><<<grid, kStepThreads, 0, s>>>(k, b, (int4*)state->core, state->cash, nullptr, n_envs,
                                                              mask, path_idx, obs, state->episode_acc);
    return check_launch("env_reset_kernel");
}

static int env_step_impl(const cantor_env_params* params, const cantor_replay_book* book,
                         const cantor_env_state* state, int64_t n_envs, int32_t precision, const float* actions,
                         float* obs, void* reward, uint8_t* done, float* terminal_obs, int32_t auto_reset,
                         const cantor_reset_rule* reset_rule, const cantor_info_out* info, void* stream,
                         int32_t n_steps) {
    StepConsts k;
    Book b;
    int rc = make_consts(params, book, &k, &b);
    if (rc) return rc;
    rc = check_state(state, precision);
    if (rc) return rc;
    CANTOR_REQUIRE(n_envs >= 0, "n_envs < 0");
    CANTOR_REQUIRE(actions && obs && reward && done, "actions/obs/reward/done is NULL");
    CANTOR_REQUIRE((reinterpret_cast<uintptr_t>(actions) & 7u) == 0, "actions must be 8-byte aligned");
    ResetRule rr{CANTOR_RESET_SAME_PATH, nullptr, 0ull, 0ll, 0ll};
    if (reset_rule != nullptr) {
        CANTOR_REQUIRE(reset_rule->mode >= CANTOR_RESET_SAME_PATH && reset_rule->mode <= CANTOR_RESET_PHILOX, "reset mode");
        CANTOR_REQUIRE(reset_rule->mode != CANTOR_RESET_FROM_ARRAY || reset_rule->next_path != nullptr,
                       "reset_rule.next_path is NULL");
        rr = ResetRule{reset_rule->mode, reset_rule->next_path, reset_rule->seed, reset_rule->env_offset,
                       reset_rule->episode_counter};
    }
    InfoOut io{nullptr, nullptr};
    if (info != nullptr) {
        CANTOR_REQUIRE(info->f64 != nullptr && info->i32 != nullptr, "info arrays are NULL");
        io = InfoOut{info->f64, info->i32};
    }
    Monitor mon{state->episode_acc, state->episode_return, state->episode_length, {}};
    rc = make_stats_out(state->episode_acc ? state->stats : nullptr, &mon.stats);
    if (rc) return rc;
    const bool mon_on = state->episode_acc != nullptr;
    if (n_envs == 0) return CANTOR_OK;
    const unsigned grid = (unsigned)((n_envs + kStepThreads - 1) / kStepThreads);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t reward_bytes = precision == CANTOR_F64 ? sizeof(double) : sizeof(float);
    int4* core = (int4*)state->core;
    void* cash = state->cash;
    double* pv = state->pv_prev;
    long long n = n_envs;
    for (int32_t t = 0; t < n_steps; ++t) {
        // step t of a rollout writes slab t of the caller's [n_steps, n_envs, ...] buffers
        const float2* a_t = (const float2*)actions + (size_t)t * n_envs;
        float* obs_t = obs + (size_t)t * n_envs * CANTOR_OBS_DIM;
        void* rew_t = (char*)reward + (size_t)t * n_envs * reward_bytes;
        unsigned char* done_t = done + (size_t)t * n_envs;
        int tma_ok = (aligned16(obs_t) ? 1 : 0) | ((reset_rule != nullptr && (reset_rule->flags & CANTOR_STEP_KEEP_OBS_IN_L2)) ? 2 : 0);
        void* args[] = {&k, &b, &core, &cash, &pv, &n, &a_t, &obs_t, &rew_t, &done_t, &terminal_obs, &auto_reset,
                        &rr, &io, &tma_ok, &mon};
        const void* fn;
#define PICK(F64) (info ? (mon_on ? (const void*)hedge_step_kernel<F64, true, true> : (const void*)hedge_step_kernel<F64, true, false>) \
                        : (mon_on ? (const void*)hedge_step_kernel<F64, false, true> : (const void*)hedge_step_kernel<F64, false, false>))
        if (precision == CANTOR_F64) fn = PICK(true);
        else fn = PICK(false);
#undef PICK
        rc = launch_pdl(fn, dim3(grid), dim3(kStepThreads), s, args);
        if (rc) return rc;
        rr.episode_counter += 1;
    }
    return CANTOR_OK;
}

extern "C" int cantor_env_step(const cantor_env_params* params, const cantor_replay_book* book,
                               const cantor_env_state* state, int64_t n_envs, int32_t precision,
                               const float* actions, float* obs, void* reward, uint8_t* done, float* terminal_obs,
                               int32_t auto_reset, const cantor_reset_rule* reset_rule, const cantor_info_out* info,
                               void* stream) {
    return env_step_impl(params, book, state, n_envs, precision, actions, obs, reward, done, terminal_obs, auto_reset,
                         reset_rule, info, stream, 1);
}

extern "C" int cantor_env_step_many(const cantor_env_params* params, const cantor_replay_book* book,
                                    const cantor_env_state* state, int64_t n_envs, int32_t precision, int32_t n_steps,
                                    const float* actions, float* obs, void* reward, uint8_t* done, float* terminal_obs,
                                    const cantor_reset_rule* reset_rule, void* stream) {
    CANTOR_REQUIRE(n_steps >= 0, "n_steps < 0");
    return env_step_impl(params, book, state, n_envs, precision, actions, obs, reward, done, terminal_obs, 1,
                         reset_rule, nullptr, stream, n_steps);
}
