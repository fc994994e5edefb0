// K3 -- the fused hedge step, and the env reset.
//
// One thread owns one env.  Three kernels share ONE step body (step_body below):
//
//   hedge_step_kernel        gym-style, replay mode: one launch = one env-step of every env.  Reads, once, the env's 16-byte
//                            state record, its cash, its action and the two 16-byte path records {S, v, C, P} at (t, path) and
//                            (t+1, path); writes, once, the 13-float observation, the reward, the done flag and the new state:
//                            137 algorithmic bytes (F32) / 157 (F64) per env-step, no re-reads.  Launches are chained with
//                            programmatic dependent launch so the next step's CTAs are resident before the previous grid drains.
//   hedge_step_many_kernel   cantor_env_step_many (the action tape of n_steps steps is known): ONE persistent launch, the env
//                            state stays in registers from step to step, the next step's action and path record are in flight
//                            while the current step computes.  Per env-step only the action (8 B) and the new path record (16 B)
//                            are read and the observation / reward / done (57 B) written: 81 B + 40 B / n_steps.
//   hedge_step_sim_kernel    gym-style, ON-THE-FLY mode (cantor_env_step_sim): no book in memory.  The env carries {S, v}; the
//                            kernel draws the day's Philox normals, advances the path (GBM / Heston), reprices the ATM call / put
//                            before and after the move with the SAME device functions and counters as K1 (sim_core.cuh), so its
//                            outputs are bit-identical to replaying a book that cantor_sim_paths wrote.  121 B per env-step.
//
// The observation tile of a CTA (contiguous in the caller's [n_envs, 13] array) is staged in shared memory and leaves the SM
// as one 1-D TMA bulk store (UBLKCP).
//
// Reference semantics: HedgingEnv.step / _get_observation / _calculate_greeks / reset,
// src/env/hedging_env_v2.py:175-294 / :109-143 / :79-107 / :145-173 (v1: src/env/hedging_env.py); on-the-fly path step
// src/sim/rbergomi_sim.py:454-464.
#include <stdlib.h>

#include "hedge_core.cuh"
#include "sim_core.cuh"

#ifndef CANTOR_STEP_THREADS
#define CANTOR_STEP_THREADS 128
#endif
#ifndef CANTOR_STEP_MIN_BLOCKS
#define CANTOR_STEP_MIN_BLOCKS 16   // 32 registers, no spills; round-1 sweep: 12 -> 17.97 us, 14/16 -> 17.5 us per 2^20-env launch
#endif
#ifndef CANTOR_STEP_PREFETCH          // prefetch.global.L2 of the path record the NEXT step needs.  Measured: it HURTS
#define CANTOR_STEP_PREFETCH 0        // (2^20 envs: 21.25 -> 23.63 us per launch; 2^23 envs: 174 -> 191 us), so it is off.
#endif
#ifndef CANTOR_OBS_EVICT_FIRST        // observation tiles leave through L2 with an evict-first policy: they are never re-read by
#define CANTOR_OBS_EVICT_FIRST 1      // the env, and must not displace the state / path rows that the next launch re-reads
#endif                                // (2^20 envs: 21.25 -> 18.05 us per launch; no effect once nothing fits L2: 174 vs 171 us at 2^23).
#ifndef CANTOR_STEP_VN_BLOCKS         // resident CTAs per SM of the fp32 step kernel that also produces VecNormalize's moments (40 registers)
#define CANTOR_STEP_VN_BLOCKS 12
#endif
#ifndef CANTOR_STEP_MON_BLOCKS        // resident CTAs per SM of the fp32 Monitor variants (statistics formed lazily: no float64 accumulators live
#define CANTOR_STEP_MON_BLOCKS 12     // across the step body); 2^20 envs, step + Monitor + statistics: 8 / 10 / 12 CTAs -> 24.0 / 23.4 / 22.5 us
#endif
#ifndef CANTOR_MANY_MIN_BLOCKS
#define CANTOR_MANY_MIN_BLOCKS 10     // persistent multi-step kernel: 48 registers, two 6.5 KB observation tiles per CTA
#endif                                // (2^20 envs x 252 steps: 8 CTAs / SM 3.94 ms, 10: 3.78 ms, 12: 3.76 ms with spills, 16: 4.85 ms)
#ifndef CANTOR_MANY_TMA               // 1: the persistent kernel stores its observation tiles with TMA bulk copies like the per-step
#define CANTOR_MANY_TMA 1             // kernel; 0: plain 16-byte streaming stores out of the staged tile.  The fence.proxy.async before a
#endif                                // bulk copy compiles to DEPBAR + MEMBAR.ALL.CTA and so waits for the thread's in-flight prefetch
                                      // loads once per step -- yet measured (2^20 envs x 252 steps) TMA 3.755 ms, plain stores 3.862 ms.

#ifndef CANTOR_MANY_EVICT_FIRST       // L2 policy of the persistent kernel's observation stores.  Nothing it touches is ever re-read, so there
#define CANTOR_MANY_EVICT_FIRST 0     // is nothing to protect in L2: the plain policy measured 3.689 ms against 3.771 ms with evict-first
#endif                                // (2^20 envs x 252 steps) -- evict-first lines leave L2 in smaller, less DRAM-friendly batches.
#ifndef CANTOR_MANY_PREFETCH          // how many steps ahead the persistent kernel requests a step's action and path record (1 or 2).
#define CANTOR_MANY_PREFETCH 1        // Measured (2^20 envs x 252 steps): 1 -> 3.695 ms, 2 -> 3.742 ms (same 48 registers): the first-use stalls
#endif                                // are back-pressure of the memory system, not latency that deeper prefetch could hide.
#ifndef CANTOR_MANY_WARP_STORES
#define CANTOR_MANY_WARP_STORES 0     // 1: every warp stores its own 32 rows of the observation tile (one 1 664-byte bulk store behind a
#endif                                // warp barrier): no CTA barrier in a step.  Measured: 3.743 ms against 3.658 (2^20 x 252) -- the four
                                      // times smaller bulk stores cost more than the barrier; 12 CTAs/SM on top 3.87, prefetch 2 on top 3.81
#ifndef CANTOR_MANY_CLUSTER
#define CANTOR_MANY_CLUSTER 1         // > 1: the persistent kernel's CTAs run in clusters of this size that stay within one step of each other
#endif                                // (split cluster barrier around the observation store), so a cluster's pieces of a slab leave together.
                                      // Measured (2^20 x 252): 2 / 4 / 8 -> 4.60 / 4.91 / 4.92 ms against 3.66: lockstep is what this kernel
                                      // does NOT want -- drifting CTAs are what keeps five streams flowing at once
#ifndef CANTOR_STEP_ALTERNATE
#define CANTOR_STEP_ALTERNATE 1       // the per-step replay kernel walks the env tiles forward on even global steps, backward on odd ones
#endif
#ifndef CANTOR_MANY_MON_BLOCKS
#define CANTOR_MANY_MON_BLOCKS 8      // resident CTAs per SM of the persistent kernel's fp32 Monitor variant
#endif
#ifndef CANTOR_MANY_THREADS
#define CANTOR_MANY_THREADS 128       // envs per CTA of the persistent kernel
#endif

namespace cantor {

constexpr int kStepThreads = CANTOR_STEP_THREADS;
constexpr int kManyThreads = CANTOR_MANY_THREADS;

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
    float* f32;                // F32 mode: the 17 float keys as float32 (68 B instead of 136 B per env-step) when not NULL
};

// Optional Monitor state / outputs (cantor_env_state.episode_*): per-env running sums of the current episode.
struct Monitor {
    void* acc;                 // [n * 4] float / double: {reward, pps, |pps|, cost}
    void* episode_return;      // [n] float / double, written at episode end
    int* episode_length;       // [n]
    StatsOut stats;            // stats.sums == NULL: no reduction
};

// Optional VecNormalize fusion (cantor_env_state.vecnorm): the step kernel also produces what cantor_vecnorm_step's moments
// kernel would have to re-read the batch for -- per-CTA partial sums of the observation columns (sum, sum of squares: 26 values)
// and of the updated discounted returns (2 values), written with plain stores to partial[CTA][statistic] and folded later
// in a fixed order (cantor_vecnorm_step_fused), so the running statistics stay bitwise reproducible.
struct VecNormFuse {
    double* partial;           // [n_cta, 28] (NULL = off)
    double* returns;           // [n] discounted returns, updated in place: ret <- ret * gamma + reward
    double gamma;
    int norm_obs, norm_reward;
};
constexpr int kVnFuseSums = 2 * CANTOR_OBS_DIM + 2;

// Called by every thread of the CTA after the observation tile is complete and a CTA barrier has been passed.
// `reward` = this thread's raw reward of the step, `ret_prev` = its discounted return before the step (both ignored for
// threads without an env).
// partial layout: [CTA][28] -- one contiguous, fully written 224-byte record per CTA (no partial-sector writes).
template <int THREADS>
__device__ __forceinline__ void vecnorm_partials(const VecNormFuse& vn, const float* tile, int rows, bool live, long long i,
                                                 double reward, double ret_prev, double* scratch /* [2 * 8 * 13 + 2 * THREADS / 32] */,
                                                 unsigned tile_idx) {
    constexpr int C = CANTOR_OBS_DIM, PARTS = 8;
    double* part = scratch;                            // [2][PARTS][C]
    double* wsum = scratch + 2 * PARTS * C;            // [2][THREADS / 32]
    if (vn.norm_obs && threadIdx.x < PARTS * C) {      // (part v, column c): rows v, v + 8, ... in a fixed order, two chains each
        const int v = threadIdx.x / C, c = threadIdx.x % C;
        double a0 = 0.0, q0 = 0.0, a1 = 0.0, q1 = 0.0;
        int r = v;
        for (; r + PARTS < rows; r += 2 * PARTS) {
            const double x0 = (double)tile[r * C + c], x1 = (double)tile[(r + PARTS) * C + c];
            a0 += x0; q0 = fma(x0, x0, q0);
            a1 += x1; q1 = fma(x1, x1, q1);
        }
        if (r < rows) {
            const double x0 = (double)tile[r * C + c];
            a0 += x0; q0 = fma(x0, x0, q0);
        }
        part[v * C + c] = a0 + a1;
        part[(PARTS + v) * C + c] = q0 + q1;
    }
    double rs = 0.0, rq = 0.0;
    if (vn.norm_reward && live) {
        const double ret = fma(ret_prev, vn.gamma, reward);   // ret_prev was requested with the kernel's first loads: a load here,
        vn.returns[i] = ret;                                  // at the end of the CTA's life, would add a DRAM round trip to every CTA
        rs = ret;
        rq = ret * ret;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        rs += __shfl_down_sync(0xffffffffu, rs, off);
        rq += __shfl_down_sync(0xffffffffu, rq, off);
    }
    if ((threadIdx.x & 31) == 0) {
        wsum[threadIdx.x >> 5] = rs;
        wsum[THREADS / 32 + (threadIdx.x >> 5)] = rq;
    }
    __syncthreads();
    double* out = vn.partial + (long long)tile_idx * kVnFuseSums;      // by TILE, not by CTA: the fold's summation order does not depend on the walk
    if (threadIdx.x < 2 * C) {                         // column sums: the 8 parts in order
        const int kind = threadIdx.x / C, c = threadIdx.x % C;
        double a = 0.0;
        if (vn.norm_obs) {
#pragma unroll
            for (int v = 0; v < PARTS; ++v) a += part[(kind * PARTS + v) * C + c];
        }
        out[threadIdx.x] = a;
    } else if (threadIdx.x < 2 * C + 2) {              // returns: the warps in order
        const int kind = threadIdx.x - 2 * C;
        double a = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) a += wsum[kind * (THREADS / 32) + w];
        out[threadIdx.x] = a;
    }
}

// The env state of one thread, in registers.  `path` is the replay path index (replay mode) or the episode number (on the fly).
struct EnvRegs {
    int pos_c, pos_p, step, path;
    float s0, cash_f;          // cash_f: F32 ledger
    double cash, pv_prev;      // F64 ledger
};

template <bool F64>
__device__ __forceinline__ EnvRegs load_env(const int4* __restrict__ core_arr, const void* __restrict__ cash_arr,
                                            const double* __restrict__ pv_arr, long long i) {
    EnvRegs e;
    const int4 core = core_arr[i];
    e.cash = 0.0; e.pv_prev = 0.0; e.cash_f = 0.f;
    if (F64) {
        e.cash = reinterpret_cast<const double*>(cash_arr)[i];
        e.pv_prev = pv_arr[i];
    } else {
        e.cash_f = reinterpret_cast<const float*>(cash_arr)[i];
    }
    e.pos_c = unpack_lo(core.x);
    e.pos_p = unpack_hi(core.x);
    e.step = core.y;
    e.path = core.z;
    e.s0 = __int_as_float(core.w);
    return e;
}

template <bool F64>
__device__ __forceinline__ void store_env(const EnvRegs& e, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                                          double* __restrict__ pv_arr, long long i) {
    core_arr[i] = make_int4(pack_pos(e.pos_c, e.pos_p), e.step, e.path, __float_as_int(e.s0));
    if (F64) {
        reinterpret_cast<double*>(cash_arr)[i] = e.cash;
        pv_arr[i] = e.pv_prev;
    } else {
        reinterpret_cast<float*>(cash_arr)[i] = e.cash_f;
    }
}

// hedging_env_v2.py:150-170 for one env: the reset state from the episode's first record r0 = {S0, v0, C0, P0}, and the
// reset observation.  SHARE: the observation's greeks are handed in (they came with the on-the-fly price of r0).
template <bool F64, bool SHARE>
__device__ __forceinline__ void reset_regs(const StepConsts& k, const float4& r0, int path, float* __restrict__ o, EnvRegs& e,
                                           const Greeks& g0) {
    const float s0 = (r0.x < 1e-6f) ? 1.0f : r0.x;                            // :157
    e.pos_c = 0;                                                              // no contracts
    e.pos_p = 0;
    e.step = 0;                                                               // current_step
    e.path = path;
    e.s0 = s0;
    e.cash = k.initial_cash;                                                  // :165
    e.cash_f = (float)k.initial_cash;
    // :167-168 evaluated in float32: (shares * S) + 0 + cash
    e.pv_prev = (double)__fadd_rn(__fmul_rn(k.shares_f, r0.x), k.initial_cash_f);
    if (F64) make_observation_f64(o, k, r0.x, r0.y, r0.z, r0.w, s0, 0, 0, 0, r0.x, r0.y);
    else if (SHARE) make_observation_f32(o, k, r0.x, r0.y, r0.z, r0.w, mufu_rcp(fmaxf(s0, 25.0f)), 0, 0, 0, r0.x, r0.y, g0);
    else make_observation_f32(o, k, r0.x, r0.y, r0.z, r0.w, mufu_rcp(fmaxf(s0, 25.0f)), 0, 0, 0, r0.x, r0.y);
}

__device__ __forceinline__ int next_episode_path(const ResetRule& rr, const Book& b, long long i, int current, long long counter) {
    if (rr.mode == CANTOR_RESET_FROM_ARRAY) return rr.next_path[i];
    if (rr.mode == CANTOR_RESET_PHILOX) {
        const unsigned long long gid = (unsigned long long)(rr.env_offset + i);
        const uint4 r = philox4x32_10(make_uint4((unsigned)gid, (unsigned)(gid >> 32), (unsigned)counter,
                                                 (unsigned)((unsigned long long)counter >> 32) ^ 0x52455345u),
                                      make_uint2((unsigned)rr.seed, (unsigned)(rr.seed >> 32)));
        return (int)__umulhi(r.x, (unsigned)b.n_paths);                       // uniform in [0, n_paths)
    }
    return current;
}

// ---------------------------------------------------------------------------------------------------
// One env-step of one env, everything in registers: action -> trade -> commission / slippage -> cash -> advance ->
// mark-to-market -> P&L reward -> done -> observation of the advanced state (written to `o`, a row of the CTA's shared-memory
// tile).  `prev` / `cur` are the path records at the env's step t and t+1 (row T of a packed book, and the on-the-fly
// kernel, repeat the option marks of T-1: the reference's stale terminal mark, :226-231).  Updates `e`; returns terminated.
// The caller loads the records, handles terminal_obs / auto-reset and stores the state.
template <bool F64, bool INFO, bool MON, bool SHARE>
__device__ __forceinline__ bool step_body(const StepConsts& k, EnvRegs& e, const float2 a, const float4& prev, const float4& cur,
                                          const Greeks& g_cur, float* __restrict__ o, long long i, long long n_envs,
                                          void* __restrict__ reward_slot, const InfoOut& info, const Monitor& mon,
                                          float4& fin_acc, bool& finished_episode, double* reward_out = nullptr,
                                          float4* acc_f32 = nullptr, double2* acc_f64 = nullptr) {
    const int pos_c = e.pos_c, pos_p = e.pos_p;
    const bool already_done = e.step >= k.T;                                  // only reachable with auto_reset = 0
    const int t_prev = already_done ? k.T - 1 : e.step;
    const int t_new = t_prev + 1;
    const bool terminated = t_new >= k.T;                                     // :220
    const float S_prev = prev.x, v_prev = prev.y, C_prev = prev.z, P_prev = prev.w;
    const float S_new = cur.x, v_new = cur.y, C_new = cur.z, P_new = cur.w;
    const float s0 = e.s0;

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
        const double cash_new = __dsub_rn(e.cash, costs);
        // ---- (iv) mark to market  :233-238 ----------------------------------------------------------
        const float stock_new = __fmul_rn(k.shares_f, S_new);                 // float32 stock leg (:235)
        const double opt_new = __dadd_rn(__dmul_rn(__dmul_rn((double)new_c, (double)C_new), k.mult_d),
                                         __dmul_rn(__dmul_rn((double)new_p, (double)P_new), k.mult_d));
        const double pv = __dadd_rn(__dadd_rn((double)stock_new, opt_new), cash_new);
        const double step_pnl = __dsub_rn(pv, e.pv_prev);                     // :237
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
            e.cash = cash_new;
            e.pv_prev = pv;
        }
        __stcs(reinterpret_cast<double*>(reward_slot), reward);
        if (reward_out != nullptr) *reward_out = reward;
        if (MON && !already_done) {
            // the running sums of the episode: in the caller's registers (persistent kernel) or in the state arrays (one-step kernels)
            double2* ap = acc_f64 != nullptr ? acc_f64 : reinterpret_cast<double2*>(mon.acc) + 2 * i;
            double2 a0 = ap[0], a1 = ap[1];                                   // {reward, pps}, {|pps|, cost}
            a0.x += reward; a0.y += pps; a1.x += fabs(pps); a1.y += costs;
            if (terminated) {
                if (mon.episode_return != nullptr) reinterpret_cast<double*>(mon.episode_return)[i] = a0.x;
                if (mon.episode_length != nullptr) mon.episode_length[i] = t_new;
                if (mon.stats.sums != nullptr) {                               // the caller turns these four sums into the statistics
                    fin_acc = make_float4((float)a0.x, (float)a0.y, (float)a1.x, (float)a1.y);
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
        const LedgerF32 L = ledger_f32(k, a.x, a.y, pos_c, pos_p, e.step, inv_s0, prev, cur, already_done);
        const float commission = L.commission, slippage = L.slippage, costs = L.costs, step_pnl = L.step_pnl;
        const float pps = L.pps, rpc = L.rpc, tcp = L.tcp, theta_pen = L.theta_pen, reward = L.reward;
        const float opt_new = L.opt_new;
        const float cash_new = e.cash_f - costs;
        if (INFO && !already_done) {
            const long long n = n_envs;
            const float pv = (float)((double)__fmul_rn(k.shares_f, S_new) + (double)(opt_new * k.mult_f) + (double)cash_new);
            if (info.f32 != nullptr) {                                        // float32 info arrays: 68 B instead of 136 B
                float* f = info.f32 + i;
                f[0 * n] = step_pnl;   f[1 * n] = pps;        f[2 * n] = fabsf(pps); f[3 * n] = costs;
                f[4 * n] = commission; f[5 * n] = slippage;   f[6 * n] = rpc;        f[7 * n] = tcp;
                f[8 * n] = theta_pen;  f[9 * n] = reward;     f[10 * n] = pv;        f[11 * n] = cash_new;
            } else {
                double* f = info.f64 + i;
                f[0 * n] = step_pnl;   f[1 * n] = pps;        f[2 * n] = fabsf(pps); f[3 * n] = costs;
                f[4 * n] = commission; f[5 * n] = slippage;   f[6 * n] = rpc;        f[7 * n] = tcp;
                f[8 * n] = theta_pen;  f[9 * n] = reward;     f[11 * n] = cash_new;
                f[10 * n] = (double)__fmul_rn(k.shares_f, S_new) + (double)(opt_new * k.mult_f) + (double)cash_new;
            }
        }
        if (!already_done) e.cash_f = cash_new;
        __stcs(reinterpret_cast<float*>(reward_slot), reward);
        if (reward_out != nullptr) *reward_out = (double)reward;
        if (MON && !already_done) {
            float4* ap = acc_f32 != nullptr ? acc_f32 : reinterpret_cast<float4*>(mon.acc) + i;
            float4 m = *ap;                                                   // {reward, pps, |pps|, cost}
            m.x += reward; m.y += pps; m.z += fabsf(pps); m.w += costs;
            if (terminated) {
                if (mon.episode_return != nullptr) reinterpret_cast<float*>(mon.episode_return)[i] = m.x;
                if (mon.episode_length != nullptr) mon.episode_length[i] = t_new;
                if (mon.stats.sums != nullptr) {
                    fin_acc = m;
                    finished_episode = true;
                }
                m = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            *ap = m;
        }
    }
    if (INFO && !already_done) {
        const long long n = n_envs;
        if (!F64 && info.f32 != nullptr) {
            float* f = info.f32 + i;
            f[12 * n] = a.x; f[13 * n] = a.y; f[14 * n] = cf_c; f[15 * n] = cf_p; f[16 * n] = s0;
        } else {
            double* f = info.f64 + i;
            f[12 * n] = a.x; f[13 * n] = a.y; f[14 * n] = cf_c; f[15 * n] = cf_p; f[16 * n] = s0;
        }
        int* q = info.i32 + i;
        q[0 * n] = new_c; q[1 * n] = new_p; q[2 * n] = req_c; q[3 * n] = req_p; q[4 * n] = tc; q[5 * n] = tp;
    }
    e.pos_c = new_c;
    e.pos_p = new_p;
    e.step = t_new;

    // ---- observation of the advanced state  :266 ------------------------------------------------------
    if (F64) make_observation_f64(o, k, S_new, v_new, C_new, P_new, s0, new_c, new_p, t_new, S_prev, v_prev);
    else if (SHARE) make_observation_f32(o, k, S_new, v_new, C_new, P_new, inv_s0, new_c, new_p, t_new, S_prev, v_prev, g_cur);
    else make_observation_f32(o, k, S_new, v_new, C_new, P_new, inv_s0, new_c, new_p, t_new, S_prev, v_prev);
    return terminated;
}

// Store the CTA's staged observation rows (contiguous tile) -- caller has written the rows; this is the barrier + store.
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

__device__ __forceinline__ void write_terminal_obs(float* __restrict__ terminal_obs, long long i, const float* o) {
    float* to = terminal_obs + i * CANTOR_OBS_DIM;
#pragma unroll
    for (int j = 0; j < CANTOR_OBS_DIM; ++j) to[j] = o[j];
}

// MON epilogue: finished episodes -> statistics vector (warp shuffle -> shared -> one atomic per statistic per CTA, only
// when some env of this CTA finished: block-uniform vote, so the barrier inside is safe), the env-step counter, and the
// fused all-reduce of the statistics (last CTA of the launch; no-op unless a ticket is attached).
template <int THREADS>
__device__ __forceinline__ void monitor_epilogue(const Monitor& mon, double (&stat)[11], bool finished_episode, double* red,
                                                 double env_steps) {
    if (mon.stats.sums == nullptr) return;
    if (__syncthreads_or(finished_episode)) block_accumulate<11, THREADS>(stat, mon.stats.sums, red);
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(mon.stats.sums + 11, env_steps);
    push_statistics_to_all_ranks<THREADS>(mon.stats);
}
// One-step kernels: an env finishes at most one episode per launch, so the eleven float64 statistics are formed HERE, inside the
// (block-uniform, rare) branch, from the four sums the step body handed back -- they are not live across the step body, which keeps
// the Monitor variants' registers down.
template <int THREADS>
__device__ __forceinline__ void monitor_epilogue_one_step(const Monitor& mon, const StepConsts& k, const float4& fin_acc,
                                                          bool finished_episode, double* red, double env_steps) {
    if (mon.stats.sums == nullptr) return;
    if (__syncthreads_or(finished_episode)) {
        double stat[11];
#pragma unroll
        for (int s = 0; s < 11; ++s) stat[s] = 0.0;
        if (finished_episode) episode_statistics(stat, fin_acc.x, fin_acc.y, fin_acc.z, fin_acc.w, k.inv_T_f, mon.stats);
        block_accumulate<11, THREADS>(stat, mon.stats.sums, red);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(mon.stats.sums + 11, env_steps);
    push_statistics_to_all_ranks<THREADS>(mon.stats);
}

// ---------------------------------------------------------------------------------------------------
// Gym-style step, replay mode: one launch = one env-step.
template <bool F64, bool INFO, bool MON, bool VN>
__global__ void __launch_bounds__(kStepThreads, (INFO || F64) ? 8 : (MON ? CANTOR_STEP_MON_BLOCKS : (VN ? CANTOR_STEP_VN_BLOCKS : CANTOR_STEP_MIN_BLOCKS)))
hedge_step_kernel(const StepConsts k, const Book b, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                  double* __restrict__ pv_arr, long long n_envs, const float2* __restrict__ actions,
                  float* __restrict__ obs, void* __restrict__ reward_arr, unsigned char* __restrict__ done_arr,
                  float* __restrict__ terminal_obs, int auto_reset, const ResetRule rr, const InfoOut info,
                  int obs_tma_ok, const Monitor mon, const VecNormFuse vn) {
    __shared__ __align__(128) float tile[kStepThreads * CANTOR_OBS_DIM];
    __shared__ double red[MON ? 11 * (kStepThreads / 32) : 1];
    __shared__ double vn_scratch[VN ? 2 * 8 * CANTOR_OBS_DIM + 2 * (kStepThreads / 32) : 1];
    double my_reward = 0.0, ret_prev = 0.0;
    float4 fin_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool finished_episode = false;                                            // MON: this thread's env just ended an episode
    // Bit 2 of obs_tma_ok: walk the env tiles LAST FIRST.  The host alternates the direction from launch to launch, so a launch starts
    // with the state and path records the previous launch touched last -- the part of them that is still in L2 when the population
    // is larger than the cache (at 2^20 envs everything fits either way; at 2^23 the first ~1.5 M envs of every launch hit).
    const unsigned tile_idx = (obs_tma_ok & 4) ? gridDim.x - 1 - blockIdx.x : blockIdx.x;
    const long long first_env = (long long)tile_idx * kStepThreads;
    const long long i = first_env + threadIdx.x;
    const int rows = (int)min((long long)kStepThreads, n_envs - first_env);
    float* o = tile + threadIdx.x * CANTOR_OBS_DIM;                            // stride 13 words: conflict-free

    pdl_wait_prior_grid();          // state / actions may come from the kernel launched just before this one

    if (i < n_envs) {
        // ---- independent loads first -----------------------------------------------------------------
        EnvRegs e = load_env<F64>(core_arr, cash_arr, pv_arr, i);
        const float2 a = __ldcs(actions + i);
        if (VN && vn.norm_reward) ret_prev = vn.returns[i];
        // ---- the two time records of this env's path ----------------------------------------------------
        const int t_prev = (e.step >= k.T) ? k.T - 1 : e.step;
        const float4* rp = b.rec + ((long long)t_prev * b.ld + e.path);
        const float4 prev = __ldcs(rp);                                       // last use of slab t
        const float4 cur = __ldg(rp + b.ld);                                  // slab t+1 is read again next step
        pdl_launch_dependents();
        const int path = e.path;
        const Greeks none{0.f, 0.f, 0.f};
        const size_t rb = F64 ? sizeof(double) : sizeof(float);
        const bool terminated = step_body<F64, INFO, MON, false>(k, e, a, prev, cur, none, o, i, n_envs,
                                                                 (char*)reward_arr + i * rb, info, mon, fin_acc, finished_episode,
                                                                 VN ? &my_reward : nullptr);
#if CANTOR_STEP_PREFETCH
        // the next step of this env reads rows `step` (just read: L2-resident) and `step + 1` (new): start that DRAM read now
        if (!terminated) prefetch_l2(rp + 2 * b.ld);
#endif
        if (terminated) {
            if (terminal_obs != nullptr) write_terminal_obs(terminal_obs, i, o);
            if (auto_reset) {                                                 // VecEnv convention: next episode starts now
                const int next = next_episode_path(rr, b, i, path, rr.episode_counter);
                reset_regs<F64, false>(k, b.rec[next], next, o, e, none);
            }
        }
        // ---- stores -------------------------------------------------------------------------------------
        store_env<F64>(e, core_arr, cash_arr, pv_arr, i);
        done_arr[i] = terminated ? 1 : 0;
    } else {
        pdl_launch_dependents();
    }
    store_obs_tile(obs, tile, first_env, rows, (obs_tma_ok & 1) && (rows % 4 == 0), (obs_tma_ok & 2) != 0);
    if (VN) vecnorm_partials<kStepThreads>(vn, tile, rows, i < n_envs, i, my_reward, ret_prev, vn_scratch, tile_idx);
    if (MON) monitor_epilogue_one_step<kStepThreads>(mon, k, fin_acc, finished_episode, red, (double)n_envs);
}

// ---------------------------------------------------------------------------------------------------
// cantor_env_step_many: n_steps consecutive env-steps of every env in ONE launch.  The state lives in registers; per step
// the thread reads its action and ONE new path record (both requested a step ahead) and the CTA writes one observation tile
// (double-buffered in shared memory: the TMA store of step s drains while step s + 1 computes), 128 rewards and 128 dones.
template <bool F64, bool MON>
__global__ void __launch_bounds__(kManyThreads, F64 ? (6 * 128 / kManyThreads) : ((MON ? CANTOR_MANY_MON_BLOCKS : CANTOR_MANY_MIN_BLOCKS) * 128 / kManyThreads))
hedge_step_many_kernel(const StepConsts k, const Book b, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                       double* __restrict__ pv_arr, long long n_envs, int n_steps, const float2* __restrict__ actions,
                       float* __restrict__ obs, void* __restrict__ reward_arr, unsigned char* __restrict__ done_arr,
                       float* __restrict__ terminal_obs, const ResetRule rr, int obs_tma_ok, const Monitor mon) {
    __shared__ __align__(128) float tiles[2][kManyThreads * CANTOR_OBS_DIM];
    __shared__ double red[MON ? 11 * (kManyThreads / 32) : 1];
    // Monitor: the episode's running sums stay in registers for the whole launch (one 16 / 32-byte record each way per LAUNCH, not per step),
    // and the eleven statistics of finished episodes are formed inside the rare, block-uniform branch below -- nothing float64 is
    // live across the step body (the first form kept an 11-double accumulator per thread: 6 CTAs per SM, 5.02 ms per 2^20 x 252 sweep)
    float4 macc_f = make_float4(0.f, 0.f, 0.f, 0.f);
    double2 macc_d[2] = {make_double2(0.0, 0.0), make_double2(0.0, 0.0)};
    pdl_wait_prior_grid();          // the state may come from a per-step launch, which releases its dependents before it finishes
    const long long first_env = (long long)blockIdx.x * kManyThreads;
    const long long i = first_env + threadIdx.x;
    const bool live = i < n_envs;
    const int rows = (int)max(0ll, min((long long)kManyThreads, n_envs - first_env));   // 0: a CTA that only fills up the last cluster
    const bool use_tma = rows > 0 && (obs_tma_ok & 1) && (rows % 4 == 0) && ((n_envs & 3) == 0);   // every step's tile 16-byte aligned
    const bool keep_in_l2 = (obs_tma_ok & 2) != 0;
    const size_t rb = F64 ? sizeof(double) : sizeof(float);
    const InfoOut no_info{nullptr, nullptr, nullptr};
    const Greeks none{0.f, 0.f, 0.f};

    EnvRegs e{};
    float4 prev = make_float4(0.f, 0.f, 0.f, 0.f), cur = prev, cur_n = prev, cur_n2 = prev;
    float2 a = make_float2(0.f, 0.f), a_n = a, a_n2 = a;
    const float4* rp = b.rec;                                                  // record (t + 1, path) of the step being computed
    if (live) {
        e = load_env<F64>(core_arr, cash_arr, pv_arr, i);
        if (MON) {
            if (F64) { macc_d[0] = reinterpret_cast<const double2*>(mon.acc)[2 * i]; macc_d[1] = reinterpret_cast<const double2*>(mon.acc)[2 * i + 1]; }
            else macc_f = reinterpret_cast<const float4*>(mon.acc)[i];
        }
        const int t_prev = (e.step >= k.T) ? k.T - 1 : e.step;                // cannot be >= T here (auto-reset), kept for safety
        rp = b.rec + ((long long)(t_prev + 1) * b.ld + e.path);
        prev = __ldcs(rp - b.ld);
        cur = __ldcs(rp);
        a = __ldcs(actions + i);
#if CANTOR_MANY_PREFETCH >= 2
        if (n_steps > 1) {                                                     // step 1's inputs (its record only if step 0 does not end the episode)
            a_n = __ldcs(actions + n_envs + i);
            if (t_prev + 2 <= k.T) cur_n = __ldcs(rp + b.ld);
        }
#endif
    }
#if CANTOR_MANY_CLUSTER > 1
    cluster_arrive();
#endif
    for (int s = 0; s < n_steps; ++s) {
        float* tile = tiles[s & 1];
        float* o = tile + threadIdx.x * CANTOR_OBS_DIM;
        const long long at = (long long)s * n_envs + i;
        bool terminated = false;
        float4 fin_acc = make_float4(0.f, 0.f, 0.f, 0.f);
        bool fin_now = false;
        if (live) {
            // request later steps' inputs before this step's arithmetic: CANTOR_MANY_PREFETCH steps ahead (a record only while the
            // episode it belongs to is the current one: after an episode end the new path is known only at the reset)
#if CANTOR_MANY_PREFETCH >= 2
            if (s + 2 < n_steps) {
                a_n2 = __ldcs(actions + at + 2 * n_envs);
                if (e.step + 3 <= k.T) cur_n2 = __ldcs(rp + 2 * b.ld);
            }
#else
            const bool ends = e.step + 1 >= k.T;
            if (s + 1 < n_steps) {
                a_n = __ldcs(actions + at + n_envs);
                if (!ends) cur_n = __ldcs(rp + b.ld);
            }
#endif
            const int path = e.path;
            terminated = step_body<F64, false, MON, false>(k, e, a, prev, cur, none, o, i, n_envs, (char*)reward_arr + at * rb,
                                                           no_info, mon, fin_acc, fin_now, nullptr, &macc_f, macc_d);
            if (terminated) {
                if (terminal_obs != nullptr) write_terminal_obs(terminal_obs, i, o);
                const int next = next_episode_path(rr, b, i, path, rr.episode_counter + s);
                const float4 r0 = b.rec[next];
                reset_regs<F64, false>(k, r0, next, o, e, none);
                rp = b.rec + (b.ld + next);
                prev = r0;
                if (s + 1 < n_steps) cur_n = __ldcs(rp);
#if CANTOR_MANY_PREFETCH >= 2
                if (s + 2 < n_steps && k.T >= 2) cur_n2 = __ldcs(rp + b.ld);
#endif
            } else {
                prev = cur;
                rp += b.ld;
            }
            cur = cur_n;
            a = a_n;
#if CANTOR_MANY_PREFETCH >= 2
            cur_n = cur_n2;
            a_n = a_n2;
#endif
            done_arr[at] = terminated ? 1 : 0;
        }
        if (MON && mon.stats.sums != nullptr) {                               // episodes that ended in this step -> the statistics vector
            if (__syncthreads_or(fin_now)) {
                double stat[11];
#pragma unroll
                for (int q = 0; q < 11; ++q) stat[q] = 0.0;
                if (fin_now) episode_statistics(stat, fin_acc.x, fin_acc.y, fin_acc.z, fin_acc.w, k.inv_T_f, mon.stats);
                block_accumulate<11, kManyThreads>(stat, mon.stats.sums, red);
            }
        }
        // ---- the CTA's observation tile of step s ---------------------------------------------------------------
        float* dst = obs + ((long long)s * n_envs + first_env) * CANTOR_OBS_DIM;
#if CANTOR_MANY_TMA && CANTOR_MANY_WARP_STORES && CANTOR_MANY_CLUSTER <= 1
        if (use_tma) {
            // per warp: lane 0 makes sure its earlier bulk stores have read their buffers (the one of step s - 1 is the buffer step s + 1
            // writes), every lane publishes its row to the async proxy, lane 0 stores the warp's 32 rows (32 x 52 B = 13 full lines)
            const int w = threadIdx.x >> 5;
            const int wrows = min(32, rows - 32 * w);
            if ((threadIdx.x & 31) == 0) tma_store_wait_read();
            fence_proxy_async_smem();
            __syncwarp();
            if ((threadIdx.x & 31) == 0 && wrows > 0) {
                const uint32_t bytes = (uint32_t)(wrows * CANTOR_OBS_DIM * sizeof(float));
#if CANTOR_MANY_EVICT_FIRST
                if (!keep_in_l2) tma_store_1d_evict_first(dst + w * 32 * CANTOR_OBS_DIM, tile + w * 32 * CANTOR_OBS_DIM, bytes);
                else
#endif
                tma_store_1d(dst + w * 32 * CANTOR_OBS_DIM, tile + w * 32 * CANTOR_OBS_DIM, bytes);
                tma_store_commit();
            }
            continue;
        }
#elif CANTOR_MANY_TMA
        if (use_tma) {
            // the store issued two steps ago read this step's buffer: thread 0 makes sure it has, before the barrier everybody passes
            if (threadIdx.x == 0) tma_store_wait_read();
            fence_proxy_async_smem();
            __syncthreads();
#if CANTOR_MANY_CLUSTER > 1
            cluster_wait();                       // every CTA of the cluster has issued its stores of step s - 1
#endif
            if (threadIdx.x == 0) {
#if CANTOR_MANY_EVICT_FIRST
                if (!keep_in_l2) tma_store_1d_evict_first(dst, tile, (uint32_t)(rows * CANTOR_OBS_DIM * sizeof(float)));
                else
#endif
                tma_store_1d(dst, tile, (uint32_t)(rows * CANTOR_OBS_DIM * sizeof(float)));
                tma_store_commit();
            }
#if CANTOR_MANY_CLUSTER > 1
            cluster_arrive();
#endif
            continue;
        }
#endif
        __syncthreads();                          // tile s complete; everybody finished copying tile s - 1 (the other buffer) out
#if CANTOR_MANY_CLUSTER > 1
        cluster_wait();
        cluster_arrive();
#endif
        if (use_tma) {                            // 16-byte aligned tile of whole float4s: coalesced 16-byte streaming stores
            const float4* src4 = reinterpret_cast<const float4*>(tile);
            float4* dst4 = reinterpret_cast<float4*>(dst);
            const int n4 = rows * CANTOR_OBS_DIM / 4;
#pragma unroll
            for (int j = threadIdx.x; j < kManyThreads * CANTOR_OBS_DIM / 4; j += kManyThreads) {
                if (j < n4) {
                    if (keep_in_l2) dst4[j] = src4[j];
                    else __stcs(dst4 + j, src4[j]);
                }
            }
        } else {
            for (int j = threadIdx.x; j < rows * CANTOR_OBS_DIM; j += kManyThreads) dst[j] = tile[j];
        }
    }
#if CANTOR_MANY_CLUSTER > 1
    cluster_wait();
#endif
#if CANTOR_MANY_TMA
    if (use_tma && (threadIdx.x & 31) == 0) tma_store_wait_read();              // shared memory must outlive the last bulk reads
#endif
    if (live) {
        store_env<F64>(e, core_arr, cash_arr, pv_arr, i);
        if (MON) {
            if (F64) { reinterpret_cast<double2*>(mon.acc)[2 * i] = macc_d[0]; reinterpret_cast<double2*>(mon.acc)[2 * i + 1] = macc_d[1]; }
            else reinterpret_cast<float4*>(mon.acc)[i] = macc_f;
        }
    }
    if (MON) monitor_epilogue_one_step<kManyThreads>(mon, k, make_float4(0.f, 0.f, 0.f, 0.f), false, red, (double)n_envs * (double)n_steps);
}

// ---------------------------------------------------------------------------------------------------
// Gym-style step, on-the-fly mode.  The env's path is not in memory: episode `e.path` of global env g is global path
// e.path * total_envs + g of the Philox stream that cantor_sim_paths uses, generated one day per step.  State carried between
// steps: {S, v} of day t (v unclamped, as K1 carries it).  The pre-advance option marks (needed by the slippage term, :206-209,
// and by the option leg of the P&L) are re-priced from the carried state; the post-advance marks are priced once and give the
// observation's greeks for free (SHARE).  Terminal step: stale marks (:226-231), greeks of the terminal state evaluated directly.
// With auto_reset = 0 the carried state of a finished env stays at day T - 1 (current_step = T marks it finished), so stepping it
// again reproduces the terminal observation exactly, like the replay kernel does.
struct SimSource {
    float2* __restrict__ sv;     // [n_envs] carried {S, v}
    long long total_envs;
};

template <int MODEL>
__device__ __forceinline__ void sim_day_normals(const SimConsts& sk, unsigned long long gp, int t, float& z1, float& z2) {
    constexpr int NPS = MODEL == 0 ? 1 : 2;
    const unsigned call = (unsigned)((t * NPS) >> 2);
    const int j = (t * NPS) & 3;                                               // first normal of day t inside the Philox call
    const uint4 x = philox4x32_10(make_uint4((unsigned)gp, (unsigned)(gp >> 32), call, kStreamPaths),
                                  make_uint2(sk.seed_lo, sk.seed_hi));
    float n0, n1;
    box_muller((j & 2) ? x.z : x.x, (j & 2) ? x.w : x.y, n0, n1);              // the pair this day's normals live in
    if (MODEL == 0) { z1 = (j & 1) ? n1 : n0; z2 = z1; }
    else { z1 = n0; z2 = n1; }
}

__device__ __forceinline__ float4 sim_first_record(const SimConsts& sk, Greeks& g0) {
    float4 r0;
    r0.x = sk.s0;
    r0.y = fmaxf(sk.v0, 0.f);
    const AtmQuote q = atm_quote_f32(r0.x, r0.y, sk);
    r0.z = q.call;
    r0.w = q.put;
    g0 = q.g;
    return r0;
}

template <int MODEL, bool F64, bool INFO, bool MON, bool VN>
__global__ void __launch_bounds__(kStepThreads, 8)
hedge_step_sim_kernel(const StepConsts k, const SimConsts sk, const SimSource src, int4* __restrict__ core_arr,
                      void* __restrict__ cash_arr, double* __restrict__ pv_arr, long long n_envs,
                      const float2* __restrict__ actions, float* __restrict__ obs, void* __restrict__ reward_arr,
                      unsigned char* __restrict__ done_arr, float* __restrict__ terminal_obs, int auto_reset,
                      const InfoOut info, int obs_tma_ok, int share_quote, const Monitor mon, const VecNormFuse vn) {
    __shared__ __align__(128) float tile[kStepThreads * CANTOR_OBS_DIM];
    __shared__ double red[MON ? 11 * (kStepThreads / 32) : 1];
    __shared__ double vn_scratch[VN ? 2 * 8 * CANTOR_OBS_DIM + 2 * (kStepThreads / 32) : 1];
    double my_reward = 0.0, ret_prev = 0.0;
    float4 fin_acc = make_float4(0.f, 0.f, 0.f, 0.f);
    bool finished_episode = false;
    const unsigned tile_idx = (obs_tma_ok & 4) ? gridDim.x - 1 - blockIdx.x : blockIdx.x;     // CANTOR_STEP_WALK_BACKWARD
    const long long first_env = (long long)tile_idx * kStepThreads;
    const long long i = first_env + threadIdx.x;
    const int rows = (int)min((long long)kStepThreads, n_envs - first_env);
    float* o = tile + threadIdx.x * CANTOR_OBS_DIM;

    pdl_wait_prior_grid();

    if (i < n_envs) {
        EnvRegs e = load_env<F64>(core_arr, cash_arr, pv_arr, i);
        const float2 a = __ldcs(actions + i);
        float2 sv = src.sv[i];
        if (VN && vn.norm_reward) ret_prev = vn.returns[i];
        pdl_launch_dependents();
        const int t_prev = (e.step >= k.T) ? k.T - 1 : e.step;
        const unsigned long long gp = (unsigned long long)e.path * (unsigned long long)src.total_envs +
                                      (unsigned long long)(sk.path_offset + i);
        // ---- day t: the carried state and its marks ---------------------------------------------------------------
        float4 prev;
        prev.x = sv.x;
        prev.y = fmaxf(sv.y, 0.f);
        {
            const AtmQuote q = atm_quote_f32(prev.x, prev.y, sk);
            prev.z = q.call;
            prev.w = q.put;
        }
        // ---- day t + 1 ---------------------------------------------------------------------------------------------
        float z1, z2;
        sim_day_normals<MODEL>(sk, gp, t_prev, z1, z2);
        float S = sv.x, v = sv.y;
        sim_advance<MODEL>(sk, S, v, z1, z2);
        float4 cur;
        cur.x = S;
        cur.y = fmaxf(v, 0.f);
        cur.z = prev.z;
        cur.w = prev.w;                                                       // stale marks at the terminal step
        Greeks g_cur{0.f, 0.f, 0.f};
        if (t_prev + 1 < k.T) {
            const AtmQuote q = atm_quote_f32(cur.x, cur.y, sk);
            cur.z = q.call;
            cur.w = q.put;
            g_cur = q.g;
        } else if (!F64 && share_quote) {
            g_cur = atm_greeks_f32(cur.x, rintf(cur.x), cur.y, k.g);
        }
        const size_t rb = F64 ? sizeof(double) : sizeof(float);
        bool terminated;
        if (!F64 && share_quote)
            terminated = step_body<F64, INFO, MON, true>(k, e, a, prev, cur, g_cur, o, i, n_envs, (char*)reward_arr + i * rb, info,
                                                         mon, fin_acc, finished_episode, VN ? &my_reward : nullptr);
        else
            terminated = step_body<F64, INFO, MON, false>(k, e, a, prev, cur, g_cur, o, i, n_envs, (char*)reward_arr + i * rb, info,
                                                          mon, fin_acc, finished_episode, VN ? &my_reward : nullptr);
        if (!terminated || auto_reset) sv = make_float2(S, v);                // a finished env without auto-reset stays at day T - 1
        if (terminated) {
            if (terminal_obs != nullptr) write_terminal_obs(terminal_obs, i, o);
            if (auto_reset) {
                Greeks g0;
                const float4 r0 = sim_first_record(sk, g0);
                const int episode = e.path + 1;
                if (!F64 && share_quote) reset_regs<F64, true>(k, r0, episode, o, e, g0);
                else reset_regs<F64, false>(k, r0, episode, o, e, g0);
                sv = make_float2(sk.s0, sk.v0);
            }
        }
        store_env<F64>(e, core_arr, cash_arr, pv_arr, i);
        src.sv[i] = sv;
        done_arr[i] = terminated ? 1 : 0;
    } else {
        pdl_launch_dependents();
    }
    store_obs_tile(obs, tile, first_env, rows, (obs_tma_ok & 1) && (rows % 4 == 0), (obs_tma_ok & 2) != 0);
    if (VN) vecnorm_partials<kStepThreads>(vn, tile, rows, i < n_envs, i, my_reward, ret_prev, vn_scratch, tile_idx);
    if (MON) monitor_epilogue_one_step<kStepThreads>(mon, k, fin_acc, finished_episode, red, (double)n_envs);
}

// ---------------------------------------------------------------------------------------------------
// Reset.  SIM = false: episode index path_idx[i] of the book; SIM = true: episode number episode[i] (NULL = 0) of the
// on-the-fly stream.
template <bool F64, bool SIM>
__global__ void __launch_bounds__(kStepThreads)
env_reset_kernel(const StepConsts k, const Book b, const SimConsts sk, const SimSource src, int share_quote,
                 int4* __restrict__ core_arr, void* __restrict__ cash_arr, double* __restrict__ pv_arr, long long n_envs,
                 const unsigned char* __restrict__ mask, const int* __restrict__ path_idx, float* __restrict__ obs,
                 void* __restrict__ episode_acc) {
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
    EnvRegs e{};
    Greeks g0{0.f, 0.f, 0.f};
    if (SIM) {
        const float4 r0 = sim_first_record(sk, g0);
        const int episode = path_idx != nullptr ? path_idx[i] : 0;
        if (!F64 && share_quote) reset_regs<F64, true>(k, r0, episode, o, e, g0);
        else reset_regs<F64, false>(k, r0, episode, o, e, g0);
        src.sv[i] = make_float2(sk.s0, sk.v0);
    } else {
        const int path = path_idx[i];
        reset_regs<F64, false>(k, b.rec[path], path, o, e, g0);
    }
    store_env<F64>(e, core_arr, cash_arr, pv_arr, i);
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

static int make_info(const cantor_info_out* info, int precision, InfoOut* io) {
    *io = InfoOut{nullptr, nullptr, nullptr};
    if (info == nullptr) return CANTOR_OK;
    CANTOR_REQUIRE(info->i32 != nullptr, "info.i32 is NULL");
    CANTOR_REQUIRE(info->f64 != nullptr || (precision == CANTOR_F32 && info->f32 != nullptr),
                   "info.f64 is NULL (float32 info arrays, info.f32, exist in F32 mode only)");
    *io = InfoOut{info->f64, info->i32, precision == CANTOR_F32 ? info->f32 : nullptr};
    return CANTOR_OK;
}

static int make_monitor(const cantor_env_state* state, Monitor* mon) {
    *mon = Monitor{state->episode_acc, state->episode_return, state->episode_length, {}};
    return make_stats_out(state->episode_acc ? state->stats : nullptr, &mon->stats);
}

// state->vecnorm: fused VecNormalize moments (INFO must be off: the info path keeps the two-kernel cantor_vecnorm_step).
static int make_vecnorm(const cantor_env_state* state, bool info_on, int64_t n_envs, VecNormFuse* vn) {
    *vn = VecNormFuse{nullptr, nullptr, 0.0, 0, 0};
    const cantor_vecnorm_fuse* f = state->vecnorm;
    if (f == nullptr) return CANTOR_OK;
    CANTOR_REQUIRE(!info_on, "state.vecnorm (fused VecNormalize moments) cannot be combined with info output");
    CANTOR_REQUIRE(f->partial != nullptr && f->returns != nullptr, "vecnorm.partial / vecnorm.returns is NULL");
    CANTOR_REQUIRE(f->n_partial_ctas >= (n_envs + kStepThreads - 1) / kStepThreads, "vecnorm.partial is too small for n_envs");
    *vn = VecNormFuse{f->partial, f->returns, f->gamma, f->norm_obs, f->norm_reward};
    return CANTOR_OK;
}

static int make_sim_source(const cantor_env_sim* src, SimConsts* sk, SimSource* ss) {
    CANTOR_REQUIRE(src != nullptr && src->sim != nullptr && src->sv != nullptr, "sim source / sim params / sv is NULL");
    CANTOR_REQUIRE(src->sim->model == CANTOR_MODEL_GBM || src->sim->model == CANTOR_MODEL_HESTON, "sim.model");
    CANTOR_REQUIRE(src->sim->dt > 0 && src->sim->tenor > 0, "dt and tenor must be positive");
    CANTOR_REQUIRE(src->episode_length > 0, "episode_length must be positive");
    CANTOR_REQUIRE(src->total_envs > 0 && src->sim->path_offset >= 0, "total_envs must be positive, path_offset >= 0");
    CANTOR_REQUIRE((reinterpret_cast<uintptr_t>(src->sv) & 7u) == 0, "sv must be 8-byte aligned");
    fill_sim_consts(src->sim, src->episode_length, sk);
    *ss = SimSource{reinterpret_cast<float2*>(src->sv), src->total_envs};
    return CANTOR_OK;
}

// the observation's greeks can ride on the price evaluation when the env and the simulator agree on (r, tenor)
static int share_quote_ok(const cantor_env_params* p, const StepConsts& k, const SimConsts& sk) {
    return (p->record_metrics && k.g.r_f == sk.r && k.g.T_f == sk.tenor && sk.tenor > 1e-6f) ? 1 : 0;
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
    const SimConsts sk = {};
    const SimSource ss{nullptr, 0};
    if (precision == CANTOR_F64)
        env_reset_kernel<true, false><<<grid, kStepThreads, 0, s>>>(k, b, sk, ss, 0, (int4*)state->core, state->cash, state->pv_prev,
                                                                    n_envs, mask, path_idx, obs, state->episode_acc);
    else
        env_reset_kernel<false, false><<<grid, kStepThreads, 0, s>>>(k, b, sk, ss, 0, (int4*)state->core, state->cash, nullptr, n_envs,
                                                                     mask, path_idx, obs, state->episode_acc);
    return check_launch("env_reset_kernel");
}

extern "C" int cantor_env_reset_sim(const cantor_env_params* params, const cantor_env_sim* source,
                                    const cantor_env_state* state, int64_t n_envs, int32_t precision,
                                    const uint8_t* mask, const int32_t* episode, float* obs, void* stream) {
    StepConsts k;
    SimConsts sk;
    SimSource ss;
    int rc = make_sim_source(source, &sk, &ss);
    if (rc) return rc;
    rc = make_step_consts(params, source->episode_length, &k);
    if (rc) return rc;
    rc = check_state(state, precision);
    if (rc) return rc;
    CANTOR_REQUIRE(n_envs >= 0, "n_envs < 0");
    if (n_envs == 0) return CANTOR_OK;
    const unsigned grid = (unsigned)((n_envs + kStepThreads - 1) / kStepThreads);
    cudaStream_t s = (cudaStream_t)stream;
    const Book b{nullptr, 0, 1};
    const int share = share_quote_ok(params, k, sk);
    if (precision == CANTOR_F64)
        env_reset_kernel<true, true><<<grid, kStepThreads, 0, s>>>(k, b, sk, ss, share, (int4*)state->core, state->cash, state->pv_prev,
                                                                   n_envs, mask, episode, obs, state->episode_acc);
    else
        env_reset_kernel<false, true><<<grid, kStepThreads, 0, s>>>(k, b, sk, ss, share, (int4*)state->core, state->cash, nullptr, n_envs,
                                                                    mask, episode, obs, state->episode_acc);
    return check_launch("env_reset_kernel<sim>");
}

static bool force_per_step_launches() {       // CANTOR_STEP_MANY_LAUNCHES=1: cantor_env_step_many as n_steps chained launches (A/B runs)
    const char* v = getenv("CANTOR_STEP_MANY_LAUNCHES");
    return v != nullptr && v[0] != '\0' && v[0] != '0';
}

static int env_step_impl(const cantor_env_params* params, const cantor_replay_book* book,
                         const cantor_env_state* state, int64_t n_envs, int32_t precision, const float* actions,
                         float* obs, void* reward, uint8_t* done, float* terminal_obs, int32_t auto_reset,
                         const cantor_reset_rule* reset_rule, const cantor_info_out* info, void* stream,
                         int32_t n_steps, bool persistent) {
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
    InfoOut io;
    rc = make_info(info, precision, &io);
    if (rc) return rc;
    Monitor mon;
    rc = make_monitor(state, &mon);
    if (rc) return rc;
    const bool mon_on = state->episode_acc != nullptr;
    VecNormFuse vn;
    rc = make_vecnorm(state, info != nullptr, n_envs, &vn);
    if (rc) return rc;
    CANTOR_REQUIRE(vn.partial == nullptr || !persistent || n_steps <= 1, "cantor_env_step_many does not produce fused VecNormalize moments");
    if (n_envs == 0 || n_steps == 0) return CANTOR_OK;
    const unsigned grid = (unsigned)((n_envs + kStepThreads - 1) / kStepThreads);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t reward_bytes = precision == CANTOR_F64 ? sizeof(double) : sizeof(float);
    int4* core = (int4*)state->core;
    void* cash = state->cash;
    double* pv = state->pv_prev;
    long long n = n_envs;
    const int keep = (reset_rule != nullptr && (reset_rule->flags & CANTOR_STEP_KEEP_OBS_IN_L2)) ? 2 : 0;
    if (persistent && n_steps > 1 && !force_per_step_launches()) {
        // one launch for the whole action tape: state in registers, 81 B + 40 B / n_steps per env-step
        const float2* a0 = (const float2*)actions;
        int tma_ok = (aligned16(obs) ? 1 : 0) | keep;
        int ns = n_steps;
        void* args[] = {&k, &b, &core, &cash, &pv, &n, &ns, &a0, &obs, &reward, &done, &terminal_obs, &rr, &tma_ok, &mon};
        const void* fn = precision == CANTOR_F64
            ? (mon_on ? (const void*)hedge_step_many_kernel<true, true> : (const void*)hedge_step_many_kernel<true, false>)
            : (mon_on ? (const void*)hedge_step_many_kernel<false, true> : (const void*)hedge_step_many_kernel<false, false>);
        unsigned grid_many = (unsigned)((n_envs + kManyThreads - 1) / kManyThreads);
        grid_many = (grid_many + CANTOR_MANY_CLUSTER - 1) / CANTOR_MANY_CLUSTER * CANTOR_MANY_CLUSTER;   // CTAs beyond the envs only keep the barriers company
        return launch_pdl(fn, dim3(grid_many), dim3(kManyThreads), s, args, 0, CANTOR_MANY_CLUSTER);
    }
    // The alternating walk pays wherever a launch's footprint presses on L2: beyond ~1.5 M envs, or when the observations are kept in L2 /
    // info, Monitor or VecNormalize records ride along.  The plain fast path at 2^20 envs (everything it re-reads fits) is 2-3 % faster
    // walking forward every time (17.3 against 17.9 us per launch), so it keeps doing that.
    const bool alternate = CANTOR_STEP_ALTERNATE && (n_envs > 1500000 || keep != 0 || info != nullptr || mon_on || vn.partial != nullptr);
    for (int32_t t = 0; t < n_steps; ++t) {
        // step t of a rollout writes slab t of the caller's [n_steps, n_envs, ...] buffers
        const float2* a_t = (const float2*)actions + (size_t)t * n_envs;
        float* obs_t = obs + (size_t)t * n_envs * CANTOR_OBS_DIM;
        void* rew_t = (char*)reward + (size_t)t * n_envs * reward_bytes;
        unsigned char* done_t = done + (size_t)t * n_envs;
        int tma_ok = (aligned16(obs_t) ? 1 : 0) | keep | ((alternate && (rr.episode_counter & 1)) ? 4 : 0);
        void* args[] = {&k, &b, &core, &cash, &pv, &n, &a_t, &obs_t, &rew_t, &done_t, &terminal_obs, &auto_reset,
                        &rr, &io, &tma_ok, &mon, &vn};
        const void* fn;
#define PICK(F64) (info ? (mon_on ? (const void*)hedge_step_kernel<F64, true, true, false> : (const void*)hedge_step_kernel<F64, true, false, false>) \
                        : (vn.partial ? (mon_on ? (const void*)hedge_step_kernel<F64, false, true, true> : (const void*)hedge_step_kernel<F64, false, false, true>) \
                                      : (mon_on ? (const void*)hedge_step_kernel<F64, false, true, false> : (const void*)hedge_step_kernel<F64, false, false, false>)))
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
                         reset_rule, info, stream, 1, false);
}

extern "C" int cantor_env_step_many(const cantor_env_params* params, const cantor_replay_book* book,
                                    const cantor_env_state* state, int64_t n_envs, int32_t precision, int32_t n_steps,
                                    const float* actions, float* obs, void* reward, uint8_t* done, float* terminal_obs,
                                    const cantor_reset_rule* reset_rule, void* stream) {
    CANTOR_REQUIRE(n_steps >= 0, "n_steps < 0");
    return env_step_impl(params, book, state, n_envs, precision, actions, obs, reward, done, terminal_obs, 1,
                         reset_rule, nullptr, stream, n_steps, true);
}

extern "C" int cantor_env_step_sim(const cantor_env_params* params, const cantor_env_sim* source,
                                   const cantor_env_state* state, int64_t n_envs, int32_t precision,
                                   const float* actions, float* obs, void* reward, uint8_t* done, float* terminal_obs,
                                   int32_t auto_reset, const cantor_info_out* info, int32_t flags, void* stream) {
    StepConsts k;
    SimConsts sk;
    SimSource ss;
    int rc = make_sim_source(source, &sk, &ss);
    if (rc) return rc;
    rc = make_step_consts(params, source->episode_length, &k);
    if (rc) return rc;
    rc = check_state(state, precision);
    if (rc) return rc;
    CANTOR_REQUIRE(n_envs >= 0, "n_envs < 0");
    CANTOR_REQUIRE(actions && obs && reward && done, "actions/obs/reward/done is NULL");
    CANTOR_REQUIRE((reinterpret_cast<uintptr_t>(actions) & 7u) == 0, "actions must be 8-byte aligned");
    InfoOut io;
    rc = make_info(info, precision, &io);
    if (rc) return rc;
    Monitor mon;
    rc = make_monitor(state, &mon);
    if (rc) return rc;
    const bool mon_on = state->episode_acc != nullptr;
    VecNormFuse vn;
    rc = make_vecnorm(state, info != nullptr, n_envs, &vn);
    if (rc) return rc;
    if (n_envs == 0) return CANTOR_OK;
    const unsigned grid = (unsigned)((n_envs + kStepThreads - 1) / kStepThreads);
    int4* core = (int4*)state->core;
    void* cash = state->cash;
    double* pv = state->pv_prev;
    long long n = n_envs;
    const float2* a = (const float2*)actions;
    int tma_ok = (aligned16(obs) ? 1 : 0) | ((flags & CANTOR_STEP_KEEP_OBS_IN_L2) ? 2 : 0) | ((flags & CANTOR_STEP_WALK_BACKWARD) ? 4 : 0);
    int share = share_quote_ok(params, k, sk);
    void* args[] = {&k, &sk, &ss, &core, &cash, &pv, &n, &a, &obs, &reward, &done, &terminal_obs, &auto_reset, &io, &tma_ok,
                    &share, &mon, &vn};
    const void* fn;
#define PICK2(MODEL, F64) (info ? (mon_on ? (const void*)hedge_step_sim_kernel<MODEL, F64, true, true, false> : (const void*)hedge_step_sim_kernel<MODEL, F64, true, false, false>) \
                                : (vn.partial ? (mon_on ? (const void*)hedge_step_sim_kernel<MODEL, F64, false, true, true> : (const void*)hedge_step_sim_kernel<MODEL, F64, false, false, true>) \
                                              : (mon_on ? (const void*)hedge_step_sim_kernel<MODEL, F64, false, true, false> : (const void*)hedge_step_sim_kernel<MODEL, F64, false, false, false>)))
    if (source->sim->model == CANTOR_MODEL_GBM) fn = precision == CANTOR_F64 ? PICK2(0, true) : PICK2(0, false);
    else fn = precision == CANTOR_F64 ? PICK2(1, true) : PICK2(1, false);
#undef PICK2
    return launch_pdl(fn, dim3(grid), dim3(kStepThreads), (cudaStream_t)stream, args);
}
