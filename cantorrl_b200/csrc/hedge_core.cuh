// Device-side core of the hedging environment, shared by the per-step kernel (hedge_step.cu) and the
// episode-fused rollout kernel (rollout.cu): constants, packed state helpers, the float32 ledger of one
// env-step and the observation builders.
//
// Reference semantics: HedgingEnv.step / _get_observation / reset,
// src/env/hedging_env_v2.py:175-294 / :109-143 / :145-173 (v1: src/env/hedging_env.py).
#pragma once
#include "bs_math.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace cantor {

// Everything derived from cantor_env_params once per launch, passed by value (lives in the constant bank).
struct StepConsts {
    // float64 ledger (F64 kernels)
    double cost_per_contract, lambda_cost, neg_w, theta_weight, bps_frac, initial_cash, mult_d, shares_d;
    // float32 ledger (F32 kernels): the same quantities, reciprocals instead of divisions
    float cost_f, lambda_f, neg_w_f, theta_per_step_f, slip_f, inv_shares_f, mult_f, inv_mc_f, inv_T_f;
    float shares_f, max_trade_f, initial_cash_f;
    int max_trade, max_contracts, shares, loss_mse, T;
    GreekConsts g;
};

struct Book {
    const float4* __restrict__ rec;   // {S, v, C, P} at [t * ld + path]; row T carries the marks of row T-1
    long long ld;
    int n_paths;
};

__device__ __forceinline__ int unpack_lo(int packed) { return (int)(short)(packed & 0xffff); }
__device__ __forceinline__ int unpack_hi(int packed) { return packed >> 16; }
__device__ __forceinline__ int pack_pos(int c, int p) { return (c & 0xffff) | (p << 16); }

// hedging_env_v2.py:181-188: float32 product, rint (half to even), int cast, clip.
// NaN / +-inf / |x| >= 2^63 take the x86 "integer indefinite" value INT64_MIN in the reference's
// astype(int), which the clip turns into -max_trade; reproduced here explicitly.
__device__ __forceinline__ int requested_trade(float scaled, float max_trade_f) {
    const float r = fminf(fmaxf(rintf(scaled), -max_trade_f), max_trade_f);
    return (int)((fabsf(scaled) < 9.2233720368547758e18f) ? r : -max_trade_f);
}

__device__ __forceinline__ float clip_unit(float x) {          // np.clip(x, -1, 1); NaN stays NaN
    const float c = fminf(fmaxf(x, -1.f), 1.f);
    return (x != x) ? x : c;
}

// ---- float32 ledger of one env-step (hedging_env_v2.py:178-262) --------------------------------------
// The portfolio value is never formed: the step P&L is the sum of three differences (stock leg, option
// legs, cost), each small, so nothing cancels catastrophically and the float32 stock leg of the reference
// (:235) enters exactly.  `prev` / `cur` are the path records at the env's step t and t+1.
struct LedgerF32 {
    int new_c, new_p, req_c, req_p;
    float cf_c, cf_p, commission, slippage, costs, step_pnl, pps, rpc, tcp, theta_pen, reward, opt_new;
};

__device__ __forceinline__ LedgerF32 ledger_f32(const StepConsts& k, float ax, float ay, int pos_c, int pos_p, int step,
                                                float inv_s0, const float4& prev, const float4& cur, bool frozen) {
    LedgerF32 L;
    L.cf_c = __fmul_rn(ax, k.max_trade_f);                                    // :181-182
    L.cf_p = __fmul_rn(ay, k.max_trade_f);
    L.req_c = requested_trade(L.cf_c, k.max_trade_f);
    L.req_p = requested_trade(L.cf_p, k.max_trade_f);
    L.new_c = frozen ? pos_c : max(-k.max_contracts, min(k.max_contracts, pos_c + L.req_c));   // :193-197
    L.new_p = frozen ? pos_p : max(-k.max_contracts, min(k.max_contracts, pos_p + L.req_p));
    const int atc = abs(L.new_c - pos_c), atp = abs(L.new_p - pos_p);
    L.commission = (float)(atc + atp) * k.cost_f;                             // :203-204
    L.slippage = fmaf((float)atc, prev.z, (float)atp * prev.w) * k.slip_f;    // :206-210 PRE-advance option prices
    L.costs = L.commission + L.slippage;
    const float d_stock = __fmul_rn(k.shares_f, cur.x) - __fmul_rn(k.shares_f, prev.x);
    L.opt_new = fmaf((float)L.new_c, cur.z, (float)L.new_p * cur.w);
    const float opt_old = fmaf((float)pos_c, prev.z, (float)pos_p * prev.w);
    L.step_pnl = fmaf(L.opt_new - opt_old, k.mult_f, d_stock) - L.costs;      // :233-237
    if (step == 0 && k.initial_cash_f != 0.f) {
        // the reference's first portfolio_value_t_minus_1 is float32(stock + cash) (:167-168)
        const float stock0 = __fmul_rn(k.shares_f, prev.x);
        L.step_pnl += (stock0 - __fadd_rn(stock0, k.initial_cash_f)) + k.initial_cash_f;
    }
    L.pps = k.shares != 0 ? L.step_pnl * k.inv_shares_f : L.step_pnl;         // :238
    const float term = k.loss_mse ? (L.pps * inv_s0) * (L.pps * inv_s0) : fabsf(L.pps) * inv_s0;   // :246-253
    L.rpc = k.neg_w_f * term;
    L.tcp = k.lambda_f * L.costs;
    L.theta_pen = k.theta_per_step_f * (float)(k.T - (step + 1));             // :259-260
    L.reward = frozen ? 0.f : (L.rpc - L.tcp) - L.theta_pen;                  // :262
    return L;
}

// hedging_env_v2.py:109-143, float64 ledger.
__device__ __forceinline__ void make_observation_f64(float* __restrict__ o, const StepConsts& k, float S, float v,
                                                     float C, float P, float s0, int pos_c, int pos_p, int step,
                                                     float S_prev, float v_prev) {
    const float s0_safe = fmaxf(s0, 25.0f);                                   // :116
    o[0] = __fdiv_rn(S, s0_safe);
    o[1] = __fdiv_rn(C, s0_safe);
    o[2] = __fdiv_rn(P, s0_safe);
    o[3] = k.max_contracts != 0 ? (float)((double)pos_c / (double)k.max_contracts) : 0.f;   // :120 int64 / int
    o[4] = k.max_contracts != 0 ? (float)((double)pos_p / (double)k.max_contracts) : 0.f;
    o[5] = v;
    o[6] = k.T != 0 ? (float)((double)(k.T - step) / (double)k.T) : 0.f;      // :122
    const Greeks g = atm_greeks_f64(S, rintf(S), v, k.g);                     // :124-127, np.round = half to even
    o[7] = g.call_delta;
    o[8] = g.gamma;
    o[9] = g.put_delta;
    o[10] = g.gamma;
    float ret = 0.f, dv = 0.f;
    if (step != 0 && S_prev != 0.f) {                                         // :129-134
        ret = __fdiv_rn(__fsub_rn(S, S_prev), S_prev);
        dv = __fsub_rn(v, v_prev);
    }
    o[11] = clip_unit(ret);                                                   // :135-136
    o[12] = clip_unit(dv);
}

// The same observation on the float32 throughput path (reciprocals from the SFU, no divisions).
// `o` may point to shared memory (stride-13 rows) or to registers.
// `g` = the ATM greeks of (S, v), computed by the caller (atm_greeks_f32, or shared with the price: atm_quote_f32).
__device__ __forceinline__ void make_observation_f32(float* __restrict__ o, const StepConsts& k, float S, float v,
                                                     float C, float P, float inv_s0, int pos_c, int pos_p, int step,
                                                     float S_prev, float v_prev, const Greeks& g) {
    o[0] = S * inv_s0;
    o[1] = C * inv_s0;
    o[2] = P * inv_s0;
    o[3] = (float)pos_c * k.inv_mc_f;
    o[4] = (float)pos_p * k.inv_mc_f;
    o[5] = v;
    o[6] = (float)(k.T - step) * k.inv_T_f;
    o[7] = g.call_delta;
    o[8] = g.gamma;
    o[9] = g.put_delta;
    o[10] = g.gamma;
    float ret = 0.f, dv = 0.f;
    if (step != 0 && S_prev != 0.f) {
        ret = (S - S_prev) * mufu_rcp(S_prev);
        dv = v - v_prev;
    }
    o[11] = clip_unit(ret);
    o[12] = clip_unit(dv);
}
__device__ __forceinline__ void make_observation_f32(float* __restrict__ o, const StepConsts& k, float S, float v,
                                                     float C, float P, float inv_s0, int pos_c, int pos_p, int step,
                                                     float S_prev, float v_prev) {
    make_observation_f32(o, k, S, v, C, P, inv_s0, pos_c, pos_p, step, S_prev, v_prev, atm_greeks_f32(S, rintf(S), v, k.g));
}

// ---- episode statistics (include/cantor_hedge.h: cantor_stats_out) ------------------------------------------------
struct StatsOut {
    double* sums;               // [CANTOR_STATS_LEN]
    unsigned long long* hist;   // [hist_bins] of b = |sum pps| / T, or NULL
    double* hist_sum;           // [hist_bins] sum of b per bin, or NULL
    float* episode_b;           // [n_episodes_per_env, n_envs] or NULL: per-episode b for an exact CVaR
    float hist_scale;           // bins / hist_max
    int hist_bins;
    long long episode_slots;    // capacity of episode_b in episodes per env
    // fused all-reduce (ticket == NULL: off)
    void* mc_global;            // NVLS multicast address of the global block {sums[16], hist[bins], hist_sum[bins]}, or NULL
    void* peer_global[8];       // unicast addresses of every rank's global block
    int n_peers;
    unsigned* ticket;
};

// ---- fused statistics all-reduce --------------------------------------------------------------------------------------
// multimem.red: one instruction adds a value into the same offset of every GPU's copy of a multicast-mapped buffer; the
// reduction is performed by the NVSwitch (NVLS), the SM only issues the request.
__device__ __forceinline__ void multimem_add_f64(double* mc, double v) {
    asm volatile("multimem.red.relaxed.sys.global.add.f64 [%0], %1;" :: "l"(mc), "d"(v) : "memory");
}
__device__ __forceinline__ void multimem_add_u64(unsigned long long* mc, unsigned long long v) {
    asm volatile("multimem.red.relaxed.sys.global.add.u64 [%0], %1;" :: "l"(mc), "l"(v) : "memory");
}

// Called by every thread of every CTA at the very end of a kernel that accumulated into st.sums / hist / hist_sum.
// The last CTA to arrive (atomic ticket) owns the finished local statistics: it pushes them to all ranks and clears them.
template <int THREADS>
__device__ __forceinline__ void push_statistics_to_all_ranks(const StatsOut& st) {
    if (st.ticket == nullptr) return;
    __shared__ int is_last;
    __threadfence();                                   // this CTA's atomics are visible before its ticket
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(st.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    const int n_words = CANTOR_STATS_LEN + 2 * st.hist_bins;           // global block, 8-byte words
    for (int w = threadIdx.x; w < n_words; w += THREADS) {
        const bool is_count = w >= CANTOR_STATS_LEN && w < CANTOR_STATS_LEN + st.hist_bins;
        unsigned long long* lp;                                         // local word
        if (w < CANTOR_STATS_LEN) lp = reinterpret_cast<unsigned long long*>(st.sums + w);
        else if (is_count) lp = st.hist != nullptr ? st.hist + (w - CANTOR_STATS_LEN) : nullptr;
        else lp = st.hist_sum != nullptr ? reinterpret_cast<unsigned long long*>(st.hist_sum + (w - CANTOR_STATS_LEN - st.hist_bins)) : nullptr;
        if (lp == nullptr) continue;
        const unsigned long long bits = __ldcg(lp);
        if (bits == 0ull) continue;                                     // +0.0 / count 0: nothing to add
        *lp = 0ull;                                                     // local accumulators are per-launch
        if (st.mc_global != nullptr) {
            if (is_count) multimem_add_u64(reinterpret_cast<unsigned long long*>(st.mc_global) + w, bits);
            else multimem_add_f64(reinterpret_cast<double*>(st.mc_global) + w, __longlong_as_double((long long)bits));
        } else {
            for (int r = 0; r < st.n_peers; ++r) {
                if (is_count) atomicAdd(reinterpret_cast<unsigned long long*>(st.peer_global[r]) + w, bits);
                else atomicAdd(reinterpret_cast<double*>(st.peer_global[r]) + w, __longlong_as_double((long long)bits));
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) *st.ticket = 0u;                              // ready for the next launch
}

// Block reduction of NS doubles per thread, then one atomicAdd per value per block.  smem: [NS * THREADS / 32] doubles.
template <int NS, int THREADS>
__device__ __forceinline__ void block_accumulate(double (&v)[NS], double* __restrict__ global, double* smem) {
#pragma unroll
    for (int s = 0; s < NS; ++s) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v[s] += __shfl_down_sync(0xffffffffu, v[s], off);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0 && warp < THREADS / 32) {                    // a CTA may carry extra warps that hold no values
#pragma unroll
        for (int s = 0; s < NS; ++s) smem[s * (THREADS / 32) + warp] = v[s];
    }
    __syncthreads();
    if (threadIdx.x < NS) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < THREADS / 32; ++w) t += smem[threadIdx.x * (THREADS / 32) + w];
        atomicAdd(global + threadIdx.x, t);
    }
    __syncthreads();
}

// The 11 per-episode statistics of one finished episode (sums[0..10]) from its running sums; also the histogram.
// acc = {sum reward, sum pps, sum |pps|, sum cost}.  Returns b = |sum pps| / T.
__device__ __forceinline__ float episode_statistics(double (&stat)[11], float acc_reward, float acc_pps, float acc_abs,
                                                    float acc_cost, float inv_T, const StatsOut& st) {
    const double ea = (double)(acc_abs * inv_T);                // baselines.py:54
    const float ebf = fabsf(acc_pps) * inv_T;
    const double eb = (double)ebf;                              // train_ppo_v2.py:520
    const double ec = (double)(acc_cost * inv_T);               // :521 / baselines.py:55
    const double er = (double)acc_reward, es = (double)acc_pps;
    stat[0] += 1.0;
    stat[1] += ea; stat[2] += ea * ea;
    stat[3] += eb; stat[4] += eb * eb;
    stat[5] += ec; stat[6] += ec * ec;
    stat[7] += er; stat[8] += er * er;
    stat[9] += es; stat[10] += es * es;
    if (st.hist != nullptr) {
        const int bin = min(st.hist_bins - 1, max(0, (int)(ebf * st.hist_scale)));
        atomicAdd(st.hist + bin, 1ull);
        if (st.hist_sum != nullptr) atomicAdd(st.hist_sum + bin, eb);
    }
    return ebf;
}

inline int make_stats_out(const cantor_stats_out* stats, StatsOut* so) {
    *so = StatsOut{};
    if (stats == nullptr) return CANTOR_OK;
    CANTOR_REQUIRE(stats->sums != nullptr, "stats.sums is NULL");
    CANTOR_REQUIRE(stats->hist == nullptr || (stats->hist_bins > 0 && stats->hist_max > 0), "histogram needs bins and a range");
    so->sums = stats->sums;
    so->hist = (unsigned long long*)stats->hist;
    so->hist_sum = stats->hist ? stats->hist_sum : nullptr;
    so->episode_b = stats->episode_b;
    so->hist_scale = stats->hist ? (float)(stats->hist_bins / stats->hist_max) : 0.f;
    so->hist_bins = stats->hist ? stats->hist_bins : 0;
    so->episode_slots = stats->episode_slots;
    if (stats->ticket != nullptr) {
        CANTOR_REQUIRE(stats->mc_global != nullptr || (stats->n_peers >= 1 && stats->n_peers <= 8), "fused all-reduce needs a multicast address or 1..8 peers");
        so->mc_global = stats->mc_global;
        so->n_peers = stats->mc_global != nullptr ? 0 : stats->n_peers;
        for (int r = 0; r < so->n_peers; ++r) {
            CANTOR_REQUIRE(stats->peer_global[r] != nullptr, "peer_global entry is NULL");
            so->peer_global[r] = stats->peer_global[r];
        }
        so->ticket = stats->ticket;
    }
    return CANTOR_OK;
}

// Host side: cantor_env_params -> StepConsts.  T = episode length.
inline int make_step_consts(const cantor_env_params* p, int T, StepConsts* k) {
    CANTOR_REQUIRE(p != nullptr, "params is NULL");
    CANTOR_REQUIRE(T > 0, "episode_length must be positive");
    CANTOR_REQUIRE(p->max_contracts_held >= 0 && p->max_contracts_held <= 32767, "max_contracts_held must be in [0, 32767]");
    CANTOR_REQUIRE(p->max_trade_per_step >= 0 && p->max_trade_per_step <= 32767, "max_trade_per_step must be in [0, 32767]");
    CANTOR_REQUIRE(p->loss_type == CANTOR_LOSS_ABS || p->loss_type == CANTOR_LOSS_MSE, "loss_type");
    k->cost_per_contract = p->transaction_cost_per_contract;
    k->lambda_cost = p->lambda_cost;
    k->neg_w = -p->pnl_penalty_weight;
    k->theta_weight = p->theta_weight;
    k->bps_frac = p->slippage_bps / 10000.0;                                  // hedging_env_v2.py:207
    k->initial_cash = p->initial_cash;
    k->mult_d = (double)p->option_contract_multiplier;
    k->shares_d = (double)p->shares_to_hedge;
    k->cost_f = (float)k->cost_per_contract;
    k->lambda_f = (float)k->lambda_cost;
    k->neg_w_f = (float)k->neg_w;
    k->theta_per_step_f = (float)(p->theta_weight / 252.0);
    k->slip_f = (float)(k->mult_d * k->bps_frac);
    k->inv_shares_f = p->shares_to_hedge != 0 ? (float)(1.0 / k->shares_d) : 1.0f;
    k->mult_f = (float)k->mult_d;
    k->inv_mc_f = p->max_contracts_held != 0 ? (float)(1.0 / (double)p->max_contracts_held) : 0.0f;
    k->inv_T_f = (float)(1.0 / (double)T);
    k->shares_f = (float)p->shares_to_hedge;
    k->max_trade_f = (float)p->max_trade_per_step;
    k->initial_cash_f = (float)p->initial_cash;
    k->max_trade = p->max_trade_per_step;
    k->max_contracts = p->max_contracts_held;
    k->shares = p->shares_to_hedge;
    k->loss_mse = p->loss_type == CANTOR_LOSS_MSE;
    k->T = T;
    k->g.r_f = (float)p->risk_free_rate;
    k->g.T_f = (float)p->option_tenor_years;
    k->g.T_d = p->option_tenor_years;
    k->g.sqrtT_d = sqrt(p->option_tenor_years);
    k->g.inv_sqrtT_f = p->option_tenor_years > 0 ? (float)(1.0 / k->g.sqrtT_d) : 0.f;
    k->g.record_metrics = p->record_metrics;
    return CANTOR_OK;
}

inline int make_book(const cantor_replay_book* book, Book* b) {
    CANTOR_REQUIRE(book != nullptr && book->svcp != nullptr, "book is NULL");
    CANTOR_REQUIRE(aligned16(book->svcp), "book.svcp must be 16-byte aligned");
    CANTOR_REQUIRE(book->n_paths > 0 && book->episode_length > 0, "empty book");
    CANTOR_REQUIRE(book->ld >= book->n_paths, "ld < n_paths");
    b->rec = reinterpret_cast<const float4*>(book->svcp);
    b->ld = book->ld;
    b->n_paths = book->n_paths;
    return CANTOR_OK;
}

}  // namespace cantor
