// K3 -- the fused hedge step in replay mode, and the env reset.
//
// One thread owns one env.  A step reads, once, the env's 16-byte state record, its cash, its action and
// the two time slabs (t, t+1) of its path, and writes, once, the 13-float observation, the reward, the
// done flag and the new state: 137 algorithmic bytes (F32) / 157 (F64) per env-step, no re-reads.  The
// observation tile of a CTA (256 envs x 52 B = 13 KB, contiguous in the caller's [n_envs, 13] array) is
// staged in shared memory and leaves the SM as one 1-D TMA bulk store.
//
// Reference semantics: HedgingEnv.step / _get_observation / _calculate_greeks / reset,
// src/env/hedging_env_v2.py:175-294 / :109-143 / :79-107 / :145-173 (v1: src/env/hedging_env.py).
#include "bs_math.cuh"
#include "common.cuh"
#include "philox.cuh"

namespace cantor {

constexpr int kStepThreads = 256;

// Everything derived from cantor_env_params once per launch, passed by value (lives in constant bank).
struct StepConsts {
    // float64 ledger
    double cost_per_contract, lambda_cost, neg_w, theta_weight, bps_frac, initial_cash, mult_d, shares_d;
    double inv_mc_d;        // unused in F64 (true division there), 1/max_contracts for F32
    // float32 ledger
    float shares_f, max_trade_f, initial_cash_f;
    int max_trade, max_contracts, shares, loss_mse, T;
    GreekConsts g;
};

struct Book {
    const float* __restrict__ S;
    const float* __restrict__ v;
    const float* __restrict__ C;
    const float* __restrict__ P;
    long long ld;
    int n_paths;
};

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

__device__ __forceinline__ int unpack_lo(int packed) { return (int)(short)(packed & 0xffff); }
__device__ __forceinline__ int unpack_hi(int packed) { return packed >> 16; }
__device__ __forceinline__ int pack_pos(int c, int p) { return (c & 0xffff) | (p << 16); }

// hedging_env_v2.py:181-188: float32 product, rint (half to even), int cast, clip.
// NaN / +-inf / |x| >= 2^63 take the x86 "integer indefinite" value INT64_MIN in the reference's
// astype(int), which the clip turns into -max_trade; reproduced here explicitly.
__device__ __forceinline__ int requested_trade(float scaled, int max_trade) {
    if (!(fabsf(scaled) < 9.2233720368547758e18f)) return -max_trade;
    const float r = rintf(scaled);
    const float m = (float)max_trade;
    return (int)fminf(fmaxf(r, -m), m);
}

// hedging_env_v2.py:109-143.  `lag_valid` = (current_step != 0 && S_prev != 0).
template <bool F64>
__device__ __forceinline__ void make_observation(float* __restrict__ o, const StepConsts& k, float S, float v, float C,
                                                 float P, float s0, int pos_c, int pos_p, int step, float S_prev,
                                                 float v_prev) {
    const float s0_safe = fmaxf(s0, 25.0f);                                   // :116
    if (F64) {
        o[0] = __fdiv_rn(S, s0_safe);
        o[1] = __fdiv_rn(C, s0_safe);
        o[2] = __fdiv_rn(P, s0_safe);
        o[3] = k.max_contracts != 0 ? (float)((double)pos_c / (double)k.max_contracts) : 0.f;   // :120 int64 / int
        o[4] = k.max_contracts != 0 ? (float)((double)pos_p / (double)k.max_contracts) : 0.f;
        o[6] = k.T != 0 ? (float)((double)(k.T - step) / (double)k.T) : 0.f;  // :122
    } else {
        const float inv = __frcp_rn(s0_safe);
        o[0] = S * inv;
        o[1] = C * inv;
        o[2] = P * inv;
        o[3] = (float)pos_c * (float)k.inv_mc_d;
        o[4] = (float)pos_p * (float)k.inv_mc_d;
        o[6] = k.T != 0 ? (float)(k.T - step) / (float)k.T : 0.f;
    }
    o[5] = v;
    const Greeks g = atm_greeks<F64>(S, rintf(S), v, k.g);                    // :124-127, np.round = half to even
    o[7] = g.call_delta;
    o[8] = g.gamma;
    o[9] = g.put_delta;
    o[10] = g.gamma;
    float ret = 0.f, dv = 0.f;
    if (step != 0 && S_prev != 0.f) {                                         // :129-134
        ret = F64 ? __fdiv_rn(__fsub_rn(S, S_prev), S_prev) : (S - S_prev) / S_prev;
        dv = __fsub_rn(v, v_prev);
    }
    o[11] = fminf(fmaxf(ret, -1.f), 1.f);                                     // :135-136 (NaN propagates like np.clip)
    o[12] = fminf(fmaxf(dv, -1.f), 1.f);
    if (ret != ret) o[11] = ret;
    if (dv != dv) o[12] = dv;
}

// hedging_env_v2.py:150-170 for one env; returns the reset state and fills the reset observation.
template <bool F64>
__device__ __forceinline__ void reset_one(const StepConsts& k, const Book& b, int path, float* __restrict__ o,
                                          int4& core, double& cash, double& pv_prev) {
    const float S0raw = b.S[path];
    const float v0 = b.v[path];
    const float C0 = b.C[path];
    const float P0 = b.P[path];
    const float s0 = (S0raw < 1e-6f) ? 1.0f : S0raw;                          // :157
    core.x = 0;                                                               // no contracts
    core.y = 0;                                                               // current_step
    core.z = path;
    core.w = __float_as_int(s0);
    cash = k.initial_cash;                                                    // :165
    // :167-168 evaluated in float32: (shares * S) + 0 + cash
    pv_prev = (double)__fadd_rn(__fmul_rn(k.shares_f, S0raw), k.initial_cash_f);
    make_observation<F64>(o, k, S0raw, v0, C0, P0, s0, 0, 0, 0, S0raw, v0);
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
                                               int rows, bool use_tma) {
    if (use_tma) {
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
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
template <bool F64, bool INFO>
__global__ void __launch_bounds__(kStepThreads)
hedge_step_kernel(const StepConsts k, const Book b, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                  double* __restrict__ pv_arr, long long n_envs, const float2* __restrict__ actions,
                  float* __restrict__ obs, void* __restrict__ reward_arr, unsigned char* __restrict__ done_arr,
                  float* __restrict__ terminal_obs, int auto_reset, const ResetRule rr, const InfoOut info,
                  int obs_tma_ok) {
    __shared__ __align__(128) float tile[kStepThreads * CANTOR_OBS_DIM];
    const long long first_env = (long long)blockIdx.x * kStepThreads;
    const long long i = first_env + threadIdx.x;
    const int rows = (int)min((long long)kStepThreads, n_envs - first_env);
    float* o = tile + threadIdx.x * CANTOR_OBS_DIM;                            // stride 13 words: conflict-free

    if (i < n_envs) {
        // ---- independent loads first -----------------------------------------------------------------
        int4 core = ldg_stream(core_arr + i);
        const float2 a = ldg_stream(actions + i);
        double cash = F64 ? __ldcs(reinterpret_cast<const double*>(cash_arr) + i)
                          : (double)__ldcs(reinterpret_cast<const float*>(cash_arr) + i);
        double pv_prev = F64 ? __ldcs(pv_arr + i) : 0.0;

        int pos_c = unpack_lo(core.x), pos_p = unpack_hi(core.x);
        int step = core.y;
        const int path = core.z;
        const float s0 = __int_as_float(core.w);
        const bool already_done = step >= k.T;                                // only reachable with auto_reset = 0

        // ---- the two time slabs of this env's path ------------------------------------------------------
        const int t_prev = already_done ? k.T - 1 : step;
        const int t_new = already_done ? k.T : step + 1;
        const bool terminated = t_new >= k.T;                                 // :220
        const int t_opt = terminated ? t_new - 1 : t_new;                     // :226-231 stale option marks at the end
        const float S_prev = ldg_stream(b.S + t_prev * b.ld + path);
        const float v_prev = ldg_stream(b.v + t_prev * b.ld + path);
        const float C_prev = ldg_stream(b.C + t_prev * b.ld + path);
        const float P_prev = ldg_stream(b.P + t_prev * b.ld + path);
        const float S_new = ldg_stream(b.S + t_new * b.ld + path);
        const float v_new = ldg_stream(b.v + t_new * b.ld + path);
        const float C_new = terminated ? C_prev : ldg_stream(b.C + t_opt * b.ld + path);
        const float P_new = terminated ? P_prev : ldg_stream(b.P + t_opt * b.ld + path);

        double reward = 0.0;
        if (!already_done) {
            // ---- (i) action -> trade  :178-200 ----------------------------------------------------------
            const float cf_c = __fmul_rn(a.x, k.max_trade_f);
            const float cf_p = __fmul_rn(a.y, k.max_trade_f);
            const int req_c = requested_trade(cf_c, k.max_trade);
            const int req_p = requested_trade(cf_p, k.max_trade);
            const int new_c = max(-k.max_contracts, min(k.max_contracts, pos_c + req_c));
            const int new_p = max(-k.max_contracts, min(k.max_contracts, pos_p + req_p));
            const int tc = new_c - pos_c, tp = new_p - pos_p;
            const int atc = abs(tc), atp = abs(tp);

            // ---- (ii) commission, slippage on the PRE-advance option prices, cash  :203-213 ------------
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
            double step_pnl;
            if (F64) {
                step_pnl = __dsub_rn(pv, pv_prev);                            // :237
            } else {
                // F32 state carries no portfolio value: rebuild last step's from the slab at t (same float32
                // stock leg, same marks), so pv - pv_prev keeps the reference's roundings.
                const float stock_prev = __fmul_rn(k.shares_f, S_prev);
                double pv_old;
                if (step == 0) {
                    pv_old = (double)__fadd_rn(stock_prev, k.initial_cash_f);  // reset computes it in float32 (:167)
                } else {
                    const double opt_old = __dadd_rn(__dmul_rn(__dmul_rn((double)pos_c, (double)C_prev), k.mult_d),
                                                     __dmul_rn(__dmul_rn((double)pos_p, (double)P_prev), k.mult_d));
                    pv_old = __dadd_rn(__dadd_rn((double)stock_prev, opt_old), cash);
                }
                step_pnl = pv - pv_old;
            }
            const double pps = k.shares != 0 ? __ddiv_rn(step_pnl, k.shares_d) : step_pnl;   // :238

            // ---- (v) reward  :243-262 -------------------------------------------------------------------
            const float s0_floor = fmaxf(s0, 25.0f);
            double term;
            if (k.loss_mse) term = __ddiv_rn(__dmul_rn(pps, pps), (double)__fadd_rn(__fmul_rn(s0_floor, s0_floor), 1e-9f));
            else term = __ddiv_rn(fabs(pps), (double)__fadd_rn(s0_floor, 1e-9f));
            const double rpc = __dmul_rn(k.neg_w, term);
            const double tcp = __dmul_rn(k.lambda_cost, costs);
            const double theta_pen = __dmul_rn(k.theta_weight, __ddiv_rn((double)(k.T - t_new), 252.0));
            reward = __dsub_rn(__dsub_rn(rpc, tcp), theta_pen);

            if (INFO) {
                double* f = info.f64 + i;
                const long long n = n_envs;
                f[0 * n] = step_pnl;   f[1 * n] = pps;        f[2 * n] = fabs(pps);  f[3 * n] = costs;
                f[4 * n] = commission; f[5 * n] = slippage;   f[6 * n] = rpc;        f[7 * n] = tcp;
                f[8 * n] = theta_pen;  f[9 * n] = reward;     f[10 * n] = pv;        f[11 * n] = cash_new;
                f[12 * n] = a.x;       f[13 * n] = a.y;       f[14 * n] = cf_c;      f[15 * n] = cf_p;
                f[16 * n] = s0;
                int* q = info.i32 + i;
                q[0 * n] = new_c; q[1 * n] = new_p; q[2 * n] = req_c; q[3 * n] = req_p; q[4 * n] = tc; q[5 * n] = tp;
            }
            pos_c = new_c;
            pos_p = new_p;
            cash = cash_new;
            pv_prev = pv;
            step = t_new;
        }

        // ---- observation of the advanced state  :266 ------------------------------------------------------
        make_observation<F64>(o, k, S_new, v_new, C_new, P_new, s0, pos_c, pos_p, step, S_prev, v_prev);
        core.x = pack_pos(pos_c, pos_p);
        core.y = step;

        if (terminated) {
            if (terminal_obs != nullptr) {
                float* to = terminal_obs + i * CANTOR_OBS_DIM;
#pragma unroll
                for (int j = 0; j < CANTOR_OBS_DIM; ++j) to[j] = o[j];
            }
            if (auto_reset) {                                                 // VecEnv convention: next episode starts now
                const int next = next_episode_path(rr, b, i, path);
                reset_one<F64>(k, b, next, o, core, cash, pv_prev);
            }
        }

        // ---- stores -------------------------------------------------------------------------------------
        __stcs(core_arr + i, core);
        if (F64) {
            __stcs(reinterpret_cast<double*>(cash_arr) + i, cash);
            __stcs(pv_arr + i, pv_prev);
            __stcs(reinterpret_cast<double*>(reward_arr) + i, reward);
        } else {
            __stcs(reinterpret_cast<float*>(cash_arr) + i, (float)cash);
            __stcs(reinterpret_cast<float*>(reward_arr) + i, (float)reward);
        }
        done_arr[i] = terminated ? 1 : 0;
    }
    store_obs_tile(obs, tile, first_env, rows, obs_tma_ok && (rows % 4 == 0));
}

// ---------------------------------------------------------------------------------------------------
template <bool F64>
__global__ void __launch_bounds__(kStepThreads)
env_reset_kernel(const StepConsts k, const Book b, int4* __restrict__ core_arr, void* __restrict__ cash_arr,
                 double* __restrict__ pv_arr, long long n_envs, const unsigned char* __restrict__ mask,
                 const int* __restrict__ path_idx, float* __restrict__ obs) {
    const long long i = (long long)blockIdx.x * kStepThreads + threadIdx.x;
    if (i >= n_envs) return;
    if (mask != nullptr && mask[i] == 0) return;
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
    CANTOR_REQUIRE(book->S && book->v && book->C && book->P, "book array is NULL");
    CANTOR_REQUIRE(book->n_paths > 0 && book->episode_length > 0, "empty book");
    CANTOR_REQUIRE(book->ld >= book->n_paths, "ld < n_paths");
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
    k->inv_mc_d = p->max_contracts_held != 0 ? 1.0 / (double)p->max_contracts_held : 0.0;
    k->shares_f = (float)p->shares_to_hedge;
    k->max_trade_f = (float)p->max_trade_per_step;
    k->initial_cash_f = (float)p->initial_cash;
    k->max_trade = p->max_trade_per_step;
    k->max_contracts = p->max_contracts_held;
    k->shares = p->shares_to_hedge;
    k->loss_mse = p->loss_type == CANTOR_LOSS_MSE;
    k->T = book->episode_length;
    k->g.r_f = (float)p->risk_free_rate;
    k->g.T_f = (float)p->option_tenor_years;
    k->g.T_d = p->option_tenor_years;
    k->g.sqrtT_d = sqrt(p->option_tenor_years);
    k->g.sqrtT_f = (float)k->g.sqrtT_d;
    k->g.record_metrics = p->record_metrics;
    b->S = book->S; b->v = book->v; b->C = book->C; b->P = book->P;
    b->ld = book->ld;
    b->n_paths = book->n_paths;
    return CANTOR_OK;
}

static int check_state(const cantor_env_state* st, int precision) {
    CANTOR_REQUIRE(st != nullptr && st->core != nullptr && st->cash != nullptr, "state array is NULL");
    CANTOR_REQUIRE(aligned16(st->core), "state.core must be 16-byte aligned");
    CANTOR_REQUIRE(precision == CANTOR_F32 || precision == CANTOR_F64, "precision must be 32 or 64");
    CANTOR_REQUIRE(precision == CANTOR_F32 || st->pv_prev != nullptr, "state.pv_prev is required in F64 mode");
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
                                                             n_envs, mask, path_idx, obs);
    else
        env_reset_kernel<false><<<grid, kStepThreads, 0, s>>>(k, b, (int4*)state->core, state->cash, nullptr, n_envs,
                                                              mask, path_idx, obs);
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
    if (n_envs == 0) return CANTOR_OK;
    const unsigned grid = (unsigned)((n_envs + kStepThreads - 1) / kStepThreads);
    cudaStream_t s = (cudaStream_t)stream;
    const size_t reward_bytes = precision == CANTOR_F64 ? sizeof(double) : sizeof(float);
    for (int32_t t = 0; t < n_steps; ++t) {
        // step t of a rollout writes slab t of the caller's [n_steps, n_envs, ...] buffers
        const float2* a_t = (const float2*)actions + (size_t)t * n_envs;
        float* obs_t = obs + (size_t)t * n_envs * CANTOR_OBS_DIM;
        void* rew_t = (char*)reward + (size_t)t * n_envs * reward_bytes;
        uint8_t* done_t = done + (size_t)t * n_envs;
        const int tma_ok = aligned16(obs_t) ? 1 : 0;
#define LAUNCH(F64, INFO)                                                                                          \
    hedge_step_kernel<F64, INFO><<<grid, kStepThreads, 0, s>>>(k, b, (int4*)state->core, state->cash, state->pv_prev, \
                                                               n_envs, a_t, obs_t, rew_t, done_t, terminal_obs,     \
                                                               auto_reset, rr, io, tma_ok)
        if (precision == CANTOR_F64) {
            if (info) LAUNCH(true, true); else LAUNCH(true, false);
        } else {
            if (info) LAUNCH(false, true); else LAUNCH(false, false);
        }
#undef LAUNCH
        rr.episode_counter += 1;
    }
    return check_launch("hedge_step_kernel");
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
