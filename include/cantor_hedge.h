/*
 * cantor_hedge.h  --  C ABI of libcantor_hedge.so (hand-written sm_100a CUDA kernels)
 *
 * Drop-in boundary for the ONE data-parallel hot path of bcosm/CantorRL:
 *
 *   src/env/hedging_env_v2.py   HedgingEnv.reset / step / _get_observation / _calculate_greeks
 *   src/env/hedging_env.py      (v1 = v2 with slippage_bps = 0, theta_weight = 0, commission 0.05)
 *   src/sim/rbergomi_sim.py     outer log-Euler path step + output schema (:454-464, :528)
 *   src/sim/option_price_assignment.py   Black-Scholes repricing along paths (:10-52)
 *   src/tools/bs_delta.py       single-call Black-Scholes delta hedge (:11-55)
 *
 * The reference has no FFI of its own (it is pure Python); the entry points below are what a
 * ctypes/cffi binding of that path binds instead of the Python methods cited at each function.
 *
 * Conventions
 *   - Every function returns 0 (CANTOR_OK) or a CANTOR_ERR_* code; cantor_last_error() returns the
 *     message of the last failure on the calling thread.  Kernels never trap.
 *   - Unless a function name ends in _host, every pointer is a DEVICE pointer owned by the caller
 *     (e.g. torch.Tensor.data_ptr()); nothing is allocated, freed or synchronised inside a call.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream); calls are asynchronous.
 *   - Path/option arrays are TIME-MAJOR: element (t, path) lives at [t * ld + path], so the 32 envs of a
 *     warp read consecutive memory of one time slab.
 *   - precision = CANTOR_F32 : float cash / reward, float Black-Scholes; 137 algorithmic bytes per env-step.
 *     precision = CANTOR_F64 : double cash / portfolio value / reward with the reference's exact
 *     float32/float64 operation ledger (bit-exact integers, <= 1e-6 relative on every float); 157 B.
 */
#ifndef CANTOR_HEDGE_H
#define CANTOR_HEDGE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CANTOR_ABI_VERSION 2
#define CANTOR_OBS_DIM 13            /* hedging_env_v2.py:138-142 */

enum {
    CANTOR_OK = 0,
    CANTOR_ERR_INVALID = 1,          /* bad argument (NULL pointer, size, alignment, enum) */
    CANTOR_ERR_CUDA = 2,             /* a CUDA runtime call failed; message has the CUDA error string */
    CANTOR_ERR_NO_DEVICE = 3,        /* no usable sm_100 device */
    CANTOR_ERR_SHAPE = 4             /* "Data shapes are inconsistent." (hedging_env_v2.py:45-48) */
};

enum { CANTOR_F32 = 32, CANTOR_F64 = 64 };
enum { CANTOR_LOSS_ABS = 0, CANTOR_LOSS_MSE = 1 };   /* "cvar"/unknown strings use ABS (hedging_env_v2.py:250-253) */

/* How a finished env picks its next episode (hedging_env_v2.py:150 draws np_random.integers(num_episodes)). */
enum {
    CANTOR_RESET_SAME_PATH = 0,      /* keep the current path index (env i replays path i: sharded-by-path runs) */
    CANTOR_RESET_FROM_ARRAY = 1,     /* next_path[i], supplied by the host (e.g. the reference's PCG64 draws) */
    CANTOR_RESET_PHILOX = 2          /* Philox4x32-10(seed; env_offset + i, episode counter) mod n_paths */
};

/* HedgingEnv.__init__ keyword arguments (hedging_env_v2.py:10-22) and constants (:56-58). */
typedef struct cantor_env_params {
    double transaction_cost_per_contract;   /* 0.65 (v2) / 0.05 (v1) */
    double lambda_cost;                     /* 1.0 */
    double pnl_penalty_weight;              /* 0.01 */
    double theta_weight;                    /* 0.0 */
    double slippage_bps;                    /* 0.0 */
    double initial_cash;                    /* 0.0 */
    double risk_free_rate;                  /* 0.04      (:57) */
    double option_tenor_years;              /* 30 / 252  (:58) */
    int32_t loss_type;                      /* CANTOR_LOSS_* */
    int32_t shares_to_hedge;                /* 10000 */
    int32_t max_contracts_held;             /* 200, must be <= 32767 */
    int32_t max_trade_per_step;             /* 15 */
    int32_t option_contract_multiplier;     /* 100 (:56) */
    int32_t record_metrics;                 /* 1; 0 zeroes obs[7:11] (:80-81) */
} cantor_env_params;

/* The four env-schema arrays of the reference npz (hedging_env_v2.py:36-48) as ONE packed float32 array in HBM:
 * record (t, path) = {S, v, C, P} = {'paths', 'volatilities' (instantaneous VARIANCE, :84), 'call_prices_atm',
 * 'put_prices_atm'} at svcp[(t * ld + path) * 4 .. +3], t = 0..T.  An env-step is two aligned 16-byte loads.
 * Row T has no option columns in the reference (they are (n, T)); it must repeat the C, P of row T-1, which is
 * exactly the stale mark the reference uses at the terminal step (:226-231).  cantor_pack_book builds it. */
typedef struct cantor_replay_book {
    const float* svcp;     /* [(T+1) * ld * 4], 16-byte aligned */
    int64_t ld;            /* leading dimension in records, >= n_paths */
    int32_t n_paths;       /* num_episodes   (:50) */
    int32_t episode_length;/* T = paths.shape[1] - 1 (:51) */
} cantor_replay_book;

/* Per-env state, struct of arrays, caller-owned.  Packed so one env-step moves 20 B (F32) of state each way. */
struct cantor_stats_out;
/* Optional fusion of VecNormalize's batch moments into the step kernel (see cantor_vecnorm_step_fused): while the CTA's
 * observation tile is still in shared memory the step kernel writes per-CTA partial sums of the observation columns and of
 * the updated discounted returns, so the wrapper does not re-read the batch to get them. */
typedef struct cantor_vecnorm_fuse {
    double* partial;          /* scratch of 28 * (n_partial_ctas + ceil(n_partial_ctas / 128)) doubles: one 28-double record per step-kernel
                                 CTA {obs sum [13], obs sum of squares [13], return sum, sumsq}, then the fold's level-1 records */
    double* returns;          /* [n_envs] discounted returns (VecNormalize.returns): ret <- ret * gamma + reward, in place */
    double gamma;
    int64_t n_partial_ctas;   /* capacity of `partial` in CTAs: >= ceil(n_envs / 128) */
    int32_t norm_obs, norm_reward;
} cantor_vecnorm_fuse;
typedef struct cantor_env_state {
    int32_t* core;         /* [n_envs * 4] 16-byte records {pos, step, path, s0}:
                              pos  = call contracts (low int16) | put contracts (high int16)
                              step = current_step, path = current_episode_idx,
                              s0   = float bits of initial_S0_for_episode (after the < 1e-6 -> 1.0 rule, :157) */
    void* cash;            /* [n_envs] float (F32) or double (F64): cash_balance */
    double* pv_prev;       /* [n_envs] F64 only: portfolio_value_t_minus_1; NULL in F32 mode */
    /* Optional Monitor (SB3 `Monitor` + the evaluation statistics of train_ppo_v2.py:482-530) fused into the step kernel;
     * all NULL = off (the 137-byte fast path).  episode_acc is state: zeroed by cantor_env_reset and at every episode end. */
    void* episode_acc;     /* [n_envs * 4] float (F32) / double (F64): running sums over the current episode of
                              {reward, per_share_step_pnl, |per_share_step_pnl|, transaction_costs_total} */
    void* episode_return;  /* [n_envs] float / double, written when an env finishes: Monitor's info["episode"]["r"] */
    int32_t* episode_length;               /* [n_envs], written when an env finishes: info["episode"]["l"] */
    const struct cantor_stats_out* stats;  /* or NULL: finished episodes are reduced (warp -> block -> atomics) into it */
    const cantor_vecnorm_fuse* vecnorm;    /* or NULL: cantor_env_step / cantor_env_step_sim (without info) also write VecNormalize's
                                              batch moments for cantor_vecnorm_step_fused */
} cantor_env_state;

/* cantor_reset_rule.flags */
#define CANTOR_STEP_KEEP_OBS_IN_L2 1   /* store the observations with the normal L2 policy because a device-side consumer (policy
                                          network, cantor_vecnorm_step) reads them next; default: evict-first, which is faster for
                                          the step itself when the observations leave for the host or for a later kernel */
#define CANTOR_STEP_WALK_BACKWARD 2    /* cantor_env_step_sim: walk the env tiles last-first in this launch.  A caller that alternates the flag from
                                          step to step lets every launch start with the state the previous one touched last -- the part that is
                                          still in L2 when the population is larger than the cache (results do not depend on it).
                                          cantor_env_step alternates by itself on the parity of reset_rule.episode_counter. */
typedef struct cantor_reset_rule {
    int32_t mode;              /* CANTOR_RESET_* */
    int32_t flags;             /* CANTOR_STEP_* bits */
    const int32_t* next_path;  /* [n_envs] for FROM_ARRAY, else NULL */
    uint64_t seed;             /* PHILOX key */
    int64_t env_offset;        /* global index of local env 0 (rank * n_envs): results independent of GPU count */
    int64_t episode_counter;   /* PHILOX: a number that changes between steps (e.g. the global step count) */
} cantor_reset_rule;

/* Optional per-step diagnostics = the numeric keys of the info dict (hedging_env_v2.py:268-293).
 * Key-major: value of key k for env i at [k * n_envs + i].  Pass NULL to skip (the fast path). */
#define CANTOR_INFO_F64_KEYS 17
#define CANTOR_INFO_I32_KEYS 6
typedef struct cantor_info_out {
    double* f64;   /* [17 * n_envs]: step_pnl_total, per_share_step_pnl, raw_pnl_deviation_abs, transaction_costs_total,
                      commission_cost, slippage_cost, reward_pnl_component, transaction_cost_penalty, theta_penalty,
                      reward_step, portfolio_value, cash, raw_action_call, raw_action_put, scaled_float_call,
                      scaled_float_put, initial_S0_for_episode */
    int32_t* i32;  /* [6 * n_envs]: call_contracts, put_contracts, requested_calls_rounded_clipped,
                      requested_puts_rounded_clipped, actual_calls_traded, actual_puts_traded */
    float* f32;    /* CANTOR_F32 only, optional: the same 17 keys as float32 [17 * n_envs]; when not NULL the float keys go
                      here and f64 is not written (92 instead of 160 info bytes per env-step).  Ignored in CANTOR_F64 mode. */
} cantor_info_out;

/* ---- library ------------------------------------------------------------------------------------ */
int cantor_abi_version(void);
const char* cantor_last_error(void);
/* Writes "sm_XY" style facts about device `device`; returns CANTOR_ERR_NO_DEVICE without a GPU. */
int cantor_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem);

/* ---- on-disk formats <-> packed book ------------------------------------------------------------------
 * Path-major device copies of the reference npz arrays (paths, volatilities: [n_paths, T+1]; call_prices_atm,
 * put_prices_atm: [n_paths, T]; src_dtype CANTOR_F32 or CANTOR_F64 = the npz dtype, rbergomi_sim.py:528) ->
 * packed float32 book; does the `.astype(np.float32)` of hedging_env_v2.py:38-41 and the transposition. */
int cantor_pack_book(const void* paths, const void* vols, const void* calls, const void* puts,
                     int32_t src_dtype, int32_t n_paths, int32_t episode_length, float* svcp, int64_t ld,
                     void* stream);
/* The inverse (e.g. to np.savez a simulated book in the reference schema). */
int cantor_unpack_book(const float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length,
                       int32_t dst_dtype, void* paths, void* vols, void* calls, void* puts, void* stream);

/* ---- K1: path simulation (+ fused ATM repricing) -------------------------------------------------------
 * Replaces the outer path loop of generate_paths_and_options (src/sim/rbergomi_sim.py:454-464) and its output
 * schema (:528) with GBM / Heston log-Euler paths driven by counter-based Philox4x32-10 normals:
 * counter = (global path index, call number, "PATH"), key = seed, so a path does not depend on the launch
 * geometry or on how paths are sharded over GPUs.  With reprice != 0 the option columns are the closed-form
 * Black-Scholes ATM call / put (K = round(S_t), :418; tenor, :19) that the north star substitutes for the
 * nested-MC pricer (:246-306).  Writes the packed book (see cantor_replay_book), float32. */
enum { CANTOR_MODEL_GBM = 0, CANTOR_MODEL_HESTON = 1 };
typedef struct cantor_sim_params {
    int32_t model;          /* CANTOR_MODEL_* */
    int32_t reprice;        /* 1: fill C, P with Black-Scholes ATM prices; 0: leave them 0 */
    double s0;              /* 100.0  S0_DEFAULT (rbergomi_sim.py:27) */
    double v0;              /* 0.04   XI_DEFAULT (:23): GBM variance / Heston initial variance */
    double r;               /* 0.04   (:13) */
    double dt;              /* 1/252  (:14) */
    double kappa, theta, sigma_v, rho;   /* Heston: dv = kappa (theta - v+) dt + sigma_v sqrt(v+ dt) z_v, corr(z_S, z_v) = rho */
    double tenor;           /* 30/252 (:19) */
    uint64_t seed;          /* 42     (:17) */
    int64_t path_offset;    /* global index of path 0 of this shard */
} cantor_sim_params;
int cantor_sim_paths(const cantor_sim_params* params, int32_t n_paths, int32_t episode_length, float* svcp,
                     int64_t ld, void* stream);
/* Fills the C, P columns of a packed book that already holds S and v (paths loaded from disk). */
int cantor_reprice_atm(float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length, double r, double tenor,
                       void* stream);
/* The reference's own float64 step (rbergomi_sim.py:454-464) on exported draws, for parity: time-major
 * v [(T+1) * ld], dW1 / dW2 [T * ld] (unscaled N(0,1)), per-path S0 / rho [n_paths] -> paths [(T+1) * ld]. */
int cantor_euler_from_normals(const double* S0, const double* v, const double* dW1, const double* dW2,
                              const double* rho, int32_t n_paths, int32_t episode_length, int64_t ld,
                              double r, double dt, double* paths, void* stream);

/* ---- rough-Bergomi generator + nested-Monte-Carlo ATM pricer (the reference's own data generator) -----------------
 * Replaces generate_paths_and_options of src/sim/rbergomi_sim.py (:309-499) and price_rbergomi_option_gpu (:246-306):
 * per-path perturbed parameters (:363-367, clips :35-40), Brownian increments, the "fractional" variance driver
 * (:206-243; evaluated as the circular FIR filter it is equal to, no FFT), the log-Euler step (:454-464), and for every
 * (path, day) the ATM call and put priced by n_mc inner paths x 30 steps from (S_t, K = round(S_t), xi := v_t, H, eta, rho).
 * Defaults in the comments are the reference's module constants. */
typedef struct cantor_rbergomi_params {
    double s0, xi, H, eta, rho;              /* base parameters: estimate_base_params (:171-193); defaults 100, .04, .1, 1, -.7 (:23-27) */
    double perturb_s0, perturb_xi, perturb_H, perturb_eta, perturb_rho;      /* .01 .20 .20 .20 .10 (:29-33) */
    double min_xi_factor, min_eta_factor;    /* .5 .5 (:35-36) */
    double clip_H_min, clip_H_max, clip_rho_min, clip_rho_max;               /* .01 .49 -.99 -.01 (:37-40) */
    double r, dt, tenor;                     /* .04, 1/252, 30/252 (:13-14, :19) */
    int32_t n_mc;                            /* 5000 (:20) */
    int32_t shared_draws;                    /* 0 = call and put on independent draws like the reference (:437-446); 1 = the same draws */
    int32_t tensor_cores;                    /* 1 = the 30-tap filter as split-TF32 tcgen05.mma (throughput form); 0 = float32 FFMA */
    int32_t reserved;
    uint64_t seed;                           /* 42 (:17) */
    int64_t path_offset;                     /* global index of path 0 of this shard */
} cantor_rbergomi_params;
/* Outer generator.  Writes S, v (C = P = 0) of the packed book svcp (or NULL), the per-path parameters path_params
 * [5, n_paths] = {S0, xi, H, eta, rho} (float64; needed by the pricer), and optionally float64 paths64 / v64
 * [n_paths, T + 1].  For parity runs the draws can be supplied: params_in [5, n_paths] and / or the unscaled increments
 * dW1_in, dW2_in [n_paths, M_in] with M_in = next_power_of_two(T + 1) (:200-204, :380-382). */
int cantor_rbergomi_paths(const cantor_rbergomi_params* params, int32_t n_paths, int32_t episode_length,
                          const double* params_in, const double* dW1_in, const double* dW2_in, int32_t M_in, float* svcp,
                          int64_t ld, double* path_params, double* paths64, double* v64, void* stream);
/* Nested-MC ATM prices of days [t_begin, t_end) (t_end <= T) into the C, P columns of the packed book (row T repeats
 * row T - 1).  Any day range can be run at any time and on any shard: the draws depend on (seed, global path, day) only. */
int cantor_rbergomi_price_atm(const cantor_rbergomi_params* params, float* svcp, int64_t ld, int32_t n_paths,
                              int32_t episode_length, const double* path_params, int32_t t_begin, int32_t t_end, void* stream);
/* The same inner-path arithmetic on exported increments dW1, dW2 [batch, n_mc, 32] (parity with the reference's draws).
 * H, eta, rho must be consecutive rows of one [3, batch] float64 array. */
int cantor_rbergomi_price_from_increments(const cantor_rbergomi_params* params, const double* S0, const double* K,
                                          const double* xi, const double* H, const double* eta, const double* rho,
                                          const double* dW1, const double* dW2, int32_t batch, int32_t n_mc, int32_t M,
                                          int32_t is_put, double* price, void* stream);

/* ---- K2: Black-Scholes repricing in float64 (src/sim/option_price_assignment.py, src/tools/bs_delta.py) ---
 * black_scholes_vectorized (:10-21), elementwise over n outputs; stride 0 broadcasts a scalar input. */
int cantor_bs_price(const double* S, const double* K, const double* T, const double* sigma, int64_t n,
                    int32_t stride_S, int32_t stride_K, int32_t stride_T, int32_t stride_sigma, double r,
                    double epsilon, double* call, double* put, void* stream);
/* process_price_paths (:33-52) on time-major float64 paths [(T+1) * ld]: realised volatility of the path prefix
 * (:23-31; column 0 = 0, column 1 = NaN), maturity to the episode end T_t = clip(1 - t/252, 0), strikes
 * K_m = round(S_0) * strike_mult[m] (reference: one strike, multiplier 1).  vols [(T+1) * ld] may be NULL;
 * calls / puts are [n_strikes * (T+1) * ld] ("schema B", generalised to a strike ladder). */
int cantor_schema_b_book(const double* paths, int32_t n_paths, int32_t episode_length, int64_t ld, double r,
                         const double* strike_mult, int32_t n_strikes, double* vols, double* calls, double* puts,
                         void* stream);
/* The same book in float32, throughput form, straight from a packed book (S, v) in HBM: price, call delta and gamma of
 * the strike ladder K_m = round(S_0) * strike_mult[m] along every path.  sigma_source selects the reference's realised
 * volatility of the path prefix (:23-31) or the book's own instantaneous variance, sigma = sqrt(max(v_t, 0)) (Heston);
 * fixed_tenor = 0 runs maturities to the episode end (:38), > 0 prices a constant tenor (the env's 30/252).
 * calls / puts / deltas / gammas are [n_strikes * (T+1) * ld] float32, time-major per strike; deltas and gammas may
 * be NULL.  Delta / gamma follow HedgingEnv._calculate_greeks (src/env/hedging_env_v2.py:94-106). */
enum { CANTOR_SIGMA_REALISED = 0, CANTOR_SIGMA_BOOK_VARIANCE = 1 };
int cantor_reprice_book(const float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length, double r,
                        const float* strike_mult, int32_t n_strikes, int32_t sigma_source, double fixed_tenor,
                        float* calls, float* puts, float* deltas, float* gammas, void* stream);
/* The same with its own leading dimension for the outputs ([n_strikes * (T+1) * out_ld], out_ld >= n_paths): a slice of the
 * paths of a large book (svcp + 4 * first_path, n_paths = slice width, ld = the book's) is priced into slice-sized arrays --
 * BASELINE configs[2]'s 2^24-path book x 8 strikes would need 259 GB of outputs in one piece. */
int cantor_reprice_book_strided(const float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length, double r,
                                const float* strike_mult, int32_t n_strikes, int32_t sigma_source, double fixed_tenor,
                                int64_t out_ld, float* calls, float* puts, float* deltas, float* gammas, void* stream);
/* bs_delta_hedge (src/tools/bs_delta.py:36-55): per-path delta-hedge P&L, time-major paths -> pnl [(T+1) * ld]. */
int cantor_bs_delta_hedge(const double* paths, int32_t n_paths, int32_t episode_length, int64_t ld, double r,
                          double dt, double* pnl, void* stream);

/* ---- K3: fused hedge step (replay mode) ------------------------------------------------------------
 * Replaces HedgingEnv.reset (hedging_env_v2.py:145-173) for the envs with mask[i] != 0 (mask NULL = all).
 * The episode index of env i is path_idx[i] (the value the reference draws at :150).
 * Writes the reset observation to obs[i*13 .. i*13+12] for those envs; other rows are left untouched. */
int cantor_env_reset(const cantor_env_params* params, const cantor_replay_book* book,
                     const cantor_env_state* state, int64_t n_envs, int32_t precision,
                     const uint8_t* mask, const int32_t* path_idx, float* obs, void* stream);

/* Replaces HedgingEnv.step (hedging_env_v2.py:175-294) + _get_observation (:109-143) + _calculate_greeks
 * (:79-107) for n_envs envs in ONE kernel: action -> trade -> commission/slippage -> cash -> advance ->
 * mark-to-market -> P&L reward -> done -> (auto_reset) reset of finished envs -> observation.
 *   actions      [n_envs, 2] float32 (call, put) in [-1, 1] (unclipped values behave as in the reference)
 *   obs          [n_envs, 13] float32; with auto_reset the row of a finished env is its RESET observation
 *   reward       [n_envs] float (F32) / double (F64)
 *   done         [n_envs] uint8: terminated (truncated is always False, :221)
 *   terminal_obs [n_envs, 13] or NULL: pre-reset observation, written only for finished envs
 * With auto_reset = 0 a finished env stays at current_step = T and reports done again if stepped. */
int cantor_env_step(const cantor_env_params* params, const cantor_replay_book* book,
                    const cantor_env_state* state, int64_t n_envs, int32_t precision,
                    const float* actions, float* obs, void* reward, uint8_t* done, float* terminal_obs,
                    int32_t auto_reset, const cantor_reset_rule* reset_rule, const cantor_info_out* info,
                    void* stream);

/* n_steps consecutive cantor_env_step calls (auto_reset on, no info) for open-loop action sequences / rollout storage:
 * step t reads actions[t] and writes obs[t], reward[t], done[t] of [n_steps, n_envs, ...] buffers; results are identical
 * to n_steps cantor_env_step calls.  Because the action tape is known, the whole call is ONE persistent launch that keeps
 * the env state in registers (81 + 40 / n_steps algorithmic bytes per env-step instead of 137); CANTOR_STEP_MANY_LAUNCHES=1
 * in the environment selects n_steps chained launches of the per-step kernel instead.  terminal_obs is [n_envs, 13] or NULL. */
int cantor_env_step_many(const cantor_env_params* params, const cantor_replay_book* book,
                         const cantor_env_state* state, int64_t n_envs, int32_t precision, int32_t n_steps,
                         const float* actions, float* obs, void* reward, uint8_t* done, float* terminal_obs,
                         const cantor_reset_rule* reset_rule, void* stream);

/* ---- K3, on-the-fly mode: the fused hedge step with the path generated inside the kernel -------------------------
 * No book in memory: the env carries {S, v} of its current day; each step draws the day's Philox normals (the counters of
 * cantor_sim_paths), advances the path (src/sim/rbergomi_sim.py:454-464), re-prices the ATM call / put before and after the move
 * (the pre-advance marks feed the slippage term, hedging_env_v2.py:206-209; the terminal step keeps the stale marks, :226-231)
 * and runs the same step body as cantor_env_step.  Outputs are bit-identical to cantor_env_step over a book that
 * cantor_sim_paths wrote with the same parameters.  121 bytes per env-step (SURVEY 8(d) counts 129 with a 24-byte state).
 * Episode e of global env g (= sim->path_offset + local index) runs global path e * total_envs + g, like cantor_rollout;
 * state.core[...].path holds the episode number e. */
typedef struct cantor_env_sim {
    const cantor_sim_params* sim;   /* dynamics, seed, tenor; sim->path_offset = global index of env 0 of this shard */
    float* sv;                      /* [n_envs * 2] carried {S, v} (v unclamped), caller-owned state, 8-byte aligned */
    int64_t total_envs;             /* global env population (>= n_envs) */
    int32_t episode_length;         /* T */
    int32_t reserved;
} cantor_env_sim;
/* episode [n_envs] = episode number each (masked) env starts, or NULL = 0. */
int cantor_env_reset_sim(const cantor_env_params* params, const cantor_env_sim* source, const cantor_env_state* state,
                         int64_t n_envs, int32_t precision, const uint8_t* mask, const int32_t* episode, float* obs,
                         void* stream);
/* flags: CANTOR_STEP_* bits.  With auto_reset a finished env continues with its next episode (e + 1). */
int cantor_env_step_sim(const cantor_env_params* params, const cantor_env_sim* source, const cantor_env_state* state,
                        int64_t n_envs, int32_t precision, const float* actions, float* obs, void* reward, uint8_t* done,
                        float* terminal_obs, int32_t auto_reset, const cantor_info_out* info, int32_t flags, void* stream);

/* ---- VecNormalize on the device -------------------------------------------------------------------------------
 * Replaces Stable-Baselines3's VecNormalize as the reference uses it around its env (src/agents/train_ppo_v2.py:204-208,
 * 305-309; statistics consumed at quantconnect/model_wrapper.py:131): RunningMeanStd of observations and of discounted
 * returns, normalisation + clipping in place.  rms is CANTOR_VECNORM_DOUBLES doubles of device memory (layout: obs mean
 * [0,13), obs var [13,26), obs count [26], return mean / var / count [27..29], then library scratch incl. per-CTA partial sums); returns is
 * [n_envs] doubles.  cantor_vecnorm_init sets mean 0, var 1, count 1e-4, returns 0 (it synchronises the stream).
 * cantor_vecnorm_step = VecNormalize.step_wait() on the arrays a cantor_env_step just wrote: obs [n_envs, 13] and reward
 * [n_envs] (float / double by reward_precision) are normalised IN PLACE, terminal_obs (or NULL) where done. */
#define CANTOR_VECNORM_DOUBLES 16704      /* 64 + 28 * 592 (per-CTA partial sums of the moments kernel) rounded up */
int cantor_vecnorm_init(double* rms, double* returns, int64_t n_envs, void* stream);
int cantor_vecnorm_step(double* rms, double* returns, int64_t n_envs, float* obs, void* reward, int32_t reward_precision,
                        const uint8_t* done, float* terminal_obs, double gamma, double clip_obs, double clip_reward,
                        double epsilon, int32_t training, int32_t norm_obs, int32_t norm_reward, void* stream);
/* The same step_wait() when the step kernel that just ran had state.vecnorm = fuse attached (training mode): the batch moments
 * are already in fuse->partial and the returns already updated, so this only folds the per-CTA partials in a fixed order (one
 * small two-level kernel), does the running-statistics update, and normalises obs / reward in place (returns[done] <- 0,
 * terminal_obs where done).  Same results as cantor_vecnorm_step up to the summation order of the batch moments. */
int cantor_vecnorm_step_fused(double* rms, const cantor_vecnorm_fuse* fuse, int64_t n_envs, float* obs, void* reward,
                              int32_t reward_precision, const uint8_t* done, float* terminal_obs, double clip_obs,
                              double clip_reward, double epsilon, void* stream);

/* ---- episode-fused rollout + statistics ------------------------------------------------------------------
 * Replaces the evaluation loops around the env: evaluate_baseline_policy (src/agents/baselines.py:32-72),
 * run_benchmark_strategy (src/benchmark/delta_and_nothing.py:34-114) and the statistics of run_evaluation
 * (src/agents/train_ppo_v2.py:482-530).  One kernel runs n_steps consecutive env-steps per env with the state in
 * registers: path step (replayed from `book` or simulated on the fly from `sim`, same Philox counters as
 * cantor_sim_paths) -> ATM repricing -> observation -> policy -> fused hedge step -> auto-reset, and reduces
 * per-episode statistics on the fly.  Episode e of global env g uses global path e * total_envs + g
 * (mod n_paths when replaying), so results do not depend on how envs are sharded over GPUs. */
enum {
    CANTOR_POLICY_NO_HEDGE = 0,          /* baselines.py:74-75 */
    CANTOR_POLICY_RANDOM = 1,            /* action_space.sample(): uniform(-1, 1), Philox stream "ACTN" */
    CANTOR_POLICY_DELTA_BASELINES = 2,   /* policy_delta_every_step, baselines.py:77-103 */
    CANTOR_POLICY_DELTA_BENCHMARK = 3,   /* delta_hedging_action_selector, delta_and_nothing.py:122-163 */
    CANTOR_POLICY_MLP = 4,               /* ReLU MLP 13-64-64-2 on the normalised obs, output clipped to [-1, 1] */
    CANTOR_POLICY_ACTIONS = 5,           /* open-loop: actions[step, env, 2] */
    CANTOR_POLICY_LSTM = 6               /* LSTM(13 -> 128) -> ReLU MLP(128 -> 64 -> 64) -> 2 on the tensor cores (bf16): the policy the
                                            reference trained (quantconnect/model_wrapper.py:167-204); cantor_policy.mlp = the
                                            CANTOR_LSTM_IMAGE_BYTES weight image (cantorrl_b200/rollout.py: pack_lstm) */
};
#define CANTOR_LSTM_IMAGE_BYTES 178816   /* 4 gate tiles [128 x 144] + W1 [64 x 144] + W2 [64 x 80] + W3 [16 x 80] bf16 + mean / inv_std.
                                            Gate tile p = hidden units 32 p .. 32 p + 31 as two 64-row halves, rows ordered
                                            [half][gate i, f, g, o][16 units] (one tcgen05.ld.x64 per half in the epilogue); columns =
                                            {w_ih (13), b_ih + b_hh, 0, 0 | w_hh (128)}; the i / f / o rows hold HALF the torch.nn.LSTM
                                            weights and biases (sigmoid(x) is evaluated as 0.5 + 0.5 tanh(x / 2)).  All tiles in the
                                            canonical K-major no-swizzle UMMA layout; pack_lstm builds the image. */
#define CANTOR_MLP_FLOATS 5212           /* W1[13][64] b1[64] W2[64][64] b2[64] W3[64][2] b3[2] obs_mean[13] obs_inv_std[13] */
typedef struct cantor_policy {
    int32_t kind;                        /* CANTOR_POLICY_* */
    int32_t put_leg_disabled;            /* 1: action[1] forced to 0 (one European call) */
    const float* mlp;                    /* [CANTOR_MLP_FLOATS] for CANTOR_POLICY_MLP */
    const float* actions;                /* [n_steps, n_envs, 2] for CANTOR_POLICY_ACTIONS */
    uint64_t seed;                       /* CANTOR_POLICY_RANDOM */
    int32_t mlp_tensor_cores;            /* CANTOR_POLICY_MLP: 0 = float32 FFMA (parity form), 1 = bf16 tcgen05.mma (throughput form) */
    int32_t action_squash;               /* network policies: CANTOR_SQUASH_CLIP = clip the action means to [-1, 1] (what SB3 does with the
                                            Box bounds when it steps the env), CANTOR_SQUASH_TANH = tanh (quantconnect/model_wrapper.py:202) */
    float obs_clip;                      /* network policies: the normalised observation is clipped to +-obs_clip before the first layer:
                                            10 = SB3 VecNormalize's clip_obs (train_ppo_v2.py:204-208); +inf (or <= 0) = no clip, which is
                                            what the deployment wrapper does (quantconnect/model_wrapper.py:131) */
    int32_t reserved;
} cantor_policy;
enum { CANTOR_SQUASH_CLIP = 0, CANTOR_SQUASH_TANH = 1 };

/* sums[] layout; every entry is a plain sum over finished episodes, so shards combine by addition (all-reduce):
 *  0 n_episodes
 *  1,2  sum, sum of squares of a = mean_t |per_share_step_pnl|        (baselines.py:49-54)
 *  3,4  ... of b = |sum_t per_share_step_pnl| / T                       (train_ppo_v2.py:482,520)
 *  5,6  ... of c = sum_t transaction_costs_total / T                    (train_ppo_v2.py:483,521; baselines.py:50,55)
 *  7,8  ... of R = sum_t reward                                         (train_ppo_v2.py:484)
 *  9,10 ... of s = sum_t per_share_step_pnl (signed)
 *  11   env-steps executed
 *  15   error flag: > 0 if a tensor-core MLP launch timed out waiting for an MMA (results invalid) */
#define CANTOR_STATS_LEN 16
typedef struct cantor_stats_out {
    double* sums;                        /* [CANTOR_STATS_LEN], accumulated into (zero it first) */
    uint64_t* hist;                      /* [hist_bins] histogram of b over [0, hist_max) for CVaR95 (:527-530), or NULL */
    double* hist_sum;                    /* [hist_bins] sum of b per bin (tail means exact up to the boundary bin), or NULL */
    float* episode_b;                    /* [episode_slots, n_envs] per-episode b for an exact CVaR, or NULL */
    double hist_max;
    int32_t hist_bins;
    int32_t reserved;
    int64_t episode_slots;
    /* Fused all-reduce over NVLink / NVSwitch (optional; ticket == NULL = off).  The LAST CTA of a launch adds the launch's
     * local statistics into EVERY rank's copy of a symmetric "global" block {double sums[CANTOR_STATS_LEN]; uint64 hist[hist_bins];
     * double hist_sum[hist_bins]} and zeroes the local accumulators: through the NVLS multicast address with multimem.red
     * (the reduction happens in the switch) when mc_global != NULL, else with one atomic per peer on the unicast addresses.
     * The global blocks hold the all-reduced totals once every rank's launch has finished (a barrier, no NCCL call). */
    void* mc_global;                     /* multicast address of the global block, or NULL */
    void* peer_global[8];                /* unicast address of each rank's global block (n_peers entries, own rank included) */
    int32_t n_peers;
    int32_t reserved2;
    uint32_t* ticket;                    /* [1] device counter, zero before the first launch; reset by the kernel */
} cantor_stats_out;

typedef struct cantor_rollout_out {      /* optional rollout storage, time-major */
    float* obs;                          /* [n_steps, n_envs, 13] observation the policy acted on */
    float* actions;                      /* [n_steps, n_envs, 2] */
    float* reward;                       /* [n_steps, n_envs] */
    uint8_t* done;                       /* [n_steps, n_envs] */
} cantor_rollout_out;

/* Exactly one of book / sim is non-NULL (episode_length is taken from the book when replaying).
 * first_episode: episode number the envs start with -- a launch runs episodes first_episode, first_episode + 1, ... of every
 * env, so consecutive calls (e.g. PPO collection phases) see fresh paths and fresh random actions when the caller advances it by
 * the number of episodes already played; 0 reproduces the same rollout.  Statistics count finished episodes only: steps of a
 * trailing partial episode (n_steps % T) run but are not reported. */
int cantor_rollout(const cantor_env_params* params, const cantor_replay_book* book, const cantor_sim_params* sim,
                   int32_t episode_length, const cantor_policy* policy, int64_t n_envs, int64_t env_offset,
                   int64_t total_envs, int64_t first_episode, int32_t n_steps, const cantor_stats_out* stats,
                   const cantor_rollout_out* out, void* stream);

/* ---- host-buffer face (what a NumPy / SubprocVecEnv-style caller binds) ---------------------------------
 * These are the only functions that take HOST pointers, allocate, and synchronise.  The handle owns the device
 * copies; a step copies the actions in, runs the fused hedge-step kernel and copies obs / reward / done back,
 * chunked over streams so the copies overlap the kernel, and returns when the results are on the host.
 * Replaces HedgingEnv.__init__ / reset / step as driven by SB3's VecEnv (src/agents/train_ppo_v2.py:127-141). */
typedef struct cantor_vecenv cantor_vecenv;
int cantor_vecenv_create(cantor_vecenv** out, const cantor_env_params* params, int32_t precision, int64_t n_envs,
                         int32_t device, int32_t n_chunks /* 0 = auto */);
int cantor_vecenv_destroy(cantor_vecenv* env);
/* np.load()-layout host arrays (path-major; paths, vols [n_paths, T+1]; calls, puts [n_paths, T]) -> packed book. */
int cantor_vecenv_load_book_host(cantor_vecenv* env, const void* paths, const void* vols, const void* calls,
                                 const void* puts, int32_t src_dtype, int32_t n_paths, int32_t episode_length);
/* or simulate it in HBM (cantor_sim_paths). */
int cantor_vecenv_simulate_book(cantor_vecenv* env, const cantor_sim_params* sim, int32_t n_paths, int32_t episode_length);
int cantor_vecenv_set_reset_rule(cantor_vecenv* env, int32_t mode, uint64_t seed, int64_t env_offset);
/* path_idx [n_envs] or NULL (env i starts on path (env_offset + i) mod n_paths); obs_host [n_envs, 13] or NULL. */
int cantor_vecenv_reset_host(cantor_vecenv* env, const int32_t* path_idx, float* obs_host);
/* actions_host [n_envs, 2] -> obs_host [n_envs, 13], reward_host [n_envs] (float / double by precision),
 * done_host [n_envs]; next_path_host [n_envs] or NULL supplies the episode each env takes if it finishes now. */
int cantor_vecenv_step_host(cantor_vecenv* env, const float* actions_host, float* obs_host, void* reward_host,
                            uint8_t* done_host, const int32_t* next_path_host);
int cantor_vecenv_episode_length(const cantor_vecenv* env);
int cantor_vecenv_num_paths(const cantor_vecenv* env);
/* cudaHostRegister / cudaHostUnregister of a caller-owned buffer (page-locked copies run at full PCIe speed). */
int cantor_host_register(void* ptr, size_t bytes);
int cantor_host_unregister(void* ptr);
/* Raw ceiling of the host-buffer face: only the copies of cantor_vecenv_step_host, no kernel.  One round = h2d_bytes in and
 * d2h_bytes out between page-locked host memory and HBM, in n_chunks pieces over three streams; sync_each_round = 1 waits for
 * the streams after every round like a gym step has to.  Runs for >= seconds; reports GB/s per direction and rounds/s. */
int cantor_host_copy_probe(int32_t device, int64_t d2h_bytes, int64_t h2d_bytes, int32_t n_chunks, int32_t sync_each_round,
                           double seconds, double* d2h_gbs, double* h2d_gbs, double* rounds_per_s);
/* Measurement aid of the tensor-core actors: clocks per tcgen05.mma (cta_group::1, kind::f16, bf16, M = 128, K = 16) at
 * width n, issued as `reps` back-to-back K-loops of `k_steps` instructions by one thread of one CTA per SM (n_ctas CTAs);
 * a_in_tmem = 0: both operands in shared memory (what the actors use), 1: A in tensor memory.  Synchronous; allocates a few bytes. */
int cantor_umma_probe(int32_t n, int32_t k_steps, int32_t reps, int32_t a_in_tmem, int32_t n_ctas,
                      double* issue_clk_per_mma, double* total_clk_per_mma);

#ifdef __cplusplus
}
#endif
#endif /* CANTOR_HEDGE_H */
