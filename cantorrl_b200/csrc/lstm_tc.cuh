// Recurrent actor for the rollout kernel on the tensor cores: the policy the reference actually trained,
//   LSTM(13 -> 128) -> ReLU MLP(128 -> 64 -> 64) -> Linear(64 -> 2)       (quantconnect/model_wrapper.py:167-204,
//   shapes in quantconnect/model_files/policy_weights.pth; SB3 RecurrentPPO "MlpLstmPolicy", train_ppo_v2.py:46-47, 220-228)
// evaluated with tcgen05.mma (bf16 operands, float32 accumulation in tensor memory).
//
// One CTA = 256 envs in TWO GROUPS of 128 plus an issuer warp (288 threads).  Warps 0-3 own group 0, warps 4-7 group 1:
// thread m of a group owns env m, TMEM lane m and row m of the group's A tile.  Warp 8 is the ISSUER: it alone talks to
// the tensor core and the TMA unit.  The two sides meet only on mbarriers (env threads arrive with count 128 after
// publishing their shared-memory rows / finishing their TMEM reads; the issuer signals with tcgen05.commit); there is no
// __syncthreads in a step.  The groups ping-pong: the SFU-bound epilogue of one runs while the other's MMAs, head chain or
// env step are in flight, which is what a second resident CTA would give -- shared and tensor memory allow only one.
//
//   A tile [128 x 144] bf16 per group = { normalised obs (13), 1.0, 0, 0 | h (128) }   (biases ride on the ones column)
//   gates: four passes of 32 hidden units each, D[128 x 128] = A * Wg_p^T with Wg_p rows = {i, f, g, o} x 32 units.
//          The 147 KB of gate weights do not fit next to the rest, so the 36 KB tile of pass p + 2 is streamed L2 -> shared
//          by ONE 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) into the buffer pass p just released; both
//          groups use the same tile (group 0's MMAs, then group 1's), so the weights cross the L2 -> SM link once per 256 envs.
//   epilogue of a pass: tcgen05.ld of the thread's own lane (five loads in flight, one wait), sigmoid / tanh on the SFU
//          (tanh.approx), the cell state c in float32 in 128 TMEM columns per group.  h_t is held back in registers as
//          bf16 (64 registers) until the group's last pass has read h_{t-1}, then written over it in the A tile -- one A
//          tile per group instead of two is what makes room for the second group.
//   head:  D[128 x 64] = A * W1^T (K = 144; the x-columns weigh 0, the ones column carries the bias), ReLU,
//          D = A2 * W2^T, ReLU, D[128 x 16] = A2 * W3^T; accumulators reuse the group's gate columns.
//   episode starts (h = c = 0; the envs of a CTA reset in lockstep, every T steps): the passes are x-part only.
// TMEM (all 512 columns, one CTA per SM): [0, 128) / [128, 256) gate + head accumulators of group 0 / 1, [256, 384) /
// [384, 512) cell states.  Weight images arrive pre-arranged in the canonical K-major no-swizzle core-matrix layout
// (cantorrl_b200/rollout.py: pack_lstm), so a tile is one contiguous copy.
#pragma once
#include "mlp_tc.cuh"

namespace cantor {
namespace lstmtc {

using mlptc::kLbo;
constexpr int kRows = 128, kIn = 13, kH = 128;
constexpr int kGroups = 2;
constexpr int kEnvs = kGroups * kRows;               // 256 envs per CTA
constexpr int kThreads = kEnvs + 32;                 // 8 env warps + the issuer warp
constexpr int kKA = 16 + kH;                         // 144: A-tile width
constexpr int kSboA = (kKA / 8) * 128;               // 2304
constexpr int kPassN = 128;                          // gate columns per pass: 4 gates x 32 units
constexpr int kUnitsPerPass = 32;
constexpr int kPasses = kH / kUnitsPerPass;          // 4
constexpr int kWgBytes = kPassN * kKA * 2;           // 36864 per pass tile
constexpr int kABytes = kRows * kKA * 2;             // 36864
constexpr int kW1Bytes = 64 * kKA * 2;               // 18432
constexpr int kK2 = mlptc::kK2, kSbo2 = mlptc::kSbo2;
constexpr int kW2Bytes = mlptc::kW2Bytes, kW3Bytes = mlptc::kW3Bytes, kA2Bytes = mlptc::kA2Bytes;
// global weight image: 4 gate tiles, W1, W2, W3 (bytes), then mean[16], inv_std[16] (floats)
constexpr int kImgGate = 0, kImgW1 = kPasses * kWgBytes, kImgW2 = kImgW1 + kW1Bytes, kImgW3 = kImgW2 + kW2Bytes;
constexpr int kImgNorm = kImgW3 + kW3Bytes, kImgBytes = kImgNorm + 128;
// shared memory: wg[2] | A[2] | W1 W2 W3 | A2[2] | 14 mbarriers + TMEM slot (128 B) | mean / inv_std (128 B)
constexpr int kOffA = 2 * kWgBytes, kOffW1 = kOffA + kGroups * kABytes, kOffA2 = kOffW1 + kW1Bytes + kW2Bytes + kW3Bytes;
constexpr int kOffBars = kOffA2 + kGroups * kA2Bytes, kSmemBytes = kOffBars + 128 + 128;
constexpr int kNumBars = 14;
constexpr int kTmemCols = 512, kColGates = 0, kColCell = 256;                     // group g: + 128 g
constexpr int kColHeadOut = 64;                                                   // the two actions, inside the group's gate columns

__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
// ties the 16 registers of an earlier tcgen05.ld to this point of the instruction stream: arithmetic on them cannot be
// scheduled above the tcgen05.wait::ld that precedes this call
__device__ __forceinline__ void tmem_ld_fence16(uint32_t (&r)[16]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&f)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])),
                    "r"(__float_as_uint(f[3])), "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])),
                    "r"(__float_as_uint(f[7])), "r"(__float_as_uint(f[8])), "r"(__float_as_uint(f[9])), "r"(__float_as_uint(f[10])),
                    "r"(__float_as_uint(f[11])), "r"(__float_as_uint(f[12])), "r"(__float_as_uint(f[13])), "r"(__float_as_uint(f[14])),
                    "r"(__float_as_uint(f[15])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 1-D TMA bulk load global -> shared::cta, completion counted in bytes on `mbar` (SASS: UBLKCP + SYNCS.ARRIVE.TRANS64)
__device__ __forceinline__ void tma_load_1d(uint32_t dst_smem, const void* gsrc, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst_smem), "l"(__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t mbar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(mbar) : "memory"); }
// One lane of a converged warp.  Evaluated at every issue site (as CUTLASS does), not kept in a register: with a plain
// per-thread predicate (`threadIdx.x == 0`) ptxas if-converted the bulk copy into a lane-predicated UBLKCP and
// rollout_kernel<*, 3, true> died with "illegal instruction" on the B200; behind elect.sync it stays a one-lane branch.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    __syncwarp();
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// the env warps' own barrier (the issuer warp never joins it)
__device__ __forceinline__ void env_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kEnvs) : "memory"); }

#ifdef LSTM_TRACE
// debugging aid (variant builds only): clock64 timestamps of CTA 0's three actors during steps [LSTM_TRACE, LSTM_TRACE + 2)
__device__ long long g_lstm_trace[3][256][2];
__device__ int g_lstm_trace_n[3];
#define LSTM_TR(actor, tag) do { if (blockIdx.x == 0 && trace_on && tr_n < 256) { g_lstm_trace[actor][tr_n][0] = (tag); g_lstm_trace[actor][tr_n][1] = clock64(); ++tr_n; g_lstm_trace_n[actor] = tr_n; } } while (0)
#else
#define LSTM_TR(actor, tag) do { } while (0)
#endif

struct Actor {
    unsigned char* smem;        // base of the actor's shared memory (offsets above)
    const unsigned char* img;   // global weight image
    const float* norm;          // shared: mean[16], inv_std[16]
    // mbarriers, [group] where it says so.  issuer -> env threads (tcgen05.commit, count 1): bar_g gate pass complete,
    // bar_h head layer complete; TMA -> issuer (count 1 + bytes): bar_w[buffer] weight tile landed; env threads -> issuer
    // (count 128): bar_x x rows published, bar_gfree gate columns read, bar_hready h rows published, bar_a2 A2 rows published
    uint32_t bars;              // shared address of barrier 0
    uint32_t ph_g, ph_h;        // phase parities of the barriers this env thread waits on (its own group's)
    uint32_t tmem;
    int grp;                    // this thread's group (env threads)
#ifdef LSTM_TRACE
    int n_forward, tr_n;        // fire-and-forget stores only: the trace must not stall the actor it watches
#endif
    bool timed_out;

    __device__ __forceinline__ uint32_t bar_g(int g) const { return bars + 8 * g; }
    __device__ __forceinline__ uint32_t bar_h(int g) const { return bars + 8 * (2 + g); }
    __device__ __forceinline__ uint32_t bar_w(int b) const { return bars + 8 * (4 + b); }
    __device__ __forceinline__ uint32_t bar_x(int g) const { return bars + 8 * (6 + g); }
    __device__ __forceinline__ uint32_t bar_gfree(int g) const { return bars + 8 * (8 + g); }
    __device__ __forceinline__ uint32_t bar_hready(int g) const { return bars + 8 * (10 + g); }
    __device__ __forceinline__ uint32_t bar_a2(int g) const { return bars + 8 * (12 + g); }
    __device__ __forceinline__ unsigned char* a_tile(int g) const { return smem + kOffA + g * kABytes; }
    __device__ __forceinline__ unsigned char* a2_tile(int g) const { return smem + kOffA2 + g * kA2Bytes; }
    // [128 x 13] float staging tile of group g for the rollout's observation store: aliases the group's OWN A2 tile, which is
    // idle between the group's last head layer and its next one (head_epilogue rewrites the constant tail of its rows every
    // time).  Per group, because the groups are not in lockstep: one may still be in its head while the other stores.
    __device__ __forceinline__ float* obs_staging(int g) const { return reinterpret_cast<float*>(a2_tile(g)); }

    // CTA-collective (all kThreads threads).
    __device__ __forceinline__ void setup(unsigned char* smem_base, const unsigned char* image, float obs_clip) {
        smem = smem_base;
        img = image;
        uint64_t* bar_mem = reinterpret_cast<uint64_t*>(smem + kOffBars);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mem + kNumBars);
        float* norm_s = reinterpret_cast<float*>(smem + kOffBars + 128);
        bars = mlptc::smem_u32(bar_mem);
        ph_g = ph_h = 0;
        timed_out = false;
#ifdef LSTM_TRACE
        n_forward = 0;
        tr_n = 0;
#endif
        const int tid = threadIdx.x;
        grp = tid >> 7;
        if (tid == 0) {
#pragma unroll
            for (int j = 0; j < kNumBars; ++j)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bars + 8 * j), "r"(j < 6 ? 1 : kRows) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        {   // the head's three weight tiles are contiguous in both places
            const uint4* s = reinterpret_cast<const uint4*>(img + kImgW1);
            uint4* d = reinterpret_cast<uint4*>(smem + kOffW1);
            for (int j = tid; j < (kW1Bytes + kW2Bytes + kW3Bytes) / 16; j += kThreads) d[j] = __ldg(s + j);
        }
        if (tid < 32) norm_s[tid] = tid == 15 ? obs_clip : reinterpret_cast<const float*>(img + kImgNorm)[tid];   // [15] = clip of the normalised obs
        norm = norm_s;
        for (int j = tid; j < kGroups * kABytes / 16; j += kThreads) reinterpret_cast<uint4*>(smem + kOffA)[j] = make_uint4(0u, 0u, 0u, 0u);
        if (tid < 32) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(mlptc::smem_u32(tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_proxy_async_smem();
        mlptc::fence_before_sync();
        __syncthreads();
        mlptc::fence_after_sync();
        tmem = *tmem_slot;
        if (tid < kEnvs) reset_state();
    }

    // CTA-collective; everything the issuer queued has completed by now (it drains its barriers before it gets here)
    __device__ __forceinline__ void teardown() {
        mlptc::fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }

    __device__ __forceinline__ uint32_t lane_base() const { return tmem + ((uint32_t)(threadIdx.x & 96) << 16); }   // this warp's 32 TMEM lanes

    // c = 0 for this thread's env (SB3 resets the LSTM state at an episode start).  Env warps, together (tcgen05.st is
    // warp-collective): the envs of a rollout finish their episodes in lockstep.  h needs no clearing: the issuer knows
    // the schedule (every T steps) and issues x-part-only passes for the first step of an episode.
    __device__ __forceinline__ void reset_state() {
        const uint32_t cell = lane_base() + kColCell + kH * grp;
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0.f;
#pragma unroll
        for (int c = 0; c < kH / 16; ++c) tmem_st16(cell + 16 * c, z);
        tmem_st_wait();
    }

    // `flag`: the caller's timed-out flag (a register copy on the hot paths; the member lives in local memory)
    static __device__ __forceinline__ void wait_on(uint32_t bar, uint32_t& phase, bool& flag) {
        uint32_t done = 0;
        unsigned spins = 0;
        while (!done && !flag) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            if (!done && ++spins > mlptc::kSpinLimit) flag = true;            // never hang the GPU: bail out, the caller reports it
        }
        phase ^= 1;
    }
    __device__ __forceinline__ void wait(uint32_t bar, uint32_t& phase) { wait_on(bar, phase, timed_out); }
    // wait, then order the tcgen05 operations that follow behind what the barrier stands for
    __device__ __forceinline__ void wait_tc(uint32_t bar, uint32_t& phase) {
        wait(bar, phase);
        mlptc::fence_after_sync();
    }

    // ---- issuer warp ---------------------------------------------------------------------------------------------------
    // The whole rollout from the tensor core's side: `n_steps` policy steps, episodes of `T` steps in lockstep.
    // The pass / group loops are unrolled so that every phase bit and descriptor has a static index and lives in a register:
    // this warp shares its scheduler with two epilogue warps, and what it executes between two MMA batches is latency the
    // second group of a pass waits for.
    __device__ __noinline__ void issuer_loop(int n_steps, int T) {
        // register copies of what lives in the (local-memory) actor object
        const uint32_t B = bars, tm = tmem;
        const unsigned char* const image = img;
        bool to = timed_out;
        auto bg = [B](int g) { return B + 8 * g; };
        auto bh = [B](int g) { return B + 8 * (2 + g); };
        auto bw = [B](int b) { return B + 8 * (4 + b); };
        auto bx = [B](int g) { return B + 8 * (6 + g); };
        auto bgfree = [B](int g) { return B + 8 * (8 + g); };
        auto bhready = [B](int g) { return B + 8 * (10 + g); };
        auto ba2 = [B](int g) { return B + 8 * (12 + g); };
        uint32_t ph_gi[kGroups] = {0, 0}, ph_w[2] = {0, 0}, ph_x[kGroups] = {0, 0}, ph_gfree[kGroups] = {0, 0};
        uint32_t ph_hready[kGroups] = {0, 0}, ph_a2[kGroups] = {0, 0};
        const uint32_t wg_s[2] = {mlptc::smem_u32(smem), mlptc::smem_u32(smem + kWgBytes)};
        const uint64_t d_wg[2] = {mlptc::smem_desc(wg_s[0], kLbo, kSboA), mlptc::smem_desc(wg_s[1], kLbo, kSboA)};
        const uint64_t d_a[kGroups] = {mlptc::smem_desc(mlptc::smem_u32(a_tile(0)), kLbo, kSboA), mlptc::smem_desc(mlptc::smem_u32(a_tile(1)), kLbo, kSboA)};
        const uint64_t d_a2[kGroups] = {mlptc::smem_desc(mlptc::smem_u32(a2_tile(0)), kLbo, kSbo2), mlptc::smem_desc(mlptc::smem_u32(a2_tile(1)), kLbo, kSbo2)};
        const uint64_t d_w1 = mlptc::smem_desc(mlptc::smem_u32(smem + kOffW1), kLbo, kSboA);
        const uint64_t d_w2 = mlptc::smem_desc(mlptc::smem_u32(smem + kOffW1 + kW1Bytes), kLbo, kSbo2);
        const uint64_t d_w3 = mlptc::smem_desc(mlptc::smem_u32(smem + kOffW1 + kW1Bytes + kW2Bytes), kLbo, kSbo2);
        const uint32_t idesc_g = mlptc::instr_desc(kRows, kPassN), idesc_h = mlptc::instr_desc(kRows, 64), idesc_o = mlptc::instr_desc(kRows, mlptc::kN3);
        if (elect_one()) {                                                     // passes 0 and 1 of the first step
            tma_load_1d(wg_s[0], image + kImgGate, kWgBytes, bw(0));
            tma_load_1d(wg_s[1], image + kImgGate + kWgBytes, kWgBytes, bw(1));
        }
        __syncwarp();
        int t = 0;                                                             // step within the episode
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
#ifdef LSTM_TRACE
            const bool trace_on = threadIdx.x == kEnvs && s >= LSTM_TRACE && s < LSTM_TRACE + 2;
#endif
            const int kend = t == 0 ? 1 : kKA / 16;                            // h_{t-1} = 0 at an episode start: x-part only
#pragma unroll
            for (int p = 0; p < kPasses; ++p) {
                const int b = p & 1;
                wait_on(bw(b), ph_w[b], to);                                       // tile p is in weight buffer b
#pragma unroll
                for (int g = 0; g < kGroups; ++g) {
                    // pass 0 needs the group's x rows (which also says: its head accumulators of the last step are read);
                    // later passes need the group's epilogue of the previous pass to have read the gate columns
                    if (p == 0) wait_on(bx(g), ph_x[g], to);
                    else wait_on(bgfree(g), ph_gfree[g], to);
                    mlptc::fence_after_sync();
                    LSTM_TR(2, 100 + 10 * p + g);
                    if (elect_one()) {
                        if (kend == 1) mlptc::umma_batch<1>(d_a[g], d_wg[b], idesc_g, tm + kColGates + kPassN * g);
                        else mlptc::umma_batch<kKA / 16>(d_a[g], d_wg[b], idesc_g, tm + kColGates + kPassN * g);
                        mlptc::umma_commit(bg(g));
                    }
                    __syncwarp();
                    LSTM_TR(2, 105 + 10 * p + g);
                }
                // both groups' pass p done (in-order pipe: the second commit covers the first): weight buffer b is free
                wait_on(bg(0), ph_gi[0], to);
                LSTM_TR(2, 140 + p);
                wait_on(bg(1), ph_gi[1], to);
                LSTM_TR(2, 150 + p);
                if (elect_one()) tma_load_1d(wg_s[b], image + kImgGate + ((p + 2) % kPasses) * kWgBytes, kWgBytes, bw(b));
                __syncwarp();
            }
            // head: three small layers per group, interleaved
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                wait_on(bhready(g), ph_hready[g], to);
                    mlptc::fence_after_sync();                          // h_t rows are in the A tile; gate columns read
                LSTM_TR(2, 160 + g);
                if (elect_one()) {
                    mlptc::umma_batch<kKA / 16>(d_a[g], d_w1, idesc_h, tm + kColGates + kPassN * g);
                    mlptc::umma_commit(bh(g));
                }
                __syncwarp();
            }
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                wait_on(ba2(g), ph_a2[g], to);
                    mlptc::fence_after_sync();
                if (elect_one()) {
                    mlptc::umma_batch<kK2 / 16>(d_a2[g], d_w2, idesc_h, tm + kColGates + kPassN * g);
                    mlptc::umma_commit(bh(g));
                }
                __syncwarp();
            }
#pragma unroll
            for (int g = 0; g < kGroups; ++g) {
                wait_on(ba2(g), ph_a2[g], to);
                    mlptc::fence_after_sync();
                if (elect_one()) {
                    mlptc::umma_batch<kK2 / 16>(d_a2[g], d_w3, idesc_o, tm + kColGates + kPassN * g + kColHeadOut);
                    mlptc::umma_commit(bh(g));
                }
                __syncwarp();
            }
            t = t + 1 == T ? 0 : t + 1;
        }
        wait_on(bw(0), ph_w[0], to);                                               // the two tiles requested in the last passes
        wait_on(bw(1), ph_w[1], to);
        timed_out = to;
    }

    // ---- env warps ----------------------------------------------------------------------------------------------------
    // gates of hidden units 32 p .. 32 p + 31 are in the group's gate columns as {i | f | g | o} x 32: update c, return h as bf16 pairs
    __device__ __forceinline__ void gate_epilogue(uint32_t gates, uint32_t cell, uint32_t (&hp)[16]) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t gi[16], gf[16], gg[16], go[16], cc[16];
            mlptc::tmem_ld16(gates + 0 * kUnitsPerPass + 16 * half, gi);
            mlptc::tmem_ld16(gates + 1 * kUnitsPerPass + 16 * half, gf);
            mlptc::tmem_ld16(gates + 2 * kUnitsPerPass + 16 * half, gg);
            mlptc::tmem_ld16(gates + 3 * kUnitsPerPass + 16 * half, go);
            mlptc::tmem_ld16(cell + 16 * half, cc);
            mlptc::tmem_ld_wait();
            tmem_ld_fence16(gi);
            tmem_ld_fence16(gf);
            tmem_ld_fence16(gg);
            tmem_ld_fence16(go);
            tmem_ld_fence16(cc);
            float c[16], h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                c[j] = fmaf(sigmoid_approx(__uint_as_float(gf[j])), __uint_as_float(cc[j]),
                            sigmoid_approx(__uint_as_float(gi[j])) * tanh_approx(__uint_as_float(gg[j])));
                h[j] = sigmoid_approx(__uint_as_float(go[j])) * tanh_approx(c[j]);
            }
            tmem_st16(cell + 16 * half, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) hp[8 * half + j] = mlptc::pack_bf16(h[2 * j], h[2 * j + 1]);
        }
        tmem_st_wait();
    }

    // 64 head accumulators -> ReLU -> bf16 -> this thread's A2 row (constant tail included: the rollout's observation
    // staging tile aliases the A2 tiles between steps)
    __device__ __forceinline__ void head_epilogue(uint32_t acc, unsigned char* row) {
        uint32_t r[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) mlptc::tmem_ld16(acc + c * 16, r[c]);
        mlptc::tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            tmem_ld_fence16(r[c]);
            float f[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r[c][j]);
            *reinterpret_cast<uint4*>(row + (2 * c) * kLbo) = make_uint4(mlptc::relu_pack_bf16(f[0], f[1]), mlptc::relu_pack_bf16(f[2], f[3]),
                                                                         mlptc::relu_pack_bf16(f[4], f[5]), mlptc::relu_pack_bf16(f[6], f[7]));
            *reinterpret_cast<uint4*>(row + (2 * c + 1) * kLbo) = make_uint4(mlptc::relu_pack_bf16(f[8], f[9]), mlptc::relu_pack_bf16(f[10], f[11]),
                                                                             mlptc::relu_pack_bf16(f[12], f[13]), mlptc::relu_pack_bf16(f[14], f[15]));
        }
        *reinterpret_cast<uint4*>(row + 8 * kLbo) = make_uint4(0x00003F80u, 0u, 0u, 0u);          // column 64 = bf16(1.0): the bias column
        *reinterpret_cast<uint4*>(row + 9 * kLbo) = make_uint4(0u, 0u, 0u, 0u);
    }

    // publish this thread's shared-memory rows (generic proxy -> async proxy) and its finished TMEM reads, then tell the issuer
    __device__ __forceinline__ void publish(uint32_t bar) {
        mlptc::fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(bar);
    }

    // One policy step on this thread's observation; collective over the 128 env threads of a group.  NOT inlined: with so
    // few warps per scheduler there is nobody to hide an instruction-cache miss behind, and the first, fully inlined and
    // 4x-unrolled form (19 k SASS instructions, 300 KB) spent 40 % of its stall samples in `no_instruction`.
    __device__ __noinline__ float2 forward(const float* o) {
        const int m = threadIdx.x & (kRows - 1);
        // this object lives in local memory (its address is taken by this call) and every asm below clobbers memory: read
        // what the step needs into registers once instead of once per use
        const int g = grp;
        const uint32_t b_g = bar_g(g), b_h = bar_h(g), b_x = bar_x(g), b_gfree = bar_gfree(g), b_hready = bar_hready(g), b_a2 = bar_a2(g);
        const uint32_t gates = lane_base() + kColGates + kPassN * g, cell0 = lane_base() + kColCell + kH * g;
        unsigned char* const row2 = a2_tile(g) + (m >> 3) * kSbo2 + (m & 7) * 16;
        uint32_t phg = ph_g, phh = ph_h;
        bool to = timed_out;
        const float* const nrm = norm;
#ifdef LSTM_TRACE
        const bool trace_on = m == 0 && n_forward >= LSTM_TRACE && n_forward < LSTM_TRACE + 2;
        ++n_forward;
#endif
        LSTM_TR(grp, 1);
        float x[16];
        const float clip = nrm[15];
#pragma unroll
        for (int i = 0; i < kIn; ++i) x[i] = fminf(fmaxf((o[i] - nrm[i]) * nrm[16 + i], -clip), clip);
        x[13] = 1.0f;
        x[14] = 0.f;
        x[15] = 0.f;
        unsigned char* const row = a_tile(g) + (m >> 3) * kSboA + (m & 7) * 16;
        *reinterpret_cast<uint4*>(row) = make_uint4(mlptc::pack_bf16(x[0], x[1]), mlptc::pack_bf16(x[2], x[3]), mlptc::pack_bf16(x[4], x[5]), mlptc::pack_bf16(x[6], x[7]));
        *reinterpret_cast<uint4*>(row + kLbo) = make_uint4(mlptc::pack_bf16(x[8], x[9]), mlptc::pack_bf16(x[10], x[11]), mlptc::pack_bf16(x[12], x[13]), mlptc::pack_bf16(x[14], x[15]));
        publish(b_x);
        LSTM_TR(grp, 2);
        uint32_t hp[kPasses][16];                                              // h_t as bf16 pairs, held until every pass has read h_{t-1}
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
            wait_on(b_g, phg, to);                                             // pass p is in the group's gate columns
            mlptc::fence_after_sync();
            LSTM_TR(grp, 10 + p);
            gate_epilogue(gates, cell0 + kUnitsPerPass * p, hp[p]);
            LSTM_TR(grp, 20 + p);
            if (p + 1 < kPasses) {                                             // pass p + 1 reuses them
                mlptc::fence_before_sync();
                mbar_arrive(b_gfree);
            }
        }
        // the last pass's MMAs are done (its accumulator was just read): h_t may overwrite h_{t-1} in the A tile
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(row + (2 + 4 * p + q) * kLbo) = make_uint4(hp[p][4 * q], hp[p][4 * q + 1], hp[p][4 * q + 2], hp[p][4 * q + 3]);
        }
        publish(b_hready);
        LSTM_TR(grp, 30);
        wait_on(b_h, phh, to);
        mlptc::fence_after_sync();
        LSTM_TR(grp, 31);
        head_epilogue(gates, row2);
        publish(b_a2);
        LSTM_TR(grp, 32);
        wait_on(b_h, phh, to);
        mlptc::fence_after_sync();
        LSTM_TR(grp, 33);
        head_epilogue(gates, row2);
        publish(b_a2);
        LSTM_TR(grp, 34);
        wait_on(b_h, phh, to);
        mlptc::fence_after_sync();
        LSTM_TR(grp, 35);
        uint32_t r0, r1;
        mlptc::tmem_ld2(gates + kColHeadOut, r0, r1);
        mlptc::tmem_ld_wait();
        ph_g = phg;
        ph_h = phh;
        timed_out = to;
        return make_float2(__uint_as_float(r0), __uint_as_float(r1));         // action means; the caller squashes them
    }
};

}  // namespace lstmtc
}  // namespace cantor
