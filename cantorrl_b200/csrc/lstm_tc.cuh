// Recurrent actor for the rollout kernel on the tensor cores: the policy the reference actually trained,
//   LSTM(13 -> 128) -> ReLU MLP(128 -> 64 -> 64) -> Linear(64 -> 2)       (quantconnect/model_wrapper.py:167-204,
//   shapes in quantconnect/model_files/policy_weights.pth; SB3 RecurrentPPO "MlpLstmPolicy", train_ppo_v2.py:46-47, 220-228)
// evaluated with tcgen05.mma (bf16 operands, float32 accumulation in tensor memory).
//
// One CTA = 256 envs in TWO GROUPS of 128 plus an issuer warp (288 threads).  Warps 0-3 own group 0, warps 4-7 group 1:
// thread m of a group owns env m, TMEM lane m and row m of the group's A tile.  Warp 8 is the ISSUER: it alone talks to
// the tensor core and the TMA unit.  The two sides meet only on mbarriers (env threads arrive with count 128 after
// publishing their shared-memory rows / finishing their TMEM reads; the issuer signals with tcgen05.commit); there is no
// __syncthreads in a step.
//
// The groups run HALF A STEP APART: the issuer's program is a fixed interleaving of one group's eight gate products with
// the other group's three head layers, so while one group is in its SFU-bound gate epilogues the other is in its head
// chain / env step, where the SFU and the tensor pipe would otherwise idle (round 1 ran the groups side by side: both
// fought for the SFU in the gate phase and both left it idle afterwards).
//
//   A tile [128 x 144] bf16 per group = { normalised obs (13), 1.0, 0, 0 | h (128) }   (biases ride on the ones column)
//   gates: EIGHT half-passes of 16 hidden units each, D[128 x 64] = A * Wg_q^T with Wg_q rows = {i, f, g, o} x 16 units, into
//          the two 64-column halves of the group's 128 gate columns in turn: the tensor core fills one half while the
//          group's threads work on the other.  A thread hands a half back as soon as its 64 + 16 values are in REGISTERS
//          (not after the arithmetic), which is what hides the next product's round trip behind the epilogue.
//          The 147 KB of gate weights do not fit next to the rest, so the 18 KB tile of half-pass q + 3 is streamed L2 ->
//          shared by ONE 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) into a ring of four buffers.
//   epilogue of a half-pass: one tcgen05.ld.x64 + one .x16 (cell state) of the thread's own lane, sigmoid / tanh on the SFU
//          (tanh.approx; the i / f / o rows are pre-halved by pack_lstm so that sigmoid(x) = 0.5 + 0.5 tanh(acc)), the cell
//          state c in float32 in 128 TMEM columns per group.  h_t is held back in registers as bf16 (64 registers) until
//          the group's last product has read h_{t-1}, then written over it in the A tile -- one A tile per group.
//   head:  D[128 x 64] = A * W1^T (K = 144; the x-columns weigh 0, the ones column carries the bias), ReLU,
//          D = A2 * W2^T, ReLU, D[128 x 16] = A2 * W3^T; accumulators reuse the group's gate columns.
//   episode starts (h = c = 0; the envs of a CTA reset in lockstep, every T steps): the products are x-part only.
// TMEM (all 512 columns, one CTA per SM): [0, 128) / [128, 256) gate halves + head accumulators of group 0 / 1, [256, 384) /
// [384, 512) cell states.  Weight images arrive pre-arranged in the canonical K-major no-swizzle core-matrix layout
// (cantorrl_b200/rollout.py: pack_lstm), so a tile is one contiguous copy.
#pragma once
#include "mlp_tc.cuh"

#ifndef CANTOR_LSTM_HEAD_FIRST
#define CANTOR_LSTM_HEAD_FIRST 0         // issuer: a head layer that is ready when its slot starts goes before the slot's gate product
#endif                                   // (measured: 64.1 -> 67.5 ms -- it delays the gate product the other group's epilogue chain waits for)
#ifndef CANTOR_LSTM_SLOT0_REFILL_LATE
#define CANTOR_LSTM_SLOT0_REFILL_LATE 0  // issuer: slot 0 requests the next weight tile after its gate product, not before
#endif                                   // (measured: 64.1 -> 67.6 ms, both: 68.7 -- the tile then lands late for slot 1)
#ifndef CANTOR_LSTM_FULL_LOAD_PASSES
#define CANTOR_LSTM_FULL_LOAD_PASSES 0   // the first n passes of a step read all 128 gate columns before any arithmetic (the h_t backlog is
#endif                                   // still small enough for 128 + 16 input registers); the others release after the second half's load.
                                         // Measured (2^20 envs x 252 steps): 0 -> 65.4 ms, 1 -> 65.5, 2 -> 70.5, 3 -> 74.0: at the 168-register
                                         // cap of a 9-warp CTA the up-front loads spill inside the epilogue, which costs more than they hide

namespace cantor {
namespace lstmtc {

using mlptc::kLbo;
constexpr int kRows = 128, kIn = 13, kH = 128;
constexpr int kGroups = 2;
constexpr int kEnvs = kGroups * kRows;               // 256 envs per CTA
constexpr int kThreads = kEnvs + 32;                 // 8 env warps + the issuer warp
constexpr int kKA = 16 + kH;                         // 144: A-tile width
constexpr int kSboA = (kKA / 8) * 128;               // 2304
constexpr int kPassN = 128;                          // gate columns per pass: two halves of {i, f, g, o} x 16 units
constexpr int kHalfN = 64, kUnitsPerHalf = 16;
constexpr int kUnitsPerPass = 32;
constexpr int kPasses = kH / kUnitsPerPass;          // 4
constexpr int kWgBufs = 2;                           // ring of weight-tile buffers
constexpr int kWgBytes = kPassN * kKA * 2;           // 36864 per pass tile (two 64-row halves, contiguous)
constexpr int kABytes = kRows * kKA * 2;             // 36864
constexpr int kW1Bytes = 64 * kKA * 2;               // 18432
constexpr int kK2 = mlptc::kK2, kSbo2 = mlptc::kSbo2;
constexpr int kW2Bytes = mlptc::kW2Bytes, kW3Bytes = mlptc::kW3Bytes, kA2Bytes = mlptc::kA2Bytes;
// global weight image: 4 gate tiles (each two 64-row halves), W1, W2, W3 (bytes), then mean[16], inv_std[16] (floats)
constexpr int kImgGate = 0, kImgW1 = kPasses * kWgBytes, kImgW2 = kImgW1 + kW1Bytes, kImgW3 = kImgW2 + kW2Bytes;
constexpr int kImgNorm = kImgW3 + kW3Bytes, kImgBytes = kImgNorm + 128;
// shared memory: wg[2] | A[2] | W1 W2 W3 | A2[2] | 16 mbarriers + TMEM slot (256 B) | mean / inv_std (128 B)
constexpr int kOffA = kWgBufs * kWgBytes, kOffW1 = kOffA + kGroups * kABytes, kOffA2 = kOffW1 + kW1Bytes + kW2Bytes + kW3Bytes;
constexpr int kOffBars = kOffA2 + kGroups * kA2Bytes, kSmemBytes = kOffBars + 256 + 128;
constexpr int kNumBars = 16, kNumBarsCount1 = 8;
constexpr int kGroupCols = 128;                                                   // TMEM columns of a group's gate halves / head accumulators
constexpr int kTmemCols = 512, kColGates = 0, kColCell = 256;                     // group g: + 128 g
constexpr int kColHeadOut = 64;                                                   // the two actions, inside the group's gate columns

__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// sigmoid(2 xh) = 0.5 + 0.5 tanh(xh): the i / f / o gate rows of the weight image are PRE-HALVED by pack_lstm (exact in bf16),
// so the accumulator already holds xh = x / 2
__device__ __forceinline__ float sigmoid_of_half(float xh) { return fmaf(0.5f, tanh_approx(xh), 0.5f); }
// ties the 16 registers of an earlier tcgen05.ld to this point of the instruction stream: arithmetic on them cannot be
// scheduled above the tcgen05.wait::ld that precedes this call
__device__ __forceinline__ void tmem_ld_fence16(uint32_t (&r)[16]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                      "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]) :: "memory");
}
// 64 consecutive TMEM columns of this thread's lane in ONE instruction (ptxas cannot spread it through the arithmetic that follows)
__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&r)[64]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_fence64(uint32_t (&r)[64]) {
    asm volatile("" : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31]), "+r"(r[32]), "+r"(r[33]), "+r"(r[34]), "+r"(r[35]), "+r"(r[36]), "+r"(r[37]), "+r"(r[38]), "+r"(r[39]), "+r"(r[40]), "+r"(r[41]), "+r"(r[42]), "+r"(r[43]), "+r"(r[44]), "+r"(r[45]), "+r"(r[46]), "+r"(r[47]), "+r"(r[48]), "+r"(r[49]), "+r"(r[50]), "+r"(r[51]), "+r"(r[52]), "+r"(r[53]), "+r"(r[54]), "+r"(r[55]), "+r"(r[56]), "+r"(r[57]), "+r"(r[58]), "+r"(r[59]), "+r"(r[60]), "+r"(r[61]), "+r"(r[62]), "+r"(r[63]) :: "memory");
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
// one group's 128 env threads (barrier ids 2 / 3)
__device__ __forceinline__ void group_sync(int grp) { asm volatile("bar.sync %0, %1;" :: "r"(2 + grp), "n"(kRows) : "memory"); }

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
    // mbarriers.  issuer -> env threads (tcgen05.commit, count 1): bar_g[group] gate product complete, bar_h[group] head
    // layer complete; TMA -> issuer (count 1 + bytes): bar_w[buffer] weight tile landed; tensor pipe -> issuer (commit, count 1):
    // bar_wfree[buffer] the product that read the tile is complete; env threads -> issuer (count 128): bar_x x rows published,
    // bar_gfree[group] gate columns read into registers, bar_hready h rows published, bar_a2 A2 rows published
    uint32_t bars;              // shared address of barrier 0
    uint32_t ph_h;              // phase parity of this env thread's bar_h (bar_g completes four times per step: always 0 at entry)
    uint32_t tmem;
    int grp;                    // this thread's group (env threads)
#ifdef LSTM_TRACE
    int n_forward, tr_n;        // fire-and-forget stores only: the trace must not stall the actor it watches
#endif
    bool timed_out;

    static constexpr int kBarG = 0, kBarH = 2, kBarW = 4, kBarWfree = 6, kBarX = 8, kBarGfree = 10, kBarHready = 12, kBarA2 = 14;
    __device__ __forceinline__ uint32_t bar_g(int g) const { return bars + 8 * (kBarG + g); }
    __device__ __forceinline__ uint32_t bar_h(int g) const { return bars + 8 * (kBarH + g); }
    __device__ __forceinline__ uint32_t bar_x(int g) const { return bars + 8 * (kBarX + g); }
    __device__ __forceinline__ uint32_t bar_gfree(int g) const { return bars + 8 * (kBarGfree + g); }
    __device__ __forceinline__ uint32_t bar_hready(int g) const { return bars + 8 * (kBarHready + g); }
    __device__ __forceinline__ uint32_t bar_a2(int g) const { return bars + 8 * (kBarA2 + g); }
    __device__ __forceinline__ unsigned char* a_tile(int g) const { return smem + kOffA + g * kABytes; }
    __device__ __forceinline__ unsigned char* a2_tile(int g) const { return smem + kOffA2 + g * kA2Bytes; }
    // [128 x 13] float staging tile of group g for the rollout's observation store: aliases the group's OWN A2 tile, which is
    // idle between the group's last head layer and its next one (head_epilogue rewrites the constant tail of its rows every
    // time).  Per group, because the groups are half a step apart: one is in its head while the other stores.
    __device__ __forceinline__ float* obs_staging(int g) const { return reinterpret_cast<float*>(a2_tile(g)); }

    // CTA-collective (all kThreads threads).
    __device__ __forceinline__ void setup(unsigned char* smem_base, const unsigned char* image, float obs_clip) {
        smem = smem_base;
        img = image;
        uint64_t* bar_mem = reinterpret_cast<uint64_t*>(smem + kOffBars);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mem + kNumBars);
        float* norm_s = reinterpret_cast<float*>(smem + kOffBars + 256);
        bars = mlptc::smem_u32(bar_mem);
        ph_h = 0;
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
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bars + 8 * j), "r"(j < kNumBarsCount1 ? 1 : kRows) : "memory");
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
    // the schedule (every T steps) and issues x-part-only products for the first step of an episode.
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

    // ---- issuer warp ---------------------------------------------------------------------------------------------------
    // Everything the issuer keeps between two issue sites: phase parities and descriptors, all statically indexed (the loops
    // below are unrolled), so that they live in registers -- this warp shares its scheduler with two epilogue warps, and what
    // it executes between two products is latency somebody waits for.
    struct Issuer {
        uint32_t B, tm;
        const unsigned char* image;
        uint32_t wg_s[kWgBufs];
        uint64_t d_wg[kWgBufs], d_a[kGroups], d_a2[kGroups], d_w1, d_w2, d_w3;
        uint32_t ph_w[kWgBufs], ph_wfree[kWgBufs], ph_x[kGroups], ph_gfree[kGroups], ph_hready[kGroups], ph_a2[kGroups];
        bool to;
#ifdef LSTM_TRACE
        bool trace_on;
        int tr_n;
#endif
    };

    // One HALF-CYCLE of the issuer's program: the four gate products of group GA (`gates_on`), interleaved with the three head
    // layers of the other group (`head_on`), whose gate products were issued in the previous half-cycle.  Every wait refers to
    // something issued earlier in this fixed order, so the order cannot deadlock.
    //   slot p:  refill                 the buffer of product p - 1 (complete by now) with the tile of product p + 1
    //            gate product p of GA   <- weight tile p landed; x rows published (p = 0) / gate columns of p - 1 read (p >= 1)
    //            head layer p + 1 of GB <- h rows published (L1) / A2 rows published (L2, L3)            (p < 3; first if ready first)
    template <int GA>
    __device__ __forceinline__ void half_cycle(Issuer& I, bool gates_on, bool head_on, bool x_only, bool first) {
        constexpr int GB = GA ^ 1;
        constexpr uint32_t idesc_g = mlptc::instr_desc(kRows, kPassN), idesc_h = mlptc::instr_desc(kRows, 64), idesc_o = mlptc::instr_desc(kRows, mlptc::kN3);
        const uint32_t B = I.B;
#ifdef LSTM_TRACE
        const bool trace_on = I.trace_on;
        int& tr_n = I.tr_n;
#endif
        auto ready = [](uint32_t bar, uint32_t phase) {                                              // non-blocking look at a barrier
            uint32_t done;
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            return done != 0;
        };
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
            const int buf = p % kWgBufs;
            auto refill = [&]() {                       // the tile of product p + 1 into the buffer of product p - 1, once that product is complete
                if (gates_on && !(first && p == 0)) {
                    const int r = buf ^ 1;
                    wait_on(B + 8 * (kBarWfree + r), I.ph_wfree[r], I.to);
                    if (elect_one()) tma_load_1d(I.wg_s[r], I.image + kImgGate + ((p + 1) % kPasses) * kWgBytes, kWgBytes, B + 8 * (kBarW + r));
                    __syncwarp();
                }
            };
            auto head_layer = [&]() {
                mlptc::fence_after_sync();
                LSTM_TR(2, 120 + p);
                if (elect_one()) {
                    const uint32_t d = I.tm + kColGates + kGroupCols * GB;
                    if (p == 0) mlptc::umma_batch<kKA / 16>(I.d_a[GB], I.d_w1, idesc_h, d);
                    else if (p == 1) mlptc::umma_batch<kK2 / 16>(I.d_a2[GB], I.d_w2, idesc_h, d);
                    else mlptc::umma_batch<kK2 / 16>(I.d_a2[GB], I.d_w3, idesc_o, d + kColHeadOut);
                    mlptc::umma_commit(B + 8 * (kBarH + GB));
                }
                __syncwarp();
            };
            const bool head_slot = head_on && p < 3;
            const uint32_t head_bar = B + 8 * ((p == 0 ? kBarHready : kBarA2) + GB);
            uint32_t& head_phase = p == 0 ? I.ph_hready[GB] : I.ph_a2[GB];
            // the refill first, so that the copy has this whole slot to land
            if (p != 0 || !CANTOR_LSTM_SLOT0_REFILL_LATE) refill();
            bool head_done = !head_slot;
            if (CANTOR_LSTM_HEAD_FIRST && head_slot && ready(head_bar, head_phase)) {                   // the short head layer goes first when it is ready first
                head_phase ^= 1;
                head_layer();
                head_done = true;
            }
            if (gates_on) {
                wait_on(B + 8 * (kBarW + buf), I.ph_w[buf], I.to);
                if (p == 0) wait_on(B + 8 * (kBarX + GA), I.ph_x[GA], I.to);
                else wait_on(B + 8 * (kBarGfree + GA), I.ph_gfree[GA], I.to);
                mlptc::fence_after_sync();
                LSTM_TR(2, 100 + 2 * p);
                if (elect_one()) {
                    const uint32_t d = I.tm + kColGates + kGroupCols * GA;
                    if (x_only) mlptc::umma_batch<1>(I.d_a[GA], I.d_wg[buf], idesc_g, d);
                    else mlptc::umma_batch<kKA / 16>(I.d_a[GA], I.d_wg[buf], idesc_g, d);
                    mlptc::umma_commit(B + 8 * (kBarG + GA));
                    mlptc::umma_commit(B + 8 * (kBarWfree + buf));
                }
                __syncwarp();
                LSTM_TR(2, 101 + 2 * p);
            }
            if (!head_done) {
                wait_on(head_bar, head_phase, I.to);                           // L1: h_t rows in the A tile, gate columns read; L2 / L3: A2 rows
                head_layer();
            }
            if (p == 0 && CANTOR_LSTM_SLOT0_REFILL_LATE) refill();
        }
    }

    // The whole rollout from the tensor core's side: `n_steps` policy steps, episodes of `T` steps in lockstep.
    __device__ __noinline__ void issuer_loop(int n_steps, int T) {
        Issuer I;
        I.B = bars;
        I.tm = tmem;
        I.image = img;
        I.to = timed_out;
#pragma unroll
        for (int b = 0; b < kWgBufs; ++b) {
            I.wg_s[b] = mlptc::smem_u32(smem + b * kWgBytes);
            I.d_wg[b] = mlptc::smem_desc(I.wg_s[b], kLbo, kSboA);
            I.ph_w[b] = I.ph_wfree[b] = 0;
        }
#pragma unroll
        for (int g = 0; g < kGroups; ++g) {
            I.d_a[g] = mlptc::smem_desc(mlptc::smem_u32(a_tile(g)), kLbo, kSboA);
            I.d_a2[g] = mlptc::smem_desc(mlptc::smem_u32(a2_tile(g)), kLbo, kSbo2);
            I.ph_x[g] = I.ph_gfree[g] = I.ph_hready[g] = I.ph_a2[g] = 0;
        }
        I.d_w1 = mlptc::smem_desc(mlptc::smem_u32(smem + kOffW1), kLbo, kSboA);
        I.d_w2 = mlptc::smem_desc(mlptc::smem_u32(smem + kOffW1 + kW1Bytes), kLbo, kSbo2);
        I.d_w3 = mlptc::smem_desc(mlptc::smem_u32(smem + kOffW1 + kW1Bytes + kW2Bytes), kLbo, kSbo2);
#ifdef LSTM_TRACE
        I.tr_n = 0;
#endif
        if (elect_one()) {                                                     // the first two tiles
#pragma unroll
            for (int b = 0; b < kWgBufs; ++b) tma_load_1d(I.wg_s[b], I.image + kImgGate + b * kWgBytes, kWgBytes, I.B + 8 * (kBarW + b));
        }
        __syncwarp();
        int t = 0;                                                             // step within the episode
#pragma unroll 1
        for (int s = 0; s <= n_steps; ++s) {
#ifdef LSTM_TRACE
            I.trace_on = threadIdx.x == kEnvs && s >= LSTM_TRACE && s < LSTM_TRACE + 2;
#endif
            const bool more = s < n_steps;
            // group 0's gates of step s | group 1's head of step s - 1;   then group 1's gates of step s | group 0's head of step s
            half_cycle<0>(I, more, s >= 1, t == 0, s == 0);
            if (more) half_cycle<1>(I, true, true, t == 0, false);
            t = t + 1 == T ? 0 : t + 1;
        }
        wait_on(I.B + 8 * (kBarW + 0), I.ph_w[0], I.to);                       // the tile requested by the last slot (tile 0), or the first two
        if (n_steps <= 0) wait_on(I.B + 8 * (kBarW + 1), I.ph_w[1], I.to);
        timed_out = I.to;
    }

    // ---- env warps ----------------------------------------------------------------------------------------------------
    // gates of hidden units 32 p .. 32 p + 31 are in the group's gate columns as two halves of {i | f | g | o} x 16: update c,
    // return h as bf16 pairs.  `release`: the columns go back to the issuer as soon as the SECOND half is in registers -- one
    // tcgen05.ld.x64, which ptxas cannot spread through the arithmetic -- so the group's next product runs on the tensor core under
    // the second half's arithmetic.
    __device__ __forceinline__ void gate_epilogue(uint32_t gates, uint32_t cell, uint32_t (&hp)[16], uint32_t b_gfree, bool release) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t ga[64], cc[16];
            tmem_ld64(gates + kHalfN * half, ga);
            if (half == 0) mlptc::tmem_ld16(cell, cc);
            mlptc::tmem_ld_wait();
            tmem_ld_fence64(ga);
            if (half == 1) {
                if (release) {                                                 // before anything else: the issuer is waiting for this
                    mlptc::fence_before_sync();
                    mbar_arrive(b_gfree);
                }
                mlptc::tmem_ld16(cell + 16, cc);
                mlptc::tmem_ld_wait();
            }
            tmem_ld_fence16(cc);
            float c[16], h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                c[j] = fmaf(sigmoid_of_half(__uint_as_float(ga[16 + j])), __uint_as_float(cc[j]),
                            sigmoid_of_half(__uint_as_float(ga[j])) * tanh_approx(__uint_as_float(ga[32 + j])));
                h[j] = sigmoid_of_half(__uint_as_float(ga[48 + j])) * tanh_approx(c[j]);
            }
            tmem_st16(cell + 16 * half, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) hp[8 * half + j] = mlptc::pack_bf16(h[2 * j], h[2 * j + 1]);
        }
    }

    // The same pass with BOTH halves read up front and handed back at once: the group's next product runs on the tensor core under the
    // whole of this pass's arithmetic.  128 + 16 registers of inputs next to the h_t backlog: for the first passes of a step only.
    __device__ __forceinline__ void gate_epilogue_all(uint32_t gates, uint32_t cell, uint32_t (&hp)[16], uint32_t b_gfree) {
        uint32_t ga[64], gb[64];
        tmem_ld64(gates, ga);
        tmem_ld64(gates + kHalfN, gb);
        mlptc::tmem_ld_wait();
        tmem_ld_fence64(ga);
        tmem_ld_fence64(gb);
        mlptc::fence_before_sync();
        mbar_arrive(b_gfree);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t cc[16];
            mlptc::tmem_ld16(cell + 16 * half, cc);
            mlptc::tmem_ld_wait();
            tmem_ld_fence16(cc);
            float c[16], h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t gi = half ? gb[j] : ga[j], gf = half ? gb[16 + j] : ga[16 + j];
                const uint32_t gg = half ? gb[32 + j] : ga[32 + j], go = half ? gb[48 + j] : ga[48 + j];
                c[j] = fmaf(sigmoid_of_half(__uint_as_float(gf)), __uint_as_float(cc[j]),
                            sigmoid_of_half(__uint_as_float(gi)) * tanh_approx(__uint_as_float(gg)));
                h[j] = sigmoid_of_half(__uint_as_float(go)) * tanh_approx(c[j]);
            }
            tmem_st16(cell + 16 * half, c);
#pragma unroll
            for (int j = 0; j < 8; ++j) hp[8 * half + j] = mlptc::pack_bf16(h[2 * j], h[2 * j + 1]);
        }
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
    // Split in two so that the caller can do work that does not depend on the action (the next path record) between them, under the
    // head's first product: forward_gates = x rows .. h_t published, forward_head = the three head layers.
    __device__ __forceinline__ float2 forward(const float* o) {
        forward_gates(o);
        return forward_head();
    }

    __device__ __noinline__ void forward_gates(const float* o) {
        const int m = threadIdx.x & (kRows - 1);
        // this object lives in local memory (its address is taken by this call) and every asm below clobbers memory: read
        // what the step needs into registers once instead of once per use
        const int g = grp;
        const uint32_t b_g = bar_g(g), b_x = bar_x(g), b_gfree = bar_gfree(g), b_hready = bar_hready(g);
        const uint32_t gates = lane_base() + kColGates + kGroupCols * g, cell0 = lane_base() + kColCell + kH * g;
        uint32_t phg = 0;
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
        uint32_t hp[kPasses][16];                                              // h_t as bf16 pairs, held until every product has read h_{t-1}
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
            wait_on(b_g, phg, to);                                             // product p is in the group's gate columns
            mlptc::fence_after_sync();
            LSTM_TR(grp, 10 + p);
            if (p < CANTOR_LSTM_FULL_LOAD_PASSES) gate_epilogue_all(gates, cell0 + kUnitsPerPass * p, hp[p], b_gfree);
            else gate_epilogue(gates, cell0 + kUnitsPerPass * p, hp[p], b_gfree, p + 1 < kPasses);
            LSTM_TR(grp, 20 + p);
        }
        tmem_st_wait();
        // the last product is done (its accumulator was just read): h_t may overwrite h_{t-1} in the A tile
#pragma unroll
        for (int p = 0; p < kPasses; ++p) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(row + (2 + 4 * p + q) * kLbo) = make_uint4(hp[p][4 * q], hp[p][4 * q + 1], hp[p][4 * q + 2], hp[p][4 * q + 3]);
        }
        publish(b_hready);
        LSTM_TR(grp, 30);
        timed_out = to;
    }

    __device__ __noinline__ float2 forward_head() {
        const int m = threadIdx.x & (kRows - 1);
        const int g = grp;
        const uint32_t b_h = bar_h(g), b_a2 = bar_a2(g);
        const uint32_t gates = lane_base() + kColGates + kGroupCols * g;
        unsigned char* const row2 = a2_tile(g) + (m >> 3) * kSbo2 + (m & 7) * 16;
        uint32_t phh = ph_h;
        bool to = timed_out;
#ifdef LSTM_TRACE
        const bool trace_on = m == 0 && n_forward > LSTM_TRACE && n_forward <= LSTM_TRACE + 2;
#endif
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
        ph_h = phh;
        timed_out = to;
        return make_float2(__uint_as_float(r0), __uint_as_float(r1));         // action means; the caller squashes them
    }
};

}  // namespace lstmtc
}  // namespace cantor
