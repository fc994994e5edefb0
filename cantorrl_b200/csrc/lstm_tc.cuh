// Recurrent actor for the rollout kernel on the tensor cores: the policy the reference actually trained,
//   LSTM(13 -> 128) -> ReLU MLP(128 -> 64 -> 64) -> Linear(64 -> 2)       (quantconnect/model_wrapper.py:167-204,
//   shapes in quantconnect/model_files/policy_weights.pth; SB3 RecurrentPPO "MlpLstmPolicy", train_ppo_v2.py:46-47, 220-228)
// evaluated for one CTA's 128 envs per env-step with tcgen05.mma (bf16 operands, float32 accumulation in tensor memory).
//
// Warp-specialised: warps 0-3 (128 threads) own the envs -- thread m owns env m, TMEM lane m and row m of every A tile --
// and warp 4 is the ISSUER: it alone talks to the tensor core and the TMA unit, so building descriptors and issuing ~60
// MMAs per env-step runs beside the epilogues instead of in front of them.  The two sides meet only on mbarriers
// (env threads arrive with count 128 after publishing their shared-memory rows / finishing their TMEM reads; the issuer
// signals with tcgen05.commit); there is no __syncthreads in the step.
//
//   A tile [128 x 144] bf16 = { normalised obs (13), 1.0, 0, 0 | h (128) }          (biases ride on the ones column)
//   gates: four passes of 32 hidden units each, D[128 x 128] = A * Wg_p^T with Wg_p rows = {i, f, g, o} x 32 units.
//          The 147 KB of gate weights do not fit next to the rest, so the 36 KB tile of pass p + 2 is streamed L2 -> shared
//          by ONE 1-D TMA bulk copy (cp.async.bulk ... mbarrier::complete_tx) into the buffer pass p just released.
//          The gate accumulators are double-buffered in TMEM: the MMAs of pass p + 1 run under the epilogue of pass p.
//   epilogue of a pass: tcgen05.ld of the thread's own lane (five loads in flight, one wait), sigmoid / tanh on the SFU
//          (tanh.approx), the cell state c in float32 in 128 TMEM columns, h -> bf16 -> the OTHER A tile (all four passes
//          still read the old h), which then feeds the MLP head as its layer-1 operand and becomes the next step's A tile
//   head:  D[128 x 64] = A * W1^T (K = 144, x-columns weigh 0), ReLU, D = A2 * W2^T, ReLU, D[128 x 16] = A2 * W3^T
//   recurrence ahead of the observation: the h-part of the next step's first two passes (K = 128 of 144) depends only on
//          h_t, so the issuer queues it right behind the head's last MMA and it runs under the env step; when the next
//          observation is ready only one K = 16 MMA per pass is left before the first epilogue.
//   episode starts (h = c = 0; the envs of a CTA reset in lockstep, every T steps): the passes are x-part only.
// TMEM (512 columns, one CTA per SM -- which the ~200 KB of shared memory imply anyway): [0, 128) and [128, 256) gate
// accumulators, [256, 384) cell state, [384, 448) head accumulators.  Weight images arrive pre-arranged in the canonical
// K-major no-swizzle core-matrix layout (cantorrl_b200/rollout.py: pack_lstm), so a tile is one contiguous copy.
#pragma once
#include "mlp_tc.cuh"

namespace cantor {
namespace lstmtc {

using mlptc::kLbo;
constexpr int kRows = 128, kIn = 13, kH = 128;
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
constexpr int kSmemBytes = 2 * kWgBytes + 2 * kABytes + kW1Bytes + kW2Bytes + kW3Bytes + kA2Bytes + 128 + 128;
constexpr int kTmemCols = 512, kColGates = 0, kColCell = 256, kColHead = 384;     // gate buffer b at kColGates + 128 b
constexpr int kThreads = kRows + 32;                                              // 4 env warps + the issuer warp

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
__device__ __forceinline__ void env_sync() { asm volatile("bar.sync 1, %0;" :: "n"(kRows) : "memory"); }

struct Actor {
    unsigned char* wg[2];       // gate-weight buffers (filled by TMA)
    unsigned char* a[2];        // A tiles; a[cur] = {x_t, h_{t-1}}
    unsigned char* w1;
    unsigned char* w2;
    unsigned char* w3;
    unsigned char* a2;
    const unsigned char* img;   // global weight image
    const float* norm;          // shared: mean[16], inv_std[16]
    // mbarriers.  issuer -> env threads (tcgen05.commit, count 1): gate buffer b complete, head layer complete;
    // TMA -> issuer (count 1 + bytes): weight buffer b landed; env threads -> issuer (count 128): x rows published,
    // gate buffer b read, h rows published (and both gate buffers read), A2 rows published
    uint32_t bar_g[2], bar_h, bar_w[2], bar_x, bar_gfree[2], bar_hready, bar_a2;
    uint32_t ph_g[2], ph_h, ph_w[2];      // phase parities of the barriers this thread waits on
    uint32_t tmem;
    int cur;
    bool timed_out;

    __device__ __forceinline__ void copy_tile(unsigned char* dst, const unsigned char* src, int bytes) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (int j = threadIdx.x; j < bytes / 16; j += kThreads) d[j] = __ldg(s + j);
    }

    // CTA-collective (all kThreads threads).
    __device__ __forceinline__ void setup(unsigned char* smem, const unsigned char* image) {
        img = image;
        wg[0] = smem;
        wg[1] = wg[0] + kWgBytes;
        a[0] = wg[1] + kWgBytes;
        a[1] = a[0] + kABytes;
        w1 = a[1] + kABytes;
        w2 = w1 + kW1Bytes;
        w3 = w2 + kW2Bytes;
        a2 = w3 + kW3Bytes;
        uint64_t* bars = reinterpret_cast<uint64_t*>(a2 + kA2Bytes);         // 10 mbarriers, then the TMEM address slot
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
        float* norm_s = reinterpret_cast<float*>(a2 + kA2Bytes + 128);
        const uint32_t b0 = mlptc::smem_u32(bars);
        bar_g[0] = b0;
        bar_g[1] = b0 + 8;
        bar_h = b0 + 16;
        bar_w[0] = b0 + 24;
        bar_w[1] = b0 + 32;
        bar_x = b0 + 40;
        bar_gfree[0] = b0 + 48;
        bar_gfree[1] = b0 + 56;
        bar_hready = b0 + 64;
        bar_a2 = b0 + 72;
        ph_g[0] = ph_g[1] = ph_h = ph_w[0] = ph_w[1] = 0;
        cur = 0;
        timed_out = false;
        const int tid = threadIdx.x;
        if (tid == 0) {
#pragma unroll
            for (int j = 0; j < 10; ++j)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(b0 + 8 * j), "r"(j < 5 ? 1 : kRows) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        copy_tile(w1, img + kImgW1, kW1Bytes + kW2Bytes + kW3Bytes);      // the head's three tiles are contiguous in both places
        if (tid < 32) norm_s[tid] = reinterpret_cast<const float*>(img + kImgNorm)[tid];
        norm = norm_s;
        // zero both A tiles (h = 0) and the constant tail of the A2 rows
        for (int j = tid; j < 2 * kABytes / 16; j += kThreads) reinterpret_cast<uint4*>(a[0])[j] = make_uint4(0u, 0u, 0u, 0u);
        if (tid < kRows) {
            unsigned char* row = a2 + (tid >> 3) * kSbo2 + (tid & 7) * 16;
            *reinterpret_cast<uint4*>(row + 8 * kLbo) = make_uint4(0x00003F80u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(row + 9 * kLbo) = make_uint4(0u, 0u, 0u, 0u);
        }
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
        if (tid < kRows) reset_state();
    }

    // CTA-collective; everything the issuer queued has completed by now (it drains its barriers before it gets here)
    __device__ __forceinline__ void teardown() {
        mlptc::fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }

    // h = c = 0 for this thread's env (SB3 resets the LSTM state at an episode start).  Env warps, together (tcgen05.st is
    // warp-collective): the envs of a rollout finish their episodes in lockstep.  The issuer knows the schedule (every
    // T steps) and issues x-part-only passes for the first step of an episode.
    __device__ __forceinline__ void reset_state() {
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16);
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0.f;
#pragma unroll
        for (int c = 0; c < kH / 16; ++c) tmem_st16(lane_addr + kColCell + 16 * c, z);
        tmem_st_wait();
        unsigned char* row = a[cur] + (m >> 3) * kSboA + (m & 7) * 16;
#pragma unroll
        for (int c = 2; c < kKA / 8; ++c) *reinterpret_cast<uint4*>(row + c * kLbo) = make_uint4(0u, 0u, 0u, 0u);
    }

    __device__ __forceinline__ void wait(uint32_t bar, uint32_t& phase) {
        uint32_t done = 0;
        unsigned spins = 0;
        while (!done && !timed_out) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar), "r"(phase) : "memory");
            if (!done && ++spins > mlptc::kSpinLimit) timed_out = true;       // never hang the GPU: bail out, the caller reports it
        }
        phase ^= 1;
    }
    // wait, then order the tcgen05 operations that follow behind what the barrier stands for
    __device__ __forceinline__ void wait_tc(uint32_t bar, uint32_t& phase) {
        wait(bar, phase);
        mlptc::fence_after_sync();
    }

    // ---- issuer warp ---------------------------------------------------------------------------------------------------
    // K-steps [k0, k1) of D[128 x n] (+)= A * B^T into TMEM column `dcol`; the first one overwrites unless `acc`.
    // a_desc / b_desc: descriptors of K-step 0; one K-step = two 128-byte core-matrix columns = 16 in the address field.
    __device__ __forceinline__ void mma(uint64_t a_desc, uint64_t b_desc, int k0, int k1, uint32_t idesc, int dcol, bool acc) {
#pragma unroll 1
        for (int k = k0; k < k1; ++k)
            mlptc::umma_bf16(tmem + dcol, a_desc + (uint64_t)(k * (2 * kLbo / 16)), b_desc + (uint64_t)(k * (2 * kLbo / 16)), idesc, (acc || k > k0) ? 1u : 0u);
    }

    // The whole rollout from the tensor core's side: `n_steps` policy steps, episodes of `T` steps in lockstep.
    __device__ __noinline__ void issuer_loop(int n_steps, int T) {
        uint32_t ph_x = 0, ph_gfree[2] = {0, 0}, ph_hready = 0, ph_a2 = 0;
        const uint32_t wg_s[2] = {mlptc::smem_u32(wg[0]), mlptc::smem_u32(wg[1])};
        const uint64_t d_wg[2] = {mlptc::smem_desc(wg_s[0], kLbo, kSboA), mlptc::smem_desc(wg_s[1], kLbo, kSboA)};
        const uint64_t d_a[2] = {mlptc::smem_desc(mlptc::smem_u32(a[0]), kLbo, kSboA), mlptc::smem_desc(mlptc::smem_u32(a[1]), kLbo, kSboA)};
        const uint64_t d_w1 = mlptc::smem_desc(mlptc::smem_u32(w1), kLbo, kSboA);
        const uint64_t d_w2 = mlptc::smem_desc(mlptc::smem_u32(w2), kLbo, kSbo2), d_w3 = mlptc::smem_desc(mlptc::smem_u32(w3), kLbo, kSbo2);
        const uint64_t d_a2 = mlptc::smem_desc(mlptc::smem_u32(a2), kLbo, kSbo2);
        const uint32_t idesc_g = mlptc::instr_desc(kRows, kPassN), idesc_h = mlptc::instr_desc(kRows, 64), idesc_o = mlptc::instr_desc(kRows, mlptc::kN3);
        if (elect_one()) {                                                     // passes 0 and 1 of the first step
            tma_load_1d(wg_s[0], img + kImgGate, kWgBytes, bar_w[0]);
            tma_load_1d(wg_s[1], img + kImgGate + kWgBytes, kWgBytes, bar_w[1]);
        }
        __syncwarp();
        int c = 0, t = 0;                                                      // A tile of this step, step within the episode
        bool pre = false;                                                      // the h-part of passes 0, 1 is already queued
#pragma unroll 1
        for (int g = 0; g < n_steps; ++g) {
            const bool hz = t == 0;                                            // h_{t-1} = 0: x-part only, overwrite
            const int kend = hz ? 1 : kKA / 16;
            const uint64_t d_cur = d_a[c], d_next = d_a[c ^ 1];
            wait_tc(bar_x, ph_x);                                              // the env threads' x rows are in a[c]
#pragma unroll 1
            for (int b = 0; b < 2; ++b) {
                if (!pre) wait(bar_w[b], ph_w[b]);                             // tiles 0, 1 (else waited for when the h-part was queued)
                if (elect_one()) {
                    mma(d_cur, d_wg[b], 0, 1, idesc_g, kColGates + kPassN * b, pre);
                    mlptc::umma_commit(bar_g[b]);
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int p = 0; p < kPasses; ++p) {
                const int b = p & 1;
                wait(bar_g[b], ph_g[b]);                                       // pass p done: weight buffer b is free
                if (elect_one()) tma_load_1d(wg_s[b], img + kImgGate + ((p + 2) % kPasses) * kWgBytes, kWgBytes, bar_w[b]);
                __syncwarp();
                if (p + 2 < kPasses) {
                    wait_tc(bar_gfree[b], ph_gfree[b]);                        // every env thread has read gate buffer b
                    wait(bar_w[b], ph_w[b]);
                    if (elect_one()) {
                        mma(d_cur, d_wg[b], 0, kend, idesc_g, kColGates + kPassN * b, false);
                        mlptc::umma_commit(bar_g[b]);
                    }
                    __syncwarp();
                }
            }
            wait_tc(bar_hready, ph_hready);                                    // h_t rows are in a[c ^ 1]; both gate buffers read
            if (elect_one()) {
                mma(d_next, d_w1, 0, kKA / 16, idesc_h, kColHead, false);
                mlptc::umma_commit(bar_h);
            }
            __syncwarp();
            wait_tc(bar_a2, ph_a2);
            if (elect_one()) {
                mma(d_a2, d_w2, 0, kK2 / 16, idesc_h, kColHead, false);
                mlptc::umma_commit(bar_h);
            }
            __syncwarp();
            wait_tc(bar_a2, ph_a2);
            if (elect_one()) {
                mma(d_a2, d_w3, 0, kK2 / 16, idesc_o, kColHead, false);
                mlptc::umma_commit(bar_h);
            }
            __syncwarp();
            t = t + 1 == T ? 0 : t + 1;
            // the recurrence of the NEXT step, ahead of its observation: h_t x (tiles 0, 1) -> gate buffers 0, 1, K-steps
            // 1 .. 8; runs under the env step and is covered by the next step's commits.  Not before an episode start.
            pre = g + 1 < n_steps && t != 0;
            if (pre) {
#pragma unroll 1
                for (int b = 0; b < 2; ++b) {
                    wait(bar_w[b], ph_w[b]);
                    if (elect_one()) mma(d_next, d_wg[b], 1, kKA / 16, idesc_g, kColGates + kPassN * b, false);
                    __syncwarp();
                }
            }
            c ^= 1;
        }
        if (!pre) {                                                            // the two tiles requested in the last passes
            wait(bar_w[0], ph_w[0]);
            wait(bar_w[1], ph_w[1]);
        }
    }

    // ---- env warps ----------------------------------------------------------------------------------------------------
    // gates of hidden units 32 p .. 32 p + 31 are in TMEM columns [gcol, gcol + 128) as {i | f | g | o} x 32: update c, write h
    __device__ __forceinline__ void gate_epilogue(int p, int gcol, unsigned char* a_next) {
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16);
        unsigned char* row = a_next + (m >> 3) * kSboA + (m & 7) * 16;
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
            uint32_t gi[16], gf[16], gg[16], go[16], cc[16];
            mlptc::tmem_ld16(lane_addr + gcol + 0 * kUnitsPerPass + 16 * half, gi);
            mlptc::tmem_ld16(lane_addr + gcol + 1 * kUnitsPerPass + 16 * half, gf);
            mlptc::tmem_ld16(lane_addr + gcol + 2 * kUnitsPerPass + 16 * half, gg);
            mlptc::tmem_ld16(lane_addr + gcol + 3 * kUnitsPerPass + 16 * half, go);
            mlptc::tmem_ld16(lane_addr + kColCell + kUnitsPerPass * p + 16 * half, cc);
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
            tmem_st16(lane_addr + kColCell + kUnitsPerPass * p + 16 * half, c);
            const int chunk = (16 + kUnitsPerPass * p + 16 * half) / 8;                         // 8 bf16 per 16-byte chunk
            *reinterpret_cast<uint4*>(row + chunk * kLbo) = make_uint4(mlptc::pack_bf16(h[0], h[1]), mlptc::pack_bf16(h[2], h[3]),
                                                                        mlptc::pack_bf16(h[4], h[5]), mlptc::pack_bf16(h[6], h[7]));
            *reinterpret_cast<uint4*>(row + (chunk + 1) * kLbo) = make_uint4(mlptc::pack_bf16(h[8], h[9]), mlptc::pack_bf16(h[10], h[11]),
                                                                              mlptc::pack_bf16(h[12], h[13]), mlptc::pack_bf16(h[14], h[15]));
        }
        tmem_st_wait();
    }

    __device__ __forceinline__ void head_epilogue() {                       // 64 head accumulators -> ReLU -> bf16 -> A2 row
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16) + kColHead;
        unsigned char* row = a2 + (m >> 3) * kSbo2 + (m & 7) * 16;
        uint32_t r[4][16];
#pragma unroll
        for (int c = 0; c < 4; ++c) mlptc::tmem_ld16(lane_addr + c * 16, r[c]);
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
    }

    // publish this thread's shared-memory rows (generic proxy -> async proxy) and its finished TMEM reads, then tell the issuer
    __device__ __forceinline__ void publish(uint32_t bar) {
        mlptc::fence_before_sync();
        fence_proxy_async_smem();
        mbar_arrive(bar);
    }

    // One policy step on this thread's observation; collective over the 128 env threads.  Deliberately NOT inlined and
    // with its pass loops rolled: with one warp per scheduler there is nobody to hide an instruction-cache miss behind, and
    // the fully unrolled form (19 k SASS instructions, 300 KB) spent 40 % of its stall samples in `no_instruction`.
    __device__ __noinline__ float2 forward(const float* o) {
        const int m = threadIdx.x;
        unsigned char* a_cur = a[cur];
        unsigned char* a_next = a[cur ^ 1];
        float x[16];
#pragma unroll
        for (int i = 0; i < kIn; ++i) x[i] = fminf(fmaxf((o[i] - norm[i]) * norm[16 + i], -10.f), 10.f);
        x[13] = 1.0f;
        x[14] = 0.f;
        x[15] = 0.f;
        const uint4 x_lo = make_uint4(mlptc::pack_bf16(x[0], x[1]), mlptc::pack_bf16(x[2], x[3]), mlptc::pack_bf16(x[4], x[5]), mlptc::pack_bf16(x[6], x[7]));
        const uint4 x_hi = make_uint4(mlptc::pack_bf16(x[8], x[9]), mlptc::pack_bf16(x[10], x[11]), mlptc::pack_bf16(x[12], x[13]), mlptc::pack_bf16(x[14], x[15]));
        unsigned char* row = a_cur + (m >> 3) * kSboA + (m & 7) * 16;
        *reinterpret_cast<uint4*>(row) = x_lo;
        *reinterpret_cast<uint4*>(row + kLbo) = x_hi;
        // the head reads a_next as {0 (x columns weigh nothing but the ones column carries its bias), h_t}
        unsigned char* nrow = a_next + (m >> 3) * kSboA + (m & 7) * 16;
        *reinterpret_cast<uint4*>(nrow) = make_uint4(0u, 0u, 0u, 0u);
        *reinterpret_cast<uint4*>(nrow + kLbo) = make_uint4(0u, 0u, 0x3F800000u, 0u);            // column 13 = bf16(1.0) (high half of word 2)
        publish(bar_x);
#pragma unroll 1
        for (int p = 0; p < kPasses; ++p) {
            const int b = p & 1;
            wait_tc(bar_g[b], ph_g[b]);                                        // pass p is in gate buffer b
            gate_epilogue(p, kColGates + kPassN * b, a_next);
            if (p + 2 < kPasses) {                                             // pass p + 2 reuses gate buffer b
                mlptc::fence_before_sync();
                mbar_arrive(bar_gfree[b]);
            }
        }
        publish(bar_hready);                                                   // h_t complete
        wait_tc(bar_h, ph_h);
        head_epilogue();
        publish(bar_a2);
        wait_tc(bar_h, ph_h);
        head_epilogue();
        publish(bar_a2);
        wait_tc(bar_h, ph_h);
        uint32_t r0, r1;
        mlptc::tmem_ld2(tmem + ((uint32_t)(m & ~31) << 16) + kColHead, r0, r1);
        mlptc::tmem_ld_wait();
        cur ^= 1;                                                         // a_next = {., h_t} is the next step's A tile
        return make_float2(fminf(fmaxf(__uint_as_float(r0), -1.f), 1.f), fminf(fmaxf(__uint_as_float(r1), -1.f), 1.f));
    }
};

}  // namespace lstmtc
}  // namespace cantor
