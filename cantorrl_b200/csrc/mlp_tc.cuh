// Tensor-core actor for the rollout kernel: the ReLU MLP 13 -> 64 -> 64 -> 2 of one CTA's 128 envs as three
// tcgen05.mma products per env-step, bf16 operands staged in shared memory, float32 accumulators in tensor memory.
//
//   rows (M = 128)   = the CTA's envs: thread m owns env m, TMEM lane m holds its accumulator row
//   layer 1  D[128 x 64] = A1[128 x 16] * W1^T     A1 = {normalised obs (13), 1.0, 0, 0}; W1 column 13 = b1 (bias folded)
//   layer 2  D[128 x 64] = A2[128 x 80] * W2^T     A2 = {ReLU(h1) (64), 1.0, 0 x 15};     W2 column 64 = b2
//   layer 3  D[128 x 16] = A2[128 x 80] * W3^T     A2 = {ReLU(h2) (64), 1.0, ...};        W3 rows 0, 1 = the two actions, column 64 = b3
//
// Operands use the canonical K-major, no-swizzle UMMA layout ("core matrices" of 8 rows x 16 bytes, cute
// Layout_K_INTER_Atom): element (row, k) of a [rows x K] bf16 tile sits at
//   (row / 8) * SBO + (k / 8) * LBO + (row % 8) * 16 + (k % 8) * 2      with LBO = 128, SBO = (K / 8) * 128
// so thread m writes its row as K/8 sixteen-byte stores; eight consecutive threads fill one 128-byte core matrix
// (bank-conflict free).  Each epilogue is: tcgen05.ld of the thread's own lane (16 columns at a time) -> ReLU fused
// into cvt.rn.relu.bf16x2.f32 -> st.shared into the next layer's A tile.  One elected thread issues the MMAs and
// commits them to an mbarrier that all 128 threads wait on.
//
// Reference: the actor head of the shipped policy, quantconnect/model_wrapper.py:177-185 (ReLU MLP), :131 (observation
// normalisation), with SB3's clip of the action to the Box bounds.  bf16 operands make this the THROUGHPUT form of the
// policy (~3 significant digits); the float32 FFMA form (rollout.cu: policy_mlp) is the parity form.
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

// 1: all four warps poll the MMA mbarrier; 0: warp 0 polls, the others park at the CTA barrier.  Measured (2^20 envs x 252
// steps, GBM on the fly): 6.88 ms against 7.01 ms -- the polls are 23 % of the executed instructions but fill issue slots nobody
// else wants: the kernel is bound by the latency of its three serial MMA round trips per step, not by issue bandwidth.
#ifndef CANTOR_MLP_BIAS_IN_EPILOGUE    // 1: layers 2 / 3 of the MLP actor as K = 64 products (4 instead of 5 MMAs each), their biases added by
#define CANTOR_MLP_BIAS_IN_EPILOGUE 0  // the threads in the epilogue instead of riding on a ones column that costs a fifth K-step.
#endif                                 // Measured (2^20 x 252, GBM): 7.57 ms against 6.86 (7 CTAs/SM at 72 registers: 7.35): the 64 FADD +
                                       // 16 LDS per thread-step cost more issue slots than the two MMAs give back
#ifndef CANTOR_MLP_ALL_WARPS_POLL
#define CANTOR_MLP_ALL_WARPS_POLL 1
#endif

namespace cantor {
namespace mlptc {

constexpr int kRows = 128;                       // envs per CTA = UMMA M
constexpr int kIn = 13, kHidden = 64, kOut = 2;
constexpr int kK1 = 16;                          // layer-1 K: 13 inputs + ones column + 2 zero columns
constexpr int kK2 = 80;                          // layer-2/3 K: 64 activations + ones column + 15 zero columns
constexpr int kN3 = 16;                          // layer-3 N padded to the UMMA minimum for M = 128
constexpr int kLbo = 128;                        // bytes between the core matrices of consecutive 8-element K chunks
constexpr int kSbo1 = (kK1 / 8) * 128;           // bytes between 8-row groups, K = 16
constexpr int kSbo2 = (kK2 / 8) * 128;           // K = 80
constexpr int kW1Bytes = kHidden * kK1 * 2, kW2Bytes = kHidden * kK2 * 2, kW3Bytes = kN3 * kK2 * 2;
constexpr int kA1Bytes = kRows * kK1 * 2, kA2Bytes = kRows * kK2 * 2;
// The layer-1 operand tile A1 (4 KB) ALIASES the head of the A2 tile: A2 is first written by the layer-1 epilogue, after the
// layer-1 MMA has read A1, and A1 is next written after the layer-3 MMA has read A2.  36 KB per CTA instead of 40: six CTAs
// per SM fit instead of five.  (The A2 rows' constant tail, which A1 overwrites for some rows, is rewritten by the epilogue.)
// the MLP actor's own layer-2 / 3 geometry (the recurrent actor's head keeps the K = 80 tiles above)
constexpr int kKa = CANTOR_MLP_BIAS_IN_EPILOGUE ? 64 : kK2;
constexpr int kSboA2 = (kKa / 8) * 128;
constexpr int kW2aBytes = kHidden * kKa * 2, kW3aBytes = kN3 * kKa * 2, kA2aBytes = kRows * kKa * 2;
constexpr int kBiasBytes = CANTOR_MLP_BIAS_IN_EPILOGUE ? (kHidden + 16) * 4 : 0;      // b2[64], b3[2] as floats
constexpr int kSmemBytes = kW1Bytes + kW2aBytes + kW3aBytes + kA2aBytes + 16 + 128 + kBiasBytes;   // + mbarrier (8) + TMEM address (4) + mean / inv_std (2 x 16 floats)
constexpr int kTmemCols = 64;
constexpr unsigned kSpinLimit = 1u << 24;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1 = Blackwell).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M x N (cute::UMMA::InstrDescriptor).
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// NK consecutive K-steps of D[128 x n] = A * B^T into TMEM address `d` (the first one overwrites), as ONE straight-line
// block: the descriptors of K-step k are those of K-step 0 plus 16 k in the address field (two 128-byte core-matrix
// columns), formed by chained 64-bit adds that ptxas keeps on the uniform datapath (one R2UR per operand, then
// UIADD3.64 / UTCHMMA pairs).  The rolled loop it replaces re-derived both descriptors from vector registers for every
// MMA -- ~100 clocks each on the recurrent actor's issuer warp, which shares its scheduler with two epilogue warps, i.e.
// ~1 000 clocks between "the epilogue released the gate columns" and "the next pass is in the tensor pipe" on the critical
// path of every pass (lstm_tc.cuh: 85 -> 80.8 ms).  For the plain MLP actor, with five CTAs per SM to hide it, it is neutral.
#define CANTOR_UMMA_STEP(ACC) "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, " ACC ";\n\tadd.s64 da, da, 16;\n\tadd.s64 db, db, 16;\n\t"
#define CANTOR_UMMA_HEAD "{\n\t.reg .pred pt, pf;\n\t.reg .b64 da, db;\n\tsetp.eq.u32 pt, 0, 0;\n\tsetp.ne.u32 pf, 0, 0;\n\tmov.b64 da, %1;\n\tmov.b64 db, %2;\n\t"
template <int NK>
__device__ __forceinline__ void umma_batch(uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t d) {
    static_assert(NK == 1 || NK == 4 || NK == 5 || NK == 9, "K-step counts in use: K = 16 (x-part, MLP layer 1), 64 / 80 (layers 2 / 3), 144 (LSTM A tile)");
    if (NK == 1)
        asm volatile(CANTOR_UMMA_HEAD CANTOR_UMMA_STEP("pf") "}" :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
    else if (NK == 4)
        asm volatile(CANTOR_UMMA_HEAD CANTOR_UMMA_STEP("pf") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt")
                     "}" :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
    else if (NK == 5)
        asm volatile(CANTOR_UMMA_HEAD CANTOR_UMMA_STEP("pf") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt")
                     CANTOR_UMMA_STEP("pt") "}" :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
    else
        asm volatile(CANTOR_UMMA_HEAD CANTOR_UMMA_STEP("pf") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt")
                     CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt") CANTOR_UMMA_STEP("pt")
                     "}" :: "r"(d), "l"(a_desc), "l"(b_desc), "r"(idesc) : "memory");
}
#undef CANTOR_UMMA_STEP
#undef CANTOR_UMMA_HEAD

__device__ __forceinline__ void umma_commit(uint32_t mbar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(mbar) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld2(uint32_t taddr, uint32_t& r0, uint32_t& r1) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// {lo, hi} = {bf16(max(a, 0)), bf16(max(b, 0))}: `a` lands at the lower address
__device__ __forceinline__ uint32_t relu_pack_bf16(float a, float b) {
    uint32_t d;
    asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t d;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(b), "f"(a));
    return d;
}

struct Actor {
    unsigned char* w1;      // shared-memory tiles (bf16, core-matrix layout)
    unsigned char* w2;
    unsigned char* w3;
    unsigned char* a1;
    unsigned char* a2;
    uint32_t mbar, tmem, phase;
    const float* norm;      // shared: mean[16], inv_std[16]
    const float* bias;      // shared: b2[64], b3[2] (CANTOR_MLP_BIAS_IN_EPILOGUE)
    bool timed_out;

    // One-time CTA setup: converts the float32 weight block to bf16 tiles, allocates 64 TMEM columns, arms the mbarrier.
    // `w` = W1[13][64] b1[64] W2[64][64] b2[64] W3[64][2] b3[2] mean[13] inv_std[13] (cantor_policy.mlp), in global memory.
    __device__ __forceinline__ void setup(unsigned char* smem, const float* __restrict__ w, float obs_clip) {
        w1 = smem;
        w2 = w1 + kW1Bytes;
        w3 = w2 + kW2aBytes;
        a2 = w3 + kW3aBytes;
        a1 = a2;
        uint64_t* bar = reinterpret_cast<uint64_t*>(a2 + kA2aBytes);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
        mbar = smem_u32(bar);
        phase = 0;
        timed_out = false;
        const float* W1 = w;
        const float* b1 = W1 + kIn * kHidden;
        const float* W2 = b1 + kHidden;
        const float* b2 = W2 + kHidden * kHidden;
        const float* W3 = b2 + kHidden;
        const float* b3 = W3 + kHidden * kOut;
        const int tid = threadIdx.x;
        float* norm_s = reinterpret_cast<float*>(a2 + kA2aBytes + 16);
        float* bias_s = norm_s + 32;
        bias = bias_s;
        if (CANTOR_MLP_BIAS_IN_EPILOGUE && tid < kHidden + kOut) bias_s[tid] = tid < kHidden ? b2[tid] : b3[tid - kHidden];
        if (tid < 32) norm_s[tid] = (tid & 15) < kIn ? (b3 + kOut)[(tid >> 4) * kIn + (tid & 15)] : (tid == 15 ? obs_clip : 0.f);   // [15] = clip
        norm = norm_s;
        __nv_bfloat16* w1h = reinterpret_cast<__nv_bfloat16*>(w1);
        for (int e = tid; e < kHidden * kK1; e += kRows) {           // B1(n, k) = W1[k][n], k = 13 -> b1[n]
            const int n = e / kK1, k = e % kK1;
            const float v = k < kIn ? W1[k * kHidden + n] : (k == kIn ? b1[n] : 0.f);
            w1h[((n >> 3) * kSbo1 + (k >> 3) * kLbo + (n & 7) * 16 + (k & 7) * 2) >> 1] = __float2bfloat16_rn(v);
        }
        __nv_bfloat16* w2h = reinterpret_cast<__nv_bfloat16*>(w2);
        for (int e = tid; e < kHidden * kKa; e += kRows) {           // B2(n, k) = W2[k][n], k = 64 -> b2[n]
            const int n = e / kKa, k = e % kKa;
            const float v = k < kHidden ? W2[k * kHidden + n] : (k == kHidden ? b2[n] : 0.f);
            w2h[((n >> 3) * kSboA2 + (k >> 3) * kLbo + (n & 7) * 16 + (k & 7) * 2) >> 1] = __float2bfloat16_rn(v);
        }
        __nv_bfloat16* w3h = reinterpret_cast<__nv_bfloat16*>(w3);
        for (int e = tid; e < kN3 * kKa; e += kRows) {               // B3(n, k) = W3[k][n] for n < 2, k = 64 -> b3[n]
            const int n = e / kKa, k = e % kKa;
            float v = 0.f;
            if (n < kOut) v = k < kHidden ? W3[k * kOut + n] : (k == kHidden ? b3[n] : 0.f);
            w3h[((n >> 3) * kSboA2 + (k >> 3) * kLbo + (n & 7) * 16 + (k & 7) * 2) >> 1] = __float2bfloat16_rn(v);
        }
        if (!CANTOR_MLP_BIAS_IN_EPILOGUE) {                           // constant tail of this thread's A2 row: column 64 = 1.0 (bias), 65..79 = 0
            const int m = tid;
            unsigned char* row = a2 + (m >> 3) * kSboA2 + (m & 7) * 16;
            *reinterpret_cast<uint4*>(row + 8 * kLbo) = make_uint4(0x00003F80u, 0u, 0u, 0u);     // bf16(1.0) = 0x3F80
            *reinterpret_cast<uint4*>(row + 9 * kLbo) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        if (tid < 32) {                                               // warp 0 owns the TMEM allocation
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         :: "r"(smem_u32(tmem_slot)), "r"((uint32_t)kTmemCols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
        fence_proxy_async_smem();                                     // weight tiles -> visible to the tensor core (async proxy)
        fence_before_sync();
        __syncthreads();
        fence_after_sync();
        tmem = *tmem_slot;
    }

    __device__ __forceinline__ void teardown() {
        fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }

    // Only warp 0 polls the mbarrier; the other three warps sleep in the hardware CTA barrier until it joins them.  With all
    // four warps polling, the try_wait loop was 23 % of the kernel's executed instructions (~23 iterations per wait per warp:
    // profiles/r01_rollout_mlp_bf16_tcgen05_ncu_summary.txt) on an issue-bound kernel; a warp parked at bar.sync issues nothing.
    __device__ __forceinline__ void wait_mma() {
#if CANTOR_MLP_ALL_WARPS_POLL
        const bool poller = true;
#else
        const bool poller = threadIdx.x < 32;
#endif
        if (poller) {
            uint32_t done = 0;
            unsigned spins = 0;
            while (!done && !timed_out) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
                if (!done && ++spins > kSpinLimit) timed_out = true;  // never hang the GPU: bail out, the caller reports it
            }
        }
#if !CANTOR_MLP_ALL_WARPS_POLL
        __syncthreads();                  // (a timed-out warp 0 still arrives; its flag reaches the statistics through thread 0)
#endif
        phase ^= 1;
        fence_after_sync();
    }

    // publish this thread's freshly written A rows, then let one thread issue `ksteps` MMAs of K = 16 each
    template <int KSTEPS>
    __device__ __forceinline__ void run_layer(uint32_t a_addr, uint32_t a_sbo, uint32_t b_addr, uint32_t b_sbo, uint32_t idesc) {
        fence_before_sync();          // earlier tcgen05.ld of the accumulator this layer overwrites
        fence_proxy_async_smem();     // st.shared of the A tile -> async proxy
        __syncthreads();
        if (threadIdx.x == 0) {
            fence_after_sync();
            umma_batch<KSTEPS>(smem_desc(a_addr, kLbo, a_sbo), smem_desc(b_addr, kLbo, b_sbo), idesc, tmem);
            umma_commit(mbar);
        }
        wait_mma();
    }

    // accumulator row of this thread (64 columns) -> ReLU -> bf16 -> this thread's A2 row; two TMEM loads in flight per wait
    // (all four at once cost 30 more registers and a resident CTA per SM: 7.99 ms instead of 6.9 for the 2^20 x 252 sweep)
    // ADD_BIAS (layer 2 with CANTOR_MLP_BIAS_IN_EPILOGUE): b2 is added here, in float32, before the ReLU
    template <bool ADD_BIAS>
    __device__ __forceinline__ void hidden_epilogue() {
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16);
        unsigned char* row = a2 + (m >> 3) * kSboA2 + (m & 7) * 16;
#pragma unroll
        for (int c2 = 0; c2 < kHidden / 32; ++c2) {
            uint32_t r[2][16];
            tmem_ld16(lane_addr + (2 * c2) * 16, r[0]);
            tmem_ld16(lane_addr + (2 * c2 + 1) * 16, r[1]);
            tmem_ld_wait();
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = 2 * c2 + h;
                // ties the registers of the asynchronous loads to this point: nothing below may be scheduled above the wait
                asm volatile("" : "+r"(r[h][0]), "+r"(r[h][1]), "+r"(r[h][2]), "+r"(r[h][3]), "+r"(r[h][4]), "+r"(r[h][5]), "+r"(r[h][6]), "+r"(r[h][7]),
                                  "+r"(r[h][8]), "+r"(r[h][9]), "+r"(r[h][10]), "+r"(r[h][11]), "+r"(r[h][12]), "+r"(r[h][13]), "+r"(r[h][14]), "+r"(r[h][15]) :: "memory");
                if (ADD_BIAS) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 bq = *reinterpret_cast<const float4*>(bias + 16 * c + 4 * q);
                        r[h][4 * q + 0] = __float_as_uint(__uint_as_float(r[h][4 * q + 0]) + bq.x);
                        r[h][4 * q + 1] = __float_as_uint(__uint_as_float(r[h][4 * q + 1]) + bq.y);
                        r[h][4 * q + 2] = __float_as_uint(__uint_as_float(r[h][4 * q + 2]) + bq.z);
                        r[h][4 * q + 3] = __float_as_uint(__uint_as_float(r[h][4 * q + 3]) + bq.w);
                    }
                }
                uint4 lo, hi;
                lo.x = relu_pack_bf16(__uint_as_float(r[h][0]), __uint_as_float(r[h][1]));
                lo.y = relu_pack_bf16(__uint_as_float(r[h][2]), __uint_as_float(r[h][3]));
                lo.z = relu_pack_bf16(__uint_as_float(r[h][4]), __uint_as_float(r[h][5]));
                lo.w = relu_pack_bf16(__uint_as_float(r[h][6]), __uint_as_float(r[h][7]));
                hi.x = relu_pack_bf16(__uint_as_float(r[h][8]), __uint_as_float(r[h][9]));
                hi.y = relu_pack_bf16(__uint_as_float(r[h][10]), __uint_as_float(r[h][11]));
                hi.z = relu_pack_bf16(__uint_as_float(r[h][12]), __uint_as_float(r[h][13]));
                hi.w = relu_pack_bf16(__uint_as_float(r[h][14]), __uint_as_float(r[h][15]));
                *reinterpret_cast<uint4*>(row + (2 * c) * kLbo) = lo;
                *reinterpret_cast<uint4*>(row + (2 * c + 1) * kLbo) = hi;
            }
        }
        if (!CANTOR_MLP_BIAS_IN_EPILOGUE) {
            *reinterpret_cast<uint4*>(row + 8 * kLbo) = make_uint4(0x00003F80u, 0u, 0u, 0u);     // column 64 = bf16(1.0): the bias column
            *reinterpret_cast<uint4*>(row + 9 * kLbo) = make_uint4(0u, 0u, 0u, 0u);
        }
    }

    // The actor on this thread's observation; CTA-collective (every thread of the CTA must call it each step).
    __device__ __forceinline__ float2 forward(const float* o) {
        const int m = threadIdx.x;
        float x[kK1];
        const float clip = norm[15];                                  // VecNormalize's clip_obs (+inf: the deployment wrapper clips nothing)
#pragma unroll
        for (int i = 0; i < kIn; ++i) x[i] = fminf(fmaxf((o[i] - norm[i]) * norm[16 + i], -clip), clip);
        x[13] = 1.0f;
        x[14] = 0.f;
        x[15] = 0.f;
        unsigned char* row = a1 + (m >> 3) * kSbo1 + (m & 7) * 16;
        *reinterpret_cast<uint4*>(row) = make_uint4(pack_bf16(x[0], x[1]), pack_bf16(x[2], x[3]), pack_bf16(x[4], x[5]), pack_bf16(x[6], x[7]));
        *reinterpret_cast<uint4*>(row + kLbo) = make_uint4(pack_bf16(x[8], x[9]), pack_bf16(x[10], x[11]), pack_bf16(x[12], x[13]), pack_bf16(x[14], x[15]));
        run_layer<kK1 / 16>(smem_u32(a1), kSbo1, smem_u32(w1), kSbo1, instr_desc(kRows, kHidden));
        hidden_epilogue<false>();
        run_layer<kKa / 16>(smem_u32(a2), kSboA2, smem_u32(w2), kSboA2, instr_desc(kRows, kHidden));
        hidden_epilogue<CANTOR_MLP_BIAS_IN_EPILOGUE != 0>();
        run_layer<kKa / 16>(smem_u32(a2), kSboA2, smem_u32(w3), kSboA2, instr_desc(kRows, kN3));
        uint32_t r0, r1;
        tmem_ld2(tmem + ((uint32_t)(m & ~31) << 16), r0, r1);
        tmem_ld_wait();
        if (CANTOR_MLP_BIAS_IN_EPILOGUE) return make_float2(__uint_as_float(r0) + bias[kHidden], __uint_as_float(r1) + bias[kHidden + 1]);
        return make_float2(__uint_as_float(r0), __uint_as_float(r1));         // action means; the caller squashes them
    }
};

}  // namespace mlptc
}  // namespace cantor
