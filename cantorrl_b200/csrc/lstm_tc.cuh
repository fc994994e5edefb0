// Recurrent actor for the rollout kernel on the tensor cores: the policy the reference actually trained,
//   LSTM(13 -> 128) -> ReLU MLP(128 -> 64 -> 64) -> Linear(64 -> 2)       (quantconnect/model_wrapper.py:167-204,
//   shapes in quantconnect/model_files/policy_weights.pth; SB3 RecurrentPPO "MlpLstmPolicy", train_ppo_v2.py:46-47, 220-228)
// evaluated for one CTA's 128 envs per env-step with tcgen05.mma (bf16 operands, float32 accumulation in tensor memory).
//
//   rows (M = 128) = the CTA's envs; thread m owns env m, TMEM lane m, and row m of every A tile
//   A tile [128 x 144] bf16 = { normalised obs (13), 1.0, 0, 0 | h (128) }          (biases ride on the ones column)
//   gates: four passes of 32 hidden units each, D[128 x 128] = A * Wg_p^T with Wg_p rows = {i, f, g, o} x 32 units;
//          the 36 KB weight tile of pass p + 1 is copied global(L2) -> shared by the threads while pass p's MMAs run
//          (the 147 KB of gate weights do not fit next to the rest, so they stream; two buffers)
//   epilogue of a pass: tcgen05.ld of the thread's own lane, sigmoid / tanh on the SFU (tanh.approx), the cell state c in
//          float32 in 128 TMEM columns (tcgen05.ld / tcgen05.st), h -> bf16 -> the OTHER A tile (all four passes still
//          read the old h), which then feeds the MLP head as its layer-1 operand and becomes the next step's A tile
//   head:  D[128 x 64] = A * W1^T (K = 144, x-columns weigh 0), ReLU, D = A2 * W2^T, ReLU, D[128 x 16] = A2 * W3^T
// TMEM: columns [0, 128) gate accumulators, [128, 256) cell state, [256, 320) head accumulators (512 allocated: one CTA per
// SM, which the ~200 KB of shared memory imply anyway).  Weight images arrive pre-arranged in the canonical K-major
// no-swizzle core-matrix layout (cantorrl_b200/rollout.py: pack_lstm), so staging them is a straight 16-byte copy.
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
constexpr int kSmemBytes = 2 * kWgBytes + 2 * kABytes + kW1Bytes + kW2Bytes + kW3Bytes + kA2Bytes + 16 + 128;
constexpr int kTmemCols = 512, kColGates = 0, kColCell = 128, kColHead = 256;

__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sigmoid_approx(float x) { return fmaf(0.5f, tanh_approx(0.5f * x), 0.5f); }
__device__ __forceinline__ void tmem_ld16f(uint32_t taddr, float (&f)[16]) {
    uint32_t r[16];
    mlptc::tmem_ld16(taddr, r);
    mlptc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void tmem_st16f(uint32_t taddr, const float (&f)[16]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(__float_as_uint(f[0])), "r"(__float_as_uint(f[1])), "r"(__float_as_uint(f[2])),
                    "r"(__float_as_uint(f[3])), "r"(__float_as_uint(f[4])), "r"(__float_as_uint(f[5])), "r"(__float_as_uint(f[6])),
                    "r"(__float_as_uint(f[7])), "r"(__float_as_uint(f[8])), "r"(__float_as_uint(f[9])), "r"(__float_as_uint(f[10])),
                    "r"(__float_as_uint(f[11])), "r"(__float_as_uint(f[12])), "r"(__float_as_uint(f[13])), "r"(__float_as_uint(f[14])),
                    "r"(__float_as_uint(f[15])) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

struct Actor {
    unsigned char* wg[2];       // gate-weight buffers
    unsigned char* a[2];        // A tiles; a[cur] = {x_t, h_{t-1}}
    unsigned char* w1;
    unsigned char* w2;
    unsigned char* w3;
    unsigned char* a2;
    const unsigned char* img;   // global weight image
    const float* norm;          // shared: mean[16], inv_std[16]
    uint32_t mbar, tmem, phase;
    int cur;
    bool timed_out;

    __device__ __forceinline__ void copy_tile(unsigned char* dst, const unsigned char* src, int bytes) {
        const uint4* s = reinterpret_cast<const uint4*>(src);
        uint4* d = reinterpret_cast<uint4*>(dst);
        for (int j = threadIdx.x; j < bytes / 16; j += kRows) d[j] = __ldg(s + j);
    }

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
        uint64_t* bar = reinterpret_cast<uint64_t*>(a2 + kA2Bytes);
        uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
        float* norm_s = reinterpret_cast<float*>(a2 + kA2Bytes + 16);
        mbar = mlptc::smem_u32(bar);
        phase = 0;
        cur = 0;
        timed_out = false;
        const int tid = threadIdx.x;
        copy_tile(wg[0], img + kImgGate, kWgBytes);                       // pass 0 of the first step
        copy_tile(w1, img + kImgW1, kW1Bytes + kW2Bytes + kW3Bytes);      // the head's three tiles are contiguous in both places
        if (tid < 32) norm_s[tid] = reinterpret_cast<const float*>(img + kImgNorm)[tid];
        norm = norm_s;
        // zero both A tiles (h = 0) and the constant tail of the A2 rows
        for (int j = tid; j < 2 * kABytes / 16; j += kRows) reinterpret_cast<uint4*>(a[0])[j] = make_uint4(0u, 0u, 0u, 0u);
        {
            unsigned char* row = a2 + (tid >> 3) * kSbo2 + (tid & 7) * 16;
            *reinterpret_cast<uint4*>(row + 8 * kLbo) = make_uint4(0x00003F80u, 0u, 0u, 0u);
            *reinterpret_cast<uint4*>(row + 9 * kLbo) = make_uint4(0u, 0u, 0u, 0u);
        }
        if (tid == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(mbar) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
        reset_state();
    }

    __device__ __forceinline__ void teardown() {
        mlptc::fence_before_sync();
        __syncthreads();
        if (threadIdx.x < 32)
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(tmem), "r"((uint32_t)kTmemCols) : "memory");
    }

    // h = c = 0 for this thread's env (SB3 resets the LSTM state at an episode start).  Warp-collective (tcgen05.st):
    // every thread of the CTA calls it together -- the envs of a rollout finish their episodes in lockstep.
    __device__ __forceinline__ void reset_state() {
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16);
        float z[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) z[j] = 0.f;
#pragma unroll
        for (int c = 0; c < kH / 16; ++c) tmem_st16f(lane_addr + kColCell + 16 * c, z);
        unsigned char* row = a[cur] + (m >> 3) * kSboA + (m & 7) * 16;
#pragma unroll
        for (int c = 2; c < kKA / 8; ++c) *reinterpret_cast<uint4*>(row + c * kLbo) = make_uint4(0u, 0u, 0u, 0u);
    }

    __device__ __forceinline__ void wait_mma() {
        uint32_t done = 0;
        unsigned spins = 0;
        while (!done && !timed_out) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(mbar), "r"(phase) : "memory");
            if (!done && ++spins > mlptc::kSpinLimit) timed_out = true;
        }
        phase ^= 1;
        mlptc::fence_after_sync();
    }

    // all threads: publish shared-memory writes; one thread: `ksteps` MMAs of K = 16 into TMEM column `dcol`, then commit
    __device__ __forceinline__ void issue(uint32_t a_addr, uint32_t a_sbo, uint32_t b_addr, uint32_t b_sbo, int ksteps, uint32_t idesc, int dcol) {
        mlptc::fence_before_sync();
        fence_proxy_async_smem();
        __syncthreads();
        if (threadIdx.x == 0) {
            mlptc::fence_after_sync();
            for (int k = 0; k < ksteps; ++k)
                mlptc::umma_bf16(tmem + dcol, mlptc::smem_desc(a_addr + k * 2 * kLbo, kLbo, a_sbo), mlptc::smem_desc(b_addr + k * 2 * kLbo, kLbo, b_sbo),
                                 idesc, k > 0 ? 1u : 0u);
            mlptc::umma_commit(mbar);
        }
    }

    // gates of hidden units 32 p .. 32 p + 31 are in TMEM columns [0, 128) as {i | f | g | o} x 32: update c, write h
    __device__ __forceinline__ void gate_epilogue(int p, unsigned char* a_next) {
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16);
        unsigned char* row = a_next + (m >> 3) * kSboA + (m & 7) * 16;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            float gi[16], gf[16], gg[16], go[16], c[16];
            tmem_ld16f(lane_addr + kColGates + 0 * kUnitsPerPass + 16 * half, gi);
            tmem_ld16f(lane_addr + kColGates + 1 * kUnitsPerPass + 16 * half, gf);
            tmem_ld16f(lane_addr + kColGates + 2 * kUnitsPerPass + 16 * half, gg);
            tmem_ld16f(lane_addr + kColGates + 3 * kUnitsPerPass + 16 * half, go);
            tmem_ld16f(lane_addr + kColCell + kUnitsPerPass * p + 16 * half, c);
            float h[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                c[j] = fmaf(sigmoid_approx(gf[j]), c[j], sigmoid_approx(gi[j]) * tanh_approx(gg[j]));
                h[j] = sigmoid_approx(go[j]) * tanh_approx(c[j]);
            }
            tmem_st16f(lane_addr + kColCell + kUnitsPerPass * p + 16 * half, c);
            const int chunk = (16 + kUnitsPerPass * p + 16 * half) / 8;                         // 8 bf16 per 16-byte chunk
            *reinterpret_cast<uint4*>(row + chunk * kLbo) = make_uint4(mlptc::pack_bf16(h[0], h[1]), mlptc::pack_bf16(h[2], h[3]),
                                                                        mlptc::pack_bf16(h[4], h[5]), mlptc::pack_bf16(h[6], h[7]));
            *reinterpret_cast<uint4*>(row + (chunk + 1) * kLbo) = make_uint4(mlptc::pack_bf16(h[8], h[9]), mlptc::pack_bf16(h[10], h[11]),
                                                                              mlptc::pack_bf16(h[12], h[13]), mlptc::pack_bf16(h[14], h[15]));
        }
    }

    __device__ __forceinline__ void head_epilogue() {                       // 64 head accumulators -> ReLU -> bf16 -> A2 row
        const int m = threadIdx.x;
        const uint32_t lane_addr = tmem + ((uint32_t)(m & ~31) << 16) + kColHead;
        unsigned char* row = a2 + (m >> 3) * kSbo2 + (m & 7) * 16;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            float r[16];
            tmem_ld16f(lane_addr + c * 16, r);
            *reinterpret_cast<uint4*>(row + (2 * c) * kLbo) = make_uint4(mlptc::relu_pack_bf16(r[0], r[1]), mlptc::relu_pack_bf16(r[2], r[3]),
                                                                         mlptc::relu_pack_bf16(r[4], r[5]), mlptc::relu_pack_bf16(r[6], r[7]));
            *reinterpret_cast<uint4*>(row + (2 * c + 1) * kLbo) = make_uint4(mlptc::relu_pack_bf16(r[8], r[9]), mlptc::relu_pack_bf16(r[10], r[11]),
                                                                             mlptc::relu_pack_bf16(r[12], r[13]), mlptc::relu_pack_bf16(r[14], r[15]));
        }
    }

    // One policy step on this thread's observation; CTA-collective.
    __device__ __forceinline__ float2 forward(const float* o) {
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

        const uint32_t idesc_g = mlptc::instr_desc(kRows, kPassN);
        issue(mlptc::smem_u32(a_cur), kSboA, mlptc::smem_u32(wg[0]), kSboA, kKA / 16, idesc_g, kColGates);
#pragma unroll 1
        for (int p = 0; p < kPasses; ++p) {
            // stream the next pass's weights (pass 0 of the next step after the last one) while this pass's MMAs run
            copy_tile(wg[(p + 1) & 1], img + kImgGate + ((p + 1) % kPasses) * kWgBytes, kWgBytes);
            wait_mma();
            gate_epilogue(p, a_next);
            if (p + 1 < kPasses)
                issue(mlptc::smem_u32(a_cur), kSboA, mlptc::smem_u32(wg[(p + 1) & 1]), kSboA, kKA / 16, idesc_g, kColGates);
        }
        // head on h_t
        issue(mlptc::smem_u32(a_next), kSboA, mlptc::smem_u32(w1), kSboA, kKA / 16, mlptc::instr_desc(kRows, 64), kColHead);
        wait_mma();
        head_epilogue();
        issue(mlptc::smem_u32(a2), kSbo2, mlptc::smem_u32(w2), kSbo2, kK2 / 16, mlptc::instr_desc(kRows, 64), kColHead);
        wait_mma();
        head_epilogue();
        issue(mlptc::smem_u32(a2), kSbo2, mlptc::smem_u32(w3), kSbo2, kK2 / 16, mlptc::instr_desc(kRows, mlptc::kN3), kColHead);
        wait_mma();
        uint32_t r0, r1;
        mlptc::tmem_ld2(tmem + ((uint32_t)(m & ~31) << 16) + kColHead, r0, r1);
        mlptc::tmem_ld_wait();
        cur ^= 1;                                                         // a_next = {., h_t} is the next step's A tile
        return make_float2(fminf(fmaxf(__uint_as_float(r0), -1.f), 1.f), fminf(fmaxf(__uint_as_float(r1), -1.f), 1.f));
    }
};

}  // namespace lstmtc
}  // namespace cantor
