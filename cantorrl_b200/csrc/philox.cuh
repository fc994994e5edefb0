// Philox4x32-10 counter-based generator (Salmon et al., SC'11) and the normal transforms built on it.
// Counter-based: the draw for (path, time block) is a pure function of (seed, path, block), so
// trajectories do not depend on the launch geometry or on how paths are sharded across GPUs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cantor {

constexpr unsigned kPhiloxM0 = 0xD2511F53u, kPhiloxM1 = 0xCD9E8D57u;
constexpr unsigned kPhiloxW0 = 0x9E3779B9u, kPhiloxW1 = 0xBB67AE85u;

__host__ __device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        // one 32 x 32 -> 64 multiply per lane pair (IMAD.WIDE.U32 on the device), not separate hi / lo products
        const unsigned long long p0 = (unsigned long long)kPhiloxM0 * c.x, p1 = (unsigned long long)kPhiloxM1 * c.z;
        const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0, hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += kPhiloxW0;
        k.y += kPhiloxW1;
    }
    return c;
}

// uint32 -> uniform in (0, 1]: (x + 1) * 2^-32 rounded to float stays > 0, so log() is finite.
__host__ __device__ __forceinline__ float u32_to_unit_open0(unsigned x) {
    return ((float)(x >> 8) + 1.0f) * 5.9604644775390625e-08f;     // 24-bit mantissa grid, (0, 1]
}
__host__ __device__ __forceinline__ double u32x2_to_unit_open0(unsigned hi, unsigned lo) {
    const unsigned long long m = (((unsigned long long)hi << 32) | lo) >> 11;   // 53 bits
    return ((double)m + 1.0) * 1.1102230246251565404e-16;                         // (0, 1]
}

}  // namespace cantor
