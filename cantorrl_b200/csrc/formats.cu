// Layout conversion between the reference's on-disk arrays and the packed time-major book in HBM.
//
// Reference formats: env-schema npz {paths, volatilities (n, T+1); call_prices_atm, put_prices_atm (n, T)},
// path-major, float64 on disk / float32 in the env (src/sim/rbergomi_sim.py:528, src/env/hedging_env_v2.py:36-48).
// Packed book: float4 {S, v, C, P} at [t * ld + path], t = 0..T, row T repeating the marks of row T-1.
//
// 32 x 32 tiles through shared memory: reads are contiguous along time (source rows), writes are contiguous
// along paths (512 B of float4 records per warp).  Tiles are numbered TIME-FASTEST (1-D grid): the CTAs that run next to each
// other touch neighbouring 128 / 256-byte pieces of the same path-major rows, whose 253-element rows start at every alignment --
// the pieces meet in L2 and leave as whole lines instead of as partially written sectors.
#include "common.cuh"

namespace cantor {

template <typename T>
__global__ void __launch_bounds__(256)
pack_book_kernel(const T* __restrict__ paths, const T* __restrict__ vols, const T* __restrict__ calls,
                 const T* __restrict__ puts, int n_paths, int Tlen, float4* __restrict__ rec, long long ld) {
    __shared__ float tile[4][32][33];
    const unsigned n_t = (unsigned)(Tlen + 1 + 31) / 32;
    const int p0 = (int)(blockIdx.x / n_t) * 32, t0 = (int)(blockIdx.x % n_t) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;                    // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, t = t0 + tx;
        if (p < n_paths && t <= Tlen) {
            const int to = min(t, Tlen - 1);                                   // stale marks in row T
            tile[0][r][tx] = (float)paths[(long long)p * (Tlen + 1) + t];
            tile[1][r][tx] = (float)vols[(long long)p * (Tlen + 1) + t];
            tile[2][r][tx] = (float)calls[(long long)p * Tlen + to];
            tile[3][r][tx] = (float)puts[(long long)p * Tlen + to];
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, p = p0 + tx;
        if (p < n_paths && t <= Tlen)
            rec[(long long)t * ld + p] = make_float4(tile[0][tx][r], tile[1][tx][r], tile[2][tx][r], tile[3][tx][r]);
    }
}

template <typename T>
__global__ void __launch_bounds__(256)
unpack_book_kernel(const float4* __restrict__ rec, long long ld, int n_paths, int Tlen, T* __restrict__ paths,
                   T* __restrict__ vols, T* __restrict__ calls, T* __restrict__ puts) {
    __shared__ float tile[4][32][33];
    const unsigned n_t = (unsigned)(Tlen + 1 + 31) / 32;
    const int p0 = (int)(blockIdx.x / n_t) * 32, t0 = (int)(blockIdx.x % n_t) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int t = t0 + r, p = p0 + tx;
        if (p < n_paths && t <= Tlen) {
            const float4 v = rec[(long long)t * ld + p];
            tile[0][r][tx] = v.x; tile[1][r][tx] = v.y; tile[2][r][tx] = v.z; tile[3][r][tx] = v.w;
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int p = p0 + r, t = t0 + tx;
        if (p < n_paths && t <= Tlen) {
            paths[(long long)p * (Tlen + 1) + t] = (T)tile[0][tx][r];
            vols[(long long)p * (Tlen + 1) + t] = (T)tile[1][tx][r];
            if (t < Tlen) {
                calls[(long long)p * Tlen + t] = (T)tile[2][tx][r];
                puts[(long long)p * Tlen + t] = (T)tile[3][tx][r];
            }
        }
    }
}

}  // namespace cantor

using namespace cantor;

extern "C" int cantor_pack_book(const void* paths, const void* vols, const void* calls, const void* puts,
                                int32_t src_dtype, int32_t n_paths, int32_t episode_length, float* svcp, int64_t ld,
                                void* stream) {
    CANTOR_REQUIRE(paths && vols && calls && puts && svcp, "array is NULL");
    CANTOR_REQUIRE(src_dtype == CANTOR_F32 || src_dtype == CANTOR_F64, "src_dtype must be 32 or 64");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths, "bad shape");
    CANTOR_REQUIRE(aligned16(svcp), "svcp must be 16-byte aligned");
    const long long n_tiles = (long long)((n_paths + 31) / 32) * ((episode_length + 1 + 31) / 32);
    CANTOR_REQUIRE(n_tiles <= 0x7fffffffLL, "book too large for one launch");
    const dim3 grid((unsigned)n_tiles);
    cudaStream_t s = (cudaStream_t)stream;
    if (src_dtype == CANTOR_F64)
        pack_book_kernel<double><<<grid, 256, 0, s>>>((const double*)paths, (const double*)vols, (const double*)calls,
                                                      (const double*)puts, n_paths, episode_length, (float4*)svcp, ld);
    else
        pack_book_kernel<float><<<grid, 256, 0, s>>>((const float*)paths, (const float*)vols, (const float*)calls,
                                                     (const float*)puts, n_paths, episode_length, (float4*)svcp, ld);
    return check_launch("pack_book_kernel");
}

extern "C" int cantor_unpack_book(const float* svcp, int64_t ld, int32_t n_paths, int32_t episode_length,
                                  int32_t dst_dtype, void* paths, void* vols, void* calls, void* puts, void* stream) {
    CANTOR_REQUIRE(paths && vols && calls && puts && svcp, "array is NULL");
    CANTOR_REQUIRE(dst_dtype == CANTOR_F32 || dst_dtype == CANTOR_F64, "dst_dtype must be 32 or 64");
    CANTOR_REQUIRE(n_paths > 0 && episode_length > 0 && ld >= n_paths, "bad shape");
    const long long n_tiles = (long long)((n_paths + 31) / 32) * ((episode_length + 1 + 31) / 32);
    CANTOR_REQUIRE(n_tiles <= 0x7fffffffLL, "book too large for one launch");
    const dim3 grid((unsigned)n_tiles);
    cudaStream_t s = (cudaStream_t)stream;
    if (dst_dtype == CANTOR_F64)
        unpack_book_kernel<double><<<grid, 256, 0, s>>>((const float4*)svcp, ld, n_paths, episode_length, (double*)paths,
                                                        (double*)vols, (double*)calls, (double*)puts);
    else
        unpack_book_kernel<float><<<grid, 256, 0, s>>>((const float4*)svcp, ld, n_paths, episode_length, (float*)paths,
                                                       (float*)vols, (float*)calls, (float*)puts);
    return check_launch("unpack_book_kernel");
}
