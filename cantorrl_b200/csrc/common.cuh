// Shared host/device helpers for libcantor_hedge.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/cantor_hedge.h"

namespace cantor {

// ---- per-thread last error ---------------------------------------------------------------------
inline char* error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}
inline int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
    snprintf(error_buffer(), 512, fmt, a, b);
    return code;
}
inline int cuda_fail(cudaError_t e, const char* where) {
    return fail(CANTOR_ERR_CUDA, "%s: %s", where, cudaGetErrorString(e));
}
#define CANTOR_REQUIRE(cond, msg) \
    do { if (!(cond)) return ::cantor::fail(CANTOR_ERR_INVALID, "%s: %s", __func__, msg); } while (0)
#define CANTOR_CUDA(call) \
    do { cudaError_t e_ = (call); if (e_ != cudaSuccess) return ::cantor::cuda_fail(e_, #call); } while (0)

inline int check_launch(const char* kernel) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, kernel);
    return CANTOR_OK;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- streaming loads/stores: every byte of the hot path is touched once per step ------------------
__device__ __forceinline__ float ldg_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ int4 ldg_stream(const int4* p) { return __ldcs(p); }
__device__ __forceinline__ float2 ldg_stream(const float2* p) { return __ldcs(p); }

// ---- 1-D TMA bulk store shared::cta -> global (SASS: UBLKCP) ---------------------------------------
// Caller guarantees 16-byte aligned src/dst and bytes % 16 == 0.
__device__ __forceinline__ void tma_store_1d(void* gdst, const void* ssrc, uint32_t bytes) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 :: "l"(gdst), "r"(s), "r"(bytes) : "memory");
}
// the same with an L2 evict-first policy: data that nobody on the device reads again should not displace what is re-read
__device__ __forceinline__ void tma_store_1d_evict_first(void* gdst, const void* ssrc, uint32_t bytes) {
    uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(ssrc));
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;"
                 :: "l"(gdst), "r"(s), "r"(bytes), "l"(policy) : "memory");
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); }
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until the bulk copies have finished READING shared memory (the CTA may then exit / reuse it)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// generic-proxy writes to shared memory -> visible to the async (TMA) proxy
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- programmatic dependent launch (PDL) -----------------------------------------------------------
// A kernel launched with launch_pdl() may start while the previous kernel in the stream is still
// draining; it must call pdl_wait_prior_grid() before touching anything that kernel wrote.
// pdl_launch_dependents() tells the scheduler that the NEXT kernel's CTAs may be made resident.
__device__ __forceinline__ void pdl_wait_prior_grid() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// split cluster barrier: every thread of every CTA of the cluster arrives, later waits (all threads, converged warps)
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

// `cluster` > 1: thread-block clusters of that many CTAs along x (the grid must be a multiple of it)
inline int launch_pdl(const void* kernel, dim3 grid, dim3 block, cudaStream_t stream, void** args, size_t smem = 0, unsigned cluster = 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (cluster > 1) {
        attr[1].id = cudaLaunchAttributeClusterDimension;
        attr[1].val.clusterDim.x = cluster;
        attr[1].val.clusterDim.y = 1;
        attr[1].val.clusterDim.z = 1;
        cfg.numAttrs = 2;
    }
    cudaError_t e = cudaLaunchKernelExC(&cfg, kernel, args);
    if (e != cudaSuccess) return cuda_fail(e, "cudaLaunchKernelExC");
    return CANTOR_OK;
}

}  // namespace cantor
