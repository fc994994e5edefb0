// Library-level entry points of libcantor_hedge.so.
#include "common.cuh"

extern "C" int cantor_abi_version(void) { return CANTOR_ABI_VERSION; }

extern "C" const char* cantor_last_error(void) { return cantor::error_buffer(); }

extern "C" int cantor_device_info(int device, int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0 || device < 0 || device >= n)
        return cantor::fail(CANTOR_ERR_NO_DEVICE, "cantor_device_info: %s%s", "no usable CUDA device: ",
                            e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range");
    cudaDeviceProp prop;
    CANTOR_CUDA(cudaGetDeviceProperties(&prop, device));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return CANTOR_OK;
}
