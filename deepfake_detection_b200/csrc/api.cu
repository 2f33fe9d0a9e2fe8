// Error reporting, version and small runtime helpers of the C-ABI library.
#include <stdio.h>
#include <string.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

int dfd_set_error(int code, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "");
    return code;
}
int dfd_set_cuda_error(cudaError_t e, const char* file, int line) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) at %s:%d", (int)e, cudaGetErrorString(e), file, line);
    return DFD_ERR_CUDA;
}

extern "C" {

const char* dfd_last_error(void) { return g_err; }

int dfd_abi_version(void) { return 1; }

int dfd_stat_slots(void) { return DFD_STAT_SLOTS; }

int dfd_memset_async(void* p, int value, long long bytes, void* stream) {
    cudaError_t e = cudaMemsetAsync(p, value, (size_t)bytes, (cudaStream_t)stream);
    if (e != cudaSuccess) return dfd_set_cuda_error(e, __FILE__, __LINE__);
    return DFD_OK;
}

}  // extern "C"
