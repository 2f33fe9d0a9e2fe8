// Micro-benchmark: tcgen05.ld round-trip latency as seen by an epilogue warpgroup (cycles, clock64).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ldtm_bench tools/ldtm_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

template <int NLD>
__global__ void k(long long* out, int iters) {
    __shared__ uint32_t tptr;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(&tptr)), "r"(256) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t acc = 0;
    long long best = 1ll << 60, tot = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t v[NLD][16];
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < NLD; i++) tmem_ld16(base + i * 16, v[i]);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < NLD; i++)
#pragma unroll
            for (int j = 0; j < 16; j++) acc ^= v[i][j];
        long long t1 = clock64();
        if (t1 - t0 < best) best = t1 - t0;
        tot += t1 - t0;
    }
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x * 8 + warp * 2] = best; out[blockIdx.x * 8 + warp * 2 + 1] = tot / iters + (acc == 0x12345 ? 1 : 0); }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(256) : "memory");
}

template <int NLD> void run(int warps, int blocks) {
    long long* d; cudaMalloc(&d, blocks * 8 * sizeof(long long)); cudaMemset(d, 0, blocks * 8 * sizeof(long long));
    k<NLD><<<blocks, warps * 32>>>(d, 200);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("loads=%d (x16) warps=%d blocks=%d: best %lld avg %lld cycles (%s)\n", NLD, warps, blocks, h[0], h[1], cudaGetErrorString(e));
    cudaFree(d);
}

int main() {
    run<1>(1, 1); run<1>(4, 1); run<4>(1, 1); run<4>(4, 1); run<8>(4, 1); run<8>(4, 148); run<6>(4, 148);
    return 0;
}
