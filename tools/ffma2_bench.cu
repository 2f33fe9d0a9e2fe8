// Microbenchmark: issue rate of 3-register FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and mixed with the
// shared-memory loads + bf16 unpacks of the depthwise strip loops.  Build: nvcc -gencode arch=compute_100a,code=sm_100a
// -O3 -o tools/ffma2_bench tools/ffma2_bench.cu ; prints lane-FMAs per clock per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, const float* in) {
    __shared__ uint32_t sm[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += 256) sm[i] = 0x3f803f80u + i;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    float2 a[8], w[4];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = make_float2(in[i], in[i + 1]);
#pragma unroll
    for (int i = 0; i < 4; i++) w[i] = make_float2(in[8 + i + (lane & 1)], in[9 + i + (lane & 3)]);
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {            // scalar FFMA, register operands only
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    a[i].x = fmaf(a[i].x, w[r & 3].x, w[(r + 1) & 3].x);
                    a[i].y = fmaf(a[i].y, w[r & 3].y, w[(r + 1) & 3].y);
                }
        } else if (MODE == 1) {     // packed FFMA2
#pragma unroll
            for (int r = 0; r < 8; r++)
#pragma unroll
                for (int i = 0; i < 8; i++) a[i] = __ffma2_rn(a[i], w[r & 3], w[(r + 1) & 3]);
        } else if (MODE == 2 || MODE == 3) {   // depthwise-like: 12 LDS + unpack feed 8 outputs x 5 taps (x2 channels)
            const uint32_t* row = sm + ((it & 7) * 64) + lane;
            float2 x[12];
#pragma unroll
            for (int j = 0; j < 12; j++) {
                uint32_t u = row[j * 32];
                x[j] = make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
            }
#pragma unroll
            for (int j = 0; j < 12; j++)
#pragma unroll
                for (int kw = 0; kw < 5; kw++) {
                    const int p = j - kw;
                    if (p >= 0 && p < 8) {
                        if (MODE == 2) {
                            a[p].x = fmaf(x[j].x, w[kw & 3].x, a[p].x);
                            a[p].y = fmaf(x[j].y, w[kw & 3].y, a[p].y);
                        } else {
                            a[p] = __ffma2_rn(x[j], w[kw & 3], a[p]);
                        }
                    }
                }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += a[i].x + a[i].y;
    out[blockIdx.x * 256 + threadIdx.x] = s;
}

template <int MODE>
static void run(const char* name, double fma_per_iter_per_thread) {
    float *out, *in;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMalloc(&in, 64 * 4);
    cudaMemset(in, 0, 64 * 4);
    const int iters = 20000;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int occ = 1; occ <= 8; occ *= 2) {
        k<MODE><<<148 * occ, 256>>>(out, 100, in);
        cudaEventRecord(e0);
        k<MODE><<<148 * occ, 256>>>(out, iters, in);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        double fma = fma_per_iter_per_thread * iters * 256.0 * 148 * occ;
        printf("%-28s ctas/SM %d  %.3f ms  %.2f TFMA/s  (%.1f lane-FMA/clk/SM at 1.965 GHz)\n", name, occ, ms, fma / ms * 1e-9,
               fma / (ms * 1e-3) / 148 / 1.965e9);
    }
}

int main() {
    run<0>("FFMA 3-reg", 128);
    run<1>("FFMA2", 128);
    run<2>("strip k5 P8 FFMA", 80);
    run<3>("strip k5 P8 FFMA2", 80);
    return 0;
}
