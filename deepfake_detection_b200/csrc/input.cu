// The step's inputs that are not activations: the prefetcher's image normalisation and the stochastic-regularisation
// masks.
//
// Reference semantics restated here:
//   PrefetchLoader.__iter__       dfd/timm/data/loader.py:243-256  (uint8 NCHW batch -> float, (x - mean*255) / (std*255),
//                                 mean / std repeated per frame: img_num x RGB, dfd/params.py:24-27)
//   drop_path                     dfd/timm/models/layers/drop.py:84-100 (per-sample mask floor(keep + U[0,1)), x / keep * mask),
//                                 applied before the residual add, efficientnet_blocks.py:343-346
//   classifier dropout            F.dropout(x, p=drop_rate, training), dfd/timm/models/efficientnet.py:346-347
//
// The masks come from a counter-based generator: value(seed, step, stream, index) is a pure function, the (seed, step) pair
// lives in device memory and `dfd_rng_tick` advances the step inside the captured CUDA graph, so every replay draws fresh
// masks without any host involvement.  torch's Philox stream cannot be reproduced bit for bit (it depends on torch's
// launch geometry); parity tests therefore read the masks back and hand the SAME masks to the oracle (exact comparison)
// and check the keep rate statistically.
#include "common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// uint8 NCHW -> 16-bit NCHW, out = (x - mean255[c]) / std255[c]  (fp32 arithmetic, IEEE division, one rounding to T)
// grid: (chunks of a plane, N*C planes); a thread converts 16 consecutive pixels (one 16-byte load, two 16-byte stores)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
input_normalize_kernel(const unsigned char* __restrict__ x, const float* __restrict__ mean255,
                       const float* __restrict__ std255, T* __restrict__ out, int C, long long plane) {
    const long long pl = blockIdx.y;
    const int c = (int)(pl % C);
    const float m = mean255[c], s = std255[c];
    const unsigned char* src = x + pl * plane;
    T* dst = out + pl * plane;
    const bool vec_ok = (plane & 15) == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
    if (vec_ok) {
        const long long nvec = plane >> 4;
        for (long long v = (long long)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
            const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src) + v);
            const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
            float f[16];
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) f[i * 4 + j] = ((float)((w[i] >> (8 * j)) & 0xffu) - m) / s;
            stg16(dst + v * 16, pack8<T>(f));
            stg16(dst + v * 16 + 8, pack8<T>(f + 8));
        }
    } else {
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < plane; i += (long long)gridDim.x * blockDim.x)
            dst[i] = from_f<T>(((float)src[i] - m) / s);
    }
}

// ---------------------------------------------------------------------------------------------
// counter-based uniform generator (splitmix64 finaliser over a 64-bit counter built from step / stream / index)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ float uniform01(unsigned long long seed, unsigned long long step, unsigned stream, unsigned long long idx) {
    unsigned long long h = mix64(seed + 0x9E3779B97F4A7C15ull * (step + 1));
    h = mix64(h ^ (0xD1B54A32D192ED03ull * (stream + 1)));
    h = mix64(h ^ (idx * 0x8CB92BA72F3D8DD7ull + 0x2545F4914F6CDD1Dull));
    return (float)(h >> 40) * (1.0f / 16777216.0f);          // 24 random bits -> [0, 1)
}

struct MaskDesc {
    float* out;          // [rows, width]
    long long rows;
    int width;           // values per random draw (drop path: the channel count, one draw per sample; dropout: 1)
    float keep_prob;
    int stream;          // generator stream id (one per mask tensor)
    int _pad;
};

// out[r, :] = floor(keep + u(r)) / keep      (drop.py:95-99: random_tensor.floor_(); x.div(keep_prob) * random_tensor)
__global__ void rng_masks_kernel(const MaskDesc* __restrict__ table, const long long* __restrict__ state) {
    const MaskDesc d = table[blockIdx.y];
    const unsigned long long seed = (unsigned long long)state[0], step = (unsigned long long)state[1];
    const long long total = d.rows * d.width;
    const float inv = 1.0f / d.keep_prob;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / d.width;
        const float u = uniform01(seed, step, (unsigned)d.stream, (unsigned long long)r);
        d.out[i] = floorf(d.keep_prob + u) * inv;
    }
}

__global__ void rng_tick_kernel(long long* __restrict__ state) { state[1] += 1; }

__global__ void mul_f32_kernel(float* __restrict__ a, const float* __restrict__ b, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) a[i] *= b[i];
}

}  // namespace

extern "C" {

int dfd_input_normalize(const void* x_u8, const float* mean255, const float* std255, void* out, int N, int C, int H, int W,
                        int dt, void* stream) {
    if (N <= 0 || C <= 0 || H <= 0 || W <= 0 || !mean255 || !std255) return dfd_set_error(DFD_ERR_ARG, "dfd_input_normalize: sizes");
    const long long plane = (long long)H * W;
    const long long planes = (long long)N * C;
    if (planes > 65535LL * 32768) return dfd_set_error(DFD_ERR_ARG, "dfd_input_normalize: too many planes");
    int bx = (int)((plane / 16 + 255) / 256);
    if (bx < 1) bx = 1;
    if (bx > 64) bx = 64;
    // planes go on grid.y (<= 65535): fold larger batches into several launches
    for (long long p0 = 0; p0 < planes; p0 += 65535) {
        const int py = (int)((planes - p0 < 65535) ? planes - p0 : 65535);
        if (p0 % C) return dfd_set_error(DFD_ERR_UNSUPPORTED, "dfd_input_normalize: plane split");
        dim3 grid(bx, py);
        const unsigned char* src = (const unsigned char*)x_u8 + p0 * plane;
        if (dt == DFD_DT_BF16)
            input_normalize_kernel<bf16><<<grid, 256, 0, (cudaStream_t)stream>>>(src, mean255, std255, (bf16*)out + p0 * plane, C, plane);
        else if (dt == DFD_DT_FP16)
            input_normalize_kernel<__half><<<grid, 256, 0, (cudaStream_t)stream>>>(src, mean255, std255, (__half*)out + p0 * plane, C, plane);
        else
            return dfd_set_error(DFD_ERR_ARG, "dfd_input_normalize: dtype");
    }
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

// table: device array of { float* out; long long rows; int width; float keep_prob; int stream; int _pad; }
// state: device int64 [seed, step]
int dfd_rng_masks(const void* table, int count, const long long* state, void* stream) {
    if (count <= 0) return DFD_OK;
    if (!table || !state) return dfd_set_error(DFD_ERR_ARG, "dfd_rng_masks: operands");
    rng_masks_kernel<<<dim3(8, count), 256, 0, (cudaStream_t)stream>>>((const MaskDesc*)table, state);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

int dfd_rng_tick(long long* state, void* stream) {
    rng_tick_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

int dfd_mul_f32(float* a, const float* b, long long n, void* stream) {
    if (n <= 0) return DFD_OK;
    long long blocks = (n + 255) / 256;
    if (blocks > 1184) blocks = 1184;
    mul_f32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(a, b, (size_t)n);
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

}  // extern "C"
