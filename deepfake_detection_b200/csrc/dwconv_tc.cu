// Depthwise k x k convolution (stride 1) on the Blackwell tensor cores.
//
// Idea: with the input tile held in shared memory as 8-channel PLANES ([plane][pixel][8 ch], 16 B per pixel), a tap
// (kh, kw) of a depthwise convolution is "the same plane, shifted by kh*IW + kw pixels, times a diagonal matrix":
//
//     out[i, c] = sum_tap  in[i + shift(tap), c] * w[c, tap]          (i = linear pixel index inside the staged tile)
//
// A K-major, un-swizzled UMMA operand is exactly "rows 16 B apart, 8-row groups SBO apart, the two 8-element K halves
// LBO apart", so for a group of 16 channels (two planes) the A operand of tap t is the SAME shared memory, addressed by
// a descriptor whose start address is advanced by shift(t) pixels (LBO = plane stride, SBO = 128 B), and the B operand
// is a 16 x 16 diagonal tile holding w[c, t].  One tcgen05.mma (M = 128 pixels, N = 16 channels, K = 16) per tap and
// 128-pixel chunk accumulates the whole convolution in TMEM: k*k MMAs replace 128*16*k*k FFMAs (plus their shared-memory
// loads and bf16 unpacking), i.e. the CUDA cores only stage (BN + Swish once per element) and run the epilogue.
// Rows of a chunk that fall on halo columns / past the tile are computed and discarded (IW/TW - 1 = 6..12 % waste).
//
// Replaces nn.Conv2d(groups=C) at dfd/timm/models/efficientnet_blocks.py:152-153,283-285 (stride-1 instances).
// Weights are rounded to the activation type (as apex AMP O1 does for every convolution in the reference's GPU path).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int NT = 256;
constexpr int CBT = 64;                 // channels per CTA = 4 groups of 16 = 8 planes of 8
constexpr int TMEM_COLS_DW = 256;       // 3 chunks x 4 groups x 16 columns = 192 -> next power of two
constexpr int MAX_CHUNKS = 3;

__device__ __forceinline__ uint32_t s_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_nosw(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor, SWIZZLE_NONE (INTERLEAVE) K-major: ((8,m),(T,2)) : ((1T,SBO),(1,LBO))
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    return d;                           // layout type 0 = no swizzle
}
__device__ __forceinline__ void umma_f16_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}

struct DwTcGeom {
    int N, H, W, C, pad;
    int TH, TW, IH, IW;
    int tiles_x, tiles_y;
    int nchunks;          // 128-pixel chunks covering (TH-1)*IW + TW output indices
    int PL;               // pixels per plane (incl. slack read by the last chunk)
    int is_bf16;
};

// ACT/AFFINE as in dwconv.cu; FLIP: use the taps reversed (input-gradient of a stride-1 depthwise conv)
template <typename T, int K, int ACT, bool AFFINE>
__global__ void __launch_bounds__(NT, 1)
dwconv_fwd_tc_kernel(const T* __restrict__ x, const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ wgt, T* __restrict__ out, double* __restrict__ dsum,
                     double* __restrict__ dsq, DwTcGeom g) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((128u - (s_addr(smem_raw) & 127u)) & 127u);
    const uint32_t plane_bytes = (uint32_t)g.PL * 16;
    uint8_t* planes = smem;                                   // [8][PL][16 B]
    uint8_t* wtiles = planes + 8 * plane_bytes;               // [4 groups][K*K][512 B]
    uint64_t* bar = reinterpret_cast<uint64_t*>(wtiles + 4 * K * K * 512);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    float* red = reinterpret_cast<float*>(bar + 2);           // [4 q][64 ch] x 2

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tx = blockIdx.x % g.tiles_x, ty = blockIdx.x / g.tiles_x;
    const int c0 = blockIdx.y * CBT;
    const int oy0 = ty * g.TH, ox0 = tx * g.TW;
    const int iy0 = oy0 - g.pad, ix0 = ox0 - g.pad;
    const int Ho = g.H, Wo = g.W;                             // stride 1, symmetric padding

    // ---- one-time setup: diagonal weight tiles, barrier, TMEM ------------------------------------------------
    for (int i = tid; i < 4 * K * K * 512 / 16; i += NT) reinterpret_cast<uint4*>(wtiles)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int i = tid; i < CBT * K * K; i += NT) {
        int c = i / (K * K), tap = i - c * (K * K);
        int grp = c >> 4, n = c & 15;
        float wv = (c0 + c < g.C) ? wgt[(size_t)(c0 + c) * K * K + tap] : 0.f;
        // element (n, k = n) of the K-major un-swizzled 16 x 16 tile: [k/8][n/8][n%8][k%8]
        uint32_t off = (uint32_t)(n >> 3) * 256 + (uint32_t)(n >> 3) * 128 + (uint32_t)(n & 7) * 16 + (uint32_t)(n & 7) * 2;
        *reinterpret_cast<T*>(wtiles + ((size_t)grp * K * K + tap) * 512 + off) = from_f<T>(wv);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s_addr(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(s_addr(tmem_slot)), "r"(TMEM_COLS_DW) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    // cute::UMMA::InstrDescriptor: F32 accumulate, A/B = T, K-major both, N = 16, M = 128
    const uint32_t idesc = (1u << 4) | ((uint32_t)(g.is_bf16 ? 1 : 0) << 7) | ((uint32_t)(g.is_bf16 ? 1 : 0) << 10) |
                           ((uint32_t)(16 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    // staging role: thread -> (plane v, pixel slot)
    const int v = tid & 7;
    const int cbase = c0 + v * 8;
    const bool cvalid = cbase < g.C;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        sc[i] = (AFFINE && cvalid) ? scale[cbase + i] : 1.f;
        sh[i] = (AFFINE && cvalid) ? shift[cbase + i] : 0.f;
    }
    // epilogue role: warp -> (TMEM lane quarter, pair of channel groups)
    const int q = warp & 3, gp = warp >> 2;
    float st_s[2][16], st_q[2][16];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int i = 0; i < 16; i++) { st_s[a][i] = 0.f; st_q[a][i] = 0.f; }

    const int npix = g.IH * g.IW;
    uint32_t phase = 0;
    for (int n = blockIdx.z; n < g.N; n += gridDim.z) {
        // ---- 1. stage act(scale*x+shift) into the planes (zero outside the image) -------------------------------
        const T* img = x + (size_t)n * g.H * g.W * g.C;
        constexpr int UNR = 4;
        for (int base = tid >> 3; base < npix; base += (NT / 8) * UNR) {
            uint4 raw[UNR];
            bool ok[UNR];
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                int pix = base + u * (NT / 8);
                int r = pix / g.IW, c = pix - r * g.IW;
                int iy = iy0 + r, ix = ix0 + c;
                ok[u] = pix < npix && cvalid && iy >= 0 && iy < g.H && ix >= 0 && ix < g.W;
                if (ok[u]) raw[u] = ldg16(img + ((size_t)iy * g.W + ix) * g.C + cbase);
            }
#pragma unroll
            for (int u = 0; u < UNR; u++) {
                int pix = base + u * (NT / 8);
                if (pix >= npix) break;
                uint4 o = make_uint4(0, 0, 0, 0);
                if (ok[u]) {
                    if (AFFINE || ACT != DFD_ACT_NONE) {
                        float f[8];
                        unpack8<T>(raw[u], f);
#pragma unroll
                        for (int i = 0; i < 8; i++) f[i] = act_fwd<ACT>(AFFINE ? fmaf(f[i], sc[i], sh[i]) : f[i]);
                        o = pack8<T>(f);
                    } else {
                        o = raw[u];
                    }
                }
                *reinterpret_cast<uint4*>(planes + (size_t)v * plane_bytes + (size_t)pix * 16) = o;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy stores -> visible to the MMA
        __syncthreads();
        // ---- 2. one thread issues every MMA of the tile ---------------------------------------------------------
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t planes_a = s_addr(planes), w_a = s_addr(wtiles);
            for (int ch = 0; ch < g.nchunks; ch++) {
                for (int grp = 0; grp < 4; grp++) {
                    const uint32_t d_tmem = tmem_base + (uint32_t)(ch * 4 + grp) * 16;
                    const uint32_t a0 = planes_a + (uint32_t)(2 * grp) * plane_bytes + (uint32_t)(ch * 128) * 16;
#pragma unroll
                    for (int tap = 0; tap < K * K; tap++) {
                        const int kh = tap / K, kw = tap - kh * K;
                        uint64_t ad = make_desc_nosw(a0 + (uint32_t)(kh * g.IW + kw) * 16, plane_bytes, 128);
                        uint64_t bd = make_desc_nosw(w_a + (uint32_t)(grp * K * K + tap) * 512, 256, 128);
                        umma_f16_ss(d_tmem, ad, bd, idesc, tap ? 1u : 0u);
                    }
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s_addr(bar)) : "memory");
        }
        // ---- 3. everyone waits for the accumulators ---------------------------------------------------------------
        {
            uint32_t ba = s_addr(bar);
            asm volatile(
                "{\n"
                ".reg .pred p;\n"
                "DW_WAIT:\n"
                "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
                "@p bra DW_DONE;\n"
                "bra DW_WAIT;\n"
                "DW_DONE:\n"
                "}\n" ::"r"(ba), "r"(phase) : "memory");
            phase ^= 1;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- 4. epilogue: TMEM -> registers -> (round, statistics) -> NHWC stores -----------------------------------
        T* oimg = out + (size_t)n * Ho * Wo * g.C;
        for (int ch = 0; ch < g.nchunks; ch++) {
            const int i = ch * 128 + q * 32 + lane;
            const int r = i / g.IW, c = i - r * g.IW;
            const int oy = oy0 + r, ox = ox0 + c;
            const bool pv = r < g.TH && c < g.TW && oy < Ho && ox < Wo;
#pragma unroll
            for (int a = 0; a < 2; a++) {
                const int grp = gp * 2 + a;
                uint32_t vv[16];
                tld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(ch * 4 + grp) * 16, vv);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                const int cg = c0 + grp * 16;
                if (pv && cg < g.C) {
                    uint4 lo, hi;
                    lo.x = pack2<T>(__uint_as_float(vv[0]), __uint_as_float(vv[1]));
                    lo.y = pack2<T>(__uint_as_float(vv[2]), __uint_as_float(vv[3]));
                    lo.z = pack2<T>(__uint_as_float(vv[4]), __uint_as_float(vv[5]));
                    lo.w = pack2<T>(__uint_as_float(vv[6]), __uint_as_float(vv[7]));
                    hi.x = pack2<T>(__uint_as_float(vv[8]), __uint_as_float(vv[9]));
                    hi.y = pack2<T>(__uint_as_float(vv[10]), __uint_as_float(vv[11]));
                    hi.z = pack2<T>(__uint_as_float(vv[12]), __uint_as_float(vv[13]));
                    hi.w = pack2<T>(__uint_as_float(vv[14]), __uint_as_float(vv[15]));
                    T* dst = oimg + ((size_t)oy * Wo + ox) * g.C + cg;
                    stg16(dst, lo);
                    if (cg + 8 < g.C) stg16(dst + 8, hi);
                    float f[16];
                    unpack8<T>(lo, f);
                    unpack8<T>(hi, f + 8);
#pragma unroll
                    for (int k2 = 0; k2 < 16; k2++) {
                        float xv = (cg + k2 < g.C) ? f[k2] : 0.f;
                        st_s[a][k2] += xv;
                        st_q[a][k2] = fmaf(xv, xv, st_q[a][k2]);
                    }
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();            // TMEM and the planes are free for the next image
    }

    // ---- statistics: lanes (pixels) -> warp, the 4 lane-quarter warps of a group pair -> shared memory -> fp64 atomics
    if (dsum) {
#pragma unroll
        for (int a = 0; a < 2; a++)
#pragma unroll
            for (int k2 = 0; k2 < 16; k2++) {
                float s = warp_sum(st_s[a][k2]), qq = warp_sum(st_q[a][k2]);
                if (lane == 0) {
                    red[(q * 64) + (gp * 2 + a) * 16 + k2] = s;
                    red[256 + (q * 64) + (gp * 2 + a) * 16 + k2] = qq;
                }
            }
        __syncthreads();
        if (tid < 64 && c0 + tid < g.C) {
            float s = red[tid] + red[64 + tid] + red[128 + tid] + red[192 + tid];
            float qq = red[256 + tid] + red[320 + tid] + red[384 + tid] + red[448 + tid];
            atomicAdd(stat_slot(dsum, g.C) + c0 + tid, (double)s);
            atomicAdd(stat_slot(dsq, g.C) + c0 + tid, (double)qq);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0)
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS_DW) : "memory");
}

}  // namespace

extern "C" {

// Same contract as dfd_dwconv_fwd (stride 1 only): out = dwconv(act_in(scale*x + shift)), BN statistics of the stored output.
int dfd_dwconv_fwd_tc(const void* x, const float* scale, const float* shift, const float* w, void* out, int N, int H,
                      int W, int C, int k, int stride, int act_in, int dt, double* dsum, double* dsq, void* stream) {
    if (C % 8 || N <= 0 || H <= 0 || W <= 0) return dfd_set_error(DFD_ERR_ARG, "dfd_dwconv_fwd_tc: sizes");
    if (stride != 1 || (k != 3 && k != 5)) return dfd_set_error(DFD_ERR_UNSUPPORTED, "dfd_dwconv_fwd_tc: stride 1, k in {3,5}");
    if ((scale != nullptr) != (act_in == DFD_ACT_SWISH)) return dfd_set_error(DFD_ERR_ARG, "dfd_dwconv_fwd_tc: BN input implies Swish");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "bad dtype");
    DwTcGeom g;
    g.N = N; g.H = H; g.W = W; g.C = C; g.pad = (k - 1) / 2;
    g.TW = W <= 8 ? 8 : (W <= 16 ? 16 : 32);
    g.TH = H < 8 ? H : 8;
    g.IW = g.TW + k - 1;
    g.IH = g.TH + k - 1;
    g.tiles_x = (W + g.TW - 1) / g.TW;
    g.tiles_y = (H + g.TH - 1) / g.TH;
    g.nchunks = ((g.TH - 1) * g.IW + g.TW + 127) / 128;
    if (g.nchunks > MAX_CHUNKS) return dfd_set_error(DFD_ERR_UNSUPPORTED, "dfd_dwconv_fwd_tc: tile too large");
    g.PL = (g.nchunks * 128 + (k - 1) * g.IW + (k - 1) + 7) / 8 * 8;
    if (g.PL < g.IH * g.IW) g.PL = (g.IH * g.IW + 7) / 8 * 8;
    g.is_bf16 = dt == DFD_DT_BF16;
    size_t smem = (size_t)8 * g.PL * 16 + (size_t)4 * k * k * 512 + 16 + 512 * 4 + 128;
    const int cbs = (C + CBT - 1) / CBT, tiles = g.tiles_x * g.tiles_y;
    int gz = (148 * 4 + tiles * cbs - 1) / (tiles * cbs);
    if (gz > N) gz = N;
    if (gz < 1) gz = 1;
    dim3 grid(tiles, cbs, gz);
    cudaStream_t st = (cudaStream_t)stream;
#define LAUNCH_TC(TT, KK, ACT, AFF)                                                                              \
    do {                                                                                                         \
        auto kf = dwconv_fwd_tc_kernel<TT, KK, ACT, AFF>;                                                        \
        static bool attr = false;                                                                                \
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; } \
        kf<<<grid, NT, smem, st>>>((const TT*)x, scale, shift, w, (TT*)out, dsum, dsq, g);                       \
    } while (0)
    if (dt == DFD_DT_BF16) {
        if (k == 3) { if (scale) LAUNCH_TC(bf16, 3, DFD_ACT_SWISH, true); else LAUNCH_TC(bf16, 3, DFD_ACT_NONE, false); }
        else { if (scale) LAUNCH_TC(bf16, 5, DFD_ACT_SWISH, true); else LAUNCH_TC(bf16, 5, DFD_ACT_NONE, false); }
    } else {
        if (k == 3) { if (scale) LAUNCH_TC(__half, 3, DFD_ACT_SWISH, true); else LAUNCH_TC(__half, 3, DFD_ACT_NONE, false); }
        else { if (scale) LAUNCH_TC(__half, 5, DFD_ACT_SWISH, true); else LAUNCH_TC(__half, 5, DFD_ACT_NONE, false); }
    }
#undef LAUNCH_TC
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

}  // extern "C"
