// Pointwise (1x1) convolution GEMM on the Blackwell tensor-core path: TMA -> shared memory (128B swizzle) ->
// tcgen05.mma (cta_group::1, kind::f16, fp32 accumulators in TMEM) -> tcgen05.ld epilogue -> swizzled smem ->
// TMA store, with the per-channel BatchNorm statistics of the stored tile reduced in the same epilogue.
//
//   C[M,N] = A[M,K] * B[N,K]^T          A: NHWC activations / output gradients (K-contiguous rows)
//                                        B: conv weight [Cout,Cin] (forward) or its transpose [Cin,Cout] (dgrad)
//
// Replaces nn.Conv2d 1x1 (cuDNN/cuBLAS in the reference: dfd/timm/models/efficientnet_blocks.py:165,277,299,
// efficientnet.py:292, resnet.py:192,199) and its input-gradient (autograd, train.py:634-636).
//
// Shape regime: M = N*H*W is 12.5k .. 3.2M, K and N are 16 .. 1280 -> every instance is
// HBM-bound (AI << 219 FLOP/B), so the design goal is to stream A and C at HBM rate, not MMA peak:
//   * persistent CTAs (one per SM), tiles 128 x BLOCK_N, BLOCK_N = whole N when N <= 256 (A is read once);
//   * K/N/M tails need no padding copies: TMA zero-fills out-of-bounds loads and clips stores;
//   * 2 TMEM accumulator stages so the epilogue of tile i overlaps the loads + MMAs of tile i+1;
//   * warp roles: w0 TMA producer, w1 MMA issuer, w2 TMEM allocator, w4-7 epilogue (TMEM lane quarter = warp%4).
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "bn_finalize.cuh"

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 x 2 B = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int SLAB = 64;             // epilogue / TMA-store column slab
constexpr int NUM_THREADS = 384;      // TMA, MMA, TMEM-alloc, spare warp + two epilogue warpgroups
constexpr int EPI_THREADS = 128;      // per epilogue warpgroup
constexpr int TMEM_COLS = 256;         // per CTA; two CTAs share an SM's 512 columns
constexpr int ACC_STRIDE = 128;      // TMEM column offset between the two accumulator stages
constexpr int MAX_BLOCK_N = 128;
constexpr int SMEM_BUDGET = 200 * 1024;   // one CTA per SM (the epilogue holds a whole 128 x 128 fp32 tile in registers)

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_addr(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t addr = smem_addr(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}

// ---- TMA ------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_addr(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(smem_addr(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// TMA reduction store: global (16-bit) += shared tile, element-wise add performed in L2 (round to nearest in the tensor's type)
__device__ __forceinline__ void tma_reduce_add_4d(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.reduce.async.bulk.tensor.4d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(smem_addr(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void tma_store_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

// ---- tcgen05 --------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_addr(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_addr(bar)) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16/fp16 inputs, fp32 accumulate
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile in shared memory, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart.
// (cute::UMMA::SmemDescriptor: start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46), version=1 [46,48), layout [61,64))
__device__ __forceinline__ uint64_t make_kmajor_sw128_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;                 // LBO: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;       // SBO = 1024 B between 8-row core-matrix groups
    d |= (uint64_t)1 << 46;                 // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                 // SWIZZLE_128B
    return d;
}
// cute::UMMA::InstrDescriptor for kind::f16: c_format F32 (1) [4,6), a/b format [7,10)/[10,13) (F16=0, BF16=1),
// a/b major K (0) bits 15/16, N>>3 [17,23), M>>4 [24,29)
__device__ __forceinline__ uint32_t make_idesc(int is_bf16, int n) {
    uint32_t d = 0;
    d |= 1u << 4;
    d |= (uint32_t)(is_bf16 ? 1 : 0) << 7;
    d |= (uint32_t)(is_bf16 ? 1 : 0) << 10;
    d |= (uint32_t)(n >> 3) << 17;
    d |= (uint32_t)(BLOCK_M >> 4) << 24;
    return d;
}

__device__ __forceinline__ void epi_barrier(int wg) { asm volatile("bar.sync %0, %1;" ::"r"(wg + 1), "n"(EPI_THREADS) : "memory"); }

struct BlockDiagDesc {
    const void* src;
    void* dst;
    int N, K, pack, _pad;
};

struct TcParams {
    int M, N, K;
    int block_n;          // multiple of 16, <= MAX_BLOCK_N
    int num_m_tiles, num_n_tiles, num_k_blocks;
    int stages;
    int is_bf16;
    double* dsum;         // optional [DFD_STAT_SLOTS][N]
    double* dsq;
    int stat_n;           // statistics channel of output column c is c % stat_n (== N unless rows are packed)
    const BnFinDesc* fin; // optional: the last CTA finalises the BatchNorm behind this convolution (bn_finalize.cuh)
    // ---- implicit-GEMM convolution mode (conv != 0): k x k, stride 1, "same" padding, NHWC ------------------------------
    // The A operand is never materialised: an M tile is a (TW x TH x TN) patch of output pixels, and for every tap (kh, kw) and
    // 64-channel block the producer issues ONE 4-D TMA load of the input box shifted by the tap - rows outside the image
    // arrive as zeros (TMA out-of-bounds fill), which IS the padding. K runs over (tap, channel block); B is the packed weight
    // [Cout][kh][kw][Cin]. The C tile goes back through a 4-D map over the output tensor (the store clips at the borders).
    int conv;
    int cv_TW, cv_TH, cv_TN;          // output patch of one M tile (cv_rows = TW * TH * TN <= 128 rows are real)
    int cv_rows;
    int cv_tiles_x, cv_tiles_y;       // patches per image row / column; M tile index = (n_tile * tiles_y + ty) * tiles_x + tx
    int cv_W, cv_H, cv_N;             // output (= input) extent
    int cv_k, cv_pad, cv_cpb;         // kernel size, padding, 64-channel blocks per tap
    int cv_S;                         // convolution stride (1 or 2): the A box starts at S * patch origin + tap - pad
    // explicit tap list (cv_ntaps > 0; the parity classes of a strided input gradient): tap t reads the A box at patch origin +
    // (cv_tox[t], cv_toy[t]) against weight k-blocks (cv_tk[t] * cv_cpb + cb)
    int cv_ntaps;
    int cv_tox[9], cv_toy[9], cv_tk[9];
    int cv_accum;                     // conv mode: the output tile is ADDED to the destination (TMA reduction store)
    int dbg;              // debug switches (DFD_DBG env): 1 = skip TMA store, 2 = skip stats pass, 4 = skip slabs, 8 = A from L2
    long long* ts;        // optional trace (DFD_TS env): CTA 0 records clock64 at 8 pipeline points for its first 32 tiles
};

template <typename T>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
               const __grid_constant__ CUtensorMap tmap_c, const TcParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries of the shared address space
    uint8_t* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    // carve-up (all tile bases 1024-byte aligned): [A stages][B stages][2 warpgroups x 2 C slabs][barriers]
    const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;
    const uint32_t b_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
    const uint32_t b_stride = (b_bytes + 1023) & ~1023u;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + (size_t)p.stages * a_bytes;
    uint8_t* smem_c = smem_b + (size_t)p.stages * b_stride;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 4 * BLOCK_M * SLAB * 2);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* tmem_full = bars + 16;
    uint64_t* tmem_empty = bars + 18;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 20);
    float* red_all = reinterpret_cast<float*>(bars + 22);    // [warpgroup][2][2][64]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmap_a);
        prefetch_tmap(&tmap_b);
        prefetch_tmap(&tmap_c);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; i++) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        for (int i = 0; i < 2; i++) { mbar_init(tmem_full + i, 1); mbar_init(tmem_empty + i, EPI_THREADS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        // ================= TMA producer =================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            int lt = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, lt++) {
                const int m_idx = tile / p.num_n_tiles, n_idx = tile - m_idx * p.num_n_tiles;
                int cx0 = 0, cy0 = 0, cn0 = 0;
                if (p.conv) {
                    const int txi = m_idx % p.cv_tiles_x, r = m_idx / p.cv_tiles_x;
                    cx0 = txi * p.cv_TW; cy0 = (r % p.cv_tiles_y) * p.cv_TH; cn0 = (r / p.cv_tiles_y) * p.cv_TN;
                }
                for (int kb = 0; kb < p.num_k_blocks; kb++) {
                    mbar_wait(empty_bar + stage, phase ^ 1);
                    if (p.ts && kb == 0 && blockIdx.x == 0 && lt < 32) p.ts[lt * 8 + 0] = clock64();
                    int bk = kb * BLOCK_K;
                    if (p.conv) {
                        const int tap = kb / p.cv_cpb, cb = kb - tap * p.cv_cpb;
                        int ax, ay;
                        if (p.cv_ntaps) {
                            ax = cx0 + p.cv_tox[tap]; ay = cy0 + p.cv_toy[tap];
                            bk = (p.cv_tk[tap] * p.cv_cpb + cb) * BLOCK_K;
                        } else {
                            const int kh = tap / p.cv_k, kw = tap - kh * p.cv_k;
                            ax = cx0 * p.cv_S + kw - p.cv_pad; ay = cy0 * p.cv_S + kh - p.cv_pad;
                        }
                        mbar_arrive_expect_tx(full_bar + stage, (uint32_t)p.cv_rows * 128u + b_bytes);
                        tma_load_4d(smem_a + (size_t)stage * a_bytes, &tmap_a, full_bar + stage, cb * BLOCK_K, ax, ay, cn0);
                    } else {
                        mbar_arrive_expect_tx(full_bar + stage, a_bytes + b_bytes);
                        tma_load_2d(smem_a + (size_t)stage * a_bytes, &tmap_a, full_bar + stage, kb * BLOCK_K, (p.dbg & 8) ? 0 : m_idx * BLOCK_M);
                    }
                    tma_load_2d(smem_b + (size_t)stage * b_stride, &tmap_b, full_bar + stage, bk, n_idx * p.block_n);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.is_bf16, p.block_n);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            int lt = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, lt++) {
                mbar_wait(tmem_empty + acc, acc_phase ^ 1);
                tc_fence_after();
                const bool rec = p.ts && blockIdx.x == 0 && lt < 32;
                if (rec) p.ts[lt * 8 + 1] = clock64();
                const uint32_t d_tmem = tmem_base + acc * ACC_STRIDE;
                for (int kb = 0; kb < p.num_k_blocks; kb++) {
                    mbar_wait(full_bar + stage, phase);
                    tc_fence_after();
                    if (rec && kb == 0) p.ts[lt * 8 + 2] = clock64();
                    const uint32_t a_addr = smem_addr(smem_a + (size_t)stage * a_bytes);
                    const uint32_t b_addr = smem_addr(smem_b + (size_t)stage * b_stride);
                    int krem = p.K - kb * BLOCK_K;
                    int nk = krem >= BLOCK_K ? BLOCK_K / UMMA_K : (krem + UMMA_K - 1) / UMMA_K;
                    for (int k = 0; k < nk; k++) {
                        uint64_t adesc = make_kmajor_sw128_desc(a_addr + k * UMMA_K * 2);
                        uint64_t bdesc = make_kmajor_sw128_desc(b_addr + k * UMMA_K * 2);
                        umma_f16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(empty_bar + stage);          // frees the smem stage when the MMAs have read it
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
                umma_commit(tmem_full + acc);                // accumulator complete -> epilogue
                if (rec) p.ts[lt * 8 + 3] = clock64();
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: two warpgroups, one per accumulator stage =================
        // The CTA's i-th tile accumulates in TMEM stage i & 1 and is drained by warpgroup i & 1, so two tiles are in the
        // epilogue at once (measured: one warpgroup needs ~2500 cycles per 128 x 96 tile - TMEM load, pack, staging,
        // TMA store - which alone caps output-heavy GEMMs at a third of the HBM rate).
        const int wg = (warp - 4) >> 2;                                // 0 / 1
        const int et = (threadIdx.x - 128) & 127;                      // 0..127 == TMEM lane == tile row
        const int q = warp & 3;                                        // TMEM lane quarter this warp may access
        const bool leader = et == 0;
        uint8_t* my_c = smem_c + (size_t)wg * (2 * BLOCK_M * SLAB * 2);
        float* red = red_all + wg * 256;
        uint32_t acc_phase = 0;
        uint32_t slab_count = 0;
        const int nslabs = (p.block_n + SLAB - 1) / SLAB;
        const bool keep = p.num_n_tiles == 1;          // column identity is fixed -> keep sums in registers
        float ks[2] = {0.f, 0.f}, kq[2] = {0.f, 0.f};
        const uint32_t t_base = tmem_base + wg * ACC_STRIDE + ((uint32_t)(q * 32) << 16);
        int lt = wg;                                    // index of the tile within this CTA's sequence
        for (int tile = blockIdx.x + wg * gridDim.x; tile < num_tiles; tile += 2 * gridDim.x, lt += 2) {
            const int m_idx = tile / p.num_n_tiles, n_idx = tile - m_idx * p.num_n_tiles;
            const bool rec = p.ts && leader && blockIdx.x == 0 && lt < 32;
            // convolution mode: this thread's row is output pixel (n0 + tn, y0 + ty, x0 + tx) of the patch; rows beyond the patch
            // or the image hold garbage accumulators and are zeroed (they would otherwise enter the BatchNorm statistics)
            int cx0 = 0, cy0 = 0, cn0 = 0;
            bool row_ok = true;
            if (p.conv) {
                const int txi = m_idx % p.cv_tiles_x, r = m_idx / p.cv_tiles_x;
                cx0 = txi * p.cv_TW; cy0 = (r % p.cv_tiles_y) * p.cv_TH; cn0 = (r / p.cv_tiles_y) * p.cv_TN;
                const int px = et % p.cv_TW, q2 = et / p.cv_TW;
                const int py = q2 % p.cv_TH, pn = q2 / p.cv_TH;
                row_ok = et < p.cv_rows && cx0 + px < p.cv_W && cy0 + py < p.cv_H && cn0 + pn < p.cv_N;
            }
            if (rec) p.ts[lt * 8 + 4] = clock64();
            mbar_wait(tmem_full + wg, acc_phase);
            tc_fence_after();
            if (rec) p.ts[lt * 8 + 5] = clock64();
            uint32_t v[SLAB / 16][16];
#pragma unroll
            for (int q4 = 0; q4 < SLAB / 16; q4++)
                if (q4 * 16 < p.block_n) tmem_ld16(t_base + q4 * 16, v[q4]);
#pragma unroll
            for (int s = 0; s < MAX_BLOCK_N / SLAB; s++) {
                if (s >= nslabs || (p.dbg & 4)) break;
                uint8_t* cbuf = my_c + (size_t)(slab_count & 1) * (BLOCK_M * SLAB * 2);
                const int ncols = min(SLAB, p.block_n - s * SLAB);
                if (leader) tma_store_wait_read<1>();       // the store that last used this buffer has drained
                tmem_ld_wait();
                epi_barrier(wg);
#pragma unroll
                for (int q4 = 0; q4 < SLAB / 16; q4++) {
                    if (q4 * 16 < ncols) {
                        uint32_t (&w)[16] = v[q4];
                        if (!row_ok) {
#pragma unroll
                            for (int z = 0; z < 16; z++) w[z] = 0u;
                        }
                        uint4 lo, hi;
                        lo.x = pack2<T>(__uint_as_float(w[0]), __uint_as_float(w[1]));
                        lo.y = pack2<T>(__uint_as_float(w[2]), __uint_as_float(w[3]));
                        lo.z = pack2<T>(__uint_as_float(w[4]), __uint_as_float(w[5]));
                        lo.w = pack2<T>(__uint_as_float(w[6]), __uint_as_float(w[7]));
                        hi.x = pack2<T>(__uint_as_float(w[8]), __uint_as_float(w[9]));
                        hi.y = pack2<T>(__uint_as_float(w[10]), __uint_as_float(w[11]));
                        hi.z = pack2<T>(__uint_as_float(w[12]), __uint_as_float(w[13]));
                        hi.w = pack2<T>(__uint_as_float(w[14]), __uint_as_float(w[15]));
                        const int j = q4 * 2;            // logical 16-byte chunk index within the 128-byte row
                        uint8_t* row = cbuf + et * 128;
                        *reinterpret_cast<uint4*>(row + ((j ^ (et & 7)) << 4)) = lo;
                        *reinterpret_cast<uint4*>(row + (((j + 1) ^ (et & 7)) << 4)) = hi;
                    }
                }
                if (s + 1 < nslabs) {
                    // the next slab's TMEM loads fly while this one is fenced, stored and reduced
#pragma unroll
                    for (int q4 = 0; q4 < SLAB / 16; q4++)
                        if ((s + 1) * SLAB + q4 * 16 < p.block_n) tmem_ld16(t_base + (s + 1) * SLAB + q4 * 16, v[q4]);
                } else {
                    // all TMEM reads of this accumulator are done: hand it back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(tmem_empty + wg);
                    if (rec) p.ts[lt * 8 + 6] = clock64();
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                epi_barrier(wg);
                if (leader && !(p.dbg & 1)) {
                    if (p.conv && p.cv_accum) tma_reduce_add_4d(&tmap_c, cbuf, n_idx * p.block_n + s * SLAB, cx0, cy0, cn0);
                    else if (p.conv) tma_store_4d(&tmap_c, cbuf, n_idx * p.block_n + s * SLAB, cx0, cy0, cn0);
                    else tma_store_2d(&tmap_c, cbuf, n_idx * p.block_n + s * SLAB, m_idx * BLOCK_M);
                    tma_store_commit();
                }
                if (p.dsum && !(p.dbg & 2)) {
                    // column statistics of the stored (rounded) slab; rows past M are exact zeros
                    const int c = et & 63, half = et >> 6;
                    float sum = 0.f, sq = 0.f;
                    if (c < ncols) {
                        const uint8_t* colp = cbuf + (c & 7) * 2;
                        const int jc = c >> 3;
                        float s1 = 0.f, q1 = 0.f;            // two chains: the adds are latency-, not throughput-bound
#pragma unroll 8
                        for (int r = half * 64; r < half * 64 + 64; r += 2) {
                            float x0 = to_f<T>(*reinterpret_cast<const T*>(colp + r * 128 + ((jc ^ (r & 7)) << 4)));
                            float x1 = to_f<T>(*reinterpret_cast<const T*>(colp + (r + 1) * 128 + ((jc ^ ((r + 1) & 7)) << 4)));
                            sum += x0; s1 += x1;
                            sq = fmaf(x0, x0, sq); q1 = fmaf(x1, x1, q1);
                        }
                        sum += s1; sq += q1;
                    }
                    red[(0 * 2 + half) * 64 + c] = sum;
                    red[(1 * 2 + half) * 64 + c] = sq;
                    epi_barrier(wg);
                    if (et < 64) {
                        float ts = red[et] + red[64 + et], tq = red[128 + et] + red[192 + et];
                        if (keep) { ks[s] += ts; kq[s] += tq; }
                        else {
                            int gc = n_idx * p.block_n + s * SLAB + et;
                            if (et < ncols && gc < p.N) {
                                atomicAdd(stat_slot(p.dsum, p.stat_n) + gc % p.stat_n, (double)ts);
                                atomicAdd(stat_slot(p.dsq, p.stat_n) + gc % p.stat_n, (double)tq);
                            }
                        }
                    }
                }
                slab_count++;
            }
            if (p.dbg & 4) { tmem_ld_wait(); tc_fence_before(); mbar_arrive(tmem_empty + wg); }
            if (rec) p.ts[lt * 8 + 7] = clock64();
            acc_phase ^= 1;
        }
        if (p.dsum && keep && et < 64) {
#pragma unroll
            for (int s = 0; s < MAX_BLOCK_N / SLAB; s++) {
                int gc = s * SLAB + et;
                if (s < nslabs && gc < p.N && gc < p.block_n) {
                    atomicAdd(stat_slot(p.dsum, p.stat_n) + gc % p.stat_n, (double)ks[s]);
                    atomicAdd(stat_slot(p.dsq, p.stat_n) + gc % p.stat_n, (double)kq[s]);
                }
            }
        }
        if (leader) tma_store_wait_read<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
    bn_finalize_tail(p.fin, threadIdx.x, NUM_THREADS);
}

// =============================================================================================
// 1x1 conv weight gradient on tcgen05: dW[Nw,Kw] (fp32, accumulated) += G[M,Nw]^T * X[M,Kw].
// The contraction runs over the NHWC rows, so BOTH operands are MN-major for the MMA: a TMA box of {64 channels, 64 rows}
// with 128-byte swizzle is already the canonical MN-major SW128 atom sequence (64 contiguous MN elements per K row, 8-row
// groups 1024 B apart = SBO; 64-channel column blocks one box apart = LBO), so the tiles go from NHWC memory to the tensor
// core without any transpose. One CTA = one (128 x block_n) tile of dW and one contiguous range of 64-row blocks (split-K);
// the fp32 accumulator lives in TMEM for the whole range and is flushed once with red.global.add.
// =============================================================================================
constexpr int WG_KP = 64;                      // rows (pixels) per pipeline stage
constexpr int WG_BOX_BYTES = WG_KP * 128;      // one {64 ch, 64 rows} box

__device__ __forceinline__ uint64_t make_mnmajor_sw128_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;     // LBO: next 64-element block along M / N
    d |= (uint64_t)(1024 >> 4) << 32;                     // SBO: next 8-row group along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}

struct WgParams {
    long long M;
    int Nw, Kw;
    int block_n;            // dW columns per tile: multiple of 16, <= 128
    long long kblocks;      // ceil(M / 64)
    long long kb_per_split;
    int stages;
    int is_bf16;
    // order-deterministic mode (part != NULL): split z stores its fp32 partial tile at part[z][Nw][Kw] with plain stores and
    // dW is left alone; dfd_ordered_reduce adds the partials in split order later. part == NULL: red.global.add from every
    // split straight into dW (the order of the fp32 additions then varies run to run).
    float* part;
    int splits;
    // ---- implicit-GEMM convolution mode (conv != 0): G = dY [N,H,W,Cout] and X = the layer input [N,H,W,Cin] through 4-D
    // tensor maps; one pipeline stage = one patch of cv_rows <= 64 output pixels (TW x TH x TN); column block j of dW
    // (64 wide) = (tap = j / cpb, channel block = j % cpb): its X box is the patch shifted by the tap, padding and rows past
    // the image arrive as zeros. Rows cv_rows..63 of every box are never written by TMA and are zeroed once at kernel start.
    int conv;
    int cv_TW, cv_TH, cv_TN, cv_rows;
    int cv_tiles_x, cv_tiles_y;
    int cv_k, cv_pad, cv_cpb, cv_S;
};

template <typename T>
__global__ void __launch_bounds__(256, 2)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                float* __restrict__ dW, const WgParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_addr(smem_raw) & 1023u)) & 1023u);
    const int nbox_b = p.block_n > 64 ? 2 : 1;
    const uint32_t a_bytes = 2 * WG_BOX_BYTES, b_bytes = (uint32_t)nbox_b * WG_BOX_BYTES;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem_a + (size_t)p.stages * a_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + (size_t)p.stages * b_bytes);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + 8;
    uint64_t* tmem_full = bars + 16;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * 128, n0 = blockIdx.y * p.block_n;
    const long long kb0 = (long long)blockIdx.z * p.kb_per_split;
    long long kb1 = kb0 + p.kb_per_split;
    if (kb1 > p.kblocks) kb1 = p.kblocks;
    if (kb0 >= kb1) return;                                   // uniform per CTA
    const bool a_box1 = m0 + 64 < p.Nw;                       // second 64-channel block of the tile exists
    const bool b_box1 = nbox_b == 2 && n0 + 64 < p.Kw;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmap_g); prefetch_tmap(&tmap_x); }
    if (p.conv && p.cv_rows < WG_KP) {
        uint4* z = reinterpret_cast<uint4*>(smem_a);
        const int n16 = (int)((size_t)p.stages * (a_bytes + b_bytes) / 16);
        for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < p.stages; i++) { mbar_init(full_bar + i, 1); mbar_init(empty_bar + i, 1); }
        mbar_init(tmem_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) tmem_alloc(tmem_ptr, 128);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t box_bytes = p.conv ? (uint32_t)p.cv_rows * 128u : (uint32_t)WG_BOX_BYTES;
            const uint32_t tx = ((a_box1 ? 2u : 1u) + (b_box1 ? 2u : 1u)) * box_bytes;
            // convolution mode: (tap, channel block) of this tile's one or two 64-column blocks
            int bt_c[2] = {0, 0}, bt_dx[2] = {0, 0}, bt_dy[2] = {0, 0};
            if (p.conv) {
#pragma unroll
                for (int i = 0; i < 2; i++) {
                    const int jb = n0 / 64 + i, tap = jb / p.cv_cpb;
                    bt_c[i] = (jb - tap * p.cv_cpb) * 64;
                    bt_dy[i] = tap / p.cv_k - p.cv_pad;
                    bt_dx[i] = tap % p.cv_k - p.cv_pad;
                }
            }
            for (long long kb = kb0; kb < kb1; kb++) {
                mbar_wait(empty_bar + stage, phase ^ 1);
                mbar_arrive_expect_tx(full_bar + stage, tx);
                uint8_t* a = smem_a + (size_t)stage * a_bytes;
                uint8_t* b = smem_b + (size_t)stage * b_bytes;
                if (p.conv) {
                    const int pt = (int)kb, txi = pt % p.cv_tiles_x, r = pt / p.cv_tiles_x;
                    const int cx0 = txi * p.cv_TW, cy0 = (r % p.cv_tiles_y) * p.cv_TH, cn0 = (r / p.cv_tiles_y) * p.cv_TN;
                    tma_load_4d(a, &tmap_g, full_bar + stage, m0, cx0, cy0, cn0);
                    if (a_box1) tma_load_4d(a + WG_BOX_BYTES, &tmap_g, full_bar + stage, m0 + 64, cx0, cy0, cn0);
                    tma_load_4d(b, &tmap_x, full_bar + stage, bt_c[0], cx0 * p.cv_S + bt_dx[0], cy0 * p.cv_S + bt_dy[0], cn0);
                    if (b_box1) tma_load_4d(b + WG_BOX_BYTES, &tmap_x, full_bar + stage, bt_c[1], cx0 * p.cv_S + bt_dx[1], cy0 * p.cv_S + bt_dy[1], cn0);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                    continue;
                }
                const int row = (int)(kb * WG_KP);
                tma_load_2d(a, &tmap_g, full_bar + stage, m0, row);
                if (a_box1) tma_load_2d(a + WG_BOX_BYTES, &tmap_g, full_bar + stage, m0 + 64, row);
                tma_load_2d(b, &tmap_x, full_bar + stage, n0, row);
                if (b_box1) tma_load_2d(b + WG_BOX_BYTES, &tmap_x, full_bar + stage, n0 + 64, row);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: fp32 accumulate, 16-bit inputs, A and B MN-major (bits 15 / 16), N, M = 128
            uint32_t idesc = make_idesc(p.is_bf16, p.block_n) | (1u << 15) | (1u << 16);
            int stage = 0;
            uint32_t phase = 0;
            for (long long kb = kb0; kb < kb1; kb++) {
                mbar_wait(full_bar + stage, phase);
                tc_fence_after();
                const uint32_t a_addr = smem_addr(smem_a + (size_t)stage * a_bytes);
                const uint32_t b_addr = smem_addr(smem_b + (size_t)stage * b_bytes);
#pragma unroll
                for (int k = 0; k < WG_KP / UMMA_K; k++) {
                    // 16 rows of K = two 8-row groups = 2048 bytes further into every box
                    const uint64_t adesc = make_mnmajor_sw128_desc(a_addr + k * (UMMA_K * 128), WG_BOX_BYTES);
                    const uint64_t bdesc = make_mnmajor_sw128_desc(b_addr + k * (UMMA_K * 128), WG_BOX_BYTES);
                    umma_f16(tmem_base, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                }
                umma_commit(empty_bar + stage);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(tmem_full);
        }
    } else if (warp >= 4) {
        const int q = warp & 3;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        const int row = m0 + q * 32 + lane;                  // dW row (output channel) of this TMEM lane
        const uint32_t t_base = tmem_base + ((uint32_t)(q * 32) << 16);
        if (p.part) {
            // ---- deterministic path: plain stores of this split's partial tile (summed later in split order) ----
            float* mine = p.part + (size_t)blockIdx.z * p.Nw * p.Kw;
            for (int c = 0; c < p.block_n; c += 16) {
                uint32_t v[16];
                tmem_ld16(t_base + c, v);
                tmem_ld_wait();
                if (row < p.Nw) {
                    float* dst = mine + (size_t)row * p.Kw + n0 + c;
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        if (n0 + c + j < p.Kw)
                            *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                                              __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                }
            }
        } else
        for (int c = 0; c < p.block_n; c += 16) {
            uint32_t v[16];
            tmem_ld16(t_base + c, v);
            tmem_ld_wait();
            if (row < p.Nw) {
                // 16-byte vector reductions (Kw % 8 == 0 keeps every group of 4 columns aligned and all-or-nothing): the
                // small-M layers are bound by the NUMBER of L2 reduction ops, not by their bytes
                float* dst = dW + (size_t)row * p.Kw + n0 + c;
#pragma unroll
                for (int j = 0; j < 16; j += 4)
                    if (n0 + c + j < p.Kw)
                        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst + j), "f"(__uint_as_float(v[j])),
                                     "f"(__uint_as_float(v[j + 1])), "f"(__uint_as_float(v[j + 2])), "f"(__uint_as_float(v[j + 3]))
                                     : "memory");
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, 128);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D row-major [rows, cols] 16-bit tensor, box = [box_rows, 64 cols], 128-byte swizzle, OOB -> zeros / clipped
static int make_map(CUtensorMap* m, const void* base, long long rows, int cols, int box_rows, int is_bf16) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return dfd_set_error(DFD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[160];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d box_rows=%d", (int)r, rows, cols, box_rows);
        return dfd_set_error(DFD_ERR_CUDA, buf);
    }
    return DFD_OK;
}


// 4-D NHWC tensor [N, H, W, C] (16-bit) seen as dims {C, W, H, N}; box = {64 channels, bw, bh, bn}, 128-byte swizzle, OOB -> zeros
struct PixelView { long long sW, sH, sN; };      // byte strides between pixels / rows / images (a parity class of a tensor)
static int make_map_nhwc(CUtensorMap* m, const void* base, int N, int H, int W, int C, int bw, int bh, int bn, int is_bf16,
                         int es = 1, const PixelView* pv = nullptr) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return dfd_set_error(DFD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
    if (pv) { strides[0] = (cuuint64_t)pv->sW; strides[1] = (cuuint64_t)pv->sH; strides[2] = (cuuint64_t)pv->sN; }
    // es = traversal stride in W and H (strided convolution): a box of (b - 1) * es + 1 tensor elements delivers b of them
    cuuint32_t box[4] = {(cuuint32_t)BLOCK_K, (cuuint32_t)((bw - 1) * es + 1), (cuuint32_t)((bh - 1) * es + 1), (cuuint32_t)bn};
    cuuint32_t estr[4] = {1, (cuuint32_t)es, (cuuint32_t)es, 1};
    CUresult r = fn(m, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        char buf[200];
        snprintf(buf, sizeof(buf), "cuTensorMapEncodeTiled (4-D) failed (%d) N=%d H=%d W=%d C=%d box=%dx%dx%d", (int)r, N, H, W, C, bw, bh, bn);
        return dfd_set_error(DFD_ERR_CUDA, buf);
    }
    return DFD_OK;
}

// Implicit-GEMM convolution on the kernel above (conv mode): y[N,H,W,Cout] = conv_{k x k, stride 1, pad (k-1)/2}(x[N,H,W,Cin]),
// wpk = packed weight [Cout][kh][kw][Cin] (K-major rows of k*k*Cin). Cin % 64 == 0 keeps every 64-channel K block inside one tap.
struct ConvTaps { int n, ox[9], oy[9], kidx[9]; };
static int launch_conv_tc(const void* x, const void* wpk, void* y, int N, int Hin, int Win, int Cin, int Cout, int k, int S,
                          int dt, double* dsum, double* dsq, const void* fin, void* stream, const ConvTaps* taps = nullptr,
                          int outH = 0, int outW = 0, const PixelView* out_view = nullptr, int wcols = 0, int accum = 0) {
    TcParams p;
    p.fin = (const BnFinDesc*)fin;
    p.conv = 1;
    p.cv_S = S;
    p.cv_accum = accum;
    p.cv_ntaps = taps ? taps->n : 0;
    if (taps) for (int t = 0; t < taps->n; t++) { p.cv_tox[t] = taps->ox[t]; p.cv_toy[t] = taps->oy[t]; p.cv_tk[t] = taps->kidx[t]; }
    // output extents (explicit tap list: the caller's output grid, e.g. one parity class of the input-gradient tensor)
    const int H = taps ? outH : (Hin + 2 * ((k - 1) / 2) - k) / S + 1, W = taps ? outW : (Win + 2 * ((k - 1) / 2) - k) / S + 1;
    // output patch of an M tile: whole rows when they fit (W <= 128), as many rows as 128 / W allows, split evenly over the
    // image height; images stacked when a whole image is smaller than half a tile (7 x 7 -> two images per tile)
    int TW = W <= 128 ? W : 128;
    int maxTH = 128 / TW; if (maxTH < 1) maxTH = 1;
    int ty = (H + maxTH - 1) / maxTH;
    int TH = (H + ty - 1) / ty;
    int TN = TH == H && TW == W ? 128 / (TW * TH) : 1;
    if (TN < 1) TN = 1;
    if (TN > N) TN = N;
    p.cv_TW = TW; p.cv_TH = TH; p.cv_TN = TN; p.cv_rows = TW * TH * TN;
    p.cv_tiles_x = (W + TW - 1) / TW; p.cv_tiles_y = (H + TH - 1) / TH;
    p.cv_W = W; p.cv_H = H; p.cv_N = N;
    p.cv_k = k; p.cv_pad = (k - 1) / 2; p.cv_cpb = Cin / BLOCK_K;
    const int K = taps ? wcols : k * k * Cin;          // columns of the packed weight matrix
    p.M = N * H * W; p.N = Cout; p.K = K;
    p.stat_n = Cout;
    p.is_bf16 = dt == DFD_DT_BF16;
    p.block_n = Cout <= MAX_BLOCK_N ? ((Cout + 15) / 16) * 16 : MAX_BLOCK_N;
    p.num_m_tiles = p.cv_tiles_x * p.cv_tiles_y * ((N + TN - 1) / TN);
    p.num_n_tiles = cdiv(Cout, p.block_n);
    p.num_k_blocks = (taps ? taps->n : k * k) * p.cv_cpb;
    p.dsum = dsum; p.dsq = dsq;
    { const char* e = getenv("DFD_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.ts = nullptr;
    const int a_bytes = BLOCK_M * BLOCK_K * 2;
    const int b_stride = ((p.block_n * BLOCK_K * 2) + 1023) & ~1023;
    const int fixed = 4 * BLOCK_M * SLAB * 2 + 22 * 8 + 2 * 4 * 64 * 4 + 1024;
    int stages = (SMEM_BUDGET - fixed) / (a_bytes + b_stride);
    if (stages > 6) stages = 6;
    if (stages < 2) return dfd_set_error(DFD_ERR_UNSUPPORTED, "dfd_conv_tc: smem");
    p.stages = stages;
    size_t smem = (size_t)stages * (a_bytes + b_stride) + fixed;
    CUtensorMap ma, mb, mc;
    int rc;
    if ((rc = make_map_nhwc(&ma, x, N, Hin, Win, Cin, TW, TH, TN, p.is_bf16, S))) return rc;
    if ((rc = make_map(&mb, wpk, Cout, K, p.block_n, p.is_bf16))) return rc;
    if ((rc = make_map_nhwc(&mc, y, N, H, W, Cout, TW, TH, TN, p.is_bf16, 1, out_view))) return rc;
    int device = 0, sms = 148;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int grid = p.num_m_tiles * p.num_n_tiles;
    if (grid > sms) grid = sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.is_bf16) {
        auto kf = gemm_tc_kernel<bf16>;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET); attr = true; }
        kf<<<grid, NUM_THREADS, smem, st>>>(ma, mb, mc, p);
    } else {
        auto kf = gemm_tc_kernel<__half>;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET); attr = true; }
        kf<<<grid, NUM_THREADS, smem, st>>>(ma, mb, mc, p);
    }
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

// C[M,N] = A[M,K] * B[N,K]^T; statistics of output column c go to channel c % stat_n
static int launch_gemm_tc(const void* A, const void* B, void* C, long long M, int N, int K, int dt, double* dsum,
                          double* dsq, int stat_n, const void* fin, void* stream) {
    TcParams p;
    p.fin = (const BnFinDesc*)fin;
    p.conv = 0;
    p.M = (int)M; p.N = N; p.K = K;
    p.stat_n = stat_n;
    p.is_bf16 = dt == DFD_DT_BF16;
    p.block_n = N <= MAX_BLOCK_N ? ((N + 15) / 16) * 16 : MAX_BLOCK_N;
    p.num_m_tiles = cdiv(M, BLOCK_M);
    p.num_n_tiles = cdiv(N, p.block_n);
    p.num_k_blocks = cdiv(K, BLOCK_K);
    p.dsum = dsum; p.dsq = dsq;
    { const char* e = getenv("DFD_DBG"); p.dbg = e ? atoi(e) : 0; }
    p.ts = nullptr;
    static long long* ts_buf = nullptr;
    const bool trace = getenv("DFD_TS") != nullptr;
    if (trace) { if (!ts_buf) cudaMalloc(&ts_buf, 32 * 8 * sizeof(long long)); cudaMemset(ts_buf, 0, 32 * 8 * sizeof(long long)); p.ts = ts_buf; }
    const int a_bytes = BLOCK_M * BLOCK_K * 2;
    const int b_stride = ((p.block_n * BLOCK_K * 2) + 1023) & ~1023;
    const int fixed = 4 * BLOCK_M * SLAB * 2 + 22 * 8 + 2 * 4 * 64 * 4 + 1024 /* alignment slack */;
    int stages = (SMEM_BUDGET - fixed) / (a_bytes + b_stride);
    if (stages > 6) stages = 6;
    if (stages < 2) return dfd_set_error(DFD_ERR_UNSUPPORTED, "dfd_gemm_tn: smem");
    p.stages = stages;
    size_t smem = (size_t)stages * (a_bytes + b_stride) + fixed;

    CUtensorMap ma, mb, mc;
    int rc;
    if ((rc = make_map(&ma, A, M, K, BLOCK_M, p.is_bf16))) return rc;
    if ((rc = make_map(&mb, B, N, K, p.block_n, p.is_bf16))) return rc;
    if ((rc = make_map(&mc, C, M, N, BLOCK_M, p.is_bf16))) return rc;

    int device = 0, sms = 148;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int grid = p.num_m_tiles * p.num_n_tiles;
    if (grid > sms) grid = sms;
    cudaStream_t st = (cudaStream_t)stream;
    if (p.is_bf16) {
        auto kf = gemm_tc_kernel<bf16>;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET); attr = true; }
        kf<<<grid, NUM_THREADS, smem, st>>>(ma, mb, mc, p);
    } else {
        auto kf = gemm_tc_kernel<__half>;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BUDGET); attr = true; }
        kf<<<grid, NUM_THREADS, smem, st>>>(ma, mb, mc, p);
    }
    DFD_LAUNCH_CHECK();
    if (trace) {
        // diagnostics only: synchronous dump of CTA 0's pipeline timestamps (cycles relative to its first event)
        long long h[32 * 8];
        cudaDeviceSynchronize();
        cudaMemcpy(h, ts_buf, sizeof(h), cudaMemcpyDeviceToHost);
        long long t0 = h[0];
        fprintf(stderr, "gemm_tc trace M=%d N=%d K=%d block_n=%d stages=%d: tile | prod_go mma_acc_free mma_data mma_commit | epi_wait epi_go epi_ld epi_end\n",
                M, N, K, p.block_n, p.stages);
        for (int t = 0; t < 32 && h[t * 8 + 1]; t++) {
            fprintf(stderr, "  %2d |", t);
            for (int j = 0; j < 8; j++) fprintf(stderr, " %7lld%s", h[t * 8 + j] - t0, j == 3 ? " |" : "");
            fprintf(stderr, "\n");
        }
    }
    return DFD_OK;
}

#define DISPATCH_16(dt, ...)                                          \
    if ((dt) == DFD_DT_FP16) { typedef __half T16; __VA_ARGS__; }     \
    else { typedef bf16 T16; __VA_ARGS__; }

template <typename T>
__global__ void blockdiag_kernel(const BlockDiagDesc* __restrict__ table) {
    const BlockDiagDesc d = table[blockIdx.y];
    const T* src = (const T*)d.src;
    T* dst = (T*)d.dst;
    const int Kp = d.K * d.pack;
    const long long total = (long long)d.N * d.pack * Kp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int col = (int)(i % Kp), row = (int)(i / Kp);
        const int j = row / d.N, n = row - j * d.N, jj = col / d.K, k = col - jj * d.K;
        dst[i] = j == jj ? src[(size_t)n * d.K + k] : from_f<T>(0.f);
    }
}

}  // namespace

extern "C" {

// C[M,N] = A[M,K] * B[N,K]^T on tcgen05; optional fp64 column statistics of the stored C ([8][N] slots).
// All pointers must be 16-byte aligned, K % 8 == 0 and N % 8 == 0 (TMA global strides are multiples of 16 B).
int dfd_gemm_tn(const void* A, const void* B, void* C, long long M, int N, int K, int dt, double* dsum, double* dsq,
                const void* fin, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8)) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_tn: N%8, K%8");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_tn: dtype");
    return launch_gemm_tc(A, B, C, M, N, K, dt, dsum, dsq, N, fin, stream);
}

// The same product for SMALL K (a pointwise conv with 16 / 24 / 32 input channels): `pack` consecutive rows of A are
// read as ONE row of pack*K values and multiplied by the block-diagonal weight Bd[pack*N, pack*K] (dfd_blockdiag_weights),
// which yields `pack` consecutive rows of C side by side - byte for byte the row-major C[M,N]. TMA fetches a tile row
// per request, so 32-byte rows (K = 16) leave the load path request-bound at a quarter of the HBM rate; the packed view
// issues 128-byte rows. The extra MMA work multiplies zeros and is free at these K.
int dfd_gemm_tn_rowpack(const void* A, const void* Bd, void* C, long long M, int N, int K, int pack, int dt,
                        double* dsum, double* dsq, const void* fin, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || (N % 8) || (K % 8)) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_tn_rowpack: N%8, K%8");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_tn_rowpack: dtype");
    if (pack < 1 || pack > 8 || (M % pack)) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_tn_rowpack: M % pack");
    return launch_gemm_tc(A, Bd, C, M / pack, N * pack, K * pack, dt, dsum, dsq, N, fin, stream);
}

// Dense k x k convolution (stride 1, padding (k-1)/2) as an IMPLICIT GEMM on tcgen05: no im2col matrix exists in memory - the
// TMA producer fetches, per tap and 64-channel block, the input box shifted by the tap through a 4-D tensor map over the NHWC
// tensor (out-of-bounds rows arrive as zeros = the padding) straight into the swizzled MMA operand buffer.
//   forward : x = input,  wpk = [Cout][kh][kw][Cin]                      (resnet.py:129-136,195-197: nn.Conv2d 3x3)
//   dgrad   : x = dY,     wpk = [Cin][kh'][kw'][Cout] with flipped taps   (autograd input gradient of the same conv)
int dfd_conv_tc(const void* x, const void* wpk, void* y, int N, int H, int W, int Cin, int Cout, int k, int stride, int dt,
                double* dsum, double* dsq, const void* fin, void* stream) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (Cin % 64) || (Cout % 64) || (k != 1 && k != 3 && k != 5 && k != 7) ||
        (stride != 1 && stride != 2))
        return dfd_set_error(DFD_ERR_ARG, "dfd_conv_tc: Cin % 64, Cout % 64, k in {1,3,5,7}, stride in {1,2}");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_conv_tc: dtype");
    return launch_conv_tc(x, wpk, y, N, H, W, Cin, Cout, k, stride, dt, dsum, dsq, fin, stream);
}

// Input gradient of a 3x3, stride-2, padding-1 convolution as FOUR implicit GEMMs, one per parity class (py, px) of the input
// pixels: dx[2a+py][2b+px] = sum over the taps whose parity matches of dY[a + oy][b + ox] * W[kh][kw] - 1, 2, 2 and 4 taps.
// Each class is a stride-1 implicit GEMM over dY (out-of-range rows / columns arrive as TMA zeros) whose output tensor map is
// a strided VIEW of dx (every other pixel of every other row, offset by the parity): every dx element is written exactly once,
// no 9 x Cin column matrix and no col2im scatter. wpkD = the tap-flipped [Cin][kh'][kw'][Cout] layout of dfd_repack_weights.
int dfd_conv_dgrad_s2_tc(const void* dy, const void* wpkD, void* dx, int N, int H, int W, int Cin, int Cout, int dt, void* stream) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (Cin % 64) || (Cout % 64))
        return dfd_set_error(DFD_ERR_ARG, "dfd_conv_dgrad_s2_tc: Cin % 64, Cout % 64");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_conv_dgrad_s2_tc: dtype");
    const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
    for (int py = 0; py < 2; py++) {
        for (int px = 0; px < 2; px++) {
            const int Ha = (H - py + 1) / 2, Wb = (W - px + 1) / 2;
            if (Ha <= 0 || Wb <= 0) continue;
            ConvTaps t;
            t.n = 0;
            for (int kh = 0; kh < 3; kh++) {
                if (((py + 1 - kh) & 1) != 0) continue;                 // iy + 1 - kh must be even
                for (int kw = 0; kw < 3; kw++) {
                    if (((px + 1 - kw) & 1) != 0) continue;
                    t.oy[t.n] = (py + 1 - kh) / 2;                      // dY row = a + (py + 1 - kh) / 2
                    t.ox[t.n] = (px + 1 - kw) / 2;
                    t.kidx[t.n] = 8 - (kh * 3 + kw);                    // position of tap (kh, kw) in the flipped layout
                    t.n++;
                }
            }
            PixelView pv = {2LL * Cin * 2, 2LL * W * Cin * 2, (long long)H * W * Cin * 2};
            void* base = (char*)dx + ((size_t)py * W + px) * Cin * 2;
            // GEMM: M = N * Ha * Wb pixels of the class, K = taps x Cout (A = dY), N = Cin
            int rc = launch_conv_tc(dy, wpkD, base, N, Ho, Wo, Cout, Cin, 3, 1, dt, nullptr, nullptr, nullptr, stream, &t, Ha, Wb, &pv,
                                    9 * Cout);
            if (rc) return rc;
        }
    }
    return DFD_OK;
}

// Input gradient of a 1x1 convolution with stride s (the downsample branch, resnet.py:249-260) ADDED into dx, which already
// holds the main-path gradient of the block input: dx[n, s*a, s*b, :] += dY[n, a, b, :] * W. One implicit GEMM (k = 1) whose
// output map is the stride-s pixel view of dx and whose epilogue is a TMA reduction store (bf16 / fp16 add in L2) - replaces
// GEMM -> scratch, then col2im scatter + add (stride 2) or add_inplace (stride 1). wT = the transposed [Cin][Cout] weight.
int dfd_conv1x1_dgrad_add(const void* dy, const void* wT, void* dx, int N, int H, int W, int Cin, int Cout, int stride, int dt,
                          void* stream) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (Cin % 64) || (Cout % 64) || (stride != 1 && stride != 2))
        return dfd_set_error(DFD_ERR_ARG, "dfd_conv1x1_dgrad_add: Cin % 64, Cout % 64, stride in {1,2}");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_conv1x1_dgrad_add: dtype");
    const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    ConvTaps t;
    t.n = 1; t.ox[0] = 0; t.oy[0] = 0; t.kidx[0] = 0;
    PixelView pv = {(long long)stride * Cin * 2, (long long)stride * W * Cin * 2, (long long)H * W * Cin * 2};
    return launch_conv_tc(dy, wT, dx, N, Ho, Wo, Cout, Cin, 1, 1, dt, nullptr, nullptr, nullptr, stream, &t, Ho, Wo, &pv, Cout, 1);
}

// table: device array of {src [N,K], dst [pack*N, pack*K], N, K, pack}; dst(j*N+n, j'*K+k) = (j == j') ? src(n,k) : 0
int dfd_blockdiag_weights(const void* table, int count, int dt, void* stream) {
    if (count <= 0) return DFD_OK;
    dim3 grid(16, count);
    DISPATCH_16(dt, (blockdiag_kernel<T16><<<grid, 256, 0, (cudaStream_t)stream>>>((const BlockDiagDesc*)table)));
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

// dW[Nw,Kw] (fp32, accumulated) += G[M,Nw]^T * X[M,Kw] on tcgen05 (MN-major operands straight from NHWC, split over M)
// split count of the workspace (order-deterministic) mode: as many row ranges as fill the GPU once, but no more than keep
// the partial-sum traffic (one write + one read of splits x Nw x Kw floats) under half of the operand traffic
static long long wgrad_ws_splits(long long M, int Nw, int Kw, int sms, long long kblocks = 0) {
    const int block_n = Kw >= 128 ? 128 : ((Kw + 15) / 16) * 16;
    if (!kblocks) kblocks = (M + WG_KP - 1) / WG_KP;
    const int tm = cdiv(Nw, 128), tn = cdiv(Kw, block_n);
    long long splits = (2LL * sms + tm * tn - 1) / (tm * tn);
    const long long max_splits = (kblocks + 3) / 4;
    if (splits > max_splits) splits = max_splits;
    const long long cap = (2 * M * ((long long)Nw + Kw)) / (2LL * 2 * 4 * Nw * Kw);
    if (splits > cap) splits = cap;
    if (splits < 1) splits = 1;
    const long long kb_per = (kblocks + splits - 1) / splits;
    return (kblocks + kb_per - 1) / kb_per;
}

// number of partial matrices [Nw, Kw] dfd_gemm_wgrad writes into its workspace for this shape (workspace bytes = that x Nw x Kw x 4)
int dfd_gemm_wgrad_splits(long long M, int Nw, int Kw) {
    if (M <= 0 || Nw <= 0 || Kw <= 0) return 0;
    return (int)wgrad_ws_splits(M, Nw, Kw, 148);
}

// patch (TW x TH x TN <= 64 output pixels) of one pipeline stage of the implicit-GEMM weight gradient: the shape that covers
// the N x H x W pixels with the fewest patches (ties: the widest, longest contiguous runs for TMA)
struct WgPatch { int TW, TH, TN; long long patches; };
static WgPatch wg_patch(int N, int H, int W) {
    WgPatch best = {1, 1, 1, -1};
    for (int tw = 1; tw <= W && tw <= WG_KP; tw++) {
        for (int th = 1; th <= H && tw * th <= WG_KP; th++) {
            int tn = (tw == W && th == H) ? WG_KP / (tw * th) : 1;
            if (tn > N) tn = N;
            long long n = (long long)cdiv(W, tw) * cdiv(H, th) * cdiv(N, tn);
            if (best.patches < 0 || n < best.patches || (n == best.patches && tw > best.TW)) best = {tw, th, tn, n};
        }
    }
    return best;
}

static int launch_wgrad_tc(const void* G, const void* X, float* dW, long long M, int Nw, int Kw, int dt, void* ws,
                           long long ws_bytes, void* stream, int cvN, int cvH, int cvW, int cvCin, int cvk, int cvS = 1);

int dfd_gemm_wgrad(const void* G, const void* X, float* dW, long long M, int Nw, int Kw, int dt, void* ws, long long ws_bytes,
                   void* stream) {
    if (M <= 0 || Nw <= 0 || Kw <= 0 || (Nw % 8) || (Kw % 8)) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_wgrad: Nw%8, Kw%8");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_wgrad: dtype");
    return launch_wgrad_tc(G, X, dW, M, Nw, Kw, dt, ws, ws_bytes, stream, 0, 0, 0, 0, 0);
}

// Weight gradient of a dense k x k convolution (stride 1, padding (k-1)/2) as an IMPLICIT GEMM: dW[Cout][kh][kw][Cin] (fp32,
// the packed order of dfd_repack_weights; accumulated, or written as split partials into `ws` like dfd_gemm_wgrad) =
// sum over output pixels of dY[pixel, co] * x[pixel shifted by the tap, ci]; no im2col matrix in memory.
static inline int conv_out_extent(int h, int k, int s) { return (h + 2 * ((k - 1) / 2) - k) / s + 1; }
int dfd_conv_wgrad_splits(int N, int H, int W, int Cin, int Cout, int k, int stride) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || stride <= 0) return 0;
    const int Ho = conv_out_extent(H, k, stride), Wo = conv_out_extent(W, k, stride);
    WgPatch pt = wg_patch(N, Ho, Wo);
    return (int)wgrad_ws_splits((long long)N * Ho * Wo, Cout, k * k * Cin, 148, pt.patches);
}

int dfd_conv_wgrad_tc(const void* dy, const void* x, float* dW, int N, int H, int W, int Cin, int Cout, int k, int stride, int dt,
                      void* ws, long long ws_bytes, void* stream) {
    if (N <= 0 || H <= 0 || W <= 0 || Cin <= 0 || Cout <= 0 || (Cin % 64) || (Cout % 8) || (k != 1 && k != 3 && k != 5 && k != 7) ||
        (stride != 1 && stride != 2))
        return dfd_set_error(DFD_ERR_ARG, "dfd_conv_wgrad_tc: Cin % 64, Cout % 8, k in {1,3,5,7}, stride in {1,2}");
    if (dt != DFD_DT_BF16 && dt != DFD_DT_FP16) return dfd_set_error(DFD_ERR_ARG, "dfd_conv_wgrad_tc: dtype");
    const int Ho = conv_out_extent(H, k, stride), Wo = conv_out_extent(W, k, stride);
    return launch_wgrad_tc(dy, x, dW, (long long)N * Ho * Wo, Cout, k * k * Cin, dt, ws, ws_bytes, stream, N, H, W, Cin, k, stride);
}

static int launch_wgrad_tc(const void* G, const void* X, float* dW, long long M, int Nw, int Kw, int dt, void* ws,
                           long long ws_bytes, void* stream, int cvN, int cvH, int cvW, int cvCin, int cvk, int cvS) {
    WgParams p;
    // cvH / cvW: INPUT extents of the convolution; the patches tile the OUTPUT pixels
    const int cvHo = cvN > 0 ? conv_out_extent(cvH, cvk, cvS) : 0, cvWo = cvN > 0 ? conv_out_extent(cvW, cvk, cvS) : 0;
    p.M = M; p.Nw = Nw; p.Kw = Kw;
    p.is_bf16 = dt == DFD_DT_BF16;
    p.block_n = Kw >= 128 ? 128 : ((Kw + 15) / 16) * 16;
    p.kblocks = (M + WG_KP - 1) / WG_KP;
    p.conv = cvN > 0;
    if (p.conv) {
        WgPatch pt = wg_patch(cvN, cvHo, cvWo);
        p.cv_TW = pt.TW; p.cv_TH = pt.TH; p.cv_TN = pt.TN; p.cv_rows = pt.TW * pt.TH * pt.TN;
        p.cv_tiles_x = cdiv(cvWo, pt.TW); p.cv_tiles_y = cdiv(cvHo, pt.TH);
        p.cv_k = cvk; p.cv_pad = (cvk - 1) / 2; p.cv_cpb = cvCin / 64; p.cv_S = cvS;
        p.kblocks = pt.patches;
    }
    const int tm = cdiv(Nw, 128), tn = cdiv(Kw, p.block_n);
    int device = 0, sms = 148;
    cudaGetDevice(&device);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    long long splits = (2LL * sms + tm * tn - 1) / (tm * tn);
    long long max_splits = (p.kblocks + 3) / 4;              // at least 4 row blocks per CTA
    if (splits > max_splits) splits = max_splits;
    if (splits < 1) splits = 1;
    if (splits > 65535) splits = 65535;
    p.part = nullptr;
    if (ws) {
        splits = wgrad_ws_splits(M, Nw, Kw, 148, p.conv ? p.kblocks : 0);
        if (splits * (long long)Nw * Kw * 4 > ws_bytes)
            return dfd_set_error(DFD_ERR_ARG, "dfd_gemm_wgrad: workspace too small (dfd_gemm_wgrad_splits x Nw x Kw floats)");
        p.part = (float*)ws;
    }
    p.kb_per_split = (p.kblocks + splits - 1) / splits;
    splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;
    const int nbox_b = p.block_n > 64 ? 2 : 1;
    const int stage_bytes = (2 + nbox_b) * WG_BOX_BYTES;
    const int fixed = 18 * 8 + 64 + 1024;
    // ~100 KB per CTA: two CTAs (128 TMEM columns each) share an SM, so 2 x SMs splits run as one wave
    int stages = (100 * 1024 - fixed) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < 2) stages = 2;
    p.stages = stages;
    size_t smem = (size_t)stages * stage_bytes + fixed;
    CUtensorMap mg, mx;
    int rc;
    if (p.conv) {
        if ((rc = make_map_nhwc(&mg, G, cvN, cvHo, cvWo, Nw, p.cv_TW, p.cv_TH, p.cv_TN, p.is_bf16))) return rc;
        if ((rc = make_map_nhwc(&mx, X, cvN, cvH, cvW, cvCin, p.cv_TW, p.cv_TH, p.cv_TN, p.is_bf16, cvS))) return rc;
    } else {
        if ((rc = make_map(&mg, G, M, Nw, WG_KP, p.is_bf16))) return rc;
        if ((rc = make_map(&mx, X, M, Kw, WG_KP, p.is_bf16))) return rc;
    }
    dim3 grid(tm, tn, (unsigned)splits);
    cudaStream_t st = (cudaStream_t)stream;
    if (p.is_bf16) {
        auto kf = wgrad_tc_kernel<bf16>;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
        kf<<<grid, 256, smem, st>>>(mg, mx, dW, p);
    } else {
        auto kf = wgrad_tc_kernel<__half>;
        static bool attr = false;
        if (!attr) { cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024); attr = true; }
        kf<<<grid, 256, smem, st>>>(mg, mx, dW, p);
    }
    DFD_LAUNCH_CHECK();
    return DFD_OK;
}

}  // extern "C"
