// Stem as a tcgen05 / TMEM implicit GEMM (uint8 crops): conv3x3 s2 p1 3->32 + folded BN + SiLU, fused with the
// tensor prep of app.py:2084-2085 (timm conv_stem + bn1 = backbone.0/.1 of pretrained_detector.py:46).
//
//   D[M = frames*OH*OW, 32] = A[M, K] * W[32, K]^T + bias,   K = 96 = 3 x (27 taps*channels padded to 32)
//
// The normalised input must NOT be rounded to 16 bits (that alone costs as much logit error as the rest of the
// trunk), so the operands are split: x = x_hi + x_lo, w = w_hi + w_lo (16-bit each) and
//   A row = [x_hi | x_lo | x_hi],  W row = [w_hi | w_hi | w_lo]   ->   x_hi*w_hi + x_lo*w_hi + x_hi*w_lo
// with fp32 accumulation in TMEM: ~22 significant bits, i.e. the fp32 result up to the dropped x_lo*w_lo term.
//
// Builder warps do the im2col: one thread = one output pixel; the 27 bytes of its window go through the
// per-CTA 3x256 table (the reference's exact fp32 prep arithmetic, already split into a packed hi|lo 16-bit
// pair; zero padding is applied AFTER normalisation, as F.conv2d does) and are stored straight into the UMMA
// canonical K-major no-swizzle layout.  One elected thread issues 6 tcgen05.mma (M=128, N=32, K=16) per tile; epilogue warps
// read TMEM, add the bias, apply SiLU and store 64 contiguous bytes per pixel (NHWC).
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace dfd {

constexpr int kStBM = 128, kStK = 96, kStN = 32;
constexpr int kStChunks = kStK / 8;                       // 12 chunks of 8 halves
constexpr uint32_t kStLboA = kStBM * 16 + 16, kStLboB = kStN * 16 + 16;
constexpr uint32_t kStAStage = kStChunks * kStLboA;       // 24768 B
constexpr int kStStages = 6, kStAcc = 8;
constexpr int kStSets = 4;                                // builder sets of 128 threads taking alternate tiles
constexpr int kStEpiWarps = 4, kStBuildWarps = 4 * kStSets;
constexpr int kStThreads = (kStEpiWarps + 1 + kStBuildWarps) * 32;

template <typename T>
__global__ void __launch_bounds__(kStThreads, 1)
stem_tc_kernel(const uint8_t* __restrict__ in, const T* __restrict__ w16, const float* __restrict__ bias,
               T* __restrict__ out, int H, int W, int OH, int OW, int64_t total) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sp = smem_raw + kStStages * kStAStage;
    uint8_t* s_b = sp;                              sp += kStChunks * kStLboB;
    uint32_t* s_lut = reinterpret_cast<uint32_t*>(sp);  sp += 768 * 4;       // [c][u8] -> (hi | lo << 16)
    float* s_bias = reinterpret_cast<float*>(sp);   sp += kStN * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sp);      // full[S], empty[S], tfull[ACC], tempty[ACC]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kStStages + 2 * kStAcc);

    const uint32_t a_base0 = smem_u32(smem_raw), b_base = smem_u32(s_b);
    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kStStages);
    const uint32_t bar_tfull = smem_u32(bars + 2 * kStStages), bar_tempty = smem_u32(bars + 2 * kStStages + kStAcc);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < 768; i += kStThreads) {
        const float v = prep_value(i >> 8, i & 255);
        const T h = Half16<T>::from_float(v);
        const T l = Half16<T>::from_float(v - Half16<T>::to_float(h));
        s_lut[i] = (uint32_t)(*reinterpret_cast<const uint16_t*>(&h)) | ((uint32_t)(*reinterpret_cast<const uint16_t*>(&l)) << 16);
    }
    if (threadIdx.x < kStN) s_bias[threadIdx.x] = bias[threadIdx.x];
    for (int i = threadIdx.x; i < kStN * kStChunks; i += kStThreads) {          // W[32][96] -> canonical layout
        const int r = i / kStChunks, q = i - r * kStChunks;
        *reinterpret_cast<uint4*>(s_b + q * kStLboB + r * 16) = __ldg(reinterpret_cast<const uint4*>(w16 + r * kStK + q * 8));
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStStages; ++s) { mbar_init(bar_full + 8 * s, 128); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < kStAcc; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 128); }
        fence_barrier_init();
    }
    if (warp == kStEpiWarps) tmem_alloc(smem_u32(s_tmem), kStAcc * kStN);
    fence_proxy_async_smem();                       // the W tile was written with st.shared
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    griddep_wait();                                 // the crops may come from a kernel (crop + resize); the output buffer is reused
    const int64_t tiles = (total + kStBM - 1) / kStBM;
    const int ohw = OH * OW;

    if (warp > kStEpiWarps) {
        // ================================ BUILDERS (im2col) =====================================
        const int bt = threadIdx.x - (kStEpiWarps + 1) * 32;
        const int set = bt >> 7, row = bt & 127;
        int64_t li = set;
        for (int64_t tile = blockIdx.x + (int64_t)set * gridDim.x; tile < tiles; tile += kStSets * (int64_t)gridDim.x, li += kStSets) {
            const int stage = (int)(li % kStStages);
            const uint32_t phase = (uint32_t)(li / kStStages) & 1u;
            const int64_t pix = tile * kStBM + row;
            uint32_t xw[28];                               // packed (hi | lo << 16) per tap; 0 = zero padding
#pragma unroll
            for (int i = 0; i < 28; ++i) xw[i] = 0u;
            if (pix < total) {
                const int64_t frame = pix / ohw;
                const int rem = (int)(pix - frame * ohw);
                const int oy = rem / OW, ox = rem - oy * OW;
                const uint8_t* base = in + (size_t)frame * H * W * 3;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const int iy = 2 * oy - 1 + ky;
                    if (iy < 0 || iy >= H) continue;
                    const uint8_t* rowp = base + ((size_t)iy * W + (2 * ox - 1)) * 3;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const int ix = 2 * ox - 1 + kx;
                        if (ix < 0 || ix >= W) continue;
#pragma unroll
                        for (int c = 0; c < 3; ++c) xw[(ky * 3 + kx) * 3 + c] = s_lut[c * 256 + __ldg(rowp + kx * 3 + c)];
                    }
                }
            }
            // K layout [hi(27) 0(5) | lo(27) 0(5) | hi(27) 0(5)]: pick the hi / lo halves of consecutive taps
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int i = 0; i < 14; ++i) {
                hi[i] = __byte_perm(xw[2 * i], xw[2 * i + 1], 0x5410);
                lo[i] = __byte_perm(xw[2 * i], xw[2 * i + 1], 0x7632);
            }
            hi[14] = hi[15] = lo[14] = lo[15] = 0u;
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            const uint32_t dst = a_base0 + stage * kStAStage + row * 16;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint4 vh = make_uint4(hi[4 * q], hi[4 * q + 1], hi[4 * q + 2], hi[4 * q + 3]);
                const uint4 vl = make_uint4(lo[4 * q], lo[4 * q + 1], lo[4 * q + 2], lo[4 * q + 3]);
                sts16(dst + q * kStLboA, vh);
                sts16(dst + (4 + q) * kStLboA, vl);
                sts16(dst + (8 + q) * kStLboA, vh);
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_full + 8 * stage);
        }
    } else if (warp == kStEpiWarps) {
        // ================================ MMA ISSUER ============================================
        const uint32_t idesc = umma_idesc(Half16<T>::kUmmaFormat, kStBM, kStN);
        int64_t li = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++li) {
            const int stage = (int)(li % kStStages), acc = (int)(li % kStAcc);
            mbar_wait(bar_tempty + 8 * acc, ((uint32_t)(li / kStAcc) & 1u) ^ 1u);
            mbar_wait(bar_full + 8 * stage, (uint32_t)(li / kStStages) & 1u);
            tc_fence_after_sync();
            if (lane == 0) {
                const uint32_t a_base = a_base0 + stage * kStAStage;
#pragma unroll
                for (int j = 0; j < kStK / 16; ++j)
                    umma_f16(tmem_base + acc * kStN, umma_smem_desc(a_base + 2 * j * kStLboA, kStLboA, 128),
                             umma_smem_desc(b_base + 2 * j * kStLboB, kStLboB, 128), idesc, j > 0 ? 1u : 0u);
                umma_commit(bar_empty + 8 * stage);
                umma_commit(bar_tfull + 8 * acc);
            }
            __syncwarp();
        }
    } else {
        // ================================ EPILOGUE ==============================================
        const int row = 32 * warp + lane;
        int64_t li = 0;
        for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x, ++li) {
            const int acc = (int)(li % kStAcc);
            mbar_wait(bar_tfull + 8 * acc, (uint32_t)(li / kStAcc) & 1u);
            tc_fence_after_sync();
            const int64_t pix = tile * kStBM + row;
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * warp) << 16) + acc * kStN;
#pragma unroll
            for (int c16 = 0; c16 < 2; ++c16) {
                uint32_t r[16];
                tmem_ld16(t_row + c16 * 16, r);
                tmem_ld_wait();
                U32x8 o;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float a = __uint_as_float(r[2 * i]) + s_bias[c16 * 16 + 2 * i];
                    const float b = __uint_as_float(r[2 * i + 1]) + s_bias[c16 * 16 + 2 * i + 1];
                    o.v[i] = Half16<T>::pack(silu_f(a), silu_f(b));
                }
                if (pix < total) stg32(out + pix * kStN + c16 * 16, o);
            }
            tc_fence_before_sync();
            mbar_arrive(bar_tempty + 8 * acc);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kStEpiWarps) { tc_fence_after_sync(); tmem_dealloc(tmem_base, kStAcc * kStN); }
}

// ---------------------------------------------------------------------------------------------------------------
// Row variant (maps up to 128 output columns, i.e. the 224x224 crops): one tile = R consecutive OUTPUT ROWS of one frame
// (R = 2 when the map height is even, else 1; see the note on hand-off cost below).
//  * the 2R+1 input rows a tile needs are (2R+1)*W*3 contiguous bytes: one thread fetches them with a single 1-D bulk
//    copy (cp.async.bulk, completion in bytes on an mbarrier) into a ring of raw-byte stages — no per-byte global
//    loads, every input byte crosses the SM boundary 1.25x (R = 2) in total;
//  * the raw uint8 values ARE the A operand: an integer 0..255 is exact in fp16 (PRMT with the 0x64 magic byte gives
//    1024 + u, one HSUB2 removes the 1024), so the tensor prep of app.py:2084-2085 moves into the weights:
//        y = sum_inb w * ((u/255 - mean_c)/std_c) + b = sum_inb (w / (255 std_c)) * u + [b - sum_inb w * mean_c/std_c]
//    with w' = 256 * w / (255 std_c) split into fp16 hi + lo (MMAs against the same A tile accumulating into the same TMEM
//    columns, ~22 significant bits; the 2^-8 is applied with the bias in the epilogue) and FOUR bias vectors for the in-bounds tap sets of the
//    interior / left column / top row / top-left corner (zero padding happens after normalisation, as F.conv2d does);
//  * a builder thread (= one output pixel) reads its 3 x 9 window bytes as 9 aligned 32-bit shared-memory words.
// K layout: k = ky*10 + kx*3 + c (9 taps + 1 zero per input row, 32 in all).
constexpr int kSrK = 32, kSrChunks = kSrK / 8;
constexpr uint32_t kSrAStage = kSrChunks * kStLboA;      // 8256 B
constexpr int kSrN = kStN;                               // MMA N = 32: the W_hi and W_lo MMAs accumulate into the SAME TMEM columns
constexpr uint32_t kSrLboB = 2 * kStN * 16 + 16;         // W_hi rows 0-31, W_lo rows 32-63 of one canonical tile
constexpr uint32_t kSrBBytes = kSrChunks * kSrLboB;
#ifndef DFD_STEM_STAGES
#define DFD_STEM_STAGES 8      // A-tile ring, raw-byte ring (tools/build_variant.py sweeps them: 8/8 0.476 ms, 8/12 0.53, 6/12 0.53, 8/16 0.46)
#endif
#ifndef DFD_STEM_RAW
#define DFD_STEM_RAW 16
#endif
constexpr int kSrStages = DFD_STEM_STAGES, kSrRaw = DFD_STEM_RAW, kSrAcc = 8;
#ifndef DFD_STEM_DBG
#define DFD_STEM_DBG 0      // timing experiments only (wrong results): 1 builders skip their work, 2 epilogue skips its work, 4 no MMAs, 8 no raw loads
#endif
// kSrEpiSets epilogue sets and kSrSets builder sets (4 warps each) take alternate tiles: both roles are latency-bound per tile.
// A tile is R consecutive output rows of one frame (R = 2 when OH is even): with every role's work skipped the barrier hand-offs
// alone cost 0.18 us per tile (0.28 ms per 2048 frames at one row per tile, profiles/r02_experimental.md), so a tile carries two
// rows' worth of work per hand-off; the 2R+1 input rows it needs are still ONE bulk copy.

__device__ __forceinline__ void bulk_load_1d(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(bar) : "memory");
}

template <typename T, int kSrEpiSets, int kSrSets, int R>
__global__ void __launch_bounds__((4 * kSrEpiSets + 2 + 4 * kSrSets) * 32, 1)
stem_row_kernel(const uint8_t* __restrict__ in, const __half* __restrict__ wrow, const float* __restrict__ bias4,
                T* __restrict__ out, int H, int W, int OH, int OW, int tiles, uint32_t raw_stride) {
    // tile = frame * (OH / R) + ty as a 32-bit counter; ty (the tile's position inside its frame) is carried along (ty += stride %
    // (OH / R), one conditional subtract) — the 64-bit `tile % OH` / `tile / OH` this replaces were a ~100-instruction subroutine
    // call per tile and thread
    constexpr int kSrEpiWarps = 4 * kSrEpiSets, kSrMmaWarp = kSrEpiWarps, kSrRawWarp = kSrEpiWarps + 1;
    constexpr int kSrThreads = (kSrEpiWarps + 2 + 4 * kSrSets) * 32;
    constexpr uint32_t kTileA = R * kSrAStage;             // A operand of a tile: R row tiles of [128 pixels][32 k]
    constexpr int kAccCols = R * kSrN;                     // TMEM columns of a tile's accumulators
    static_assert(kSrAcc * kAccCols <= 512 && ((kSrAcc * kAccCols) & (kSrAcc * kAccCols - 1)) == 0, "TMEM allocation: a power of two up to 512 columns");
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* sp = smem_raw + kSrStages * kTileA;
    uint8_t* s_rawb = sp;                           sp += kSrRaw * raw_stride;
    uint8_t* s_b = sp;                              sp += kSrBBytes;
    float* s_bias = reinterpret_cast<float*>(sp);   sp += 4 * kStN * 4;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sp);      // afull[S], aempty[S], rfull[RAW], rempty[RAW], tfull[ACC], tempty[ACC]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kSrStages + 2 * kSrRaw + 2 * kSrAcc);

    const uint32_t a_base0 = smem_u32(smem_raw), raw_base0 = smem_u32(s_rawb), b_base = smem_u32(s_b);
    const uint32_t bar_afull = smem_u32(bars), bar_aempty = smem_u32(bars + kSrStages);
    const uint32_t bar_rfull = smem_u32(bars + 2 * kSrStages), bar_rempty = smem_u32(bars + 2 * kSrStages + kSrRaw);
    const uint32_t bar_tfull = smem_u32(bars + 2 * kSrStages + 2 * kSrRaw), bar_tempty = bar_tfull + 8 * kSrAcc;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t RB = (uint32_t)W * 3;
    const int OHT = OH / R;                                // tiles per frame

    for (int i = threadIdx.x; i < 4 * kStN; i += kSrThreads) s_bias[i] = 0.5f * bias4[i];   // halved: the epilogue forms h = x / 2 directly
    for (int i = threadIdx.x; i < 2 * kStN * kSrChunks; i += kSrThreads) {        // [hi|lo][32 oc][32 k] -> canonical layout
        const int h = i / (kStN * kSrChunks), r = (i / kSrChunks) % kStN, q = i % kSrChunks;
        *reinterpret_cast<uint4*>(s_b + q * kSrLboB + (h * kStN + r) * 16) = __ldg(reinterpret_cast<const uint4*>(wrow + (h * kStN + r) * kSrK + q * 8));
    }
    for (int i = threadIdx.x; i < kSrRaw * 4; i += kSrThreads)                     // 16 zero bytes in front of every raw stage
        *reinterpret_cast<uint32_t*>(s_rawb + (i >> 2) * raw_stride + (i & 3) * 4) = 0u;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kSrStages; ++s) { mbar_init(bar_afull + 8 * s, 128); mbar_init(bar_aempty + 8 * s, 1); }
        for (int s = 0; s < kSrRaw; ++s) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, 128); }
        for (int a = 0; a < kSrAcc; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 128); }
        fence_barrier_init();
    }
    if (warp == kSrMmaWarp) tmem_alloc(smem_u32(s_tmem), kSrAcc * kAccCols);
    fence_proxy_async_smem();                       // W tiles were written with st.shared
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    griddep_wait();                                 // the crops may come from a kernel (crop + resize); the output buffer is reused

    if (warp == kSrRawWarp) {
        // ================================ RAW LOADER ============================================
        if (lane == 0) {
            int stage = 0; uint32_t ph = 0;
            for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
                mbar_wait(bar_rempty + 8 * stage, ph ^ 1u);
                const int frame = tile / OHT;
                const int ty = tile - frame * OHT;
                const uint8_t* src = in + ((size_t)frame * H + (2 * R * ty - 1)) * RB;    // input rows 2R ty - 1 .. 2R ty + 2R - 1
                uint32_t dst = raw_base0 + stage * raw_stride + 16, bytes = (2 * R + 1) * RB;
                if (ty == 0) { src += RB; dst += RB; bytes -= RB; }                        // row -1 is padding (masked by the builders)
                if (DFD_STEM_DBG & 8) mbar_arrive(bar_rfull + 8 * stage);
                else {
                    mbar_arrive_expect_tx(bar_rfull + 8 * stage, bytes);
                    bulk_load_1d(dst, src, bytes, bar_rfull + 8 * stage);
                }
                if (++stage == kSrRaw) { stage = 0; ph ^= 1u; }
            }
        }
    } else if (warp > kSrRawWarp) {
        // ================================ BUILDERS (im2col from shared memory) ==================
        const int bt = threadIdx.x - (kSrRawWarp + 1) * 32;
        const int set = bt >> 7, row = bt & 127;
        const uint32_t magic = 0x00000064u;                       // byte 4 = 0x64, bytes 5..7 = 0
        const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
        uint32_t li = set;
        const int tstep = kSrSets * (int)gridDim.x, tystep = tstep % OHT;
        int ty = (int)((blockIdx.x + (uint32_t)set * gridDim.x) % (uint32_t)OHT);
        for (int tile = blockIdx.x + set * (int)gridDim.x; tile < tiles; tile += tstep, li += kSrSets) {
            const int rstage = (int)(li % kSrRaw), astage = (int)(li % kSrStages);
            const bool top = ty == 0;
            ty += tystep; if (ty >= OHT) ty -= OHT;
            uint32_t h2[R][16];
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) h2[j][i] = 0u;
            mbar_wait(bar_rfull + 8 * rstage, (li / kSrRaw) & 1u);
            if (row < OW && !(DFD_STEM_DBG & 1)) {
                const uint32_t rb = raw_base0 + rstage * raw_stride;
#pragma unroll
                for (int j = 0; j < R; ++j) {
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        if (j == 0 && r == 0 && top) continue;                             // padding row: zeros
                        const uint32_t s0 = 16u + (2 * j + r) * RB + 6u * row - 3u;        // first byte of the 9-byte window
                        const uint32_t wa = rb + (s0 & ~3u), sh = (s0 & 3u) * 8u;
                        uint32_t x0, x1, x2;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x0) : "r"(wa));
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x1) : "r"(wa + 4));
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(x2) : "r"(wa + 8));
                        uint32_t a0 = __funnelshift_r(x0, x1, sh), a1 = __funnelshift_r(x1, x2, sh), a2 = x2 >> sh;
                        if (row == 0) a0 &= 0xff000000u;                                   // left padding column: taps kx = 0
                        uint32_t p[5];
                        p[0] = __byte_perm(a0, magic, 0x4140); p[1] = __byte_perm(a0, magic, 0x4342);
                        p[2] = __byte_perm(a1, magic, 0x4140); p[3] = __byte_perm(a1, magic, 0x4342);
                        p[4] = __byte_perm(a2, magic, 0x4540);
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            __half2 v = __hsub2(*reinterpret_cast<__half2*>(&p[i]), k1024);  // exact: 1024 + u -> u
                            h2[j][r * 5 + i] = *reinterpret_cast<uint32_t*>(&v);
                        }
                    }
                }
            }
            mbar_arrive(bar_rempty + 8 * rstage);
            mbar_wait(bar_aempty + 8 * astage, ((li / kSrStages) & 1u) ^ 1u);
            if (row < OW && !(DFD_STEM_DBG & 1)) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const uint32_t dst = a_base0 + astage * kTileA + j * kSrAStage + row * 16;
#pragma unroll
                    for (int q = 0; q < 4; ++q) sts16(dst + q * kStLboA, make_uint4(h2[j][4 * q], h2[j][4 * q + 1], h2[j][4 * q + 2], h2[j][4 * q + 3]));
                }
            }
            fence_proxy_async_smem();
            mbar_arrive(bar_afull + 8 * astage);
        }
    } else if (warp == kSrMmaWarp) {
        // ================================ MMA ISSUER ============================================
        // four MMAs per output row (K = 32 = two k-steps, against W_hi and against W_lo, N = 32) accumulate x * (w_hi + w_lo) in ONE
        // set of TMEM columns: the epilogue reads 32 columns per row and adds nothing; descriptors differ only in the start address
        const uint32_t idesc = umma_idesc(0u /* fp16 operands whatever the output type */, kStBM, kSrN);
        const uint64_t a_d0 = umma_smem_desc(a_base0, kStLboA, 128);
        const uint64_t b_d0 = umma_smem_desc(b_base, kSrLboB, 128), b_d1 = umma_smem_desc(b_base + 2 * kSrLboB, kSrLboB, 128);
        const uint64_t b_l0 = umma_smem_desc(b_base + kStN * 16, kSrLboB, 128), b_l1 = umma_smem_desc(b_base + 2 * kSrLboB + kStN * 16, kSrLboB, 128);
        const uint32_t a_hi = (uint32_t)(a_d0 >> 32), a_lo0 = (uint32_t)a_d0;
        // the whole warp runs the loop converged, `elect.sync` picks the issuing lane (uniform operands; see umma_f16_elect);
        // ring positions and parities are counters
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0, acc = 0; uint32_t sph = 0, aph = 0;
        for (int tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
            mbar_wait(bar_tempty + 8 * acc, aph ^ 1u);
            mbar_wait(bar_afull + 8 * stage, sph);
            tc_fence_after_sync();
            if (!(DFD_STEM_DBG & 4)) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const uint32_t a_lo = a_lo0 + (uint32_t)(stage * R + j) * (kSrAStage >> 4);
                    const uint64_t a_k0 = ((uint64_t)a_hi << 32) | a_lo, a_k1 = ((uint64_t)a_hi << 32) | (a_lo + ((2 * kStLboA) >> 4));
                    const uint32_t d = tmem_u + acc * kAccCols + j * kSrN;
                    umma_f16_elect(d, a_k0, b_d0, idesc, 0u);
                    umma_f16_elect(d, a_k1, b_d1, idesc, 1u);
                    umma_f16_elect(d, a_k0, b_l0, idesc, 1u);
                    umma_f16_elect(d, a_k1, b_l1, idesc, 1u);
                }
            }
            umma_commit_elect(bar_aempty + 8 * stage);
            umma_commit_elect(bar_tfull + 8 * acc);
            if (++stage == kSrStages) { stage = 0; sph ^= 1u; }
            if (++acc == kSrAcc) { acc = 0; aph ^= 1u; }
        }
    } else {
        // ================================ EPILOGUE ==============================================
        const int q = warp & 3, eset = warp >> 2;          // TMEM lane quarter, epilogue set
        const int row = 32 * q + lane;
        uint32_t li = eset;
        const int tstep = kSrEpiSets * (int)gridDim.x, tystep = tstep % OHT;
        int ty = (int)((blockIdx.x + (uint32_t)eset * gridDim.x) % (uint32_t)OHT);
        for (int tile = blockIdx.x + eset * (int)gridDim.x; tile < tiles; tile += tstep, li += kSrEpiSets) {
            const int acc = (int)(li % kSrAcc);
            mbar_wait(bar_tfull + 8 * acc, (li / kSrAcc) & 1u);
            tc_fence_after_sync();
            const bool top = ty == 0;
            ty += tystep; if (ty >= OHT) ty -= OHT;
            if (!(DFD_STEM_DBG & 2)) {
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    const float* bs = s_bias + ((top && j == 0 ? 2 : 0) + (row == 0 ? 1 : 0)) * kStN;
                    const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + acc * kAccCols + j * kSrN;
                    uint32_t r[2][16];                                  // A * (W_hi + W_lo), scaled by 256
                    tmem_ld16(t_row, r[0]);
                    tmem_ld16(t_row + 16, r[1]);
                    tmem_ld_wait();
                    T* orow = out + (((size_t)tile * R + j) * OW + row) * kStN;
#pragma unroll
                    for (int c16 = 0; c16 < 2; ++c16) {
                        U32x8 o;
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            // h = x / 2 = acc * 2^-9 + bias / 2 (exact scaling), SiLU(x) = h + h tanh(h)
                            const float a = fmaf(__uint_as_float(r[c16][2 * i]), 0.001953125f, bs[c16 * 16 + 2 * i]);
                            const float b = fmaf(__uint_as_float(r[c16][2 * i + 1]), 0.001953125f, bs[c16 * 16 + 2 * i + 1]);
                            o.v[i] = Half16<T>::pack(fmaf(a, tanh_approx(a), a), fmaf(b, tanh_approx(b), b));
                        }
                        if (row < OW) stg32(orow + c16 * 16, o);
                    }
                }
            }
            tc_fence_before_sync();
            mbar_arrive(bar_tempty + 8 * acc);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kSrMmaWarp) { tc_fence_after_sync(); tmem_dealloc(tmem_base, kSrAcc * kAccCols); }
}

cudaError_t launch_stem_tc(const uint8_t* in, const void* w16, const float* bias, const void* wrow, const float* bias4, void* out,
                           int64_t frames, int H, int W, int dtype, cudaStream_t s) {
    const int OH = H / 2, OW = W / 2;
    const int64_t total = frames * OH * OW;
    if (total <= 0) return cudaSuccess;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    if (wrow && bias4 && OW <= kStBM && (W & 15) == 0 && (H & 1) == 0) {
        // row variant: one tile = two output rows of a frame (one when OH is odd)
        const int R = (OH & 1) ? 1 : 2;
        const uint32_t raw_stride = (16u + (2u * R + 1u) * (uint32_t)W * 3u + 127u) & ~127u;
        const size_t smem = (size_t)kSrStages * R * kSrAStage + kSrRaw * raw_stride + kSrBBytes + 4 * kStN * 4 + (2 * kSrStages + 2 * kSrRaw + 2 * kSrAcc) * 8 + 16;
        if (frames * OH > 0x7fffffffLL / 4) return cudaErrorInvalidValue;             // 32-bit tile counters in the kernel
        const int tiles = (int)(frames * (OH / R));
        const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
#define DFD_STEM_ROW(TT, E, B, RR) { \
            auto kern = stem_row_kernel<TT, E, B, RR>; \
            e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; \
            e = launch_pdl(kern, dim3(grid), dim3((4 * E + 2 + 4 * B) * 32), smem, s, in, (const __half*)wrow, bias4, (TT*)out, H, W, OH, OW, tiles, raw_stride); }
#ifndef DFD_STEM_ESETS
#define DFD_STEM_ESETS 3                          // epilogue sets / builder sets (tools/build_variant.py sweeps them): 3/3 0.446 ms,
#endif
#ifndef DFD_STEM_BSETS
#define DFD_STEM_BSETS 3                          // 4/2 0.452, 2/4 0.464, 2/3 0.464, 4/3 0.476, 3/4 0.483 per 2048 frames
#endif
        if (dtype == kDtypeFP16) {
            if (R == 2) DFD_STEM_ROW(__half, DFD_STEM_ESETS, DFD_STEM_BSETS, 2) else DFD_STEM_ROW(__half, DFD_STEM_ESETS, DFD_STEM_BSETS, 1)
        } else {
            if (R == 2) DFD_STEM_ROW(__nv_bfloat16, DFD_STEM_ESETS, DFD_STEM_BSETS, 2) else DFD_STEM_ROW(__nv_bfloat16, DFD_STEM_ESETS, DFD_STEM_BSETS, 1)
        }
#undef DFD_STEM_ROW
        return e;
    }
    const size_t smem = kStStages * kStAStage + kStChunks * kStLboB + 768 * 4 + kStN * 4 + (2 * kStStages + 2 * kStAcc) * 8 + 16;
    const int64_t tiles = (total + kStBM - 1) / kStBM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (dtype == kDtypeFP16) {
        e = cudaFuncSetAttribute(stem_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
        return launch_pdl(stem_tc_kernel<__half>, dim3(grid), dim3(kStThreads), smem, s, in, (const __half*)w16, bias, (__half*)out, H, W, OH, OW, total);
    } else {
        e = cudaFuncSetAttribute(stem_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
        return launch_pdl(stem_tc_kernel<__nv_bfloat16>, dim3(grid), dim3(kStThreads), smem, s, in, (const __nv_bfloat16*)w16, bias, (__nv_bfloat16*)out, H, W, OH, OW, total);
    }
}

}  // namespace dfd
