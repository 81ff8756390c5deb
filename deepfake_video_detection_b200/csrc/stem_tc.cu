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

cudaError_t launch_stem_tc(const uint8_t* in, const void* w16, const float* bias, void* out,
                           int64_t frames, int H, int W, int dtype, cudaStream_t s) {
    const int OH = H / 2, OW = W / 2;
    const int64_t total = frames * OH * OW;
    if (total <= 0) return cudaSuccess;
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    const size_t smem = kStStages * kStAStage + kStChunks * kStLboB + 768 * 4 + kStN * 4 + (2 * kStStages + 2 * kStAcc) * 8 + 16;
    const int64_t tiles = (total + kStBM - 1) / kStBM;
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (dtype == kDtypeFP16) {
        e = cudaFuncSetAttribute(stem_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
        stem_tc_kernel<__half><<<grid, kStThreads, smem, s>>>(in, (const __half*)w16, bias, (__half*)out, H, W, OH, OW, total);
    } else {
        e = cudaFuncSetAttribute(stem_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e;
        stem_tc_kernel<__nv_bfloat16><<<grid, kStThreads, smem, s>>>(in, (const __nv_bfloat16*)w16, bias, (__nv_bfloat16*)out, H, W, OH, OW, total);
    }
    return cudaGetLastError();
}

}  // namespace dfd
