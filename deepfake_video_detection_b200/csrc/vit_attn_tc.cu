// ViT-B/16 attention on the 5th-generation tensor cores (reference: src/models.py:88-107 -> timm vit_base_patch16_224 blocks,
// softmax(Q K^T / 8) V per (image, head), 197 tokens, head dim 64).
//
//   qkv [images*197][2304] 16-bit, timm column order (which*768 + head*64 + d)  ->  o [images*197][768] (column head*64 + d)
//
// One work item = (image, head, half): 128 query rows against all 197 keys (exact softmax: the whole key range is one tile).
// Persistent CTA per SM, warp-specialised, everything through mbarriers:
//   warp 0   TMA: one stage = Q tile (box 64 x 128 rows), K and V (boxes 64 x 208 rows) straight out of the qkv matrix,
//            128-byte swizzle; rows past the image's 197 tokens belong to the next image (or are zero-filled at the end of
//            the tensor): the key columns 197..207 are masked in the softmax, the query rows past 196 are never stored.
//   warp 1   MMA issue (one lane): S = Q K^T as 4 `tcgen05.mma` (M 128, N 208, K 16) into a TMEM buffer of 208 fp32 columns;
//            after the softmax of the tile, O = P V as 13 `tcgen05.mma` (M 128, N 64, K 16) with P from shared memory (K-major,
//            no-swizzle core-matrix layout written by the softmax warps) and V as an MN-MAJOR operand (the rows TMA delivered:
//            no transpose anywhere); O overwrites the first 64 columns of the consumed S buffer.  S of tile i+1 is issued before
//            waiting for P of tile i, so the tensor pipe runs ahead of the softmax warps (two TMEM buffers).
//   warps 2-5  softmax + epilogue, one thread per query row (= TMEM lane): pass 1 row maximum, pass 2 p = 2^((s - max) * log2(e) / 8)
//            (`ex2.approx`), fp32 row sum, P rounded to the storage type into shared memory (16-byte chunks: conflict-free),
//            `fence.proxy.async`, arrive; later `tcgen05.ld` of O, times 1 / sum, 128 bytes per row to global memory.
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>

namespace dfd {

namespace {
constexpr int kTTok = 197, kTKeys = 208, kTHd = 64, kTHeads = 12, kTDim = 768;
constexpr uint32_t kTQBytes = 128 * 128;                 // Q tile: 128 rows x 128 B
constexpr uint32_t kTKVBytes = kTKeys * 128;             // K or V: 208 rows x 128 B = 26 x 1024
constexpr uint32_t kTStageBytes = kTQBytes + 2 * kTKVBytes;
constexpr int kTStages = 2;
constexpr uint32_t kTPChunk = 128 * 16;                  // P: one 16-byte key chunk of all 128 rows
constexpr uint32_t kTPBytes = (kTKeys / 8) * kTPChunk;   // 26 chunks
constexpr int kTThreads = 6 * 32;
constexpr size_t kTSmem = 1024 + (size_t)kTStages * kTStageBytes + kTPBytes + 16 * 8 + 16;
static_assert(kTStageBytes % 1024 == 0 && kTQBytes % 1024 == 0 && kTKVBytes % 1024 == 0, "swizzled tiles are 1024-byte aligned");

// instruction descriptor with an MN-major B operand (bit 16)
__device__ __host__ constexpr uint32_t umma_idesc_bmn(uint32_t fmt, uint32_t m, uint32_t n) { return umma_idesc(fmt, m, n) | (1u << 16); }
// MN-major operand in the 128-byte-swizzled layout TMA wrote (row = K index, 64 MN elements = one 128-byte row): 8-row groups
// 1024 B apart (SBO); a single 64-wide atom along MN, so the leading offset is never used
__device__ __forceinline__ uint64_t umma_smem_desc_sw128_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)(kTKVBytes >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
}  // namespace

template <typename T>
__global__ void __launch_bounds__(kTThreads, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, T* __restrict__ o,
                        int images, int tiles) {
    extern __shared__ __align__(128) uint8_t at_smem[];
    const uint32_t base = (smem_u32(at_smem) + 1023u) & ~1023u;
    const uint32_t sm_p = base + kTStages * kTStageBytes;
    const uint32_t bars = sm_p + kTPBytes;                          // 8-byte aligned
    // qkv_full[2], qkv_empty[2], s_full[2], o_full[2], tmem_empty[2], p_full, p_empty
    const uint32_t b_full = bars, b_empty = bars + 16, b_sfull = bars + 32, b_ofull = bars + 48, b_tempty = bars + 64;
    const uint32_t b_pfull = bars + 80, b_pempty = bars + 88;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(at_smem + (bars + 96 - smem_u32(at_smem)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); mbar_init(b_sfull + 8 * i, 1);
            mbar_init(b_ofull + 8 * i, 1); mbar_init(b_tempty + 8 * i, 128);
        }
        mbar_init(b_pfull, 128); mbar_init(b_pempty, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(s_tmem), 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    const int n_local = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;      // tiles blockIdx.x, + gridDim.x, ...

    if (warp == 0) {
        // ---------------------------------------------------------------- TMA
        if (lane == 0) {
            tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmKV);
            for (int i = 0; i < n_local; ++i) {
                const int t = blockIdx.x + i * gridDim.x;
                const int half = t & 1, ih = t >> 1, head = ih % kTHeads, image = ih / kTHeads;
                const int st = i & 1;
                mbar_wait(b_empty + 8 * st, ((uint32_t)(i >> 1) & 1u) ^ 1u);
                const uint32_t dst = base + st * kTStageBytes, bar = b_full + 8 * st;
                mbar_arrive_expect_tx(bar, kTStageBytes);
                tma_load_2d(dst, &tmQ, head * kTHd, image * kTTok + half * 128, bar);
                tma_load_2d(dst + kTQBytes, &tmKV, kTDim + head * kTHd, image * kTTok, bar);
                tma_load_2d(dst + kTQBytes + kTKVBytes, &tmKV, 2 * kTDim + head * kTHd, image * kTTok, bar);
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issue
        const uint32_t idesc_qk = umma_idesc(Half16<T>::kUmmaFormat, 128, kTKeys);
        const uint32_t idesc_pv = umma_idesc_bmn(Half16<T>::kUmmaFormat, 128, kTHd);
        const uint64_t p_desc = umma_smem_desc(sm_p, kTPChunk, 128);         // chunks of a K step 2048 B apart, 8-row groups 128 B apart
        auto issue_qk = [&](int i) {
            const int st = i & 1, b = i & 1;
            mbar_wait(b_full + 8 * st, (uint32_t)(i >> 1) & 1u);
            mbar_wait(b_tempty + 8 * b, ((uint32_t)(i >> 1) & 1u) ^ 1u);
            tc_fence_after_sync();
            if (lane == 0) {
                const uint64_t qd = umma_smem_desc_sw128(base + st * kTStageBytes);
                const uint64_t kd = umma_smem_desc_sw128(base + st * kTStageBytes + kTQBytes);
#pragma unroll
                for (int j = 0; j < kTHd / 16; ++j) umma_f16(tmem_base + (uint32_t)(b * 256), qd + 2u * j, kd + 2u * j, idesc_qk, j > 0 ? 1u : 0u);
                umma_commit(b_sfull + 8 * b);
            }
            __syncwarp();
        };
        if (n_local > 0) issue_qk(0);
        for (int i = 0; i < n_local; ++i) {
            const int st = i & 1, b = i & 1;
            if (i + 1 < n_local) issue_qk(i + 1);
            mbar_wait(b_pfull, (uint32_t)i & 1u);
            tc_fence_after_sync();
            if (lane == 0) {
                const uint64_t vd = umma_smem_desc_sw128_mn(base + st * kTStageBytes + kTQBytes + kTKVBytes);
#pragma unroll
                for (int j = 0; j < kTKeys / 16; ++j)
                    umma_f16(tmem_base + (uint32_t)(b * 256), p_desc + (uint64_t)((2 * kTPChunk) >> 4) * j, vd + (uint64_t)(2048 >> 4) * j, idesc_pv, j > 0 ? 1u : 0u);
                umma_commit(b_ofull + 8 * b);
                umma_commit(b_empty + 8 * st);
                umma_commit(b_pempty);
            }
            __syncwarp();
        }
    } else {
        // ---------------------------------------------------------------- softmax + epilogue, thread = query row
        const int q4 = warp & 3;                                   // TMEM lane quarter this warp may access
        const int row = q4 * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(q4 * 32) << 16);
        const uint32_t p_row = sm_p + (uint32_t)(row >> 3) * 128 + (uint32_t)(row & 7) * 16;
        constexpr float kScale = 0.125f * 1.4426950408889634f;
        for (int i = 0; i < n_local; ++i) {
            const int t = blockIdx.x + i * gridDim.x;
            const int half = t & 1, ih = t >> 1, head = ih % kTHeads, image = ih / kTHeads;
            const int b = i & 1;
            const uint32_t s_addr = lane_addr + (uint32_t)(b * 256);
            mbar_wait(b_sfull + 8 * b, (uint32_t)(i >> 1) & 1u);
            tc_fence_after_sync();
            // both passes read the S row in 16-column chunks, the load of chunk c+1 in flight while chunk c is processed
            float mx = -INFINITY;
            uint32_t r[2][16];
            tmem_ld16(s_addr, r[0]);
#pragma unroll
            for (int c = 0; c < kTKeys / 16; ++c) {
                tmem_ld_wait();
                if (c + 1 < kTKeys / 16) tmem_ld16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    if (c * 16 + j < kTTok) mx = fmaxf(mx, __uint_as_float(r[c & 1][j]));
            }
            const float moff = mx * kScale;
            mbar_wait(b_pempty, ((uint32_t)i & 1u) ^ 1u);          // P V of the previous tile has read the P buffer
            float sum = 0.f;
            tmem_ld16(s_addr, r[0]);
#pragma unroll
            for (int c = 0; c < kTKeys / 16; ++c) {
                tmem_ld_wait();
                if (c + 1 < kTKeys / 16) tmem_ld16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
                uint32_t pk[8];
#pragma unroll
                for (int j = 0; j < 16; j += 2) {
                    float p0 = ex2_approx(fmaf(__uint_as_float(r[c & 1][j]), kScale, -moff));
                    float p1 = ex2_approx(fmaf(__uint_as_float(r[c & 1][j + 1]), kScale, -moff));
                    if (c * 16 + j >= kTTok) p0 = 0.f;
                    if (c * 16 + j + 1 >= kTTok) p1 = 0.f;
                    sum += p0 + p1;
                    pk[j >> 1] = Half16<T>::pack(p0, p1);
                }
                sts16(p_row + (uint32_t)(2 * c) * kTPChunk, make_uint4(pk[0], pk[1], pk[2], pk[3]));
                sts16(p_row + (uint32_t)(2 * c + 1) * kTPChunk, make_uint4(pk[4], pk[5], pk[6], pk[7]));
            }
            tc_fence_before_sync();                                 // this thread's TMEM reads of S precede the MMA that overwrites it
            fence_proxy_async_smem();
            mbar_arrive(b_pfull);
            // epilogue of this tile
            const float inv = 1.0f / sum;
            mbar_wait(b_ofull + 8 * b, (uint32_t)(i >> 1) & 1u);
            tc_fence_after_sync();
            const int qrow = half * 128 + row;
            T* dst = o + ((size_t)image * kTTok + qrow) * kTDim + head * kTHd;
            tmem_ld16(s_addr, r[0]);
#pragma unroll
            for (int c = 0; c < kTHd / 16; ++c) {
                tmem_ld_wait();
                if (c + 1 < kTHd / 16) tmem_ld16(s_addr + (c + 1) * 16, r[(c + 1) & 1]);
                if (qrow < kTTok) {
                    U32x8 v;
#pragma unroll
                    for (int j = 0; j < 8; ++j) v.v[j] = Half16<T>::pack(__uint_as_float(r[c & 1][2 * j]) * inv, __uint_as_float(r[c & 1][2 * j + 1]) * inv);
                    stg32(dst + c * 16, v);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(b_tempty + 8 * b);
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

cudaError_t launch_vit_attention_tc(const void* qkv, void* o, int64_t images, int dtype, cudaStream_t s) {
    if (images <= 0) return cudaSuccess;
    if (images * kTHeads * 2 > 0x7fffffffLL) return cudaErrorInvalidValue;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    }
    CUtensorMap tmQ, tmKV;
    cudaError_t e = make_tmap_2d(qkv, images * kTTok, 3 * kTDim, 128, &tmQ);
    if (e != cudaSuccess) return e;
    e = make_tmap_2d(qkv, images * kTTok, 3 * kTDim, kTKeys, &tmKV);
    if (e != cudaSuccess) return e;
    const int tiles = (int)(images * kTHeads * 2);
    const unsigned grid = (unsigned)(tiles < sms ? tiles : sms);
    if (dtype == kDtypeFP16) {
        e = cudaFuncSetAttribute(vit_attention_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem);
        if (e != cudaSuccess) return e;
        vit_attention_tc_kernel<__half><<<grid, kTThreads, kTSmem, s>>>(tmQ, tmKV, (__half*)o, (int)images, tiles);
    } else {
        e = cudaFuncSetAttribute(vit_attention_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTSmem);
        if (e != cudaSuccess) return e;
        vit_attention_tc_kernel<__nv_bfloat16><<<grid, kTThreads, kTSmem, s>>>(tmQ, tmKV, (__nv_bfloat16*)o, (int)images, tiles);
    }
    return cudaGetLastError();
}

}  // namespace dfd
