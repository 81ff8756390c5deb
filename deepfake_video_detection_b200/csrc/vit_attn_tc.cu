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
//   warp 1   MMA issue (the warp stays converged, `elect.sync` picks the lane): S = Q K^T as 4 `tcgen05.mma` (M 128, N 208, K 16) into one of two TMEM buffers; after the
//            softmax of the tile, O = P V as 13 `tcgen05.mma` (M 128, N 64, K 16) with P as the TMEM A OPERAND (the softmax warps
//            write it over the S columns they have consumed: P never touches shared memory) and V as an MN-MAJOR shared-memory
//            operand (the rows TMA delivered: no transpose anywhere); O lands in columns 128..191 of the same buffer.  S of tile
//            i+1 is issued before waiting for P of tile i.
//   warps 2-5 / 6-9   two softmax + epilogue groups, one per TMEM buffer (tiles alternate between them, so one group's
//            exponentials overlap the other's TMEM round trips and output stores), one thread per query row (= TMEM lane):
//            pass 1 row maximum, pass 2 p = 2^((s - max) * log2(e) / 8) (`ex2.approx`), fp32 row sum, P rounded to the storage
//            type and stored with `tcgen05.st`; later `tcgen05.ld` of O, times 1 / sum, 128 bytes per row to global memory.
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>

namespace dfd {

namespace {
constexpr int kTTok = 197, kTKeys = 208, kTHd = 64, kTHeads = 12, kTDim = 768;
constexpr uint32_t kTQBytes = 128 * 128;                 // Q tile: 128 rows x 128 B
constexpr uint32_t kTKVBytes = kTKeys * 128;             // K or V: 208 rows x 128 B = 26 x 1024
constexpr uint32_t kTStageBytes = kTQBytes + 2 * kTKVBytes;
constexpr int kTStages = 3;
constexpr int kTThreads = 10 * 32;
constexpr uint32_t kTOCol = 128;                         // O accumulator: columns 128..191 of the tile's TMEM buffer
constexpr size_t kTSmem = 1024 + (size_t)kTStages * kTStageBytes + 16 * 8 + 16;
static_assert(kTStageBytes % 1024 == 0 && kTQBytes % 1024 == 0 && kTKVBytes % 1024 == 0, "swizzled tiles are 1024-byte aligned");
static_assert(kTSmem <= 227 * 1024, "shared memory budget");

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
// D[tmem] (+)= A[tmem: lane = row, 32-bit column = two consecutive K elements] · B[smem]
// (executed by the whole converged MMA warp: `elect.sync` picks the issuing lane, operands stay warp-uniform — see umma_f16_elect)
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// registers -> TMEM: this thread's lane, 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace

template <typename T>
__global__ void __launch_bounds__(kTThreads, 1)
vit_attention_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV, T* __restrict__ o,
                        int images, int tiles) {
    extern __shared__ __align__(128) uint8_t at_smem[];
    const uint32_t base = (smem_u32(at_smem) + 1023u) & ~1023u;
    const uint32_t bars = base + kTStages * kTStageBytes;          // 8-byte aligned
    // full[3], empty[3], s_full[2], p_full[2], o_full[2], tmem_empty[2]
    const uint32_t b_full = bars, b_empty = bars + 24, b_sfull = bars + 48, b_pfull = bars + 64, b_ofull = bars + 80, b_tempty = bars + 96;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(at_smem + (bars + 112 - smem_u32(at_smem)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kTStages; ++i) { mbar_init(b_full + 8 * i, 1); mbar_init(b_empty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(b_sfull + 8 * i, 1); mbar_init(b_pfull + 8 * i, 128); mbar_init(b_ofull + 8 * i, 1); mbar_init(b_tempty + 8 * i, 128);
        }
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
            int st = 0; uint32_t ph = 0;
            for (int i = 0; i < n_local; ++i) {
                const int t = blockIdx.x + i * gridDim.x;
                const int half = t & 1, ih = t >> 1, head = ih % kTHeads, image = ih / kTHeads;
                mbar_wait(b_empty + 8 * st, ph ^ 1u);
                const uint32_t dst = base + st * kTStageBytes, bar = b_full + 8 * st;
                mbar_arrive_expect_tx(bar, kTStageBytes);
                tma_load_2d(dst, &tmQ, head * kTHd, image * kTTok + half * 128, bar);
                tma_load_2d(dst + kTQBytes, &tmKV, kTDim + head * kTHd, image * kTTok, bar);
                tma_load_2d(dst + kTQBytes + kTKVBytes, &tmKV, 2 * kTDim + head * kTHd, image * kTTok, bar);
                if (++st == kTStages) { st = 0; ph ^= 1u; }
            }
        }
    } else if (warp == 1) {
        // ---------------------------------------------------------------- MMA issue
        const uint32_t idesc_qk = umma_idesc(Half16<T>::kUmmaFormat, 128, kTKeys);
        const uint32_t idesc_pv = umma_idesc_bmn(Half16<T>::kUmmaFormat, 128, kTHd);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int qst = 0; uint32_t qph = 0;                               // stage cursor of the S = Q K^T issues (one tile ahead)
        auto issue_qk = [&](int i) {
            const int b = i & 1;
            mbar_wait(b_full + 8 * qst, qph);
            mbar_wait(b_tempty + 8 * b, ((uint32_t)(i >> 1) & 1u) ^ 1u);
            tc_fence_after_sync();
            {
                const uint64_t qd = umma_smem_desc_sw128(base + qst * kTStageBytes);
                const uint64_t kd = umma_smem_desc_sw128(base + qst * kTStageBytes + kTQBytes);
#pragma unroll
                for (int j = 0; j < kTHd / 16; ++j) umma_f16_elect(tmem_u + (uint32_t)(b * 256), qd + 2u * j, kd + 2u * j, idesc_qk, j > 0 ? 1u : 0u);
                umma_commit_elect(b_sfull + 8 * b);
            }
            if (++qst == kTStages) { qst = 0; qph ^= 1u; }
        };
        if (n_local > 0) issue_qk(0);
        int st = 0;
        for (int i = 0; i < n_local; ++i) {
            const int b = i & 1;
            if (i + 1 < n_local) issue_qk(i + 1);
            mbar_wait(b_pfull + 8 * b, (uint32_t)(i >> 1) & 1u);
            tc_fence_after_sync();
            {
                const uint64_t vd = umma_smem_desc_sw128_mn(base + st * kTStageBytes + kTQBytes + kTKVBytes);
                const uint32_t tb = tmem_u + (uint32_t)(b * 256);
#pragma unroll
                for (int j = 0; j < kTKeys / 16; ++j)
                    umma_f16_ts(tb + kTOCol, tb + 8u * j, vd + (uint64_t)(2048 >> 4) * j, idesc_pv, j > 0 ? 1u : 0u);
                umma_commit_elect(b_ofull + 8 * b);
                umma_commit_elect(b_empty + 8 * st);
            }
            if (++st == kTStages) st = 0;
        }
    } else {
        // ---------------------------------------------------------------- softmax + epilogue, thread = query row
        const int wg = (warp - 2) >> 2;                            // group 0: even tiles, TMEM buffer 0; group 1: odd tiles, buffer 1
        const int q4 = warp & 3;                                   // TMEM lane quarter this warp may access
        const int row = q4 * 32 + lane;
        const uint32_t s_addr = tmem_base + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(wg * 256);
        constexpr float kScale = 0.125f * 1.4426950408889634f;
        uint32_t ph = 0;
        for (int i = wg; i < n_local; i += 2, ph ^= 1u) {
            const int t = blockIdx.x + i * gridDim.x;
            const int half = t & 1, ih = t >> 1, head = ih % kTHeads, image = ih / kTHeads;
            mbar_wait(b_sfull + 8 * wg, ph);
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
                // P chunk c (keys 16c..16c+15) -> columns 8c..8c+7: S columns this thread has already read (chunk c / 2 <= c;
                // the load of chunk c+1 in flight reads columns >= 16 (c+1) > 8c+7)
                tmem_st8(s_addr + (uint32_t)(8 * c), pk);
            }
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(b_pfull + 8 * wg);
            // epilogue of this tile
            const float inv = 1.0f / sum;
            mbar_wait(b_ofull + 8 * wg, ph);
            tc_fence_after_sync();
            const int qrow = half * 128 + row;
            T* dst = o + ((size_t)image * kTTok + qrow) * kTDim + head * kTHd;
            tmem_ld16(s_addr + kTOCol, r[0]);
#pragma unroll
            for (int c = 0; c < kTHd / 16; ++c) {
                tmem_ld_wait();
                if (c + 1 < kTHd / 16) tmem_ld16(s_addr + kTOCol + (c + 1) * 16, r[(c + 1) & 1]);
                if (qrow < kTTok) {
                    U32x8 v;
#pragma unroll
                    for (int j = 0; j < 8; ++j) v.v[j] = Half16<T>::pack(__uint_as_float(r[c & 1][2 * j]) * inv, __uint_as_float(r[c & 1][2 * j + 1]) * inv);
                    stg32(dst + c * 16, v);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(b_tempty + 8 * wg);
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
