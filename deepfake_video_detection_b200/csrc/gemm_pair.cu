// K3b — tensor-bound GEMMs on CTA pairs (tcgen05.mma cta_group::2): the dense contractions of the ViT-B/16 frame encoder
// (reference src/models.py:88-107 -> timm vit_base_patch16_224: patch embedding, qkv, attention projection, fc1, fc2).
//
//   D[M,N] = act(A[M,K] · W[N,K]^T + bias[N])        A, W, D 16-bit, K-major; fp32 accumulate; act: none | exact GELU
//
// Why pairs: a 128 x 256 tile per CTA (gemm_tc.cu) pulls 48 KB from L2 per 64-wide k-block = 85 FLOP per L2 byte; at the tensor
// rate that is more than the L2 can deliver (measured 1.0 PFLOP/s on the qkv GEMM).  A pair of CTAs on one TPC computes a
// 256 x 256 tile with ONE tcgen05.mma (M 256, N 256): each CTA stages its own 128 rows of A and only HALF of the W block, the
// tensor cores read the other half from the peer's shared memory — 128 FLOP per L2 byte and half the shared-memory operand
// traffic per SM.
//
// Per CTA (both CTAs of a pair run the same code; rank 0 is the leader):
//   warp 17   TMA: per k-block its 128 x 64 A box and its 128 x 64 W box (128-byte swizzle) with `.cta_group::2` copies whose
//             bytes are counted on the LEADER's stage barrier (the leader expects the bytes of both CTAs)
//   warp 16   leader only: one lane issues 4 `tcgen05.mma.cta_group::2` per k-block (accumulators: 128 lanes x 256 columns in
//             each CTA's TMEM, two buffers), `tcgen05.commit` multicast releases the stage / publishes the accumulator in BOTH CTAs
//   warps 0-15  epilogue, 4 column groups of 64: `tcgen05.ld` -> +bias -> GELU -> 16-bit -> swizzled staging panel in shared
//             memory -> ONE TMA store per warp and tile (its 32 rows x 64 columns; no per-thread global stores: a thread = one
//             row, so direct stores would touch 32 cache lines per warp instruction); the accumulator is handed back to the leader's MMA warp through a
//             cluster-scope mbarrier arrive as soon as the group's TMEM reads are done.
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>

#ifndef DFD_PAIR_DBG
#define DFD_PAIR_DBG 0             // timing experiments: 1 skip the TMA stores
#endif

namespace dfd {

namespace {
constexpr int kPM = 128, kPN = 256, kPK = 64;                         // rows per CTA, columns per pair, K per stage
constexpr uint32_t kPABytes = kPM * 128, kPBBytes = (kPN / 2) * 128, kPStage = kPABytes + kPBBytes;   // 32 KB per CTA and stage
constexpr int kPStages = 5;
constexpr int kPEpiWarps = 16, kPMmaWarp = 16, kPTmaWarp = 17, kPThreads = 18 * 32;
constexpr uint32_t kPPanel = kPM * 128;                                // staging of one column group: 128 rows x 64 columns x 2 B
constexpr size_t kPSmem = 1024 + (size_t)kPStages * kPStage + 4 * kPPanel + 256;
static_assert(kPSmem <= 227 * 1024, "shared memory budget");
}  // namespace

// RES: the output is the fp32 residual stream X [M][N], updated in place: X += A W^T + bias (no activation; tmD is an fp32 map of X)
template <typename T, int ACT, bool RES>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPThreads, 1)
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD,
                 const float* __restrict__ bias, int K, int m_tiles, int n_tiles) {
    static_assert(!RES || ACT == 0, "the residual variant has no activation");
    extern __shared__ __align__(128) uint8_t pr_smem[];
    const uint32_t base = (smem_u32(pr_smem) + 1023u) & ~1023u;
    const uint32_t sm_out = base + kPStages * kPStage;
    const uint32_t bars = sm_out + 4 * kPPanel;
    const uint32_t b_full = bars, b_empty = bars + 8 * kPStages, b_tfull = bars + 16 * kPStages, b_tempty = b_tfull + 16;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(pr_smem + (b_tempty + 16 - smem_u32(pr_smem)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = (int)cluster_id_x(), pairs = (int)cluster_count_x();
    const int units = m_tiles * n_tiles, num_kb = (K + kPK - 1) / kPK;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kPStages; ++s) { mbar_init(b_full + 8 * s, 1); mbar_init(b_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(b_tfull + 8 * a, 1); mbar_init(b_tempty + 8 * a, 2 * kPEpiWarps); }
        fence_barrier_init();
    }
    if (warp == kPMmaWarp) tmem_alloc_pair(smem_u32(s_tmem), 512);
    tc_fence_before_sync();
    cluster_sync_all();                                  // barriers of both CTAs initialised before any remote arrive / copy
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;

    if (warp == kPTmaWarp) {
        {   // the whole warp, converged: `elect.sync` inside the helpers picks the issuing lane
            if (lane == 0) { tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB); }
            const uint32_t full0 = mapa_u32(b_full, 0);
            int stage = 0; uint32_t phase = 0;
            for (int u = pair; u < units; u += pairs) {
                const int mt = u / n_tiles, nt = u - mt * n_tiles;
                const int row_a = mt * (2 * kPM) + (int)rank * kPM, row_b = nt * kPN + (int)rank * (kPN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(b_empty + 8 * stage, phase ^ 1);
                    if (rank == 0) mbar_arrive_expect_tx_elect(b_full + 8 * stage, 2 * kPStage);
                    const uint32_t dst = base + stage * kPStage;
                    tma_load_2d_pair_elect(dst, &tmA, kb * kPK, row_a, full0 + 8 * stage);
                    tma_load_2d_pair_elect(dst + kPABytes, &tmB, kb * kPK, row_b, full0 + 8 * stage);
                    if (++stage == kPStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kPMmaWarp) {
        if (rank == 0) {
            // the whole warp runs this loop converged; `elect.sync` inside the helpers picks the issuing lane (uniform operands)
            const uint32_t idesc = umma_idesc(Half16<T>::kUmmaFormat, 2 * kPM, kPN);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
            for (int u = pair; u < units; u += pairs) {
                mbar_wait_cluster(b_tempty + 8 * acc, acc_phase ^ 1);        // both CTAs' epilogues have drained this buffer
                tc_fence_after_sync();
                const uint32_t d_tmem = tmem_u + (uint32_t)(acc * kPN);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(b_full + 8 * stage, phase);
                    tc_fence_after_sync();
                    const uint64_t ad = umma_smem_desc_sw128(base + stage * kPStage);
                    const uint64_t bd = umma_smem_desc_sw128(base + stage * kPStage + kPABytes);
#pragma unroll
                    for (int j = 0; j < kPK / 16; ++j) umma_f16_pair_elect(d_tmem, ad + 2u * j, bd + 2u * j, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                    umma_commit_pair_elect(b_empty + 8 * stage);
                    if (kb == num_kb - 1) umma_commit_pair_elect(b_tfull + 8 * acc);
                    if (++stage == kPStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
        __syncwarp();
    } else {
        // ---------------------------------------------------------------- epilogue: thread = accumulator row, group = 64 columns
        const int q = warp & 3, g = warp >> 2;
        const int row = q * 32 + lane;
        const uint32_t panel = sm_out + (uint32_t)g * kPPanel;
        const uint32_t out_row = panel + (uint32_t)row * 128;
        const uint32_t sw = (uint32_t)(row & 7);
        const uint32_t tempty0 = mapa_u32(b_tempty, 0);
        int acc = 0; uint32_t acc_phase = 0;
        for (int u = pair; u < units; u += pairs) {
            const int mt = u / n_tiles, nt = u - mt * n_tiles;
            mbar_wait(b_tfull + 8 * acc, acc_phase);
            tc_fence_after_sync();
            const uint32_t t_addr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kPN + g * 64);
            const float4* bp = reinterpret_cast<const float4*>(bias + nt * kPN + g * 64);
            if constexpr (RES) {
                // X += tile: the fp32 accumulator (+ bias) goes to the warp's staging slab (32 rows x 32 columns x 4 B, swizzled) and
                // from there into the residual stream with a TMA REDUCTION (fp32 add at the L2): the SM never reads X and issues
                // no per-thread global access (a thread owns an accumulator ROW: direct accesses would touch 32 lines per instruction)
                const uint32_t slab = panel + (uint32_t)q * 4096u;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (lane == 0) bulk_wait_group_read0();                     // the previous reduction has read the slab
                    __syncwarp();
                    uint32_t r[2][16];
                    tmem_ld16(t_addr + h * 32, r[0]);
                    tmem_ld16(t_addr + h * 32 + 16, r[1]);
                    tmem_ld_wait();
                    if (h == 1) {                                               // all TMEM reads of this warp are done
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);
                    }
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 b4 = __ldg(bp + h * 8 + i);
                        const uint32_t* rr = &r[i >> 2][(i & 3) * 4];
                        uint4 v;
                        v.x = __float_as_uint(__uint_as_float(rr[0]) + b4.x); v.y = __float_as_uint(__uint_as_float(rr[1]) + b4.y);
                        v.z = __float_as_uint(__uint_as_float(rr[2]) + b4.z); v.w = __float_as_uint(__uint_as_float(rr[3]) + b4.w);
                        sts16(slab + (uint32_t)lane * 128u + (((uint32_t)i ^ (uint32_t)(lane & 7)) << 4), v);
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_reduce_add_2d(&tmD, slab, nt * kPN + g * 64 + h * 32, mt * (2 * kPM) + (int)rank * kPM + q * 32);
                        bulk_commit_group();
                    }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
                continue;
            }
            if (lane == 0) bulk_wait_group_read0();                           // this warp's previous store has read its slab
            __syncwarp();
            uint32_t r[2][16];
            tmem_ld16(t_addr, r[0]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tmem_ld_wait();
                if (c + 1 < 4) {
                    tmem_ld16(t_addr + (c + 1) * 16, r[(c + 1) & 1]);
                } else {                                                        // all TMEM reads of this warp are done
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(tempty0 + 8 * acc);
                }
                uint32_t pk[8];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 b4 = __ldg(bp + c * 4 + i);
                    uint64_t a = add2(f2_pack(__uint_as_float(r[c & 1][4 * i]), __uint_as_float(r[c & 1][4 * i + 1])), f2_pack(b4.x, b4.y));
                    uint64_t b = add2(f2_pack(__uint_as_float(r[c & 1][4 * i + 2]), __uint_as_float(r[c & 1][4 * i + 3])), f2_pack(b4.z, b4.w));
                    if (ACT == 2) { a = gelu_erfc_poly2(a); b = gelu_erfc_poly2(b); }
                    const float2 af = f2_unpack(a), bf = f2_unpack(b);
                    pk[2 * i] = Half16<T>::pack(af.x, af.y); pk[2 * i + 1] = Half16<T>::pack(bf.x, bf.y);
                }
                sts16(out_row + (((uint32_t)(2 * c) ^ sw) << 4), make_uint4(pk[0], pk[1], pk[2], pk[3]));
                sts16(out_row + (((uint32_t)(2 * c + 1) ^ sw) << 4), make_uint4(pk[4], pk[5], pk[6], pk[7]));
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0 && !(DFD_PAIR_DBG & 1)) {                            // one store per warp: its 32 rows x 64 columns (measured
                // 3 % faster than one store per 4-warp group behind a named barrier: 1306 vs 1263 TFLOP/s on the qkv shape)
                tma_store_2d(&tmD, panel + (uint32_t)q * 4096u, nt * kPN + g * 64, mt * (2 * kPM) + (int)rank * kPM + q * 32);
                bulk_commit_group();
            }
            if (++acc == 2) { acc = 0; acc_phase ^= 1; }
        }
        if (lane == 0) bulk_wait_group_read0();
    }
    tc_fence_before_sync();
    cluster_sync_all();                                  // the peer may still count on this CTA's barriers / read its operands
    if (warp == kPMmaWarp) { tc_fence_after_sync(); tmem_dealloc_pair(tmem_base, 512); }
}

bool gemm_pair_supported(int64_t M, int K, int N) { return M > 0 && N % kPN == 0 && K % 8 == 0 && K >= 8 && M < (1ll << 31) - 256; }

static cudaError_t launch_pair(const void* A, const void* W, const float* bias, void* D, float* X, int64_t M, int K, int N, int act, int dtype, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if (!gemm_pair_supported(M, K, N) || (act != 0 && act != 2) || (X && act)) return cudaErrorInvalidValue;
    CUtensorMap tmA, tmB, tmD;
    cudaError_t e = make_tmap_2d(A, M, K, kPM, &tmA);
    if (e != cudaSuccess) return e;
    e = make_tmap_2d(W, N, K, kPN / 2, &tmB);
    if (e != cudaSuccess) return e;
    e = X ? make_tmap_2d_f32(X, M, N, 32, &tmD) : make_tmap_2d(D, M, N, 32, &tmD);
    if (e != cudaSuccess) return e;
    int m_tiles = (int)((M + 2 * kPM - 1) / (2 * kPM)), n_tiles = N / kPN;
    const void* fn;
    const bool f16 = dtype == kDtypeFP16;
    if (X) fn = f16 ? (const void*)gemm_pair_kernel<__half, 0, true> : (const void*)gemm_pair_kernel<__nv_bfloat16, 0, true>;
    else if (act == 2) fn = f16 ? (const void*)gemm_pair_kernel<__half, 2, false> : (const void*)gemm_pair_kernel<__nv_bfloat16, 2, false>;
    else fn = f16 ? (const void*)gemm_pair_kernel<__half, 0, false> : (const void*)gemm_pair_kernel<__nv_bfloat16, 0, false>;
    e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPSmem);
    if (e != cudaSuccess) return e;
    static int max_pairs = 0;                            // co-resident pairs (one CTA per SM; a GPC with an odd SM count leaves one idle)
    if (max_pairs == 0) {
        cudaLaunchConfig_t qc = {};
        cudaLaunchAttribute at = {};
        at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = 2; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
        qc.gridDim = dim3(2, 1, 1); qc.blockDim = dim3(kPThreads, 1, 1); qc.dynamicSmemBytes = kPSmem; qc.attrs = &at; qc.numAttrs = 1;
        int n = 0, dev = 0, sms = 0;
        if (cudaOccupancyMaxActiveClusters(&n, fn, &qc) != cudaSuccess || n <= 0) {      // any grid is correct (static striding)
            (void)cudaGetLastError();
            e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
            e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
            n = sms / 2;
        }
        max_pairs = n;
    }
    const int64_t units = (int64_t)m_tiles * n_tiles;
    const unsigned grid = 2u * (unsigned)(units < max_pairs ? units : max_pairs);
    void* args[] = {(void*)&tmA, (void*)&tmB, (void*)&tmD, (void*)&bias, (void*)&K, (void*)&m_tiles, (void*)&n_tiles};
    return cudaLaunchKernel(fn, dim3(grid, 1, 1), dim3(kPThreads, 1, 1), args, kPSmem, s);
}

cudaError_t launch_gemm_pair(const void* A, const void* W, const float* bias, void* D, int64_t M, int K, int N, int act, int dtype, cudaStream_t s) {
    return launch_pair(A, W, bias, D, nullptr, M, K, N, act, dtype, s);
}
// X[M,N] (fp32, in place) += A[M,K] * W[N,K]^T + bias: the residual stream of the ViT encoder
cudaError_t launch_gemm_pair_residual(const void* A, const void* W, const float* bias, float* X, int64_t M, int K, int N, int dtype, cudaStream_t s) {
    if (!X) return cudaErrorInvalidValue;
    return launch_pair(A, W, bias, nullptr, X, M, K, N, 0, dtype, s);
}

}  // namespace dfd
