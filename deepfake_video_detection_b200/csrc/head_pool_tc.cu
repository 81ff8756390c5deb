// K3h — conv_head 1x1 (320 -> 1280) + folded BN + SiLU + global average pool (the tail of timm `forward_features` + `global_pool`,
// reference pretrained_detector.py:116) as a TRANSPOSED tcgen05 GEMM:
//
//   D^T[channel, pixel] = W[channel, K] · X[pixel, K]^T        feat[frame, channel] = mean over the frame's pixels of SiLU(D^T + bias)
//
// The weights are the M-side operand (128 channels = 128 TMEM lanes), the activations [pixels][K] (K-major as they are) the
// N-side operand: four frames (4 x 49 pixels -> N = 208) per tile.  (Placing every frame at a 64-column boundary, N = 256, removes
// the column predicates of the epilogue but measured slower, 0.115 vs 0.109 ms: the kernel is MMA-bound, not epilogue-bound.)
// An epilogue thread then owns ONE CHANNEL and reads the
// pixels of a frame along its TMEM row: the average pool is a serial sum in the thread, in pixel order — no shared-memory
// transpose, no barrier, no shuffle, and the order of the additions depends only on the pixel index inside the frame (results
// do not depend on where a frame sits in the batch).  The row-major GEMM of gemm_tc.cu (POOL variant) spent its time in
// exactly that transpose (two bar.sync per 16 columns, MMA warp waiting 46 % for a free accumulator: 0.20 ms per 2048 frames,
// 0.30 of the tensor roof); it remains the fallback for maps of more than 64 pixels.
//
//   warp 17  TMA, activations: the pixel tile's K blocks (208 rows x 64) stay resident for the tile's channel tiles
//   warp 18  TMA, weights: a ring of 128 x 64 blocks (the whole matrix streams from L2 once per work unit)
//   warp 16  MMA issue: per channel tile 5 K blocks x 4 `tcgen05.mma` (M 128, N 208, K 16), two TMEM accumulators
//   warps 0-15  epilogue: group g = warp / 4 takes frame g of the tile, lane quarter q = warp % 4 the channels 32 q .. 32 q + 31
#include "common.cuh"
#include "kernels.h"
#include <cuda.h>

namespace dfd {

namespace {
constexpr int kHM = 128, kHK = 64, kHFpt = 4, kHMaxKb = 5;
constexpr uint32_t kHWBytes = kHM * 128;                       // one weight block: 128 channels x 64 K
constexpr int kHThreads = 19 * 32;
}

template <typename T>
__global__ void __launch_bounds__(kHThreads, 1)
head_pool_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
                    float* __restrict__ feat, int frames, int HW, int N, int K, int np16, int wstages, int cgroups, int ptiles, float inv_hw) {
    extern __shared__ __align__(128) uint8_t hp_smem[];
    const uint32_t base = (smem_u32(hp_smem) + 1023u) & ~1023u;
    const int num_kb = (K + kHK - 1) / kHK;
    const uint32_t x_bytes = (uint32_t)np16 * 128u;
    const uint32_t sm_w = base + num_kb * x_bytes;
    const uint32_t bars = sm_w + wstages * kHWBytes;
    // xfull[5], xempty[5], wfull[8], wempty[8], tfull[2], tempty[2]
    const uint32_t b_xfull = bars, b_xempty = bars + 40, b_wfull = bars + 80, b_wempty = bars + 144, b_tfull = bars + 208, b_tempty = bars + 224;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(hp_smem + (bars + 240 - smem_u32(hp_smem)));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int units = ptiles * cgroups, cpu = (N / kHM) / cgroups;            // channel tiles per unit

    if (threadIdx.x == 0) {
        for (int i = 0; i < kHMaxKb; ++i) { mbar_init(b_xfull + 8 * i, 1); mbar_init(b_xempty + 8 * i, 1); }
        for (int i = 0; i < 8; ++i) { mbar_init(b_wfull + 8 * i, 1); mbar_init(b_wempty + 8 * i, 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(b_tfull + 8 * i, 1); mbar_init(b_tempty + 8 * i, 16); }
        fence_barrier_init();
    }
    if (warp == 16) tmem_alloc(smem_u32(s_tmem), 512);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    griddep_wait();                         // barriers and TMEM are set up; the activations come from the kernel before this one

    if (warp == 17) {
        // ---------------------------------------------------------------- TMA: activations of the pixel tile
        if (lane == 0) {
            tma_prefetch_desc(&tmX);
            uint32_t ph = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x, ph ^= 1u) {
                const int pt = u / cgroups;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(b_xempty + 8 * kb, ph ^ 1u);
                    mbar_arrive_expect_tx(b_xfull + 8 * kb, x_bytes);
                    tma_load_2d(base + kb * x_bytes, &tmX, kb * kHK, pt * kHFpt * HW, b_xfull + 8 * kb);
                }
            }
        }
        __syncwarp();
    } else if (warp == 18) {
        // ---------------------------------------------------------------- TMA: weight blocks
        if (lane == 0) {
            tma_prefetch_desc(&tmW);
            int ws = 0; uint32_t ph = 0;
            for (int u = blockIdx.x; u < units; u += gridDim.x) {
                const int cg = u % cgroups;
                for (int ct = 0; ct < cpu; ++ct)
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(b_wempty + 8 * ws, ph ^ 1u);
                        mbar_arrive_expect_tx(b_wfull + 8 * ws, kHWBytes);
                        tma_load_2d(sm_w + ws * kHWBytes, &tmW, kb * kHK, (cg * cpu + ct) * kHM, b_wfull + 8 * ws);
                        if (++ws == wstages) { ws = 0; ph ^= 1u; }
                    }
            }
        }
        __syncwarp();
    } else if (warp == 16) {
        // ---------------------------------------------------------------- MMA issue
        const uint32_t idesc = umma_idesc(Half16<T>::kUmmaFormat, kHM, (uint32_t)np16);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int ws = 0; uint32_t wph = 0, xph = 0; int acc = 0; uint32_t aph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x, xph ^= 1u) {
            for (int ct = 0; ct < cpu; ++ct) {
                mbar_wait(b_tempty + 8 * acc, aph ^ 1u);
                tc_fence_after_sync();
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(b_xfull + 8 * kb, xph);
                    mbar_wait(b_wfull + 8 * ws, wph);
                    tc_fence_after_sync();
                    {   // whole warp converged, `elect.sync` picks the issuing lane (uniform operands; see umma_f16_elect)
                        const uint64_t ad = umma_smem_desc_sw128(sm_w + ws * kHWBytes);
                        const uint64_t bd = umma_smem_desc_sw128(base + kb * x_bytes);
                        const int steps = min(kHK, K - kb * kHK) >> 4;
                        for (int j = 0; j < steps; ++j) umma_f16_elect(tmem_u + (uint32_t)(acc * 256), ad + 2u * j, bd + 2u * j, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        umma_commit_elect(b_wempty + 8 * ws);
                        if (ct == cpu - 1) umma_commit_elect(b_xempty + 8 * kb);    // the tile's last use of this K block
                        if (kb == num_kb - 1) umma_commit_elect(b_tfull + 8 * acc);
                    }
                    if (++ws == wstages) { ws = 0; wph ^= 1u; }
                }
                if (++acc == 2) { acc = 0; aph ^= 1u; }
            }
        }
    } else {
        // ---------------------------------------------------------------- epilogue: thread = channel, group = frame of the tile
        const int g = warp >> 2, q = warp & 3;
        const int c0 = HW * g;                                              // first accumulator column of this group's frame
        const int cb = c0 & ~15, nchunks = ((c0 + HW + 15) >> 4) - (c0 >> 4);
        int acc = 0; uint32_t aph = 0;
        for (int u = blockIdx.x; u < units; u += gridDim.x) {
            const int pt = u / cgroups, cg = u % cgroups;
            const int frame = pt * kHFpt + g;
            for (int ct = 0; ct < cpu; ++ct) {
                const int channel = (cg * cpu + ct) * kHM + 32 * q + lane;
                const float bc = __ldg(bias + channel);
                mbar_wait(b_tfull + 8 * acc, aph);
                tc_fence_after_sync();
                const uint32_t taddr = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * 256 + cb);
                float sum = 0.f;
                uint32_t r[2][16];
                tmem_ld16(taddr, r[0]);
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    if (i < nchunks) {                                      // warp-uniform
                        tmem_ld_wait();
                        if (i + 1 < nchunks) tmem_ld16(taddr + (i + 1) * 16, r[(i + 1) & 1]);
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const int col = cb + 16 * i + j;
                            if (col >= c0 && col < c0 + HW) sum += silu_tanh(__uint_as_float(r[i & 1][j]) + bc);     // pixel order
                        }
                    }
                }
                tc_fence_before_sync();
                __syncwarp();
                if (lane == 0) mbar_arrive(b_tempty + 8 * acc);
                if (frame < frames) feat[(size_t)frame * N + channel] = sum * inv_hw;
                if (++acc == 2) { acc = 0; aph ^= 1u; }
            }
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) { tc_fence_after_sync(); tmem_dealloc(tmem_base, 512); }
}

bool head_pool_tc_supported(int64_t M, int K, int N, int HW) {
    return HW >= 1 && HW <= 64 && M > 0 && (M % HW) == 0 && (K % 16) == 0 && K >= 16 && K <= kHMaxKb * kHK && (N % kHM) == 0 && M / HW < (1ll << 30);
}

cudaError_t launch_head_pool_tc(const void* A, const void* W, const float* bias, float* feat, int64_t M, int K, int N, int HW, int dtype, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if (!head_pool_tc_supported(M, K, N, HW)) return cudaErrorInvalidValue;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    }
    const int frames = (int)(M / HW);
    const int np16 = (kHFpt * HW + 15) & ~15;
    const int num_kb = (K + kHK - 1) / kHK;
    const size_t fixed = 1024 + 256;
    int wstages = (int)((227 * 1024 - fixed - (size_t)num_kb * np16 * 128) / kHWBytes);
    if (wstages > 8) wstages = 8;
    if (wstages < 2) return cudaErrorInvalidValue;
    const size_t smem = fixed + (size_t)num_kb * np16 * 128 + (size_t)wstages * kHWBytes;
    const int ctiles = N / kHM, cgroups = (ctiles % 2 == 0) ? 2 : 1;
    const int ptiles = (frames + kHFpt - 1) / kHFpt;
    CUtensorMap tmX, tmW;
    cudaError_t e = make_tmap_2d(A, M, K, np16, &tmX);
    if (e != cudaSuccess) return e;
    e = make_tmap_2d(W, N, K, kHM, &tmW);
    if (e != cudaSuccess) return e;
    const int units = ptiles * cgroups;
    const unsigned grid = (unsigned)(units < sms ? units : sms);
    const float inv_hw = 1.0f / (float)HW;
    if (dtype == kDtypeFP16) {
        e = cudaFuncSetAttribute(head_pool_tc_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_pdl(head_pool_tc_kernel<__half>, dim3(grid), dim3(kHThreads), smem, s, tmX, tmW, bias, feat, frames, HW, N, K, np16, wstages, cgroups, ptiles, inv_hw);
    } else {
        e = cudaFuncSetAttribute(head_pool_tc_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_pdl(head_pool_tc_kernel<__nv_bfloat16>, dim3(grid), dim3(kHThreads), smem, s, tmX, tmW, bias, feat, frames, HW, N, K, np16, wstages, cgroups, ptiles, inv_hw);
    }
}

}  // namespace dfd
