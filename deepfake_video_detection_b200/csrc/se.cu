// K2 tail — squeeze-excite gate (timm SqueezeExcite inside every MBConv block, reached from
// pretrained_detector.py:116):  gate = sigmoid(W2 · SiLU(W1 · mean_hw(x) + b1) + b2), all fp32.
// Input: the per-block partial sums written by dwconv.cu.  One CTA handles 8 frames so each FC weight is
// read once per 8 frames; partial rows are added in a fixed order (bit-reproducible).  The gate itself is
// applied to the A operand inside the project GEMM (gemm_tc.cu), so the dw output is never re-written.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace dfd {

constexpr int kSeMaxThreads = 1024;

// DFD_SE1_KERNEL_BEGIN   (tools/host_emul/ also runs this kernel, unchanged, on CPU threads)
template <int kSeFrames>
__global__ void __launch_bounds__(kSeMaxThreads)
se_kernel(const float* __restrict__ partials, int nparts, float inv_hw,
          const float* __restrict__ w1, const float* __restrict__ b1,
          const float* __restrict__ w2t, const float* __restrict__ b2,
          float* __restrict__ gate, int64_t frames, int C, int rd) {
    extern __shared__ __align__(16) float smem[];
    // frame-minor layouts: one 16-byte shared-memory load fetches 4 frames of a channel (the FC loops were bound by
    // the number of shared-memory loads, 8 scalar loads per weight)
    float* s_mean = smem;                         // [C][kSeFrames]
    float* s_r = smem + kSeFrames * C;            // [rd][kSeFrames]
    const int kSeThreads = blockDim.x;
    const int64_t f0 = (int64_t)blockIdx.x * kSeFrames;
    const int nf = (int)min((int64_t)kSeFrames, frames - f0);

#pragma unroll 4
    for (int i = threadIdx.x; i < kSeFrames * C; i += kSeThreads) {
        const int f = i / C, c = i - f * C;
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = 0.f;
        if (f < nf) {
            // 8 independent partial sums keep 8 loads in flight; they are combined in a fixed tree order
            const float* p = partials + ((size_t)(f0 + f) * nparts) * C + c;
            int q = 0;
            for (; q + 8 <= nparts; q += 8) {
#pragma unroll
                for (int u = 0; u < 8; ++u) a[u] += p[(size_t)(q + u) * C];
            }
            for (int u = 0; q < nparts; ++q, ++u) a[u] += p[(size_t)q * C];
        }
        s_mean[c * kSeFrames + f] = (((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]))) * inv_hw;
    }
    __syncthreads();

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int j = warp; j < rd; j += kSeThreads / 32) {
        float acc[kSeFrames];
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) acc[f] = 0.f;
        const float* wr = w1 + (size_t)j * C;
#pragma unroll 4
        for (int c = lane; c < C; c += 32) {
            const float wv = __ldg(wr + c);
            float m[kSeFrames];
            if constexpr (kSeFrames % 4 == 0) {
#pragma unroll
                for (int f = 0; f < kSeFrames; f += 4) *reinterpret_cast<float4*>(&m[f]) = *reinterpret_cast<const float4*>(&s_mean[c * kSeFrames + f]);
            } else {
#pragma unroll
                for (int f = 0; f < kSeFrames; ++f) m[f] = s_mean[c * kSeFrames + f];
            }
#pragma unroll
            for (int f = 0; f < kSeFrames; ++f) acc[f] = fmaf(wv, m[f], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[f] += __shfl_xor_sync(0xffffffffu, acc[f], o);
        }
        if (lane == 0) {
            const float bj = __ldg(b1 + j);
#pragma unroll
            for (int f = 0; f < kSeFrames; ++f) s_r[j * kSeFrames + f] = silu_f(acc[f] + bj);
        }
    }
    __syncthreads();

    for (int c = threadIdx.x; c < C; c += kSeThreads) {
        float acc[kSeFrames];
        const float bc = __ldg(b2 + c);
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) acc[f] = bc;
#pragma unroll 8
        for (int j = 0; j < rd; ++j) {
            const float wv = __ldg(w2t + (size_t)j * C + c);
            float r[kSeFrames];
            if constexpr (kSeFrames % 4 == 0) {
#pragma unroll
                for (int f = 0; f < kSeFrames; f += 4) *reinterpret_cast<float4*>(&r[f]) = *reinterpret_cast<const float4*>(&s_r[j * kSeFrames + f]);
            } else {
#pragma unroll
                for (int f = 0; f < kSeFrames; ++f) r[f] = s_r[j * kSeFrames + f];
            }
#pragma unroll
            for (int f = 0; f < kSeFrames; ++f) acc[f] = fmaf(wv, r[f], acc[f]);
        }
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f)
            if (f < nf) gate[(size_t)(f0 + f) * C + c] = sigmoid_f(acc[f]);
    }
}


// DFD_SE1_KERNEL_END


template <int FPB>
static cudaError_t launch_se_t(const float* partials, int nparts, float inv_hw, const float* w1, const float* b1,
                               const float* w2t, const float* b2, float* gate, int64_t frames, int C, int rd, cudaStream_t s) {
    const size_t smem = (size_t)FPB * (C + rd) * sizeof(float);
    if (smem > 64 * 1024) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(se_kernel<FPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((frames + FPB - 1) / FPB);
    // the kernel is instruction-latency-bound (IPC 0.66 at 8 warps per SM): wide layers get 32 warps per CTA
    const int threads = C >= 480 ? 1024 : (C >= 144 ? 512 : 256);
    se_kernel<FPB><<<grid, threads, smem, s>>>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd);
    return cudaGetLastError();
}

cudaError_t launch_se(const float* partials, int nparts, float inv_hw, const float* w1, const float* b1,
                      const float* w2t, const float* b2, float* gate, int64_t frames, int C, int rd, cudaStream_t s) {
    if (frames <= 0) return cudaSuccess;
    // Every CTA streams both FC matrices (2*rd*C fp32, 442 KB at C = 1152) from L2, so frames per CTA sets the L2
    // traffic: 2 frames per CTA made the wide layers L2-bound (453 MB per launch, 105 us); 8 frames per CTA = 4x less.
    return launch_se_t<8>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, s);
}

}  // namespace dfd
