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
    griddep_wait();                               // the partial sums come from the depthwise kernel before this one

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


// ---- wide layers (C >= 480: the 14x14 and 7x7 stages) ------------------------------------------------------------------
// The kernel above walks both FC matrices with dependent L2 loads (74 us per launch at C = 1152, rd = 48, 2048 frames: pure
// latency).  Here 16 frames share one CTA (128 CTAs: one wave), and both matrices stream through shared memory in chunks
// fetched with cp.async one chunk ahead of the FMAs:
//   FC1  W1 [rd][C] in chunks of CC channels (smem rows padded by 4 floats: lanes of different j hit different banks); warp w
//        owns CC/16 channels of the chunk, lane = (frame quad, j residue mod 8) accumulates JPL x 4 partial dot products;
//        the 16 warps' partial sums are added in warp order (fixed), + b1, SiLU
//   FC2  W2^T [rd][C] in chunks of JC rows; thread = channel (up to 3 per thread), 16 frames in registers; + b2, sigmoid
// Results: same fp32 arithmetic, different (still fixed) summation order than the kernel above.
constexpr int kSe3Frames = 16, kSe3Threads = 512, kSe3Warps = kSe3Threads / 32;

template <int JPL>      // j values per lane: rd <= 8 * JPL
__global__ void __launch_bounds__(kSe3Threads, 1)
se_wide_kernel(const float* __restrict__ partials, int nparts, float inv_hw,
               const float* __restrict__ w1, const float* __restrict__ b1,
               const float* __restrict__ w2t, const float* __restrict__ b2,
               float* __restrict__ gate, int64_t frames, int C, int rd, int CC, int JC, uint32_t buf_floats) {
    extern __shared__ __align__(16) float smem[];
    float* s_mean = smem;                                      // [C][16]; after FC1: partial sums [16 warps][8 * JPL][16]
    float* s_r = s_mean + (size_t)C * kSe3Frames;              // [rd][16]
    float* s_w = s_r + 8 * JPL * kSe3Frames;                   // two chunk buffers of buf_floats
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t f0 = (int64_t)blockIdx.x * kSe3Frames;
    const int nf = (int)min((int64_t)kSe3Frames, frames - f0);
    const int n1 = C / CC, n2 = rd / JC, pitch1 = CC + 4;

    // chunk k < n1: W1[:, k CC .. (k+1) CC) as [rd][pitch1]; chunk n1 + k: W2^T rows k JC .. (k+1) JC as [JC][C]
    auto load_chunk = [&](int k) {
        float* dst = s_w + (size_t)(k & 1) * buf_floats;
        if (k < n1) {
            const int per_row = CC >> 2;
            for (int i = tid; i < rd * per_row; i += kSe3Threads) {
                const int j = i / per_row, q = i - j * per_row;
                cp_async16(smem_u32(dst + j * pitch1 + q * 4), w1 + (size_t)j * C + k * CC + q * 4, true);
            }
        } else if (k < n1 + n2) {
            const float* src = w2t + (size_t)(k - n1) * JC * C;
            for (int i = tid; i < (JC * C) >> 2; i += kSe3Threads) cp_async16(smem_u32(dst + i * 4), src + i * 4, true);
        }
        cp_async_commit();                                      // (an empty group past the last chunk keeps the wait counts uniform)
    };
    load_chunk(0);
    load_chunk(1);
    griddep_wait();                                             // weights are in flight; the partial sums come from the depthwise kernel

    // means of the 16 frames, frame-minor.  Thread = channel (coalesced over the warp), the 16 frames' loads of one partial row
    // are independent and issued back to back (a load -> shared-memory store loop per element is one memory latency per
    // element); partial rows are added in order.
    for (int c = tid; c < C; c += kSe3Threads) {
        float v[kSe3Frames];
#pragma unroll
        for (int f = 0; f < kSe3Frames; ++f) v[f] = 0.f;
        for (int q = 0; q < nparts; ++q) {
            float t[kSe3Frames];
#pragma unroll
            for (int f = 0; f < kSe3Frames; ++f) t[f] = f < nf ? __ldg(partials + ((size_t)(f0 + f) * nparts + q) * C + c) : 0.f;
#pragma unroll
            for (int f = 0; f < kSe3Frames; ++f) v[f] += t[f];
        }
#pragma unroll
        for (int f = 0; f < kSe3Frames; f += 4)
            *reinterpret_cast<float4*>(s_mean + (size_t)c * kSe3Frames + f) = make_float4(v[f] * inv_hw, v[f + 1] * inv_hw, v[f + 2] * inv_hw, v[f + 3] * inv_hw);
    }

    // ---- FC1
    const int fq = lane & 3, jr = lane >> 2;                   // this lane: frames 4 fq .. 4 fq + 3, j = jr + 8 i
    float acc[JPL][4];
#pragma unroll
    for (int i = 0; i < JPL; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
    const int cpw = CC / kSe3Warps;                            // channels of a chunk per warp
    for (int k = 0; k < n1; ++k) {
        cp_async_wait<1>();
        __syncthreads();                                        // chunk k (and, for k == 0, the means) visible to all
        const float* wb = s_w + (size_t)(k & 1) * buf_floats + jr * pitch1 + warp * cpw;
        const float* mb = s_mean + (size_t)(k * CC + warp * cpw) * kSe3Frames + 4 * fq;
#pragma unroll 2
        for (int c = 0; c < cpw; ++c) {
            const float4 m = *reinterpret_cast<const float4*>(mb + c * kSe3Frames);
#pragma unroll
            for (int i = 0; i < JPL; ++i) {
                const float wv = wb[i * 8 * pitch1 + c];        // rows past rd: stale shared memory, never stored below
                acc[i][0] = fmaf(wv, m.x, acc[i][0]); acc[i][1] = fmaf(wv, m.y, acc[i][1]);
                acc[i][2] = fmaf(wv, m.z, acc[i][2]); acc[i][3] = fmaf(wv, m.w, acc[i][3]);
            }
        }
        __syncthreads();                                        // buffer k & 1 free
        load_chunk(k + 2);
    }
    // partial sums -> s_mean region (the means are dead), then one fixed-order sum per (j, frame)
    float* s_part = s_mean;
#pragma unroll
    for (int i = 0; i < JPL; ++i)
        *reinterpret_cast<float4*>(s_part + ((size_t)warp * 8 * JPL + jr + 8 * i) * kSe3Frames + 4 * fq) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
    __syncthreads();
    for (int i = tid; i < rd * kSe3Frames; i += kSe3Threads) {
        const int j = i / kSe3Frames, f = i - j * kSe3Frames;
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kSe3Warps; ++w) t += s_part[((size_t)w * 8 * JPL + j) * kSe3Frames + f];
        s_r[j * kSe3Frames + f] = silu_f(t + __ldg(b1 + j));
    }

    // ---- FC2
    float g[3][kSe3Frames];
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const int c = tid + m * kSe3Threads;
        const float bc = c < C ? __ldg(b2 + c) : 0.f;
#pragma unroll
        for (int f = 0; f < kSe3Frames; ++f) g[m][f] = bc;
    }
    for (int k = 0; k < n2; ++k) {
        cp_async_wait<1>();
        __syncthreads();                                        // chunk n1 + k (and, for k == 0, s_r) visible
        const float* wb = s_w + (size_t)((n1 + k) & 1) * buf_floats;
#pragma unroll 2
        for (int jl = 0; jl < JC; ++jl) {
            float r[kSe3Frames];
#pragma unroll
            for (int f = 0; f < kSe3Frames; f += 4) *reinterpret_cast<float4*>(&r[f]) = *reinterpret_cast<const float4*>(&s_r[(k * JC + jl) * kSe3Frames + f]);
#pragma unroll
            for (int m = 0; m < 3; ++m) {
                const int c = tid + m * kSe3Threads;
                if (c < C) {
                    const float wv = wb[jl * C + c];
#pragma unroll
                    for (int f = 0; f < kSe3Frames; ++f) g[m][f] = fmaf(wv, r[f], g[m][f]);
                }
            }
        }
        __syncthreads();
        load_chunk(n1 + k + 2);
    }
    cp_async_wait<0>();
#pragma unroll
    for (int m = 0; m < 3; ++m) {
        const int c = tid + m * kSe3Threads;
        if (c < C) {
#pragma unroll
            for (int f = 0; f < kSe3Frames; ++f)
                if (f < nf) gate[(size_t)(f0 + f) * C + c] = sigmoid_f(g[m][f]);
        }
    }
}

// chunking of the wide kernel: CC channels of W1 / JC rows of W2^T per chunk, both at most kSe3ChunkBytes
constexpr size_t kSe3ChunkBytes = 57 * 1024;
static bool se_wide_plan(int C, int rd, int* CC, int* JC, size_t* buf_floats) {
    if (C < 480 || C > 3 * kSe3Threads || rd > 48 || (C & 3)) return false;
    const int rows = 8 * ((rd + 7) / 8 <= 3 ? 3 : ((rd + 7) / 8 == 4 ? 4 : 6));        // rows a lane may touch: 8 * JPL
    int n1 = 1; while (n1 <= 16 && ((C % n1) || ((C / n1) % kSe3Warps) || ((C / n1) & 3) || (size_t)rows * (C / n1 + 4) * 4 > kSe3ChunkBytes)) ++n1;
    int n2 = 1; while (n2 <= rd && ((rd % n2) || (size_t)(rd / n2) * C * 4 > kSe3ChunkBytes)) ++n2;
    if (n1 > 16 || n2 > rd) return false;
    *CC = C / n1; *JC = rd / n2;
    const size_t b1 = (size_t)rows * (*CC + 4), b2 = (size_t)*JC * C;
    *buf_floats = ((b1 > b2 ? b1 : b2) + 3) & ~size_t(3);
    return true;
}

template <int JPL>
static cudaError_t launch_se_wide_t(const float* partials, int nparts, float inv_hw, const float* w1, const float* b1, const float* w2t,
                                    const float* b2, float* gate, int64_t frames, int C, int rd, int CC, int JC, size_t buf_floats, cudaStream_t s) {
    const size_t smem = ((size_t)C * kSe3Frames + 8 * JPL * kSe3Frames + 2 * buf_floats) * sizeof(float);
    if (smem > 227 * 1024 || (size_t)kSe3Warps * 8 * JPL > (size_t)C) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(se_wide_kernel<JPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const unsigned grid = (unsigned)((frames + kSe3Frames - 1) / kSe3Frames);
    return launch_pdl(se_wide_kernel<JPL>, dim3(grid), dim3(kSe3Threads), smem, s, partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, CC, JC, (uint32_t)buf_floats);
}

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
    return launch_pdl(se_kernel<FPB>, dim3(grid), dim3(threads), smem, s, partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd);
}

cudaError_t launch_se(const float* partials, int nparts, float inv_hw, const float* w1, const float* b1,
                      const float* w2t, const float* b2, float* gate, int64_t frames, int C, int rd, cudaStream_t s) {
    if (frames <= 0) return cudaSuccess;
#ifndef DFD_SE_WIDE
#define DFD_SE_WIDE 1        // 0: every layer through se_kernel (A/B builds)
#endif
    int CC = 0, JC = 0; size_t bufw = 0;
    if (DFD_SE_WIDE && se_wide_plan(C, rd, &CC, &JC, &bufw)) {
        const int jpl = (rd + 7) / 8;
        if (jpl <= 3) return launch_se_wide_t<3>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, CC, JC, bufw, s);
        if (jpl == 4) return launch_se_wide_t<4>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, CC, JC, bufw, s);
        return launch_se_wide_t<6>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, CC, JC, bufw, s);
    }
    // Every CTA streams both FC matrices (2*rd*C fp32, 442 KB at C = 1152) from L2, so frames per CTA sets the L2
    // traffic: 2 frames per CTA made the wide layers L2-bound (453 MB per launch, 105 us); 8 frames per CTA = 4x less.
    return launch_se_t<8>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, s);
}

}  // namespace dfd
