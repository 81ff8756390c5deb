// K4 — per-video temporal attention pool + classification head, fp32 throughout.
// Reference: pretrained_detector.py:123-131 (sigmoid-MLP score -> softmax over T -> weighted FEATURE sum),
// :132-135 (mean mode), :138-141 (fc1 / ReLU / fc2; dropout is the identity in eval).
// One kernel for both members of the reference's ensemble: feature width 1280 (efficientnet_b0) and 2048 (resnet50).
// One CTA per video (ragged T via offsets).  Segmented warp-level reductions: a warp owns whole frames for
// the score MLP (1280-long dot products reduced with shuffles), then the softmax over the video's frames,
// then the weighted sum and the two FCs.  No atomics: results are bit-reproducible.
#include "common.cuh"
#include "kernels.h"

namespace dfd {

// DFD_POOLHEAD_KERNEL_BEGIN   (tools/host_emul/ runs the kernel below, unchanged, on CPU threads)
constexpr int kPhThreads = 512;
constexpr int kFeat = 1280, kAttHidden = 64, kFc1 = 256;   // kFeat: efficientnet_b0's width (the default instantiation)
constexpr int kMaxT = 1024;

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    return x;
}
__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
    return x;
}

constexpr int kPhChunk = 32;          // frames staged in shared memory at a time (32 x 1280 fp32 = 160 KB)

// FEAT: feature width (multiple of 32); CHUNK: frames staged at a time (CHUNK x FEAT fp32 of dynamic shared memory)
template <int FEAT = kFeat, int CHUNK = kPhChunk>
__global__ void __launch_bounds__(kPhThreads)
pool_head_kernel(const HeadWeights hw, const float* __restrict__ feat, const int32_t* __restrict__ offsets,
                 int frames, int use_attention, float* __restrict__ logits, float* __restrict__ frame_scores) {
    extern __shared__ float s_f[];                 // [CHUNK][FEAT] features of the current frame chunk
    __shared__ float s_w[kMaxT];
    __shared__ float s_hid[CHUNK][kAttHidden + 1];
    __shared__ float s_pooled[FEAT];
    __shared__ float s_h1[kFc1];
    const int v = blockIdx.x;
    griddep_wait();                                // features (and possibly the offsets) come from kernels before this one
    const int f0 = offsets[v];
    const int T = offsets[v + 1] - f0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int NW = kPhThreads / 32;
    if (T <= 0 || T > kMaxT || f0 < 0 || f0 + T > frames) {      // empty / over-long video or offsets outside the feature matrix:
        if (threadIdx.x < 2) logits[(size_t)v * 2 + threadIdx.x] = __int_as_float(0x7fc00000);      // poison instead of guessing
        if (frame_scores && T > 0 && f0 >= 0)
            for (int t = threadIdx.x; t < T && f0 + t < frames; t += kPhThreads) frame_scores[f0 + t] = __int_as_float(0x7fc00000);
        return;
    }
    const float* fv = feat + (size_t)f0 * FEAT;

    if (use_attention) {
        // scores: every W1 row is read ONCE per chunk of frames and applied to all frames staged in shared memory
        const float b2 = __ldg(hw.att_b2);
        for (int c0 = 0; c0 < T; c0 += CHUNK) {
            const int tc = min(CHUNK, T - c0);
            __syncthreads();
            for (int i = threadIdx.x; i < tc * FEAT; i += kPhThreads) s_f[i] = fv[(size_t)c0 * FEAT + i];
            __syncthreads();
            for (int h = warp; h < kAttHidden; h += NW) {
                float wr[FEAT / 32];
#pragma unroll
                for (int i = 0; i < FEAT / 32; ++i) wr[i] = __ldg(hw.att_w1 + (size_t)h * FEAT + lane + 32 * i);
                const float b1 = __ldg(hw.att_b1 + h), w2 = __ldg(hw.att_w2 + h);
                for (int t = 0; t < tc; ++t) {
                    float acc = 0.f;
#pragma unroll
                    for (int i = 0; i < FEAT / 32; ++i) acc = fmaf(s_f[t * FEAT + lane + 32 * i], wr[i], acc);
                    acc = warp_sum(acc);
                    if (lane == 0) s_hid[t][h] = fmaxf(acc + b1, 0.f) * w2;          // ReLU then Linear(64,1) term
                }
            }
            __syncthreads();
            for (int t = threadIdx.x; t < tc; t += kPhThreads) {
                float score = 0.f;
                for (int h = 0; h < kAttHidden; ++h) score += s_hid[t][h];           // fixed order
                s_w[c0 + t] = 1.0f / (1.0f + expf(-(score + b2)));                   // nn.Sigmoid, :70
            }
        }
        __syncthreads();
        if (warp == 0) {                                                      // F.softmax(dim=1), :127
            float mx = -INFINITY;
            for (int t = lane; t < T; t += 32) mx = fmaxf(mx, s_w[t]);
            mx = warp_max(mx);
            float sum = 0.f;
            for (int t = lane; t < T; t += 32) { const float e = expf(s_w[t] - mx); s_w[t] = e; sum += e; }
            sum = warp_sum(sum);
            for (int t = lane; t < T; t += 32) s_w[t] = s_w[t] / sum;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < FEAT; c += kPhThreads) {               // (features * w).sum(dim=1), :131
            float acc = 0.f;
            if (T <= CHUNK) { for (int t = 0; t < T; ++t) acc += s_f[t * FEAT + c] * s_w[t]; }   // chunk still staged
            else { for (int t = 0; t < T; ++t) acc += fv[(size_t)t * FEAT + c] * s_w[t]; }
            s_pooled[c] = acc;
        }
    } else {
        for (int t = threadIdx.x; t < T; t += kPhThreads) s_w[t] = 1.0f / (float)T;   // :135
        for (int c = threadIdx.x; c < FEAT; c += kPhThreads) {               // features.mean(dim=1), :134
            float acc = 0.f;
            for (int t = 0; t < T; ++t) acc += fv[(size_t)t * FEAT + c];
            s_pooled[c] = acc / (float)T;
        }
    }
    __syncthreads();
    if (frame_scores) for (int t = threadIdx.x; t < T; t += kPhThreads) frame_scores[f0 + t] = s_w[t];

    for (int j = warp * 4; j < kFc1; j += NW * 4) {                           // relu(fc1(.)), :139 — 4 outputs per pass
        const float* wr = hw.fc1_w + (size_t)j * FEAT + lane;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 8
        for (int i = 0; i < FEAT / 32; ++i) {
            const float xv = s_pooled[lane + 32 * i];
            a0 = fmaf(xv, __ldg(wr + 32 * i), a0);
            a1 = fmaf(xv, __ldg(wr + FEAT + 32 * i), a1);
            a2 = fmaf(xv, __ldg(wr + 2 * FEAT + 32 * i), a2);
            a3 = fmaf(xv, __ldg(wr + 3 * FEAT + 32 * i), a3);
        }
        a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3);
        if (lane == 0) {
            s_h1[j] = fmaxf(a0 + __ldg(hw.fc1_b + j), 0.f);         s_h1[j + 1] = fmaxf(a1 + __ldg(hw.fc1_b + j + 1), 0.f);
            s_h1[j + 2] = fmaxf(a2 + __ldg(hw.fc1_b + j + 2), 0.f); s_h1[j + 3] = fmaxf(a3 + __ldg(hw.fc1_b + j + 3), 0.f);
        }
    }
    __syncthreads();
    if (warp < 2) {                                                           // fc2, :141
        float acc = 0.f;
        for (int i = lane; i < kFc1; i += 32) acc = fmaf(s_h1[i], __ldg(hw.fc2_w + warp * kFc1 + i), acc);
        acc = warp_sum(acc);
        if (lane == 0) logits[(size_t)v * 2 + warp] = acc + __ldg(hw.fc2_b + warp);
    }
}

// DFD_POOLHEAD_KERNEL_END

template <int FEAT, int CHUNK>
static cudaError_t launch_pool_head_t(const HeadWeights& hw, const float* feat, const int32_t* offsets, int64_t videos, int64_t frames,
                                      int use_attention, float* logits, float* frame_scores, cudaStream_t s) {
    const size_t smem = (size_t)CHUNK * FEAT * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(pool_head_kernel<FEAT, CHUNK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_pdl(pool_head_kernel<FEAT, CHUNK>, dim3((unsigned)videos), dim3(kPhThreads), smem, s, hw, feat, offsets, (int)frames, use_attention, logits, frame_scores);
}

cudaError_t launch_pool_head(const HeadWeights& hw, const float* feat, const int32_t* offsets, int64_t videos, int64_t frames,
                             int feature_dim, int use_attention, float* logits, float* frame_scores, cudaStream_t s) {
    if (videos <= 0) return cudaSuccess;
    if (feature_dim == 1280) return launch_pool_head_t<1280, 32>(hw, feat, offsets, videos, frames, use_attention, logits, frame_scores, s);   // 160 KB
    if (feature_dim == 2048) return launch_pool_head_t<2048, 16>(hw, feat, offsets, videos, frames, use_attention, logits, frame_scores, s);   // 128 KB
    return cudaErrorInvalidValue;
}

}  // namespace dfd
