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

// Second variant (EXPERIMENTAL, DFD_SE_V2=1; written without GPU access, off by default).  Same contract, same phase 1, same
// FC2 summation order; FC1 sums its channels in a different (still fixed) order.  The first variant is bound by instruction
// issue at low IPC (84 k warp instructions per 8-frame CTA at C = 1152, 73 us per launch): here the FC loops use packed
// fp32x2 FMAs (sm_100 needs the .f32x2 form for the full fp32 rate), 16-byte weight loads, two FC1 rows per warp pass (each
// staged mean is read once for two rows) and two channels per thread in FC2 (one pass instead of two half-empty ones):
// about 2.3x fewer instructions.
// DFD_SE2_KERNEL_BEGIN   (tools/host_emul/ compiles the kernel up to the END marker unchanged for the CPU)
template <int kSeFrames>
__global__ void __launch_bounds__(kSeMaxThreads)
se_kernel_v2(const float* __restrict__ partials, int nparts, float inv_hw,
             const float* __restrict__ w1, const float* __restrict__ b1,
             const float* __restrict__ w2t, const float* __restrict__ b2,
             float* __restrict__ gate, int64_t frames, int C, int rd) {
    static_assert(kSeFrames == 8, "frame pairs are packed: 8 frames = two 16-byte shared-memory loads");
    extern __shared__ __align__(16) float smem[];
    float* s_mean = smem;                         // [C][8]
    float* s_r = smem + kSeFrames * C;            // [rd][8]
    const int kSeThreads = blockDim.x;
    const int64_t f0 = (int64_t)blockIdx.x * kSeFrames;
    const int nf = (int)min((int64_t)kSeFrames, frames - f0);

    // phase 1: exactly the first variant (same partial-sum tree, same bits)
#pragma unroll 4
    for (int i = threadIdx.x; i < kSeFrames * C; i += kSeThreads) {
        const int f = i / C, c = i - f * C;
        float a[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) a[u] = 0.f;
        if (f < nf) {
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

    // FC1: a warp pass = rows (j0, j0+1); lanes take 4-channel groups q = lane, lane + 32, ...; 8 frames = 4 packed pairs per row
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cq = C >> 2;
    for (int task = warp; 2 * task < rd; task += kSeThreads / 32) {
        const int j0 = 2 * task, j1 = min(j0 + 1, rd - 1);
        uint64_t a0[4] = {0ull, 0ull, 0ull, 0ull}, a1[4] = {0ull, 0ull, 0ull, 0ull};
        const float4* wr0 = reinterpret_cast<const float4*>(w1 + (size_t)j0 * C);
        const float4* wr1 = reinterpret_cast<const float4*>(w1 + (size_t)j1 * C);
#pragma unroll 2
        for (int q = lane; q < cq; q += 32) {
            const float4 wa = __ldg(wr0 + q), wb = __ldg(wr1 + q);
            const float wav[4] = {wa.x, wa.y, wa.z, wa.w}, wbv[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const ulonglong2 mA = *reinterpret_cast<const ulonglong2*>(&s_mean[(4 * q + k) * kSeFrames]);
                const ulonglong2 mB = *reinterpret_cast<const ulonglong2*>(&s_mean[(4 * q + k) * kSeFrames + 4]);
                const uint64_t p0 = f2_pack(wav[k], wav[k]), p1 = f2_pack(wbv[k], wbv[k]);
                a0[0] = fma2(p0, mA.x, a0[0]); a0[1] = fma2(p0, mA.y, a0[1]); a0[2] = fma2(p0, mB.x, a0[2]); a0[3] = fma2(p0, mB.y, a0[3]);
                a1[0] = fma2(p1, mA.x, a1[0]); a1[1] = fma2(p1, mA.y, a1[1]); a1[2] = fma2(p1, mB.x, a1[2]); a1[3] = fma2(p1, mB.y, a1[3]);
            }
        }
        float r0[kSeFrames], r1[kSeFrames];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 u = f2_unpack(a0[i]), v = f2_unpack(a1[i]);
            r0[2 * i] = u.x; r0[2 * i + 1] = u.y; r1[2 * i] = v.x; r1[2 * i + 1] = v.y;
        }
#pragma unroll
        for (int f = 0; f < kSeFrames; ++f) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { r0[f] += __shfl_xor_sync(0xffffffffu, r0[f], o); r1[f] += __shfl_xor_sync(0xffffffffu, r1[f], o); }
        }
        if (lane == 0) {
            const float bj0 = __ldg(b1 + j0), bj1 = __ldg(b1 + j1);
#pragma unroll
            for (int f = 0; f < kSeFrames; ++f) {
                s_r[j0 * kSeFrames + f] = silu_f(r0[f] + bj0);
                if (j1 != j0) s_r[j1 * kSeFrames + f] = silu_f(r1[f] + bj1);
            }
        }
    }
    __syncthreads();

    // FC2: one thread = 2 channels x 8 frames; j ascending from the bias (the first variant's order)
    for (int cp = threadIdx.x; 2 * cp < C; cp += kSeThreads) {
        const int c = 2 * cp;
        const float2 bc = __ldg(reinterpret_cast<const float2*>(b2 + c));
        uint64_t a0[4], a1[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { a0[i] = f2_pack(bc.x, bc.x); a1[i] = f2_pack(bc.y, bc.y); }
#pragma unroll 4
        for (int j = 0; j < rd; ++j) {
            const float2 wv = __ldg(reinterpret_cast<const float2*>(w2t + (size_t)j * C + c));
            const ulonglong2 rA = *reinterpret_cast<const ulonglong2*>(&s_r[j * kSeFrames]);
            const ulonglong2 rB = *reinterpret_cast<const ulonglong2*>(&s_r[j * kSeFrames + 4]);
            const uint64_t p0 = f2_pack(wv.x, wv.x), p1 = f2_pack(wv.y, wv.y);
            a0[0] = fma2(p0, rA.x, a0[0]); a0[1] = fma2(p0, rA.y, a0[1]); a0[2] = fma2(p0, rB.x, a0[2]); a0[3] = fma2(p0, rB.y, a0[3]);
            a1[0] = fma2(p1, rA.x, a1[0]); a1[1] = fma2(p1, rA.y, a1[1]); a1[2] = fma2(p1, rB.x, a1[2]); a1[3] = fma2(p1, rB.y, a1[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 u = f2_unpack(a0[i]), v = f2_unpack(a1[i]);
            if (2 * i < nf) *reinterpret_cast<float2*>(gate + (size_t)(f0 + 2 * i) * C + c) = make_float2(sigmoid_f(u.x), sigmoid_f(v.x));
            if (2 * i + 1 < nf) *reinterpret_cast<float2*>(gate + (size_t)(f0 + 2 * i + 1) * C + c) = make_float2(sigmoid_f(u.y), sigmoid_f(v.y));
        }
    }
}

// DFD_SE2_KERNEL_END

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
    const char* env_v2 = getenv("DFD_SE_V2");                                              // experimental variant, off by default
    if (env_v2 && atoi(env_v2) != 0 && (C & 3) == 0) {
        const size_t smem = (size_t)8 * (C + rd) * sizeof(float);
        if (smem > 64 * 1024) return cudaErrorInvalidValue;
        cudaError_t e = cudaFuncSetAttribute(se_kernel_v2<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        if (e != cudaSuccess) return e;
        const int threads = C >= 480 ? 1024 : (C >= 144 ? 512 : 256);
        se_kernel_v2<8><<<(unsigned)((frames + 7) / 8), threads, smem, s>>>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd);
        return cudaGetLastError();
    }
    static const int env_fpb = getenv("DFD_SE_FPB") ? atoi(getenv("DFD_SE_FPB")) : 8;      // experiments only
    if (env_fpb == 2) return launch_se_t<2>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, s);
    if (env_fpb == 4) return launch_se_t<4>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, s);
    return launch_se_t<8>(partials, nparts, inv_hw, w1, b1, w2t, b2, gate, frames, C, rd, s);
}

}  // namespace dfd
