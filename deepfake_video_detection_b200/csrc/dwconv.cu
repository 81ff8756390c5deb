// K2 — MBConv depthwise conv kxk (k in {3,5}, stride in {1,2}, pad k/2) + folded BN + SiLU, NHWC 16-bit,
// fused with the squeeze-excite spatial reduction (timm `conv_dw` + `bn` + the `x.mean((2,3))` of
// SqueezeExcite; reference call site pretrained_detector.py:116).
//
// One thread owns 8 channels (one 128-bit vector) x TW consecutive output columns of one output row:
// each input row of the window is loaded once as (TW-1)*stride+k vectors and reused from registers for the
// TW outputs; vertical reuse is served by L1.  Consecutive threads walk channel groups first, so a warp's
// loads are runs of contiguous 16-byte vectors (coalesced NHWC).  fp32 accumulation, fp32 weights.
//
// SE squeeze: every thread sums its SiLU outputs (fp32, before the 16-bit rounding) per channel; the
// block combines them in a fixed order and writes one partial row per block — no atomics, so the result
// is bit-reproducible.  se.cu adds the partial rows up in order.
#include "common.cuh"
#include "kernels.h"

namespace dfd {

constexpr int kDwThreads = 256;
constexpr int kDwTW = 4;

static inline int dw_items(int OH, int OW, int C) { return OH * ((OW + kDwTW - 1) / kDwTW) * (C / 8); }
int dw_num_partials(int OH, int OW, int C) { return (dw_items(OH, OW, C) + kDwThreads - 1) / kDwThreads; }

template <typename T, int KS, int STRIDE>
__global__ void __launch_bounds__(kDwThreads, 2)
dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
              T* __restrict__ out, float* __restrict__ partials,
              int H, int W, int C, int OH, int OW, int strips, int items, int blocks_per_frame) {
    constexpr int TW = kDwTW;
    constexpr int PAD = KS / 2;
    constexpr int NCOL = (TW - 1) * STRIDE + KS;
    __shared__ float s_part[kDwThreads][9];     // +1 pad: conflict-free column walks

    const int C8 = C >> 3;
    const int64_t frame = blockIdx.x / blocks_per_frame;
    const int blk = blockIdx.x - (int)(frame * blocks_per_frame);
    const int item = blk * kDwThreads + threadIdx.x;
    const bool valid = item < items;

    float sums[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) sums[c] = 0.f;

    if (valid) {
        const int c8 = item % C8;
        const int t = item / C8;
        const int strip = t % strips;
        const int oy = t / strips;
        const int ox0 = strip * TW;
        const int iy0 = oy * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;
        const T* in_f = in + (size_t)frame * H * W * C + c8 * 8;
        const float* wc = w + c8 * 8;

        float acc[TW][8];
        {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
#pragma unroll
            for (int j = 0; j < TW; ++j) {
                acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
                acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
            }
        }
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= H) continue;
            const T* row = in_f + (size_t)iy * W * C;
            uint4 v[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const int ix = ix0 + j;
                v[j] = (ix >= 0 && ix < W) ? ldg16(row + (size_t)ix * C) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + (ky * KS + kx) * C));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + (ky * KS + kx) * C + 4));
#pragma unroll
                for (int j = 0; j < TW; ++j) {
                    const uint4 x = v[j * STRIDE + kx];
                    const float2 x0 = Half16<T>::unpack(x.x), x1 = Half16<T>::unpack(x.y);
                    const float2 x2 = Half16<T>::unpack(x.z), x3 = Half16<T>::unpack(x.w);
                    acc[j][0] = fmaf(x0.x, w0.x, acc[j][0]); acc[j][1] = fmaf(x0.y, w0.y, acc[j][1]);
                    acc[j][2] = fmaf(x1.x, w0.z, acc[j][2]); acc[j][3] = fmaf(x1.y, w0.w, acc[j][3]);
                    acc[j][4] = fmaf(x2.x, w1.x, acc[j][4]); acc[j][5] = fmaf(x2.y, w1.y, acc[j][5]);
                    acc[j][6] = fmaf(x3.x, w1.z, acc[j][6]); acc[j][7] = fmaf(x3.y, w1.w, acc[j][7]);
                }
            }
        }
        T* orow = out + (((size_t)frame * OH + oy) * OW) * C + c8 * 8;
#pragma unroll
        for (int j = 0; j < TW; ++j) {
            const int ox = ox0 + j;
            if (ox < OW) {
                float y[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { y[c] = silu_f(acc[j][c]); sums[c] += y[c]; }
                uint4 o;
                o.x = Half16<T>::pack(y[0], y[1]); o.y = Half16<T>::pack(y[2], y[3]);
                o.z = Half16<T>::pack(y[4], y[5]); o.w = Half16<T>::pack(y[6], y[7]);
                stg16(orow + (size_t)ox * C, o);
            }
        }
    }

    // ---- deterministic block reduction of the SE sums --------------------------------------------
#pragma unroll
    for (int c = 0; c < 8; ++c) s_part[threadIdx.x][c] = sums[c];
    __syncthreads();
    if (threadIdx.x < C8) {
        const int cg = threadIdx.x;
        const int base = (blk * kDwThreads) % C8;          // channel group of thread 0 in this block
        int t0 = cg - base; if (t0 < 0) t0 += C8;
        float tot[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) tot[c] = 0.f;
        for (int t = t0; t < kDwThreads; t += C8) {
#pragma unroll
            for (int c = 0; c < 8; ++c) tot[c] += s_part[t][c];
        }
        float* dst = partials + ((size_t)frame * blocks_per_frame + blk) * C + cg * 8;
        *reinterpret_cast<float4*>(dst) = make_float4(tot[0], tot[1], tot[2], tot[3]);
        *reinterpret_cast<float4*>(dst + 4) = make_float4(tot[4], tot[5], tot[6], tot[7]);
    }
}

template <typename T>
static cudaError_t launch_dw_t(const void* in, const float* w, const float* bias, void* out, float* partials,
                               int64_t frames, int H, int W, int C, int k, int stride, cudaStream_t s) {
    const int pad = k / 2;
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    const int strips = (OW + kDwTW - 1) / kDwTW;
    const int items = dw_items(OH, OW, C);
    const int bpf = dw_num_partials(OH, OW, C);
    if (frames <= 0) return cudaSuccess;
    if ((C & 7) || C / 8 > kDwThreads || frames * (int64_t)bpf > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(frames * bpf);
#define DFD_DW(KS, ST) dwconv_kernel<T, KS, ST><<<grid, kDwThreads, 0, s>>>((const T*)in, w, bias, (T*)out, partials, H, W, C, OH, OW, strips, items, bpf)
    if (k == 3 && stride == 1) DFD_DW(3, 1);
    else if (k == 3 && stride == 2) DFD_DW(3, 2);
    else if (k == 5 && stride == 1) DFD_DW(5, 1);
    else if (k == 5 && stride == 2) DFD_DW(5, 2);
    else return cudaErrorInvalidValue;
#undef DFD_DW
    return cudaGetLastError();
}

cudaError_t launch_dwconv(const void* in, const float* w, const float* bias, void* out, float* partials,
                          int64_t frames, int H, int W, int C, int k, int stride, int dtype, cudaStream_t s) {
    if (dtype == kDtypeFP16) return launch_dw_t<__half>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
    return launch_dw_t<__nv_bfloat16>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
}

}  // namespace dfd
