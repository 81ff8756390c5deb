// Stem: conv3x3 stride 2 pad 1, 3->32, folded BN, SiLU  (timm conv_stem + bn1 = backbone.0/backbone.1 of
// pretrained_detector.py:46).  Output NHWC 16-bit.  Three input layouts:
//   DFD_IN_U8_HWC   uint8 crops; the prep of app.py:2084-2085 is fused (per-block 3x256 fp32 table, exact
//                   reference arithmetic, NO 16-bit rounding of the input), zero padding applied after
//                   normalisation as the reference does
//   DFD_IN_F32_NCHW fp32 normalised frames (what forward() receives)
//   DFD_IN_H16_NCHW K1 output
// One thread = one output pixel x 32 channels, fp32 FMAs, weights broadcast from shared memory.
#include "common.cuh"
#include "kernels.h"

namespace dfd {

// DFD_STEM_KERNEL_BEGIN   (tools/host_emul/ runs this kernel, unchanged, on CPU threads)
template <typename T, int IN_KIND>
__global__ void __launch_bounds__(128) stem_kernel(const void* __restrict__ in_, const float* __restrict__ w,
                                                   const float* __restrict__ bias, T* __restrict__ out,
                                                   int H, int W, int OH, int OW, int64_t total) {
    __shared__ __align__(16) float sw[27 * 32];
    __shared__ float sb[32];
    __shared__ float lut[IN_KIND == 0 ? 768 : 1];
    for (int i = threadIdx.x; i < 27 * 32; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 32) sb[threadIdx.x] = bias[threadIdx.x];
    if (IN_KIND == 0) {
        for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i] = prep_value(i >> 8, i & 255);
    }
    __syncthreads();
    const int64_t pix = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= total) return;
    const int ohw = OH * OW;
    const int64_t frame = pix / ohw;
    const int rem = (int)(pix - frame * ohw);
    const int oy = rem / OW, ox = rem - oy * OW;

    float x[27];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy - 1 + ky;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * ox - 1 + kx;
            const bool ok = (iy >= 0) && (iy < H) && (ix >= 0) && (ix < W);
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                float v = 0.f;
                if (ok) {
                    if (IN_KIND == 0) {
                        const uint8_t* p = reinterpret_cast<const uint8_t*>(in_) + ((frame * H + iy) * W + ix) * 3 + c;
                        v = lut[c * 256 + __ldg(p)];
                    } else if (IN_KIND == 1) {
                        v = __ldg(reinterpret_cast<const float*>(in_) + ((frame * 3 + c) * H + iy) * W + ix);
                    } else {
                        v = Half16<T>::to_float(reinterpret_cast<const T*>(in_)[((frame * 3 + c) * H + iy) * W + ix]);
                    }
                }
                x[(ky * 3 + kx) * 3 + c] = v;
            }
        }
    }
    float acc[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) acc[o] = sb[o];
#pragma unroll
    for (int i = 0; i < 27; ++i) {
        const float xi = x[i];
#pragma unroll
        for (int o4 = 0; o4 < 8; ++o4) {
            const float4 wv = *reinterpret_cast<const float4*>(&sw[i * 32 + o4 * 4]);
            acc[o4 * 4 + 0] = fmaf(xi, wv.x, acc[o4 * 4 + 0]);
            acc[o4 * 4 + 1] = fmaf(xi, wv.y, acc[o4 * 4 + 1]);
            acc[o4 * 4 + 2] = fmaf(xi, wv.z, acc[o4 * 4 + 2]);
            acc[o4 * 4 + 3] = fmaf(xi, wv.w, acc[o4 * 4 + 3]);
        }
    }
    T* dst = out + pix * 32;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        U32x8 o;
#pragma unroll
        for (int p = 0; p < 8; ++p)
            o.v[p] = Half16<T>::pack(silu_f(acc[h * 16 + 2 * p]), silu_f(acc[h * 16 + 2 * p + 1]));
        stg32(dst + h * 16, o);
    }
}

// DFD_STEM_KERNEL_END

template <typename T>
static cudaError_t launch_stem_t(const void* in, int in_kind, const float* w, const float* bias, void* out,
                                 int64_t frames, int H, int W, cudaStream_t s) {
    const int OH = H / 2, OW = W / 2;
    const int64_t total = frames * OH * OW;
    if (total <= 0) return cudaSuccess;
    const unsigned blocks = (unsigned)((total + 127) / 128);
    if (in_kind == 0) stem_kernel<T, 0><<<blocks, 128, 0, s>>>(in, w, bias, (T*)out, H, W, OH, OW, total);
    else if (in_kind == 1) stem_kernel<T, 1><<<blocks, 128, 0, s>>>(in, w, bias, (T*)out, H, W, OH, OW, total);
    else stem_kernel<T, 2><<<blocks, 128, 0, s>>>(in, w, bias, (T*)out, H, W, OH, OW, total);
    return cudaGetLastError();
}

cudaError_t launch_stem(const void* in, int in_kind, const float* w, const float* bias, void* out,
                        int64_t frames, int H, int W, int dtype, cudaStream_t s) {
    if (dtype == kDtypeFP16) return launch_stem_t<__half>(in, in_kind, w, bias, out, frames, H, W, s);
    return launch_stem_t<__nv_bfloat16>(in, in_kind, w, bias, out, frames, H, W, s);
}

}  // namespace dfd
