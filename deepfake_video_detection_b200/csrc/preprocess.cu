// K1 — tensor prep: uint8 HWC crops -> normalised 16-bit NCHW.
// Reference: app.py:2084 (`from_numpy(faces).permute(0,3,1,2).float()/255.0`) and imagenet_normalize,
// app.py:1772-1780.  Only 3x256 distinct results exist, so every block builds the table once with the
// reference's exact fp32 operation order (IEEE divide, subtract, IEEE divide), rounds it once to the
// storage type, and the streaming part is 48-byte loads / 32-byte stores per thread (HBM-bound).
#include "common.cuh"
#include "kernels.h"

namespace dfd {

// DFD_PREP_KERNEL_BEGIN   (tools/host_emul/ runs this kernel, unchanged, on CPU threads)
template <typename T>
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ in, T* __restrict__ out,
                                                         int64_t groups, int groups_per_frame, int HW) {
    __shared__ T lut[3][256];
    for (int i = threadIdx.x; i < 768; i += blockDim.x) lut[i >> 8][i & 255] = Half16<T>::from_float(prep_value(i >> 8, i & 255));
    __syncthreads();
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // one group = 16 pixels
    if (g >= groups) return;
    const int64_t frame = g / groups_per_frame;
    const int gi = (int)(g - frame * groups_per_frame);
    const uint4* src = reinterpret_cast<const uint4*>(in + (frame * HW + (int64_t)gi * 16) * 3);
    uint32_t b[12];
    {
        uint4 a0 = ldg16_stream(src), a1 = ldg16_stream(src + 1), a2 = ldg16_stream(src + 2);
        b[0] = a0.x; b[1] = a0.y; b[2] = a0.z; b[3] = a0.w; b[4] = a1.x; b[5] = a1.y; b[6] = a1.z; b[7] = a1.w;
        b[8] = a2.x; b[9] = a2.y; b[10] = a2.z; b[11] = a2.w;
    }
    auto byte_at = [&](int i) -> int { return (b[i >> 2] >> ((i & 3) * 8)) & 0xff; };
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        U32x8 o;
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            uint16_t lo = *reinterpret_cast<const uint16_t*>(&lut[c][byte_at((2 * p) * 3 + c)]);
            uint16_t hi = *reinterpret_cast<const uint16_t*>(&lut[c][byte_at((2 * p + 1) * 3 + c)]);
            o.v[p] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        stg32(out + (frame * 3 + c) * HW + (int64_t)gi * 16, o);
    }
}

// DFD_PREP_KERNEL_END

cudaError_t launch_preprocess(const uint8_t* in, void* out, int64_t frames, int H, int W, int dtype, cudaStream_t s) {
    const int HW = H * W;
    if (frames <= 0) return cudaSuccess;
    const int gpf = HW / 16;
    const int64_t groups = frames * gpf;
    const unsigned blocks = (unsigned)((groups + 255) / 256);
    if (dtype == kDtypeFP16)
        preprocess_kernel<__half><<<blocks, 256, 0, s>>>(in, (__half*)out, groups, gpf, HW);
    else
        preprocess_kernel<__nv_bfloat16><<<blocks, 256, 0, s>>>(in, (__nv_bfloat16*)out, groups, gpf, HW);
    return cudaGetLastError();
}

}  // namespace dfd
