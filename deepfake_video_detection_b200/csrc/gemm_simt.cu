// CUDA-core GEMM with the contract of gemm_tc.cu.  Bring-up / bisecting aid only (DFD_GEMM_IMPL=simt):
// one thread per output element, fp32 accumulate, same rounding points as the tensor-core kernel
// (gated A operand re-rounded to the 16-bit type before the product).
#include "common.cuh"
#include "kernels.h"

namespace dfd {

template <typename T>
__global__ void gemm_simt_kernel(const T* __restrict__ A, const T* __restrict__ W, const float* __restrict__ bias,
                                 const float* __restrict__ gate, const T* __restrict__ R, T* __restrict__ D,
                                 int64_t M, int K, int N, int HW, int act) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * N) return;
    const int64_t m = idx / N;
    const int n = (int)(idx - m * N);
    const float* g = gate ? gate + (size_t)(m / HW) * K : nullptr;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
        float a = Half16<T>::to_float(A[(size_t)m * K + k]);
        if (g) a = Half16<T>::to_float(Half16<T>::from_float(a * g[k]));
        acc = fmaf(a, Half16<T>::to_float(W[(size_t)n * K + k]), acc);
    }
    acc += bias[n];
    if (act) acc = silu_f(acc);
    if (R) acc += Half16<T>::to_float(R[idx]);
    D[idx] = Half16<T>::from_float(acc);
}

// conv_head + SiLU + average pool: one thread per (frame, n)
template <typename T>
__global__ void gemm_simt_pool_kernel(const T* __restrict__ A, const T* __restrict__ W, const float* __restrict__ bias,
                                      float* __restrict__ feat, int64_t frames, int K, int N, int HW) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= frames * N) return;
    const int64_t f = idx / N;
    const int n = (int)(idx - f * N);
    float tot = 0.f;
    for (int r = 0; r < HW; ++r) {
        float acc = 0.f;
        const T* a = A + ((size_t)f * HW + r) * K;
        for (int k = 0; k < K; ++k) acc = fmaf(Half16<T>::to_float(a[k]), Half16<T>::to_float(W[(size_t)n * K + k]), acc);
        tot += silu_f(acc + bias[n]);
    }
    feat[idx] = tot / (float)HW;
}

cudaError_t launch_gemm_simt(const void* A, const void* W, const float* bias, const float* gate, const void* R,
                             void* D, float* pool_feat, int64_t M, int K, int N, int HW, int act, int dtype,
                             cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if (pool_feat) {
        const int64_t frames = M / HW, total = frames * N;
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (dtype == kDtypeFP16) gemm_simt_pool_kernel<__half><<<grid, 256, 0, s>>>((const __half*)A, (const __half*)W, bias, pool_feat, frames, K, N, HW);
        else gemm_simt_pool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)W, bias, pool_feat, frames, K, N, HW);
        return cudaGetLastError();
    }
    const int64_t total = M * N;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == kDtypeFP16) gemm_simt_kernel<__half><<<grid, 256, 0, s>>>((const __half*)A, (const __half*)W, bias, gate, (const __half*)R, (__half*)D, M, K, N, HW, act);
    else gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)A, (const __nv_bfloat16*)W, bias, gate, (const __nv_bfloat16*)R, (__nv_bfloat16*)D, M, K, N, HW, act);
    return cudaGetLastError();
}

}  // namespace dfd
