// Micro-probe: MUFU throughput per SM for tanh.approx.f32, tanh.approx.f16x2, ex2.approx.f32 (values per clock per SM).
#include <cstdio>
#include <cuda_fp16.h>
template <int MODE> __global__ void k(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    unsigned ha = 0x3c003800u + threadIdx.x, hb = ha + 1, hc = ha + 2, hd = ha + 3;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d)); }
        if (MODE == 1) { asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ha)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hb)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hd)); }
        if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
        if (MODE == 3) { asm volatile("fma.rn.f32x2 %0, %0, %0, %0;" : "+l"(*(unsigned long long*)&a)); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ha ^ hb ^ hc ^ hd);
}
int main() {
    float* o; cudaMalloc(&o, 148 * 1024 * 4 * 8);
    const char* names[] = {"tanh.approx.f32", "tanh.approx.f16x2 (values = 2 per op)", "ex2.approx.f32"};
    for (int m = 0; m < 3; ++m) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        const int iters = 20000, blocks = 148 * 2, threads = 1024;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (m == 0) k<0><<<blocks, threads>>>(o, iters); if (m == 1) k<1><<<blocks, threads>>>(o, iters); if (m == 2) k<2><<<blocks, threads>>>(o, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
        }
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double ops = (double)blocks * threads * iters * 4;
        printf("%-40s %.2f ops/clk/SM at 1.965 GHz (%.3f ms)\n", names[m], ops / (ms * 1e-3) / 148 / 1.965e9, ms);
    }
    return 0;
}
