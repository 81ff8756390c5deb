// Micro-probes (B200): (1) MUFU throughput per SM for tanh.approx.f32 / tanh.approx.f16x2 / ex2.approx.f32;
// (2) does a MUFU warp instruction hold the issue port for its 8 pipe cycles?  One MUFU + N independent FFMAs per loop
//     iteration, 16 warps per scheduler: cycles per iteration per scheduler = max(8, N + 1) if the pipes overlap, 8 + N if not.
#include <cstdio>
#include <cuda_fp16.h>
template <int MODE> __global__ void k(float* out, int iters) {
    float a = threadIdx.x * 1e-3f, b = a + 0.1f, c = a + 0.2f, d = a + 0.3f;
    unsigned ha = 0x3c003800u + threadIdx.x, hb = ha + 1, hc = ha + 2, hd = ha + 3;
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) { asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(b)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(c)); asm volatile("tanh.approx.f32 %0, %0;" : "+f"(d)); }
        if (MODE == 1) { asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(ha)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hb)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hc)); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(hd)); }
        if (MODE == 2) { asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(b)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(c)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(d)); }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a + b + c + d + __uint_as_float(ha ^ hb ^ hc ^ hd);
}
template <int N, bool MUFU> __global__ void mix(float* out, int iters) {
    float m = threadIdx.x * 1e-3f;
    float f[12];
#pragma unroll
    for (int j = 0; j < 12; ++j) f[j] = m + j;
    for (int i = 0; i < iters; ++i) {
        if (MUFU) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(m));
#pragma unroll
        for (int j = 0; j < N; ++j) asm volatile("fma.rn.f32 %0, %0, %1, %0;" : "+f"(f[j % 12]) : "f"(1.0001f));
    }
    float s = m;
#pragma unroll
    for (int j = 0; j < 12; ++j) s += f[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int N, bool MUFU> static void run_mix(float* o) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 20000, blocks = 148 * 2, threads = 1024;      // 64 warps per SM = 16 per scheduler
    float ms = 0;
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); mix<N, MUFU><<<blocks, threads>>>(o, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1); }
    const double warp_iters_per_sched = (double)blocks * threads / 32 * iters / (148 * 4);
    printf("%s + %2d FFMA per iteration: %.2f cycles per warp-iteration per scheduler\n", MUFU ? "1 MUFU" : "0 MUFU", N, ms * 1e-3 * 1.965e9 / warp_iters_per_sched);
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
    run_mix<0, true>(o); run_mix<2, true>(o); run_mix<4, true>(o); run_mix<6, true>(o); run_mix<8, true>(o); run_mix<12, true>(o); run_mix<16, true>(o);
    run_mix<4, false>(o); run_mix<8, false>(o); run_mix<16, false>(o);
    return 0;
}
