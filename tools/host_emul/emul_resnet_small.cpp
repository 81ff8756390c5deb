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
static float sm[64 * 1024];                                  // dynamic shared memory of rn_pool_head_kernel
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
    {   // attention pool + head, feature width 2048
        const int D = 2048;
        const std::vector<int> lens = {5, 1, 0, 9};
        std::vector<int32_t> off(1, 0);
        for (int t : lens) off.push_back(off.back() + t);
        const int V = (int)lens.size(), F = off.back();
        std::vector<float> feat((size_t)F * D), w1((size_t)64 * D), b1(64), w2(64), b2(1, 0.1f), f1((size_t)256 * D), fb1(256), f2(512), fb2(2);
        for (auto& v : feat) v = fabsf(rnd());
        for (auto& v : w1) v = rnd() * 0.04f;
        for (auto& v : b1) v = rnd() * 0.1f;
        for (auto& v : w2) v = rnd() * 0.5f;
        for (auto& v : f1) v = rnd() * 0.03f;
        for (auto& v : fb1) v = rnd() * 0.1f;
        for (auto& v : f2) v = rnd() * 0.2f;
        for (auto& v : fb2) v = rnd() * 0.1f;
        const RnHead hw{w1.data(), b1.data(), w2.data(), b2.data(), f1.data(), fb1.data(), f2.data(), fb2.data()};
        for (int att = 1; att >= 0; --att) {
            std::vector<float> logits((size_t)V * 2, 123.f), scores((size_t)F, 123.f);
            run_grid(V, 256, [&]() { rn_pool_head_kernel(hw, feat.data(), off.data(), D, att, logits.data(), scores.data()); });
            double max_l = 0, max_s = 0; bool nan_ok = true;
            for (int v = 0; v < V; ++v) {
                const int T = lens[v], f0 = off[v];
                if (T == 0) { nan_ok = std::isnan(logits[v * 2]) && std::isnan(logits[v * 2 + 1]); continue; }
                std::vector<double> wgt(T), pooled(D, 0.0), h1(256);
                if (att) {
                    double mx = -1e300, sum = 0;
                    for (int t = 0; t < T; ++t) {
                        double sc = b2[0];
                        for (int h = 0; h < 64; ++h) { double a = b1[h]; for (int c = 0; c < D; ++c) a += (double)w1[(size_t)h * D + c] * feat[(size_t)(f0 + t) * D + c]; sc += std::max(a, 0.0) * w2[h]; }
                        wgt[t] = 1 / (1 + exp(-sc)); mx = std::max(mx, wgt[t]);
                    }
                    for (int t = 0; t < T; ++t) { wgt[t] = exp(wgt[t] - mx); sum += wgt[t]; }
                    for (int t = 0; t < T; ++t) wgt[t] /= sum;
                } else for (int t = 0; t < T; ++t) wgt[t] = 1.0 / T;
                for (int t = 0; t < T; ++t) for (int c = 0; c < D; ++c) pooled[c] += wgt[t] * feat[(size_t)(f0 + t) * D + c];
                for (int j = 0; j < 256; ++j) { double a = fb1[j]; for (int c = 0; c < D; ++c) a += (double)f1[(size_t)j * D + c] * pooled[c]; h1[j] = std::max(a, 0.0); }
                for (int k = 0; k < 2; ++k) { double a = fb2[k]; for (int j = 0; j < 256; ++j) a += (double)f2[k * 256 + j] * h1[j]; max_l = fmax(max_l, fabs(a - logits[v * 2 + k])); }
                for (int t = 0; t < T; ++t) max_s = fmax(max_s, fabs(wgt[t] - scores[f0 + t]));
            }
            const bool ok = max_l < 5e-5 && max_s < 1e-6 && nan_ok;
            printf("rn_pool_head_kernel (%s, D = 2048): max |dlogit| %.2e, max |dscore| %.2e, empty video -> NaN %s -> %s\n", att ? "temporal attention" : "mean pool",
                   max_l, max_s, nan_ok ? "yes" : "NO", ok ? "ok" : "MISMATCH");
            rc |= !ok;
        }
    }
    return rc;
}
