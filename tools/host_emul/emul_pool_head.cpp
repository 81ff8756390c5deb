// CPU emulation of pool_head_kernel (deepfake_video_detection_b200/csrc/poolhead.cu; reference pretrained_detector.py:123-141):
// the GPU-verified kernel text between the DFD_POOLHEAD_KERNEL markers, unchanged (only `extern __shared__ float s_f[];` is
// redirected to a host buffer by run.py; the static __shared__ arrays become function statics), on CPU threads with
// std::barrier / per-warp shuffles.  Compared with the reference arithmetic in double on ragged videos, with and without
// temporal attention, including an empty video (NaN logits by contract).
// Build + run: python tools/host_emul/run.py poolhead
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
#define __shared__ static
#ifndef INFINITY
#define INFINITY __builtin_inff()
#endif
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
using std::min;
static std::barrier<>* g_cta_bar = nullptr;
static void __syncthreads() { g_cta_bar->arrive_and_wait(); }

namespace dfd {
struct HeadWeights { const float *att_w1, *att_b1, *att_w2, *att_b2, *fc1_w, *fc1_b, *fc2_w, *fc2_b; };
static float g_dyn_smem[32 * 1280];
template <typename P> static inline P __ldg(const P* p) { return *p; }
static inline void griddep_wait() {}      // programmatic dependent launch: nothing to wait for on the host
static inline float __int_as_float(int v) { float f; memcpy(&f, &v, 4); return f; }
struct WarpX { float f[32]; std::barrier<> bar{32}; };
static std::vector<std::unique_ptr<WarpX>> g_warps;
static inline float __shfl_xor_sync(unsigned, float v, int o) {
    WarpX& w = *g_warps[threadIdx.x >> 5]; const int lane = threadIdx.x & 31;
    w.f[lane] = v; w.bar.arrive_and_wait();
    const float r = w.f[lane ^ o]; w.bar.arrive_and_wait();
    return r;
}
#include "pool_head_kernel.inc"
}  // namespace dfd

template <int FEAT, int CHUNK> static int run() {
    using namespace dfd;
    constexpr int kFeat = FEAT;                                        // shadows the namespace constant inside this function
    const std::vector<int> lens = {8, 1, 3, 0, 40, 33, 2};              // ragged, one empty, two longer than one staged chunk
    std::vector<int32_t> off(1, 0);
    for (int t : lens) off.push_back(off.back() + t);
    const int V = (int)lens.size(), F = off.back();
    uint32_t seed = 2024;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
    std::vector<float> feat((size_t)F * kFeat), w1((size_t)64 * kFeat), b1(64), w2(64), b2(1), f1((size_t)256 * kFeat), fb1(256), f2(2 * 256), fb2(2);
    for (auto& v : feat) v = fabsf(rnd()) * 1.5f;
    for (auto& v : w1) v = rnd() * 0.05f;
    for (auto& v : b1) v = rnd() * 0.1f;
    for (auto& v : w2) v = rnd() * 0.5f;
    b2[0] = 0.1f;
    for (auto& v : f1) v = rnd() * 0.04f;
    for (auto& v : fb1) v = rnd() * 0.1f;
    for (auto& v : f2) v = rnd() * 0.2f;
    for (auto& v : fb2) v = rnd() * 0.1f;
    const HeadWeights hw{w1.data(), b1.data(), w2.data(), b2.data(), f1.data(), fb1.data(), f2.data(), fb2.data()};
    int rc = 0;
    for (int att = 1; att >= 0; --att) {
        std::vector<float> logits((size_t)V * 2, 123.f), scores((size_t)F, 123.f);
        for (int b = 0; b < V; ++b) {
            std::barrier<> bar(kPhThreads); g_cta_bar = &bar;
            g_warps.clear();
            for (int i = 0; i < kPhThreads / 32; ++i) g_warps.emplace_back(new WarpX());
            std::vector<std::thread> th;
            for (int t = 0; t < kPhThreads; ++t)
                th.emplace_back([&, t, b]() { threadIdx.x = t; blockIdx.x = b; pool_head_kernel<FEAT, CHUNK>(hw, feat.data(), off.data(), (int)(feat.size() / FEAT), att, logits.data(), scores.data()); });
            for (auto& t : th) t.join();
        }
        double max_l = 0, max_s = 0; bool nan_ok = true;
        for (int v = 0; v < V; ++v) {
            const int T = lens[v], f0 = off[v];
            if (T == 0) { nan_ok = std::isnan(logits[v * 2]) && std::isnan(logits[v * 2 + 1]); continue; }
            std::vector<double> wgt(T), pooled(kFeat, 0.0), h1(256);
            if (att) {
                double mx = -1e300, sum = 0;
                for (int t = 0; t < T; ++t) {
                    double sc = b2[0];
                    for (int h = 0; h < 64; ++h) { double a = b1[h]; for (int c = 0; c < kFeat; ++c) a += (double)w1[(size_t)h * kFeat + c] * feat[(size_t)(f0 + t) * kFeat + c]; sc += std::max(a, 0.0) * w2[h]; }
                    wgt[t] = 1 / (1 + exp(-sc)); mx = std::max(mx, wgt[t]);
                }
                for (int t = 0; t < T; ++t) { wgt[t] = exp(wgt[t] - mx); sum += wgt[t]; }
                for (int t = 0; t < T; ++t) wgt[t] /= sum;
            } else for (int t = 0; t < T; ++t) wgt[t] = 1.0 / T;
            for (int t = 0; t < T; ++t) for (int c = 0; c < kFeat; ++c) pooled[c] += wgt[t] * feat[(size_t)(f0 + t) * kFeat + c];
            for (int j = 0; j < 256; ++j) { double a = fb1[j]; for (int c = 0; c < kFeat; ++c) a += (double)f1[(size_t)j * kFeat + c] * pooled[c]; h1[j] = std::max(a, 0.0); }
            for (int k = 0; k < 2; ++k) { double a = fb2[k]; for (int j = 0; j < 256; ++j) a += (double)f2[k * 256 + j] * h1[j]; max_l = fmax(max_l, fabs(a - logits[v * 2 + k])); }
            for (int t = 0; t < T; ++t) max_s = fmax(max_s, fabs(wgt[t] - scores[f0 + t]));
        }
        const bool ok = max_l < 2e-5 && max_s < 1e-6 && nan_ok && std::isfinite(max_l);
        printf("pool_head_kernel<%d, %d> (%s): %d ragged videos, max |dlogit| %.2e, max |dscore| %.2e, empty video -> NaN %s -> %s\n",
               FEAT, CHUNK, att ? "temporal attention" : "mean pool", V, max_l, max_s, nan_ok ? "yes" : "NO", ok ? "ok" : "MISMATCH");
        rc |= ok ? 0 : 1;
    }
    return rc;
}

int main() { return run<1280, 32>() | run<2048, 16>(); }
