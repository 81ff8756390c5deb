// CPU emulation of the LogicRNNLSTM head's CUDA-core kernels (deepfake_video_detection_b200/csrc/rnn.cu, GPU-verified, unchanged
// text; reference src/RNNModel.py:24-39 LogicCell gate math with the length mask of :120-125, :128-133 attention over time +
// classifier + sigmoid) against the same arithmetic in double.
// Build + run: python tools/host_emul/run.py rnn
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
#define __shared__
#ifndef INFINITY
#define INFINITY __builtin_inff()
#endif
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim;
typedef _Float16 __half;
static std::barrier<>* g_cta_bar = nullptr;
static void __syncthreads() { g_cta_bar->arrive_and_wait(); }

namespace dfd {
static float sm[64 * 1024];
template <typename T> struct Half16;
template <> struct Half16<__half> { static __half from_float(float v) { return (_Float16)v; } };
template <typename P> static inline P __ldg(const P* p) { return *p; }
struct WarpX { float f[32]; std::barrier<> bar{32}; };
static std::vector<std::unique_ptr<WarpX>> g_warps;
static inline float __shfl_xor_sync(unsigned, float v, int o) {
    WarpX& w = *g_warps[threadIdx.x >> 5]; const int lane = threadIdx.x & 31;
    w.f[lane] = v; w.bar.arrive_and_wait();
    const float r = w.f[lane ^ o]; w.bar.arrive_and_wait();
    return r;
}
#include "rnn_kernels.inc"
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
static double sig(double x) { return 1 / (1 + exp(-x)); }

int main() {
    using namespace dfd;
    int rc = 0;
    uint32_t seed = 77;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
    {   // LogicCell gate math, one time step with a length mask
        const int B = 3, H = 64, t = 2;
        std::vector<float> g((size_t)B * 7 * H), c((size_t)B * H), c0, out((size_t)B * H, 9.f);
        std::vector<_Float16> h16((size_t)B * (H + 8));
        std::vector<int> lengths = {5, 2, 3};                       // sequence 1 is already over at t = 2
        for (auto& v : g) v = 2.f * rnd();
        for (auto& v : c) v = rnd();
        c0 = c;
        run_grid((B * H + 127) / 128, 128, [&]() { rnn_cell_kernel<__half>(g.data(), c.data(), h16.data(), H + 8, out.data(), H, lengths.data(), t, B, H); });
        double e = 0;
        for (int b = 0; b < B; ++b) for (int k = 0; k < H; ++k) {
            const float* gr = &g[(size_t)b * 7 * H + k];
            const double cn = sig(gr[2 * H]) * c0[b * H + k] + sig(gr[3 * H]) * tanh(gr[4 * H]);
            const double cl = sig(gr[0]) * cn + sig(gr[H]) * tanh(gr[6 * H]);
            const double hn = sig(gr[5 * H]) * tanh(cl);
            e = fmax(e, fabs(cl - c[b * H + k]));
            e = fmax(e, fabs(hn * (t < lengths[b] ? 1.0 : 0.0) - out[b * H + k]));
            e = fmax(e, fabs(hn - (double)(float)h16[(size_t)b * (H + 8) + k]) - 1e-3 * fabs(hn));      // fp16 copy for the next GEMM
        }
        printf("rnn_cell_kernel: max |err| %.2e -> %s\n", e, e < 2e-6 ? "ok" : "MISMATCH"); rc |= !(e < 2e-6);
    }
    {   // attention over time + classifier + sigmoid
        const int B = 3, Tn = 5, H = 64;
        std::vector<float> outs((size_t)B * Tn * H), w1t((size_t)H * H), b1(H), w2(H), b2(1, 0.05f), c1t((size_t)H * H), cb1(H), c2(H), cb2(1, -0.1f), prob(B, 9.f);
        for (auto& v : outs) v = rnd();
        for (auto& v : w1t) v = 0.2f * rnd();
        for (auto& v : b1) v = 0.1f * rnd();
        for (auto& v : w2) v = 0.5f * rnd();
        for (auto& v : c1t) v = 0.2f * rnd();
        for (auto& v : cb1) v = 0.1f * rnd();
        for (auto& v : c2) v = 0.5f * rnd();
        run_grid(B, 64, [&]() { rnn_head_kernel(outs.data(), Tn, H, w1t.data(), b1.data(), w2.data(), b2.data(), c1t.data(), cb1.data(), c2.data(), cb2.data(), prob.data()); });
        double e = 0;
        for (int b = 0; b < B; ++b) {
            std::vector<double> a(Tn), ctx(H, 0.0);
            double mx = -1e300, sum = 0;
            for (int t = 0; t < Tn; ++t) {
                double s = b2[0];
                for (int j = 0; j < H; ++j) { double acc = b1[j]; for (int k = 0; k < H; ++k) acc += (double)outs[((size_t)b * Tn + t) * H + k] * w1t[(size_t)k * H + j]; s += tanh(acc) * w2[j]; }
                a[t] = s; mx = std::max(mx, s);
            }
            for (int t = 0; t < Tn; ++t) { a[t] = exp(a[t] - mx); sum += a[t]; }
            for (int t = 0; t < Tn; ++t) for (int k = 0; k < H; ++k) ctx[k] += a[t] / sum * outs[((size_t)b * Tn + t) * H + k];
            double s = cb2[0];
            for (int j = 0; j < H; ++j) { double acc = cb1[j]; for (int k = 0; k < H; ++k) acc += ctx[k] * c1t[(size_t)k * H + j]; s += std::max(acc, 0.0) * c2[j]; }
            e = fmax(e, fabs(sig(s) - prob[b]));
        }
        printf("rnn_head_kernel: max |dprob| %.2e -> %s\n", e, e < 2e-6 ? "ok" : "MISMATCH"); rc |= !(e < 2e-6);
    }
    return rc;
}
