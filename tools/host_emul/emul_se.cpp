// CPU emulation of se_kernel (deepfake_video_detection_b200/csrc/se.cu): kernel text between the
// DFD_SE1_KERNEL markers compiled UNCHANGED; threads + barriers, warp shuffles through a per-warp exchange buffer, shared
// memory pre-filled with NaN patterns.  Compared with sigmoid(W2 silu(W1 mean + b1) + b2) in double.
// Build + run: python tools/host_emul/run.py se
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
#define __align__(x)
#define __shared__

struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim;
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct ulonglong2 { unsigned long long x, y; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
using std::min;
static std::barrier<>* g_cta_bar = nullptr;
static void __syncthreads() { g_cta_bar->arrive_and_wait(); }

namespace dfd {
alignas(16) float smem[16 * 1024];
constexpr int kSeMaxThreads = 1024;
static inline uint64_t f2_pack(float a, float b) { float2 v{a, b}; uint64_t u; memcpy(&u, &v, 8); return u; }
static inline float2 f2_unpack(uint64_t u) { float2 v; memcpy(&v, &u, 8); return v; }
static inline uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { float2 x = f2_unpack(a), y = f2_unpack(b), z = f2_unpack(c); return f2_pack(fmaf(x.x, y.x, z.x), fmaf(x.y, y.y, z.y)); }
static inline float silu_f(float x) { return x / (1.0f + expf(-x)); }
static inline float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
template <typename P> static inline P __ldg(const P* p) { return *p; }
static inline void griddep_wait() {}      // programmatic dependent launch: nothing to wait for on the host
struct WarpX { float f[32]; std::barrier<> bar{32}; };
static std::vector<std::unique_ptr<WarpX>> g_warps;
static inline float __shfl_xor_sync(unsigned, float v, int o) {
    WarpX& w = *g_warps[threadIdx.x >> 5]; const int lane = threadIdx.x & 31;
    w.f[lane] = v; w.bar.arrive_and_wait();
    const float r = w.f[lane ^ o]; w.bar.arrive_and_wait();
    return r;
}
#include "se_kernel_v1.inc"
}  // namespace dfd

template <bool V2> static int run_case(int C, int rd, int nparts, int frames) {
    using namespace dfd;
    std::vector<float> parts((size_t)frames * nparts * C), w1((size_t)rd * C), b1(rd), w2t((size_t)rd * C), b2(C), gate((size_t)frames * C, NAN);
    uint32_t seed = 7u + C + rd;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
    for (auto& v : parts) v = rnd();
    for (auto& v : w1) v = 0.1f * rnd();
    for (auto& v : b1) v = 0.1f * rnd();
    for (auto& v : w2t) v = 0.3f * rnd();
    for (auto& v : b2) v = 0.3f * rnd();
    const float inv = 1.0f / 123.0f;
    const int threads = C >= 480 ? 1024 : (C >= 144 ? 512 : 256);
    blockDim.x = threads;
    const int grid = (frames + 7) / 8;
    for (int b = 0; b < grid; ++b) {
        memset(smem, 0xff, sizeof(smem));
        std::barrier<> bar(threads); g_cta_bar = &bar;
        g_warps.clear();
        for (int i = 0; i < threads / 32; ++i) g_warps.emplace_back(new WarpX());
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t, b]() { threadIdx.x = t; blockIdx.x = b;
                if (V2) abort();
                else se_kernel<8>(parts.data(), nparts, inv, w1.data(), b1.data(), w2t.data(), b2.data(), gate.data(), (int64_t)frames, C, rd); });
        for (auto& t : th) t.join();
    }
    double max_err = 0;
    for (int f = 0; f < frames; ++f) {
        std::vector<double> mean(C), r(rd);
        for (int c = 0; c < C; ++c) { double s = 0; for (int q = 0; q < nparts; ++q) s += parts[((size_t)f * nparts + q) * C + c]; mean[c] = s * inv; }
        for (int j = 0; j < rd; ++j) { double s = b1[j]; for (int c = 0; c < C; ++c) s += (double)w1[(size_t)j * C + c] * mean[c]; r[j] = s / (1 + exp(-s)); }
        for (int c = 0; c < C; ++c) {
            double s = b2[c]; for (int j = 0; j < rd; ++j) s += (double)w2t[(size_t)j * C + c] * r[j];
            max_err = fmax(max_err, fabs(1 / (1 + exp(-s)) - gate[(size_t)f * C + c]));
        }
    }
    const bool ok = max_err < 2e-6 && std::isfinite(max_err);
    printf("%s C=%d rd=%d nparts=%d frames=%d: %d CTAs x %d threads, max |err| %.2e -> %s\n", V2 ? "se_kernel_v2" : "se_kernel   ", C, rd, nparts, frames, grid, threads, max_err, ok ? "ok" : "MISMATCH");
    return ok ? 0 : 1;
}

int main() {
    int rc = 0;
    rc |= run_case<false>(32, 8, 32, 3); rc |= run_case<false>(96, 4, 8, 9); rc |= run_case<false>(144, 6, 8, 16); rc |= run_case<false>(240, 10, 4, 5);
    rc |= run_case<false>(480, 20, 2, 8); rc |= run_case<false>(672, 28, 2, 13); rc |= run_case<false>(1152, 48, 1, 17); rc |= run_case<false>(1152, 47, 3, 1);
    return rc;
}
