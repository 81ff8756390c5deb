// CPU emulation of two GPU-verified default kernels, unchanged text: preprocess_kernel (K1, row a1: app.py:2084-2085 +
// imagenet_normalize app.py:1772-1780 — must be BIT-EXACT with the reference's fp32 arithmetic rounded once to the storage
// type) and stem_kernel (row a3 for the fp32 NCHW input forward() receives, and for uint8 crops with the prep fused).
// Build + run: python tools/host_emul/run.py prepstem
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __align__(x)
#define __shared__ static
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 blockDim;
struct uint4 { uint32_t x, y, z, w; };
struct float4 { float x, y, z, w; };
typedef _Float16 __half;
static std::barrier<>* g_cta_bar = nullptr;
static void __syncthreads() { g_cta_bar->arrive_and_wait(); }

namespace dfd {
struct U32x8 { uint32_t v[8]; };
template <typename T> struct Half16;
template <> struct Half16<__half> {
    static __half from_float(float v) { return (_Float16)v; }
    static float to_float(__half v) { return (float)v; }
    static uint32_t pack(float a, float b) { _Float16 h[2] = {(_Float16)a, (_Float16)b}; uint32_t v; memcpy(&v, h, 4); return v; }
};
static inline float imagenet_mean(int c) { return c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f); }
static inline float imagenet_std(int c) { return c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f); }
static inline float prep_value(int c, int v) {            // common.cuh: IEEE fp32 divide, subtract, divide
    volatile float x = (float)v / 255.0f; volatile float d = x - imagenet_mean(c); return d / imagenet_std(c);
}
static inline float silu_f(float x) { return x / (1.0f + expf(-x)); }
template <typename P> static inline P __ldg(const P* p) { return *p; }
static inline uint4 ldg16_stream(const void* p) { uint4 v; memcpy(&v, p, 16); return v; }
static inline void stg32(void* p, const U32x8& r) { memcpy(p, &r, 32); }
#include "preprocess_kernel.inc"
#include "stem_kernel.inc"
}  // namespace dfd

template <typename F> static void run_grid(int grid, int threads, F body) {
    blockDim.x = threads;
    for (int b = 0; b < grid; ++b) {
        std::barrier<> bar(threads); g_cta_bar = &bar;
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t) th.emplace_back([&, t, b]() { threadIdx.x = t; blockIdx.x = b; body(); });
        for (auto& t : th) t.join();
    }
}

int main() {
    using namespace dfd;
    int rc = 0;
    const int frames = 2, H = 16, W = 32, HW = H * W;
    std::vector<uint8_t> u8((size_t)frames * HW * 3);
    uint32_t seed = 31;
    for (size_t i = 0; i < u8.size(); ++i) { seed = seed * 1664525u + 1013904223u; u8[i] = i < 768 ? (uint8_t)(i & 255) : (uint8_t)(seed >> 24); }   // every byte value first
    {   // ---- K1: bit-exact
        std::vector<uint16_t> out((size_t)frames * 3 * HW, 0xdead);
        const int gpf = HW / 16; const int64_t groups = (int64_t)frames * gpf;
        run_grid((int)((groups + 255) / 256), 256, [&]() { preprocess_kernel<__half>(u8.data(), reinterpret_cast<__half*>(out.data()), groups, gpf, HW); });
        size_t bad = 0;
        for (int f = 0; f < frames; ++f) for (int c = 0; c < 3; ++c) for (int p = 0; p < HW; ++p) {
            const float mean = imagenet_mean(c), sd = imagenet_std(c);
            volatile float x = (float)u8[((size_t)f * HW + p) * 3 + c] / 255.0f; volatile float d = x - mean; const float r = d / sd;
            _Float16 h = (_Float16)r; uint16_t e; memcpy(&e, &h, 2);
            bad += e != out[((size_t)f * 3 + c) * HW + p];
        }
        printf("preprocess_kernel (uint8 HWC -> fp16 NCHW): %zu of %zu values differ from the reference arithmetic -> %s\n", bad, out.size(), bad ? "MISMATCH" : "ok");
        rc |= bad != 0;
    }
    {   // ---- stem, fp32 NCHW input (what forward() receives) and uint8 input with the prep fused
        const int OH = H / 2, OW = W / 2; const int64_t total = (int64_t)frames * OH * OW;
        std::vector<float> x32((size_t)frames * 3 * HW), w(27 * 32), b(32);
        for (int f = 0; f < frames; ++f) for (int c = 0; c < 3; ++c) for (int p = 0; p < HW; ++p) x32[((size_t)f * 3 + c) * HW + p] = prep_value(c, u8[((size_t)f * HW + p) * 3 + c]);
        auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
        for (auto& v : w) v = 0.3f * rnd();
        for (auto& v : b) v = 0.2f * rnd();
        std::vector<double> ref((size_t)total * 32);
        for (int f = 0; f < frames; ++f) for (int oy = 0; oy < OH; ++oy) for (int ox = 0; ox < OW; ++ox) for (int o = 0; o < 32; ++o) {
            double a = b[o];
            for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) for (int c = 0; c < 3; ++c) {
                const int iy = 2 * oy - 1 + ky, ix = 2 * ox - 1 + kx;
                if (iy >= 0 && iy < H && ix >= 0 && ix < W) a += (double)w[((ky * 3 + kx) * 3 + c) * 32 + o] * x32[((size_t)f * 3 + c) * HW + iy * W + ix];
            }
            ref[(((size_t)f * OH + oy) * OW + ox) * 32 + o] = a / (1 + exp(-a));
        }
        for (int kind = 0; kind < 2; ++kind) {
            std::vector<uint16_t> out((size_t)total * 32, 0xdead);
            if (kind == 0) run_grid((int)((total + 127) / 128), 128, [&]() { stem_kernel<__half, 0>(u8.data(), w.data(), b.data(), reinterpret_cast<__half*>(out.data()), H, W, OH, OW, total); });
            else run_grid((int)((total + 127) / 128), 128, [&]() { stem_kernel<__half, 1>(x32.data(), w.data(), b.data(), reinterpret_cast<__half*>(out.data()), H, W, OH, OW, total); });
            double max_err = 0, scale = 0;
            for (size_t i = 0; i < ref.size(); ++i) { _Float16 h; memcpy(&h, &out[i], 2); max_err = fmax(max_err, fabs((double)(float)h - ref[i])); scale = fmax(scale, fabs(ref[i])); }
            const bool ok = max_err <= 1.2e-3 * fmax(1.0, scale);
            printf("stem_kernel (%s input): max |err| %.2e (scale %.2f) -> %s\n", kind == 0 ? "uint8 HWC, prep fused" : "fp32 NCHW", max_err, scale, ok ? "ok" : "MISMATCH");
            rc |= !ok;
        }
    }
    return rc;
}
