// CPU emulation of vit_attention_v2_kernel (deepfake_video_detection_b200/csrc/vit.cu, DFD_VIT_ATTN_V2): the kernel text
// between the DFD_ATT2_KERNEL markers is compiled UNCHANGED; one std::thread per CUDA thread, std::barrier for
// __syncthreads, ldmatrix / mma.sync / shfl by their PTX definitions through a per-warp exchange buffer.  Shared memory is
// pre-filled with NaN patterns.  Compared with softmax(Q K^T / 8) V in fp32 on the same fp16 inputs.
// Build + run: python tools/host_emul/run.py attention
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
#ifndef INFINITY
#define INFINITY __builtin_inff()
#endif

struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
struct float2 { float x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return {a, b, c, d}; }
using std::min;
typedef _Float16 __half;
static std::barrier<>* g_cta_bar = nullptr;
static void __syncthreads() { g_cta_bar->arrive_and_wait(); }

namespace dfd {
alignas(16) uint8_t att_smem[128 * 1024];
enum : int { kDtypeBF16 = 0, kDtypeFP16 = 1 };
template <typename T> struct Half16;
template <> struct Half16<__half> {
    static constexpr int kCode = kDtypeFP16;
    static float2 unpack(uint32_t v) { _Float16 h[2]; memcpy(h, &v, 4); return {(float)h[0], (float)h[1]}; }
    static uint32_t pack(float a, float b) { _Float16 h[2] = {(_Float16)a, (_Float16)b}; uint32_t v; memcpy(&v, h, 4); return v; }
};
static inline float ex2_approx(float x) { return exp2f(x); }
static inline uint32_t smem_u32(const void* p) { return (uint32_t)((const uint8_t*)p - att_smem); }

struct WarpX { uint32_t a[32][4], b[32][2]; float c[32][4]; uint32_t addr[32]; float f[32]; std::barrier<> bar{32}; };
static std::vector<std::unique_ptr<WarpX>> g_warps;
static inline WarpX& wx() { return *g_warps[threadIdx.x >> 5]; }

static inline float __shfl_xor_sync(unsigned, float v, int o) {
    WarpX& w = wx(); const int lane = threadIdx.x & 31;
    w.f[lane] = v; w.bar.arrive_and_wait();
    const float r = w.f[lane ^ o]; w.bar.arrive_and_wait();
    return r;
}
// ldmatrix.m8n8.x4[.trans].b16: lane l supplies the address of row l % 8 of matrix l / 8; lane i receives, per matrix, the
// 32-bit word (row i/4, columns 2(i%4), 2(i%4)+1), or with .trans the elements (row 2(i%4), col i/4) and (row 2(i%4)+1, col i/4)
static inline void ldsm_impl(uint32_t (&r)[4], uint32_t addr, bool trans) {
    WarpX& w = wx(); const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    w.addr[lane] = addr; w.bar.arrive_and_wait();
    for (int m = 0; m < 4; ++m) {
        uint16_t e0, e1;
        if (!trans) { memcpy(&e0, att_smem + w.addr[m * 8 + g] + (2 * t) * 2, 2); memcpy(&e1, att_smem + w.addr[m * 8 + g] + (2 * t + 1) * 2, 2); }
        else { memcpy(&e0, att_smem + w.addr[m * 8 + 2 * t] + g * 2, 2); memcpy(&e1, att_smem + w.addr[m * 8 + 2 * t + 1] + g * 2, 2); }
        r[m] = (uint32_t)e0 | ((uint32_t)e1 << 16);
    }
    w.bar.arrive_and_wait();
}
template <typename T> static inline void ldsm_x4(uint32_t (&r)[4], uint32_t addr) { ldsm_impl(r, addr, false); }
template <typename T> static inline void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) { ldsm_impl(r, addr, true); }
template <typename T> static inline void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    WarpX& w = wx();
    for (int i = 0; i < 4; ++i) { w.a[lane][i] = a[i]; w.c[lane][i] = c[i]; }
    w.b[lane][0] = b0; w.b[lane][1] = b1;
    w.bar.arrive_and_wait();
    for (int half = 0; half < 2; ++half)
        for (int j = 0; j < 2; ++j) {
            const int col = 2 * t + j;
            float acc = w.c[lane][half * 2 + j];
            for (int tp = 0; tp < 4; ++tp) {
                const float2 alo = Half16<T>::unpack(w.a[g * 4 + tp][half]), ahi = Half16<T>::unpack(w.a[g * 4 + tp][2 + half]);
                const float2 blo = Half16<T>::unpack(w.b[col * 4 + tp][0]), bhi = Half16<T>::unpack(w.b[col * 4 + tp][1]);
                acc += alo.x * blo.x + alo.y * blo.y + ahi.x * bhi.x + ahi.y * bhi.y;
            }
            c[half * 2 + j] = acc;
        }
    w.bar.arrive_and_wait();
}
constexpr int kDim = 768, kHeads = 12, kHd = 64, kTokens = 197;
constexpr int kTokPad = 208, kQKStride = 72, kVtStride = 216;
static inline float __expf(float x) { return expf(x); }
#include "vit_attention_v2_kernel.inc"
}  // namespace dfd

template <bool V2> static int run();
int main() { return run<true>(); }

template <bool V2> static int run() {
    using namespace dfd;
    const int images = 2;
    std::vector<_Float16> qkv((size_t)images * kTokens * 3 * kDim), o((size_t)images * kTokens * kDim);
    uint32_t seed = 99;
    auto rnd = [&]() { seed = seed * 1664525u + 1013904223u; return ((seed >> 8) & 0xffff) / 32768.0f - 1.0f; };
    for (auto& v : qkv) v = (_Float16)(1.5f * rnd());
    const int heads_to_run[] = {0, 5, 11 + kHeads};               // (image 0, head 0), (image 0, head 5), (image 1, head 11)
    double max_err = 0, scale = 0;
    for (int b : heads_to_run) {
        memset(att_smem, 0xff, sizeof(att_smem));
        const int threads = V2 ? kAtt2Warps * 32 : 128;
        std::barrier<> bar(threads); g_cta_bar = &bar;
        g_warps.clear();
        for (int i = 0; i < threads / 32; ++i) g_warps.emplace_back(new WarpX());
        std::vector<std::thread> th;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t, b]() { threadIdx.x = t; blockIdx.x = b;
                if (V2) vit_attention_v2_kernel<__half>(qkv.data(), o.data()); else abort(); });
        for (auto& t : th) t.join();
        const int head = b % kHeads, img = b / kHeads;
        for (int q = 0; q < kTokens; ++q) {
            std::vector<float> p(kTokens); float m = -1e30f, l = 0;
            for (int k = 0; k < kTokens; ++k) {
                float s = 0;
                for (int d = 0; d < kHd; ++d) s += (float)qkv[((size_t)img * kTokens + q) * 3 * kDim + head * kHd + d] * (float)qkv[((size_t)img * kTokens + k) * 3 * kDim + kDim + head * kHd + d];
                p[k] = s * 0.125f; m = fmaxf(m, p[k]);
            }
            for (int k = 0; k < kTokens; ++k) { p[k] = expf(p[k] - m); l += p[k]; }
            for (int d = 0; d < kHd; ++d) {
                float a = 0;
                for (int k = 0; k < kTokens; ++k) a += p[k] * (float)qkv[((size_t)img * kTokens + k) * 3 * kDim + 2 * kDim + head * kHd + d];
                a /= l;
                max_err = fmax(max_err, fabs(a - (float)o[((size_t)img * kTokens + q) * kDim + head * kHd + d]));
                scale = fmax(scale, fabs(a));
            }
        }
    }
    const bool ok = max_err < 4e-3 && std::isfinite(max_err);
    printf("%s: 3 (image, head) CTAs x %d threads, max |err| %.2e (scale %.2f) -> %s\n", V2 ? "vit_attention_v2_kernel" : "vit_attention_kernel (GPU-verified)",
           V2 ? kAtt2Warps * 32 : 128, max_err, scale, ok ? "ok" : "MISMATCH");
    return ok ? 0 : 1;
}
