// K2 — MBConv depthwise conv kxk (k in {3,5}, stride in {1,2}, pad k/2) + folded BN + SiLU, NHWC 16-bit,
// fused with the squeeze-excite spatial reduction (timm `conv_dw` + `bn` + the `x.mean((2,3))` of
// SqueezeExcite; reference call site pretrained_detector.py:116).
//
// One thread owns 4 channels (one 64-bit vector) x TW = 8 consecutive output columns (4 for stride 2) x 7-8
// consecutive output rows (processed one after the other, so the rows it re-reads are still in L1).
// The kernel is bound by the L1/shared-memory pipe (ncu: l1tex throughput > 60 %, every other pipe < 40 %),
// so the tile shape is chosen to minimise L1 bytes per output: wide strips amortise both the window halo
// and the per-tap weight loads.
// Each input row of the window is loaded once as (TW-1)*stride+k vectors, unpacked once to packed fp32x2
// and reused from registers for the TW outputs; vertical reuse is served by L1 (consecutive work items are
// channel groups first, then columns, then rows, so the CTAs that share rows run next to each other).
// All arithmetic is packed `fma.rn.f32x2` (the full-rate fp32 form on sm_100) with fp32 weights; address
// and bounds work is amortised over 8 channels.  A warp's loads are runs of contiguous 16-byte vectors.
//
// SE squeeze: every thread sums its SiLU outputs (fp32, before the 16-bit rounding) per channel over its
// rows and writes its own 8-channel slice of a partial row (one partial row per (row group, column strip)):
// no atomics, no CTA barrier (early warps never wait for late ones), bit-reproducible.  se.cu adds the
// partial rows up in a fixed order.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>

namespace dfd {

constexpr int kDwThreads = 256;
constexpr int kDwCH = 4;              // channels per thread

// Strip width (output columns per thread) and CTAs/SM per layer shape, from a sweep on B200 (tools/sweep_dw.sh):
// big maps want more resident warps (narrow strips, 3-4 CTAs/SM), the 14x14 / 7x7 maps want wide strips.
struct DwCfg { int tw, minb; };
static inline DwCfg dw_cfg(int C, int k, int stride, int W) {
    if (stride == 1) {
        if (W <= 14) return {8, 2};
        return {4, 3};
    }
    if (k == 5 && W <= 56) return {4, 3};
    return {2, 4};
}
static inline int dw_rpt(int OH) { return OH >= 56 ? 8 : 7; }          // output rows per thread
static inline int dw_slots(int OH, int OW, int tw) {                     // (row group, column strip) pairs per frame
    const int rpt = dw_rpt(OH);
    return ((OH + rpt - 1) / rpt) * ((OW + tw - 1) / tw);
}

// Kernel choice per layer shape: the row-marching kernel (dwconv_march.cu) or the window-per-row kernel below.
// A function of the shape only, so the partial-sum layout never depends on the batch.  DFD_DW_MARCH=0|1 overrides
// it for sweeps (tools/sweep_dw.sh).
static bool dw_use_march(int H, int W, int C, int k, int stride) {
    static const char* e_march = getenv("DFD_DW_MARCH");
    if (!dw_march_supported(H, W, C, k, stride)) return false;
    if (e_march) return atoi(e_march) != 0;
    return true;
}
int dw_num_partials(int OH, int OW, int C, int k, int stride) {
    if (dw_use_march(OH * stride, OW * stride, C, k, stride)) return dw_march_slots(OH, OW);
    return dw_slots(OH, OW, dw_cfg(C, k, stride, OW * stride).tw);
}

// x * sigmoid(x) with raw MUFU ex2 + rcp (no range fix-ups: e = inf -> rcp = 0 -> -0, which is the limit)
__device__ __forceinline__ float silu_fast(float x) {
#if DFD_SILU_TANH
    return silu_tanh(x);
#endif
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + ex2_approx(-1.4426950408889634f * x)));
    return x * r;
}

// One output row strip (TW columns x 8 channels) of the window.  XINT: every window column is inside the
// image, so the loads need no column predicate; CC/WC are compile-time for the twelve layer shapes of the
// 224x224 network (addresses become immediates) and 0 for the generic fallback.
template <typename T, int KS, int STRIDE, int TW, bool XINT, int CC>
__device__ __forceinline__ void dw_row(const T* __restrict__ in_f, const float* __restrict__ wc, int C, int WCs, int H,
                                       int oy, int ix0, const int (&coloff)[(TW - 1) * STRIDE + KS],
                                       const ulonglong2& b0, uint64_t (&acc)[TW][2]) {
    constexpr int PAD = KS / 2;
    constexpr int NCOL = (TW - 1) * STRIDE + KS;
    const int Cc = CC ? CC : C;
#pragma unroll
    for (int j = 0; j < TW; ++j) { acc[j][0] = b0.x; acc[j][1] = b0.y; }
#pragma unroll
    for (int ky = 0; ky < KS; ++ky) {
        const int iy = oy * STRIDE - PAD + ky;
        const bool row_ok = (unsigned)iy < (unsigned)H;          // padding rows read as zeros (branch-free)
        const T* row = in_f + (size_t)((row_ok ? iy : 0) * WCs) + (XINT ? ix0 * Cc : 0);
        uint64_t x[NCOL][2];
#pragma unroll
        for (int j = 0; j < NCOL; ++j) {
            uint2 v = make_uint2(0, 0);
            if (XINT) { if (row_ok) v = __ldg(reinterpret_cast<const uint2*>(row + j * Cc)); }
            else      { if (row_ok && coloff[j] >= 0) v = __ldg(reinterpret_cast<const uint2*>(row + coloff[j])); }
            const float2 f0 = Half16<T>::unpack(v.x), f1 = Half16<T>::unpack(v.y);
            x[j][0] = f2_pack(f0.x, f0.y); x[j][1] = f2_pack(f1.x, f1.y);
        }
#pragma unroll
        for (int kx = 0; kx < KS; ++kx) {
            const ulonglong2 w0 = __ldg(reinterpret_cast<const ulonglong2*>(wc + (ky * KS + kx) * Cc));
#pragma unroll
            for (int j = 0; j < TW; ++j) {
                acc[j][0] = fma2(x[j * STRIDE + kx][0], w0.x, acc[j][0]);
                acc[j][1] = fma2(x[j * STRIDE + kx][1], w0.y, acc[j][1]);
            }
        }
    }
}

template <typename T, int KS, int STRIDE, int CC, int WW, int TW, int MINB>
__global__ void __launch_bounds__(kDwThreads, MINB)
dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
              T* __restrict__ out, float* __restrict__ partials,
              int H_, int W_, int C_, int OH_, int OW_, int strips, int items, int blocks_per_frame, int rpt, int slots) {
    constexpr int PAD = KS / 2;
    constexpr int NCOL = (TW - 1) * STRIDE + KS;
    // compile-time geometry for the specialised instantiations (square maps), run-time otherwise
    const int C = CC ? CC : C_, W = WW ? WW : W_, H = WW ? WW : H_;
    const int OW = WW ? (WW + 2 * PAD - KS) / STRIDE + 1 : OW_, OH = WW ? OW : OH_;
    const int C8 = C >> 2;                  // channel groups of 4
    const int64_t frame = blockIdx.x / blocks_per_frame;
    const int blk = blockIdx.x - (int)(frame * blocks_per_frame);
    const int item = blk * kDwThreads + threadIdx.x;
    if (item >= items) return;

    uint64_t sums[2] = {0ull, 0ull};
    {
        const int c8 = item % C8;
        const int t = item / C8;
        const int strip = t % strips;
        const int oyb = t / strips;
        const int ox0 = strip * TW;
        const int ix0 = ox0 * STRIDE - PAD;
        const T* in_f = in + (size_t)frame * H * W * C + c8 * kDwCH;
        const float* wc = w + c8 * kDwCH;
        const int WCs = W * C;
        const bool x_interior = (ix0 >= 0) && (ix0 + NCOL <= W) && (ox0 + TW <= OW);
        int coloff[NCOL];                                   // element offset of each window column, -1 = padding
#pragma unroll
        for (int j = 0; j < NCOL; ++j) { const int ix = ix0 + j; coloff[j] = (ix >= 0 && ix < W) ? ix * C : -1; }
        const ulonglong2 b0 = __ldg(reinterpret_cast<const ulonglong2*>(bias + c8 * kDwCH));

        for (int rr = 0; rr < rpt; ++rr) {
            const int oy = oyb * rpt + rr;
            if (oy >= OH) break;
            uint64_t acc[TW][2];
            if (x_interior) dw_row<T, KS, STRIDE, TW, true, CC>(in_f, wc, C, WCs, H, oy, ix0, coloff, b0, acc);
            else            dw_row<T, KS, STRIDE, TW, false, CC>(in_f, wc, C, WCs, H, oy, ix0, coloff, b0, acc);
            T* orow = out + (((size_t)frame * OH + oy) * OW) * C + c8 * kDwCH;
#pragma unroll
            for (int j = 0; j < TW; ++j) {
                const int ox = ox0 + j;
                if (ox < OW) {
                    uint32_t o[2];
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const float2 a = f2_unpack(acc[j][c]);
                        const float y0 = silu_fast(a.x), y1 = silu_fast(a.y);
                        sums[c] = add2(sums[c], f2_pack(y0, y1));
                        o[c] = Half16<T>::pack(y0, y1);
                    }
                    *reinterpret_cast<uint2*>(orow + (size_t)ox * C) = make_uint2(o[0], o[1]);
                }
            }
        }

        // this thread is the only writer of its (frame, slot, channel group) slice
        float* dst = partials + ((size_t)frame * slots + t) * C + c8 * kDwCH;
        const float2 s0 = f2_unpack(sums[0]), s1 = f2_unpack(sums[1]);
        *reinterpret_cast<float4*>(dst) = make_float4(s0.x, s0.y, s1.x, s1.y);
    }
}

template <typename T>
static cudaError_t launch_dw_t(const void* in, const float* w, const float* bias, void* out, float* partials,
                               int64_t frames, int H, int W, int C, int k, int stride, cudaStream_t s) {
    const int pad = k / 2;
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    const DwCfg cfg = dw_cfg(C, k, stride, W);
    const int strips = (OW + cfg.tw - 1) / cfg.tw;
    const int rpt = dw_rpt(OH), slots = dw_slots(OH, OW, cfg.tw);
    const int items = slots * (C / kDwCH);
    const int bpf = (items + kDwThreads - 1) / kDwThreads;
    if (frames <= 0) return cudaSuccess;
    if ((C & 7) || frames * (int64_t)bpf > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)(frames * bpf);
#define DFD_DW(KS, ST, CC, WW, TW, MB) dwconv_kernel<T, KS, ST, CC, WW, TW, MB><<<grid, kDwThreads, 0, s>>>((const T*)in, w, bias, (T*)out, partials, H, W, C, OH, OW, strips, items, bpf, rpt, slots)
    // the twelve depthwise shapes of EfficientNet-B0 at 224x224 get compile-time geometry;
    // TW / MB must agree with dw_cfg()
#define DFD_DW_CASE(KS, ST, CC, WW, TW, MB) if (k == KS && stride == ST && C == CC && W == WW && H == WW && cfg.tw == TW) { DFD_DW(KS, ST, CC, WW, TW, MB); return cudaGetLastError(); }
    DFD_DW_CASE(3, 1, 32, 112, 4, 3) DFD_DW_CASE(3, 2, 96, 112, 2, 4) DFD_DW_CASE(3, 1, 144, 56, 4, 3) DFD_DW_CASE(5, 2, 144, 56, 4, 3)
    DFD_DW_CASE(5, 1, 240, 28, 4, 3) DFD_DW_CASE(3, 2, 240, 28, 2, 4) DFD_DW_CASE(3, 1, 480, 14, 8, 2) DFD_DW_CASE(5, 1, 480, 14, 8, 2)
    DFD_DW_CASE(5, 1, 672, 14, 8, 2) DFD_DW_CASE(5, 2, 672, 14, 4, 3) DFD_DW_CASE(5, 1, 1152, 7, 8, 2) DFD_DW_CASE(3, 1, 1152, 7, 8, 2)
    // generic geometry (other crop sizes)
    if (k == 3 && stride == 1 && cfg.tw == 4) DFD_DW(3, 1, 0, 0, 4, 3);
    else if (k == 3 && stride == 1) DFD_DW(3, 1, 0, 0, 8, 2);
    else if (k == 5 && stride == 1 && cfg.tw == 4) DFD_DW(5, 1, 0, 0, 4, 3);
    else if (k == 5 && stride == 1) DFD_DW(5, 1, 0, 0, 8, 2);
    else if (k == 3 && stride == 2) DFD_DW(3, 2, 0, 0, 2, 4);
    else if (k == 5 && stride == 2 && cfg.tw == 4) DFD_DW(5, 2, 0, 0, 4, 3);
    else if (k == 5 && stride == 2) DFD_DW(5, 2, 0, 0, 2, 4);
    else return cudaErrorInvalidValue;
#undef DFD_DW_CASE
#undef DFD_DW
    return cudaGetLastError();
}

cudaError_t launch_dwconv(const void* in, const float* w, const float* bias, void* out, float* partials,
                          int64_t frames, int H, int W, int C, int k, int stride, int dtype, cudaStream_t s) {
    if (k != 3 && k != 5) return cudaErrorInvalidValue;
    if (dw_use_march(H, W, C, k, stride)) return launch_dwconv_march(in, w, bias, out, partials, frames, H, W, C, k, stride, dtype, s);
    if (dtype == kDtypeFP16) return launch_dw_t<__half>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
    return launch_dw_t<__nv_bfloat16>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
}

}  // namespace dfd
