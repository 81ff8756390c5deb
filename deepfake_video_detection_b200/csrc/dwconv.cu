// K2 — MBConv depthwise conv kxk (k in {3,5}, stride in {1,2}, pad k/2) + folded BN + SiLU, NHWC 16-bit,
// fused with the squeeze-excite spatial reduction (timm `conv_dw` + `bn` + the `x.mean((2,3))` of
// SqueezeExcite; reference call site pretrained_detector.py:116).
//
// One thread owns 8 channels (one 128-bit vector) x TW consecutive output columns of one output row:
// each input row of the window is loaded once as (TW-1)*stride+k vectors and reused from registers for the
// TW outputs.  A CTA owns a compact 3-D tile (TR output rows x TS column strips x TC channel groups), so
// the vertical and horizontal halo re-reads of neighbouring threads hit L1 instead of going back to L2
// (a k x k window would otherwise pull every input row k times through L2).  Channel groups are the
// fastest thread index, so a warp's loads are runs of contiguous 16-byte vectors (coalesced NHWC).
// fp32 accumulation, fp32 weights.
//
// SE squeeze: every thread sums its SiLU outputs (fp32, before the 16-bit rounding) per channel; the
// CTA combines them in a fixed order and writes its channel slice of one partial row per spatial tile —
// no atomics, so the result is bit-reproducible.  se.cu adds the partial rows up in order.
#include "common.cuh"
#include "kernels.h"

namespace dfd {

constexpr int kDwThreads = 256;
constexpr int kDwTW = 4;

struct DwPlan { int strips, TS, TR, TC, tiles_x, tiles_y, groups_c; };

static DwPlan dw_plan(int OH, int OW, int C) {
    DwPlan p;
    const int C8 = C / 8;
    p.strips = (OW + kDwTW - 1) / kDwTW;
    const int nx = (p.strips + 3) / 4;               // <= 4 strips (16 output columns) per tile, balanced
    p.TS = (p.strips + nx - 1) / nx;
    const int ny = (OH + 7) / 8;                      // <= 8 output rows per tile, balanced
    p.TR = (OH + ny - 1) / ny;
    p.TC = C8 < kDwThreads / (p.TS * p.TR) ? C8 : kDwThreads / (p.TS * p.TR);
    while (p.TC * p.TS * (p.TR * 2) <= kDwThreads && p.TR * 2 <= OH) p.TR *= 2;   // few channels: taller tiles
    p.tiles_x = (p.strips + p.TS - 1) / p.TS;
    p.tiles_y = (OH + p.TR - 1) / p.TR;
    p.groups_c = (C8 + p.TC - 1) / p.TC;
    return p;
}
int dw_num_partials(int OH, int OW, int C) { const DwPlan p = dw_plan(OH, OW, C); return p.tiles_x * p.tiles_y; }

template <typename T, int KS, int STRIDE>
__global__ void __launch_bounds__(kDwThreads, 2)
dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
              T* __restrict__ out, float* __restrict__ partials,
              int H, int W, int C, int OH, int OW, const DwPlan pl) {
    constexpr int TW = kDwTW;
    constexpr int PAD = KS / 2;
    constexpr int NCOL = (TW - 1) * STRIDE + KS;
    __shared__ float s_part[kDwThreads][9];     // +1 pad: conflict-free column walks

    const int C8 = C >> 3;
    // block -> (frame, tile_y, tile_x, channel group); channel group fastest so co-resident CTAs share halos in L2
    int bid = blockIdx.x;
    const int gc = bid % pl.groups_c; bid /= pl.groups_c;
    const int tx = bid % pl.tiles_x;  bid /= pl.tiles_x;
    const int ty = bid % pl.tiles_y;
    const int64_t frame = bid / pl.tiles_y;
    // thread -> (row, strip, channel) inside the tile; channel fastest
    const int cl = threadIdx.x % pl.TC;
    const int sl = (threadIdx.x / pl.TC) % pl.TS;
    const int rl = threadIdx.x / (pl.TC * pl.TS);
    const int c8 = gc * pl.TC + cl;
    const int strip = tx * pl.TS + sl;
    const int oy = ty * pl.TR + rl;
    const bool valid = (rl < pl.TR) && (c8 < C8) && (strip < pl.strips) && (oy < OH);

    float sums[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) sums[c] = 0.f;

    if (valid) {
        const int ox0 = strip * TW;
        const int iy0 = oy * STRIDE - PAD, ix0 = ox0 * STRIDE - PAD;
        const T* in_f = in + (size_t)frame * H * W * C + c8 * 8;
        const float* wc = w + c8 * 8;

        float acc[TW][8];
        {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
#pragma unroll
            for (int j = 0; j < TW; ++j) {
                acc[j][0] = b0.x; acc[j][1] = b0.y; acc[j][2] = b0.z; acc[j][3] = b0.w;
                acc[j][4] = b1.x; acc[j][5] = b1.y; acc[j][6] = b1.z; acc[j][7] = b1.w;
            }
        }
#pragma unroll
        for (int ky = 0; ky < KS; ++ky) {
            const int iy = iy0 + ky;
            if (iy < 0 || iy >= H) continue;
            const T* row = in_f + (size_t)iy * W * C;
            uint4 v[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const int ix = ix0 + j;
                v[j] = (ix >= 0 && ix < W) ? ldg16(row + (size_t)ix * C) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int kx = 0; kx < KS; ++kx) {
                const float4 w0 = __ldg(reinterpret_cast<const float4*>(wc + (ky * KS + kx) * C));
                const float4 w1 = __ldg(reinterpret_cast<const float4*>(wc + (ky * KS + kx) * C + 4));
#pragma unroll
                for (int j = 0; j < TW; ++j) {
                    const uint4 x = v[j * STRIDE + kx];
                    const float2 x0 = Half16<T>::unpack(x.x), x1 = Half16<T>::unpack(x.y);
                    const float2 x2 = Half16<T>::unpack(x.z), x3 = Half16<T>::unpack(x.w);
                    acc[j][0] = fmaf(x0.x, w0.x, acc[j][0]); acc[j][1] = fmaf(x0.y, w0.y, acc[j][1]);
                    acc[j][2] = fmaf(x1.x, w0.z, acc[j][2]); acc[j][3] = fmaf(x1.y, w0.w, acc[j][3]);
                    acc[j][4] = fmaf(x2.x, w1.x, acc[j][4]); acc[j][5] = fmaf(x2.y, w1.y, acc[j][5]);
                    acc[j][6] = fmaf(x3.x, w1.z, acc[j][6]); acc[j][7] = fmaf(x3.y, w1.w, acc[j][7]);
                }
            }
        }
        T* orow = out + (((size_t)frame * OH + oy) * OW) * C + c8 * 8;
#pragma unroll
        for (int j = 0; j < TW; ++j) {
            const int ox = ox0 + j;
            if (ox < OW) {
                float y[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) { y[c] = silu_f(acc[j][c]); sums[c] += y[c]; }
                uint4 o;
                o.x = Half16<T>::pack(y[0], y[1]); o.y = Half16<T>::pack(y[2], y[3]);
                o.z = Half16<T>::pack(y[4], y[5]); o.w = Half16<T>::pack(y[6], y[7]);
                stg16(orow + (size_t)ox * C, o);
            }
        }
    }

    // ---- deterministic block reduction of the SE sums --------------------------------------------
#pragma unroll
    for (int c = 0; c < 8; ++c) s_part[threadIdx.x][c] = sums[c];     // invalid threads contribute zeros
    __syncthreads();
    if (threadIdx.x < pl.TC * 8) {
        const int tc = threadIdx.x >> 3, ch = threadIdx.x & 7;
        if (gc * pl.TC + tc < C8) {
            float tot = 0.f;
            const int n = (kDwThreads / pl.TC) * pl.TC;
            for (int t = tc; t < n; t += pl.TC) tot += s_part[t][ch];   // fixed order over (row, strip)
            partials[((size_t)frame * (pl.tiles_x * pl.tiles_y) + ty * pl.tiles_x + tx) * C + (gc * pl.TC + tc) * 8 + ch] = tot;
        }
    }
}

template <typename T>
static cudaError_t launch_dw_t(const void* in, const float* w, const float* bias, void* out, float* partials,
                               int64_t frames, int H, int W, int C, int k, int stride, cudaStream_t s) {
    const int pad = k / 2;
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    const DwPlan pl = dw_plan(OH, OW, C);
    if (frames <= 0) return cudaSuccess;
    const int64_t blocks = frames * pl.tiles_y * pl.tiles_x * pl.groups_c;
    if ((C & 7) || pl.TC * 8 > kDwThreads || blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)blocks;
#define DFD_DW(KS, ST) dwconv_kernel<T, KS, ST><<<grid, kDwThreads, 0, s>>>((const T*)in, w, bias, (T*)out, partials, H, W, C, OH, OW, pl)
    if (k == 3 && stride == 1) DFD_DW(3, 1);
    else if (k == 3 && stride == 2) DFD_DW(3, 2);
    else if (k == 5 && stride == 1) DFD_DW(5, 1);
    else if (k == 5 && stride == 2) DFD_DW(5, 2);
    else return cudaErrorInvalidValue;
#undef DFD_DW
    return cudaGetLastError();
}

cudaError_t launch_dwconv(const void* in, const float* w, const float* bias, void* out, float* partials,
                          int64_t frames, int H, int W, int C, int k, int stride, int dtype, cudaStream_t s) {
    if (dtype == kDtypeFP16) return launch_dw_t<__half>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
    return launch_dw_t<__nv_bfloat16>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
}

}  // namespace dfd
