// K2 — MBConv depthwise conv kxk (k in {3,5}, stride in {1,2}, pad k/2) + folded BN + SiLU, NHWC 16-bit,
// fused with the squeeze-excite spatial reduction (timm `conv_dw` + `bn` + the `x.mean((2,3))` of
// SqueezeExcite; reference call site pretrained_detector.py:116).
//
// One thread owns ONE CHANNEL PAIR (a 32-bit half2/bf162) x a TH x TW tile of output pixels:
//   * the k*k x 2 filter taps of its pair stay in registers as packed fp32x2 for the whole tile;
//   * input rows are streamed once per thread ((TH-1)*s+k rows of (TW-1)*s+k values) and every loaded
//     value feeds up to k x min(TH,k) packed FMAs (`fma.rn.f32x2`, the only way to reach the full fp32
//     rate on sm_100) — L1 traffic per output drops ~5x against a one-row-per-thread mapping;
//   * lanes walk channel pairs, so every warp load/store is one contiguous 128-byte NHWC run.
// fp32 accumulation, fp32 weights, one rounding to the storage type at the store.
//
// SE squeeze: every thread sums its SiLU outputs (fp32, before the 16-bit rounding); a CTA covers PG channel
// pairs x TG neighbouring tiles and adds its TG tiles up in a fixed order, writing its channel slice of one
// partial row per tile group — no atomics, bit-reproducible.  se.cu adds the partial rows up in order.
#include "common.cuh"
#include "kernels.h"

namespace dfd {

constexpr int kDwThreads = 256;
constexpr int kDwTW = 4;

struct DwPlan { int TH, tiles_x, tiles, PG, TG, pair_groups, tile_groups; };

static DwPlan dw_plan(int OH, int OW, int C, int k) {
    DwPlan p;
    const int pairs = C / 2;
    p.TH = (k == 3 && OH <= 14 && OH % 7 == 0) ? 7 : 4;
    p.tiles_x = (OW + kDwTW - 1) / kDwTW;
    p.tiles = p.tiles_x * ((OH + p.TH - 1) / p.TH);
    p.pair_groups = (pairs + kDwThreads - 1) / kDwThreads;
    while (pairs % p.pair_groups) ++p.pair_groups;
    p.PG = pairs / p.pair_groups;
    p.TG = kDwThreads / p.PG;
    if (p.TG > p.tiles) p.TG = p.tiles;
    p.tile_groups = (p.tiles + p.TG - 1) / p.TG;
    return p;
}
int dw_num_partials(int OH, int OW, int C, int k) { return dw_plan(OH, OW, C, k).tile_groups; }

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t pack_f2(float2 v) { return *reinterpret_cast<uint64_t*>(&v); }
__device__ __forceinline__ float2 unpack_f2(uint64_t v) { return *reinterpret_cast<float2*>(&v); }

template <typename T, int KS, int S, int TH>
__global__ void __launch_bounds__(kDwThreads, 2)
dwconv_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
              T* __restrict__ out, float* __restrict__ partials,
              int H, int W, int C, int OH, int OW, const DwPlan pl) {
    constexpr int TW = kDwTW;
    constexpr int PAD = KS / 2;
    constexpr int NCOL = (TW - 1) * S + KS;
    constexpr int NROW = (TH - 1) * S + KS;
    __shared__ float2 s_part[kDwThreads];

    int bid = blockIdx.x;
    const int pg = bid % pl.pair_groups; bid /= pl.pair_groups;
    const int tg = bid % pl.tile_groups;
    const int64_t frame = bid / pl.tile_groups;
    const int pl_ = threadIdx.x % pl.PG, tl = threadIdx.x / pl.PG;
    const int pair = pg * pl.PG + pl_;
    const int tile = tg * pl.TG + tl;
    const bool active = (tl < pl.TG) && (tile < pl.tiles);

    float2 sum = make_float2(0.f, 0.f);
    if (active) {
        const int ty = tile / pl.tiles_x, tx = tile - ty * pl.tiles_x;
        const int oy0 = ty * TH, ox0 = tx * TW;
        const int iy0 = oy0 * S - PAD, ix0 = ox0 * S - PAD;

        uint64_t wreg[KS * KS];
#pragma unroll
        for (int i = 0; i < KS * KS; ++i) wreg[i] = pack_f2(__ldg(reinterpret_cast<const float2*>(w + (size_t)i * C + 2 * pair)));
        uint64_t acc[TH][TW];
        {
            const uint64_t b = pack_f2(__ldg(reinterpret_cast<const float2*>(bias + 2 * pair)));
#pragma unroll
            for (int t = 0; t < TH; ++t)
#pragma unroll
                for (int j = 0; j < TW; ++j) acc[t][j] = b;
        }
        const T* base = in + (size_t)frame * H * W * C + 2 * pair;
#pragma unroll
        for (int r = 0; r < NROW; ++r) {
            const int iy = iy0 + r;
            if (iy < 0 || iy >= H) continue;                  // zero padding row: contributes nothing
            const T* row = base + (size_t)iy * W * C;
            uint64_t x[NCOL];
#pragma unroll
            for (int j = 0; j < NCOL; ++j) {
                const int ix = ix0 + j;
                uint32_t v = 0u;
                if (ix >= 0 && ix < W) v = __ldg(reinterpret_cast<const uint32_t*>(row + (size_t)ix * C));
                x[j] = pack_f2(Half16<T>::unpack(v));
            }
#pragma unroll
            for (int t = 0; t < TH; ++t) {
                const int ky = r - t * S;                      // compile-time after unrolling
                if (ky >= 0 && ky < KS) {
#pragma unroll
                    for (int kx = 0; kx < KS; ++kx)
#pragma unroll
                        for (int j = 0; j < TW; ++j) acc[t][j] = ffma2(x[j * S + kx], wreg[ky * KS + kx], acc[t][j]);
                }
            }
        }
        T* obase = out + (size_t)frame * OH * OW * C + 2 * pair;
#pragma unroll
        for (int t = 0; t < TH; ++t) {
            const int oy = oy0 + t;
            if (oy < OH) {
#pragma unroll
                for (int j = 0; j < TW; ++j) {
                    const int ox = ox0 + j;
                    if (ox < OW) {
                        const float2 a = unpack_f2(acc[t][j]);
                        const float y0 = silu_f(a.x), y1 = silu_f(a.y);
                        sum.x += y0; sum.y += y1;
                        *reinterpret_cast<uint32_t*>(obase + ((size_t)oy * OW + ox) * C) = Half16<T>::pack(y0, y1);
                    }
                }
            }
        }
    }

    // ---- deterministic CTA reduction of the SE sums over the CTA's tiles ----------------------------
    s_part[threadIdx.x] = sum;                         // inactive threads contribute zeros
    __syncthreads();
    if (threadIdx.x < pl.PG) {
        float2 tot = make_float2(0.f, 0.f);
        for (int t = 0; t < pl.TG; ++t) { const float2 v = s_part[t * pl.PG + threadIdx.x]; tot.x += v.x; tot.y += v.y; }
        *reinterpret_cast<float2*>(partials + ((size_t)frame * pl.tile_groups + tg) * C + 2 * pair) = tot;
    }
}

template <typename T>
static cudaError_t launch_dw_t(const void* in, const float* w, const float* bias, void* out, float* partials,
                               int64_t frames, int H, int W, int C, int k, int stride, cudaStream_t s) {
    const int pad = k / 2;
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    if (frames <= 0) return cudaSuccess;
    if ((C & 7) || (k != 3 && k != 5) || (stride != 1 && stride != 2)) return cudaErrorInvalidValue;
    const DwPlan pl = dw_plan(OH, OW, C, k);
    const int64_t blocks = frames * pl.tile_groups * pl.pair_groups;
    if (blocks > 0x7fffffffLL) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)blocks;
#define DFD_DW(KS, ST, TH) dwconv_kernel<T, KS, ST, TH><<<grid, kDwThreads, 0, s>>>((const T*)in, w, bias, (T*)out, partials, H, W, C, OH, OW, pl)
    if (k == 3 && stride == 1 && pl.TH == 4) DFD_DW(3, 1, 4);
    else if (k == 3 && stride == 1) DFD_DW(3, 1, 7);
    else if (k == 3 && stride == 2 && pl.TH == 4) DFD_DW(3, 2, 4);
    else if (k == 3 && stride == 2) DFD_DW(3, 2, 7);
    else if (k == 5 && stride == 1) DFD_DW(5, 1, 4);
    else DFD_DW(5, 2, 4);
#undef DFD_DW
    return cudaGetLastError();
}

cudaError_t launch_dwconv(const void* in, const float* w, const float* bias, void* out, float* partials,
                          int64_t frames, int H, int W, int C, int k, int stride, int dtype, cudaStream_t s) {
    if (dtype == kDtypeFP16) return launch_dw_t<__half>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
    return launch_dw_t<__nv_bfloat16>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
}

}  // namespace dfd
