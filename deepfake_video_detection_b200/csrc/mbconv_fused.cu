// K2' — MBConv expand 1x1 + BN + SiLU fused
// into the row-marching depthwise kernel (dwconv_march.cu) for the early, HBM-bound InvertedResidual blocks
// (timm `conv_pw` + `bn1` -> `conv_dw` + `bn2` + SqueezeExcite's mean; reference call site pretrained_detector.py:116).
//
// Why: on the 112x112 / 56x56 maps the expanded tensor (6x the block input) is written by the expand GEMM and read back
// by the depthwise kernel — 8.4 MB of the 28 MB a frame moves, at kernels that already run at the HBM roofline.  Here the
// expanded rows never leave the SM: the CTA that marches down the rows of (frame, channel block) computes each expanded
// row itself, one row ahead of the depthwise stencil, from the 6x smaller block input.
//
// Structure (one CTA = frame x block of CB expanded channels, whole map width, marching down the rows):
//   x ring     block-input rows [W][Cin] 16-bit staged by cp.async kXR-1 rows ahead (rows padded to 16-pixel tiles and to
//              Cin+8 halves per pixel: conflict-free fragment loads, zero K padding)
//   expand     all warps: mma.sync m16n8k16 tiles (16 pixels x 8 channels, K = Cin padded to 16/32) of row r+1 against the
//              CTA's [CB][Cin] weight slice in shared memory, + bias, SiLU (same tanh.approx form as the GEMM epilogue),
//              rounded to the storage type — the same rounding point as the HBM round trip it replaces — and written to
//              slot (r+1)&1 of a two-row expanded ring laid out exactly like the march kernel's staged rows
//   depthwise  unchanged march: one thread = 2 channels x 7 output columns, k*k weight pairs and the accumulator ring in
//              registers, packed fma.rn.f32x2, SiLU, SE partial sums, 16-bit stores
// One __syncthreads per input row, as before: expanded row r was written during step r-1; slot (r+1)&1 was last read in
// step r-1.  Partial-sum layout = dw_march_slots, so se.cu and the project GEMM are unchanged.
// tools/host_emul/ compiles the kernel below (everything between the DFD_FUSED_KERNEL markers) for the CPU with stand-ins for
// the device helpers, one std::thread per CUDA thread, to check its indexing and barrier placement without a GPU.
#ifndef DFD_HOST_EMUL
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>
#include <type_traits>
#endif

namespace dfd {

// tuning constants (the defaults are the measured choice; `tools/build_variant.py` builds alternatives for a sweep)
#ifndef DFD_FUSED_XR
#define DFD_FUSED_XR 6               // x-row ring depth
#endif
#ifndef DFD_FUSED_REG_A
#define DFD_FUSED_REG_A 128          // register cap, block 2.1.0 (16 -> 96 @112, k3 s2)
#endif
#ifndef DFD_FUSED_REG_B
#define DFD_FUSED_REG_B 128          // block 2.1.1 (24 -> 144 @56, k3 s1)
#endif
#ifndef DFD_FUSED_REG_C
#define DFD_FUSED_REG_C 168          // block 2.2.0 (24 -> 144 @56, k5 s2)
#endif
#ifndef DFD_FUSED_CARVEOUT
#define DFD_FUSED_CARVEOUT 0         // 1: ask for the maximum shared-memory carve-out (the driver's default picked 132 KB: 2 CTAs per SM)
#endif
#ifndef DFD_FUSED_BLOCKS
#define DFD_FUSED_BLOCKS 7           // bit i: fuse block 2.1.0 / 2.1.1 / 2.2.0
#endif

namespace {
constexpr int kFTW = 7;              // output columns per thread (as dwconv_march.cu)
constexpr int kXR = DFD_FUSED_XR;

#ifndef DFD_HOST_EMUL
template <typename T>
__device__ __forceinline__ void mma16816_f(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
    if constexpr (Half16<T>::kCode == kDtypeFP16)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
    uint32_t v; asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr)); return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) { asm volatile("st.shared.b32 [%0], %1;" :: "r"(addr), "r"(v) : "memory"); }
// four 8x8 16-bit matrices = the A fragment (a0..a3) of one m16n8k16 step in ONE shared-memory instruction
__device__ __forceinline__ void ldsm_x4_f(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
// 8 packed 16-bit values times 0.5 (exact: a power of two)
template <typename T> __device__ __forceinline__ uint4 halve8(uint4 v) {
    if constexpr (Half16<T>::kCode == kDtypeFP16) {
        __half2* p = reinterpret_cast<__half2*>(&v);
        const __half2 h = __floats2half2_rn(0.5f, 0.5f);
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = __hmul2(p[i], h);
    } else {
        __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&v);
        const __nv_bfloat162 h = __floats2bfloat162_rn(0.5f, 0.5f);
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = __hmul2(p[i], h);
    }
    return v;
}
#endif
}  // namespace

// DFD_FUSED_KERNEL_BEGIN
// Geometry is compile-time: CIN block input channels, C expanded channels, W x W map, CB channels per CTA.
// Launch geometry and shared-memory carve-up: ONE definition for the kernel and for its launcher.
template <int KS, int S, int CIN, int C, int W, int CB>
struct FusedGeom {
    static constexpr int PAD = KS / 2, OW = (W + 2 * PAD - KS) / S + 1, OH = OW, strips = OW / kFTW;
    // DWT threads own the depthwise work (2 channels x 7 columns each); the CTA is rounded up to whole warps, and the few
    // extra threads (they take part in staging, in the expand MMAs and in the barriers) shadow the first depthwise threads
    // with their stores suppressed
    static constexpr int DWT = strips * (CB / 2), THREADS = (DWT + 31) / 32 * 32, WARPS = THREADS / 32;
    static constexpr int pixw = ((strips * kFTW - 1) * S + KS) > W + 2 * PAD ? ((strips * kFTW - 1) * S + KS) : W + 2 * PAD;
    static constexpr int KP = (CIN + 15) & ~15;                     // K padded to whole mma k-steps
    static constexpr int XP = KP + 8;                               // halves per staged x pixel / weight row (conflict-free)
    static constexpr int PXT = (W + 15) / 16;                       // 16-pixel tiles per row
    // halves per pixel of the expanded ring: CB rounded up so that the pitch in 32-bit words is 4 (mod 8) — the 8 pixels x 4
    // channel pairs a warp stores per MMA tile then tile all 32 banks (CB = 48: 56 halves; the unpadded pitch is a 2-way conflict)
    static constexpr int EP = 2 * ((CB / 2 + 3) / 8 * 8 + 4) >= CB ? 2 * ((CB / 2 + 3) / 8 * 8 + 4) : CB + 8;
    static constexpr uint32_t rsb = (uint32_t)pixw * EP * 2;        // bytes per expanded row slot
    static constexpr uint32_t xsb = (uint32_t)PXT * 16 * XP * 2;    // x row slot: [PXT*16 pixels][XP]
    static constexpr uint32_t wsb = (uint32_t)CB * XP * 2;          // bytes of the weight slice
    static constexpr uint32_t bias_floats = (uint32_t)CB;
    static constexpr size_t smem_bytes = (size_t)2 * rsb + (size_t)kXR * xsb + wsb + (size_t)bias_floats * 4;
    static constexpr int rps = OH > 56 ? 56 : OH;                   // = march_rps(OH)
    static constexpr int segs = (OH + rps - 1) / rps;
    static constexpr int ctas_per_frame = segs * (C / CB);
    static_assert(OW % kFTW == 0, "whole strips only");
    static_assert(CB % 8 == 0 && C % CB == 0 && CIN % 8 == 0, "channel blocks");
    static_assert(THREADS - DWT < DWT, "shadow threads map onto real ones");
    static_assert(EP >= CB && EP % 2 == 0 && (EP / 2) % 8 == 4, "conflict-free pixel pitch of the expanded ring");
};

template <typename T, int KS, int S, int CIN, int C, int W, int CB, int MAXREG>
__global__ void __launch_bounds__((FusedGeom<KS, S, CIN, C, W, CB>::THREADS), 1) __maxnreg__(MAXREG)
mbconv_fused_kernel(const void* __restrict__ xv, const void* __restrict__ wev, const float* __restrict__ be,
                    const float* __restrict__ w, const float* __restrict__ bias,
                    T* __restrict__ out, float* __restrict__ partials) {
    using G = FusedGeom<KS, S, CIN, C, W, CB>;
    const T* x = reinterpret_cast<const T*>(xv);
    const T* we = reinterpret_cast<const T*>(wev);
    constexpr int TW = kFTW, PAD = G::PAD, H = W;
    constexpr int OW = G::OW, OH = G::OH, strips = G::strips;
    constexpr int DWT = G::DWT, THREADS = G::THREADS, WARPS = G::WARPS;
    constexpr int NCOL = (TW - 1) * S + KS;
    constexpr int RING = (KS + S - 1) / S, PERIOD = S * RING;
    constexpr int KP = G::KP, KSTEPS = KP / 16, XP = G::XP, PXT = G::PXT, EP = G::EP;
    constexpr uint32_t rsb = G::rsb, xsb = G::xsb, wsb = G::wsb;
    constexpr int NTL = CB / 8;                                   // 8-channel tiles of the channel block
    constexpr int rps = G::rps, segs = G::segs;
    constexpr int XCH = W * (CIN / 8);                            // 16-byte chunks of an x row
    constexpr int XK = (XCH + THREADS - 1) / THREADS;

    extern __shared__ __align__(16) uint8_t fz_smem[];
    const uint32_t sm_e = smem_u32(fz_smem);                      // [2][pixw][CB]      expanded ring
    const uint32_t sm_x = sm_e + 2 * rsb;                         // [kXR][PXT*16][XP]  x ring
    const uint32_t sm_w = sm_x + kXR * xsb;                       // [CB][XP]           expand weights of this channel block
    float* s_be = reinterpret_cast<float*>(fz_smem + 2 * rsb + kXR * xsb + wsb);   // [CB] expand bias

    constexpr int ncb = C / CB;
    const int cb = blockIdx.x % ncb;
    const int fs = blockIdx.x / ncb;
    const int seg = fs % segs;
    const int64_t frame = fs / segs;
    constexpr int CB2 = CB >> 1;
    const bool dw_active = (int)threadIdx.x < DWT;
    const int tdw = dw_active ? (int)threadIdx.x : (int)threadIdx.x - DWT;
    const int cpl = tdw % CB2, strip = tdw / CB2;
    const int c0 = cb * CB + 2 * cpl;
    const int oy0 = seg * rps;
    const int nrows = min(rps, OH - oy0);
    const int ox0 = strip * TW;
    const int iy_start = oy0 * S - PAD;
    const int rend = S * (nrows - 1) + KS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;

    // zero everything once: padding columns of the expanded slots, K padding / tile padding of the x slots and weights
    for (uint32_t i = threadIdx.x * 16; i < 2 * rsb + kXR * xsb + wsb; i += THREADS * 16) sts16(sm_e + i, make_uint4(0, 0, 0, 0));
    __syncthreads();
    // expand weights [CB][CIN] (BN folded, 16-bit, K-major) and bias of this channel block
    for (int i = threadIdx.x; i < CB * (CIN / 8); i += THREADS) {
        const int r = i / (CIN / 8), q = i - r * (CIN / 8);
        // weights and bias are staged HALVED: the MMA then accumulates h = (x W^T + b) / 2 directly, SiLU(x) = h + h tanh(h)
        sts16(sm_w + (uint32_t)(r * XP + q * 8) * 2, halve8<T>(ldg16(we + (size_t)(cb * CB + r) * CIN + q * 8)));
    }
    for (int i = threadIdx.x; i < CB; i += THREADS) s_be[i] = 0.5f * be[cb * CB + i];

    uint64_t wr[KS * KS];
    const uint64_t half2 = f2_pack(0.5f, 0.5f);
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wr[i] = mul2(__ldg(reinterpret_cast<const unsigned long long*>(w + (size_t)i * C + c0)), half2);
    const uint64_t b2 = mul2(__ldg(reinterpret_cast<const unsigned long long*>(bias + c0)), half2);

    // x-row stager: chunk i -> pixel i / (CIN/8), 8-channel group i % (CIN/8)
    uint32_t g_off[XK], s_off[XK];
#pragma unroll
    for (int k = 0; k < XK; ++k) {
        const int i = threadIdx.x + k * THREADS;
        const int px = i / (CIN / 8), sub = i - px * (CIN / 8);
        g_off[k] = (uint32_t)(px * CIN + sub * 8) * 2;
        s_off[k] = i < XCH ? (uint32_t)(px * XP + sub * 8) * 2 : 0xffffffffu;
    }
    constexpr size_t xpitch_b = (size_t)W * CIN * 2;              // bytes between consecutive x rows
    int iy_i = iy_start, left_i = rend;
    uint32_t xs_i = sm_x;
    const char* gp_i = reinterpret_cast<const char*>(x + (size_t)frame * H * W * CIN) + (ptrdiff_t)iy_start * (ptrdiff_t)xpitch_b;
    auto issue_row = [&]() {
        if (left_i > 0 && (unsigned)iy_i < (unsigned)H) {
#pragma unroll
            for (int k = 0; k < XK; ++k)
                if (s_off[k] != 0xffffffffu)
                    cp_async16(xs_i + s_off[k], gp_i + g_off[k], true);
        }
        cp_async_commit();
        ++iy_i; --left_i; gp_i += xpitch_b;
        xs_i += xsb; if (xs_i == sm_x + kXR * xsb) xs_i = sm_x;
    };

    // Expand one input row (index k of this CTA's walk, image row iy) from its x slot into expanded slot k & 1.
    // A warp owns NTL / WARPS channel tiles (always whole) and all pixel tiles of the row (compile-time unrolled), in PHASES:
    // every fragment load (ldmatrix) first, then every MMA, then the SiLU of all accumulators, then every store.  The
    // asm-volatile shared-memory accesses and MMAs keep their program order, so writing the phases out is what lets the
    // scheduler overlap the tiles' MMA -> MUFU -> store chains (tile-by-tile code spent a third of its issue slots on
    // fixed-latency waits: 1.54 -> 1.35 ms for block 2.1.0 at 2048 frames).  Holding the packed results back until after the
    // depthwise FMAs of the same step (to overlap the MUFU and FMA pipes) measured SLOWER (register pressure): 1.42 ms.
    static_assert(NTL % WARPS == 0, "whole channel tiles per warp");
    constexpr int NPW = NTL / WARPS;
    auto expand_row = [&](int k, int iy) {
        if (k >= rend || (unsigned)iy >= (unsigned)H) return;            // CTA-uniform
        const uint32_t xs = sm_x + (uint32_t)(k % kXR) * xsb;
        const uint32_t es = sm_e + (uint32_t)(k & 1) * rsb;
        // ldmatrix row of this lane: pixel (lane & 7) + 8 * ((lane >> 3) & 1) of the tile, K columns 8 * (lane >> 4) ...
        const uint32_t ar0 = xs + (uint32_t)(((lane & 7) + ((lane >> 3) & 1) * 8) * XP + (lane >> 4) * 8) * 2;
#pragma unroll
        for (int q = 0; q < NPW; ++q) {
            const int nt = warp * NPW + q;
            const uint32_t br = sm_w + (uint32_t)((nt * 8 + g) * XP + 2 * t) * 2;
            uint32_t bf[KSTEPS][2];
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) { bf[ks][0] = lds32(br + ks * 32); bf[ks][1] = lds32(br + ks * 32 + 16); }
            const float hbx = s_be[nt * 8 + 2 * t], hby = s_be[nt * 8 + 2 * t + 1];      // accumulators start from the halved bias
            const uint32_t ea0 = es + (uint32_t)((g + PAD) * EP + nt * 8 + 2 * t) * 2;
            uint32_t a[PXT][KSTEPS][4];
#pragma unroll
            for (int pt = 0; pt < PXT; ++pt)
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) ldsm_x4_f(a[pt][ks], ar0 + (uint32_t)(pt * 16 * XP) * 2 + ks * 32);
            float c[PXT][4];
#pragma unroll
            for (int pt = 0; pt < PXT; ++pt) {
                c[pt][0] = hbx; c[pt][1] = hby; c[pt][2] = hbx; c[pt][3] = hby;
#pragma unroll
                for (int ks = 0; ks < KSTEPS; ++ks) mma16816_f<T>(c[pt], a[pt][ks], bf[ks][0], bf[ks][1]);
            }
            uint32_t o[PXT][2];
#pragma unroll
            for (int pt = 0; pt < PXT; ++pt) {
                o[pt][0] = Half16<T>::pack(fmaf(c[pt][0], tanh_approx(c[pt][0]), c[pt][0]), fmaf(c[pt][1], tanh_approx(c[pt][1]), c[pt][1]));
                o[pt][1] = Half16<T>::pack(fmaf(c[pt][2], tanh_approx(c[pt][2]), c[pt][2]), fmaf(c[pt][3], tanh_approx(c[pt][3]), c[pt][3]));
            }
#pragma unroll
            for (int pt = 0; pt < PXT; ++pt) {
                const uint32_t ea = ea0 + (uint32_t)(pt * 16 * EP) * 2;
                if (pt * 16 + 8 <= W || pt * 16 + g < W) sts32(ea, o[pt][0]);
                if (pt * 16 + 16 <= W || pt * 16 + 8 + g < W) sts32(ea + 8 * EP * 2, o[pt][1]);
            }
        }
    };

    griddep_wait();                          // everything above touched only weights and shared memory
#pragma unroll
    for (int r = 0; r < kXR - 1; ++r) issue_row();
    cp_async_wait<kXR - 2>();                // x row 0 has landed (this thread's copies) ...
    __syncthreads();                         // ... and everybody's; weights and bias are visible too
    issue_row();                             // row kXR-1 into the free slot
    expand_row(0, iy_start);

    uint64_t acc[RING][TW];
#pragma unroll
    for (int s = 0; s < RING; ++s)
#pragma unroll
        for (int j = 0; j < TW; ++j) acc[s][j] = 0ull;
    uint64_t sums = 0ull;
    T* orow = out + (((size_t)frame * OH + oy0) * OW + ox0) * C + c0;     // one 64-bit add per output row, constant store offsets
    constexpr uint32_t ro_step = (uint32_t)OW * (uint32_t)C;
    constexpr uint32_t pix_b = (uint32_t)EP * 2;
    const uint32_t sb_c0 = sm_e + (uint32_t)(ox0 * S * EP + 2 * cpl) * 2;   // window column 0 of this thread in slot 0
    int iy = iy_start;

    for (int rb = 0; rb < rend; rb += PERIOD) {
#pragma unroll
        for (int p = 0; p < PERIOD; ++p) {
            const int r = rb + p;
            if (r < rend) {
                cp_async_wait<kXR - 2>();                // x row r+1 has landed (this thread's copies) ...
                __syncthreads();                         // ... and everybody's; expanded row r is complete; row r-1 consumed
                issue_row();                             // x row r+kXR into the slot of x row r (expanded in step r-1)
                expand_row(r + 1, iy + 1);
                const uint32_t sb_c = sb_c0 + (uint32_t)(r & 1) * rsb;
                if ((unsigned)iy < (unsigned)H) {
#pragma unroll
                    for (int jj = 0; jj < NCOL; ++jj) {
                        const uint32_t raw = lds32(sb_c + jj * pix_b);
                        const float2 xf = Half16<T>::unpack(raw);
                        const uint64_t xv = f2_pack(xf.x, xf.y);
#pragma unroll
                        for (int ky = 0; ky < KS; ++ky) {
                            const int dd = p - ky + 2 * PERIOD;
                            if (dd % S != 0) continue;
                            const int slot = (dd / S) % RING;
#pragma unroll
                            for (int kx = 0; kx < KS; ++kx) {
                                const int dj = jj - kx;
                                if (dj < 0 || (dj % S) != 0 || dj / S >= TW) continue;
                                const int j = dj / S;
                                if (ky == 0 && kx == 0) acc[slot][j] = fma2(xv, wr[0], b2);
                                else acc[slot][j] = fma2(xv, wr[ky * KS + kx], acc[slot][j]);
                            }
                        }
                    }
                } else if (p % S == 0) {
                    const int slot = (p / S) % RING;
#pragma unroll
                    for (int j = 0; j < TW; ++j) acc[slot][j] = b2;
                }
                ++iy;
                if ((p - (KS - 1) + 2 * PERIOD) % S == 0) {
                    const int slot = ((p - (KS - 1) + 2 * PERIOD) / S) % RING;
                    if (r >= KS - 1) {
#pragma unroll
                        for (int j = 0; j < TW; ++j) {
                            const float2 a = f2_unpack(acc[slot][j]);
                            const float y0 = fmaf(a.x, tanh_approx(a.x), a.x), y1 = fmaf(a.y, tanh_approx(a.y), a.y);
                            sums = add2(sums, f2_pack(y0, y1));
                            if (dw_active) *reinterpret_cast<uint32_t*>(orow + j * C) = Half16<T>::pack(y0, y1);
                        }
                        orow += ro_step;
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    float* dst = partials + (((size_t)frame * segs + seg) * strips + strip) * C + c0;
    if (dw_active) *reinterpret_cast<float2*>(dst) = f2_unpack(sums);
}

// DFD_FUSED_KERNEL_END

// InvertedResidual blocks of the 224x224 network this kernel runs: the early blocks, whose expand GEMM and
// depthwise kernel both sit at the HBM roofline when run separately — (Cin, mid, map, k, stride):
//     2.1.0: 16 -> 96 @112 k3 s2     2.1.1: 24 -> 144 @56 k3 s1     2.2.0: 24 -> 144 @56 k5 s2
// Measured on B200 at 2048 frames (profiles/r02_*): the step drops from 13.99 to 12.87 ms.  The later blocks (28x28 and
// smaller maps, the stem + block 0) were measured SLOWER fused (their kernels are issue-bound, not HBM-bound) and stay apart.
bool mbconv_fused_supported(int H, int W, int cin, int mid, int k, int stride) {
    if (H != W) return false;
    return ((DFD_FUSED_BLOCKS & 1) && cin == 16 && mid == 96 && W == 112 && k == 3 && stride == 2) ||
           ((DFD_FUSED_BLOCKS & 2) && cin == 24 && mid == 144 && W == 56 && k == 3 && stride == 1) ||
           ((DFD_FUSED_BLOCKS & 4) && cin == 24 && mid == 144 && W == 56 && k == 5 && stride == 2);
}

template <typename T, int KS, int S, int CIN, int C, int W, int CB, int MAXREG>
static cudaError_t fused_go(const void* x, const void* we, const float* be, const float* w, const float* bias, void* out,
                            float* partials, int64_t frames, cudaStream_t s) {
    using G = FusedGeom<KS, S, CIN, C, W, CB>;
    constexpr int THREADS = G::THREADS;
    constexpr size_t smem = G::smem_bytes;
    auto kern = mbconv_fused_kernel<T, KS, S, CIN, C, W, CB, MAXREG>;
    if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; }
#if DFD_FUSED_CARVEOUT
    { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared); if (e != cudaSuccess) return e; }
#endif
    const int64_t grid = frames * G::ctas_per_frame;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidValue;
    return launch_pdl(kern, dim3((unsigned)grid), dim3(THREADS), smem, s, x, we, be, w, bias, (T*)out, partials);
}

template <typename T>
static cudaError_t launch_fused_t(const void* x, const void* we, const float* be, const float* w, const float* bias, void* out,
                                  float* partials, int64_t frames, int W, int cin, int k, int stride, cudaStream_t s) {
    // 48 expanded channels per CTA (the wider blocks measured slower: 4.53 vs 3.57 ms for the three launches)
    if (cin == 16 && W == 112) return fused_go<T, 3, 2, 16, 96, 112, 48, DFD_FUSED_REG_A>(x, we, be, w, bias, out, partials, frames, s);
    if (cin == 24 && k == 3) return fused_go<T, 3, 1, 24, 144, 56, 48, DFD_FUSED_REG_B>(x, we, be, w, bias, out, partials, frames, s);
    return fused_go<T, 5, 2, 24, 144, 56, 48, DFD_FUSED_REG_C>(x, we, be, w, bias, out, partials, frames, s);
}

// x [frames][H][W][cin], we [mid][cin] + be [mid] (expand conv, BN folded), w [k*k][mid] fp32 + bias [mid] (depthwise, BN
// folded) -> out [frames][OH][OW][mid], partials [frames][dw_march_slots(OH,OW)][mid]
cudaError_t launch_mbconv_fused(const void* x, const void* we, const float* be, const float* w, const float* bias, void* out,
                                float* partials, int64_t frames, int H, int W, int cin, int mid, int k, int stride, int dtype,
                                cudaStream_t s) {
    if (frames <= 0) return cudaSuccess;
    if (!mbconv_fused_supported(H, W, cin, mid, k, stride)) return cudaErrorInvalidValue;
    if (dtype == kDtypeFP16) return launch_fused_t<__half>(x, we, be, w, bias, out, partials, frames, W, cin, k, stride, s);
    return launch_fused_t<__nv_bfloat16>(x, we, be, w, bias, out, partials, frames, W, cin, k, stride, s);
}

}  // namespace dfd
