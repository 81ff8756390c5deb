// K2 — MBConv depthwise conv kxk (k in {3,5}, stride in {1,2}, pad k/2) + folded BN + SiLU,
// NHWC 16-bit, fused with the squeeze-excite spatial sums (timm `conv_dw` + `bn1` + SqueezeExcite's
// `x.mean((2,3))`; reference call site pretrained_detector.py:116).
//
// One CTA owns (frame, row segment, block of CB channels) over the full width of the map and MARCHES DOWN the
// rows.  Input rows are staged through a ring of shared-memory row buffers by `cp.async` (16-byte LDGSTS, L1
// bypass) issued kRing-1 rows ahead of the row being consumed: the memory pipeline is decoupled from the
// register file, every input byte is fetched from global memory exactly once per CTA (no halo re-reads, no
// L1 traffic), and tens of KB per SM are in flight however few warps are resident.
//
// One thread owns 2 channels (one packed fp32x2 pair) x 7 consecutive output columns.  Each input row is read
// once from shared memory (NCOL = 6*stride + k 32-bit loads, conflict-free: the lanes of a warp read
// consecutive channel pairs), converted once and scattered into a small ring of accumulator rows held in
// registers (k rows for stride 1, ceil(k/2) for stride 2).  The k*k fp32 weight pairs of the thread's channels
// stay in registers for the whole march, so the inner loop is `fma.rn.f32x2` on registers only: 12.5 (5x5) /
// 4.5 (3x3) FMA-pipe instructions per output.
//
// The row loop is unrolled over one period of the accumulator ring (PERIOD = stride * RING input rows) so that
// every ring slot index is a compile-time constant.  For every output the taps are added in (ky, kx) order
// starting from the bias.
//
// SE squeeze: a thread sums its SiLU outputs (fp32, before the 16-bit rounding) over its segment and strip and
// writes its own 2-channel slice of partial row (segment, strip): no atomics, fixed order; se.cu adds the
// partial rows up.  Segmentation and strip width depend only on the layer shape, never on the batch.
#include "common.cuh"
#include "kernels.h"
#include <cstdlib>
#include <type_traits>

namespace dfd {

constexpr int kMarchTW = 7;          // output columns per thread (odd: strips interleave over the smem banks)
constexpr int kMarchMaxThreads = 256;
constexpr int kMarchMaxK = 4;        // 16-byte chunks a thread copies per input row (upper limit)

// DFD_MARCH_KERNEL_BEGIN   (tools/host_emul/ runs this kernel on CPU threads; its one inline-asm shared-memory load is mapped
// to the harness's load by run.py, everything else is the text below)
// CC / WC / CBC: channels, (square) map size and channel block as compile-time constants for the network's own
// layer shapes (every address offset becomes an immediate), 0 = run-time geometry.
template <typename T, int KS, int S, int NR, int MAXREG, bool FULL, int MAXK, int CC = 0, int WC = 0, int CBC = 0>
__global__ void __launch_bounds__(kMarchMaxThreads, 1) __maxnreg__(MAXREG)
dwconv_march_kernel(const T* __restrict__ in, const float* __restrict__ w, const float* __restrict__ bias,
                    T* __restrict__ out, float* __restrict__ partials,
                    int H_, int W_, int C_, int OH_, int OW_, int CB_, int strips_, int rps, int segs, int pixw_) {
    constexpr int TW = kMarchTW;
    constexpr int PAD = KS / 2;
    const int C = CC ? CC : C_, CB = CBC ? CBC : CB_;
    const int W = WC ? WC : W_, H = WC ? WC : H_;
    const int OW = WC ? (WC + 2 * PAD - KS) / S + 1 : OW_, OH = WC ? OW : OH_;
    const int strips = WC ? (OW + TW - 1) / TW : strips_;
    const int pixw = WC ? (((strips * TW - 1) * S + KS) > WC + 2 * PAD ? ((strips * TW - 1) * S + KS) : WC + 2 * PAD) : pixw_;
    constexpr int NCOL = (TW - 1) * S + KS;
    constexpr int RING = (KS + S - 1) / S;
    constexpr int PERIOD = S * RING;
    extern __shared__ __align__(16) uint8_t dw_smem[];

    const int ncb = C / CB;
    const int cb = blockIdx.x % ncb;                     // channel block fastest: neighbouring CTAs read neighbouring bytes
    const int fs = blockIdx.x / ncb;
    const int seg = fs % segs;
    const int64_t frame = fs / segs;
    const int CB2 = CB >> 1;
    const int cpl = threadIdx.x % CB2, strip = threadIdx.x / CB2;
    const int c0 = cb * CB + 2 * cpl;                    // first of this thread's two channels
    const int oy0 = seg * rps;
    const int nrows = min(rps, OH - oy0);
    const int ox0 = strip * TW;
    const int iy_start = oy0 * S - PAD;
    const int rend = S * (nrows - 1) + KS;               // input rows this CTA walks over (CTA-uniform)
    const uint32_t rsb = (uint32_t)pixw * CB * 2;        // bytes per staged row: [pixw pixels][CB channels]
    const uint32_t sm0 = smem_u32(dw_smem);

    // zero the staging ring once: left/right padding columns stay zero for the whole march
    for (uint32_t i = threadIdx.x * 16; i < NR * rsb; i += blockDim.x * 16) sts16(sm0 + i, make_uint4(0, 0, 0, 0));

    uint64_t wr[KS * KS];
    // weights and bias are halved on load (exact), so the accumulators hold h = x/2 and SiLU(x) = h + h*tanh(h)
    // needs no extra multiply; the bits are those of the unscaled sum
    const uint64_t half2 = f2_pack(0.5f, 0.5f);
#pragma unroll
    for (int i = 0; i < KS * KS; ++i) wr[i] = mul2(__ldg(reinterpret_cast<const unsigned long long*>(w + (size_t)i * C + c0)), half2);
    const uint64_t b2 = mul2(__ldg(reinterpret_cast<const unsigned long long*>(bias + c0)), half2);

    // this thread's 16-byte chunks of a row: chunk i -> pixel i / (CB/8), 8-channel group i % (CB/8)
    const int cpp = CB >> 3, chunks = W * cpp;
    uint32_t g_off[MAXK], s_off[MAXK];
#pragma unroll
    for (int k = 0; k < MAXK; ++k) {
        const int i = threadIdx.x + k * blockDim.x;
        const int px = i / cpp, sub = i - px * cpp;
        g_off[k] = (uint32_t)(px * C + sub * 8) * 2;                          // bytes inside a global row
        s_off[k] = i < chunks ? (uint32_t)((px + PAD) * CB + sub * 8) * 2 : 0xffffffffu;   // bytes inside a staged row
    }
    const size_t rowpitch_b = (size_t)W * C * 2;
    __syncthreads();
    griddep_wait();                                      // everything above touched only weights and shared memory

    // running state of the stager: next input row to issue, its global address and ring slot
    int iy_i = iy_start, left_i = rend;
    uint32_t sb_i = sm0;
    const char* gp_i = reinterpret_cast<const char*>(in + (size_t)frame * H * W * C + (size_t)cb * CB) + (ptrdiff_t)iy_start * (ptrdiff_t)rowpitch_b;
    auto issue_row = [&]() {
        if (left_i > 0 && (unsigned)iy_i < (unsigned)H) {
#pragma unroll
            for (int k = 0; k < MAXK; ++k)
                if (s_off[k] != 0xffffffffu) cp_async16(sb_i + s_off[k], gp_i + g_off[k], true);
        }
        cp_async_commit();
        ++iy_i; --left_i; gp_i += rowpitch_b;
        sb_i += rsb; if (sb_i == sm0 + NR * rsb) sb_i = sm0;
    };
#pragma unroll
    for (int r = 0; r < NR - 1; ++r) issue_row();

    uint64_t acc[RING][TW];
#pragma unroll
    for (int s = 0; s < RING; ++s)
#pragma unroll
        for (int j = 0; j < TW; ++j) acc[s][j] = 0ull;
    uint64_t sums = 0ull;
    // pointer to this thread's first output of the next output row: ONE 64-bit add per row, the 7 stores use constant offsets
    T* orow = out + (((size_t)frame * OH + oy0) * OW + ox0) * C + c0;
    const uint32_t ro_step = (uint32_t)OW * (uint32_t)C;
    const uint32_t pix_b = (uint32_t)CB * 2;
    uint32_t sb_c = sm0 + (uint32_t)(ox0 * S * CB + 2 * cpl) * 2;    // window column 0 of this thread in the current ring slot
    const uint32_t sb_c_end = sb_c + NR * rsb;
    int iy = iy_start;

    for (int rb = 0; rb < rend; rb += PERIOD) {
#pragma unroll
        for (int p = 0; p < PERIOD; ++p) {
            const int r = rb + p;
            if (r < rend) {
                cp_async_wait<NR - 2>();                 // this thread's copies of row r have landed ...
                __syncthreads();                         // ... and everybody's; row r-1 is consumed by all
                issue_row();                             // refills the slot of row r-1 with row r+NR-1
                if ((unsigned)iy < (unsigned)H) {
#pragma unroll
                    for (int jj = 0; jj < NCOL; ++jj) {
                        // a map one strip wide (7x7): the window's first and last PAD columns are zero padding for EVERY thread, so
                        // their loads and FMAs are dropped at compile time (17 % of the 5x5 taps); an output whose leftmost tap is
                        // padding starts from the bias alone
                        if (WC != 0 && WC <= TW && S == 1 && (jj < PAD || jj >= PAD + WC)) {
                            if (jj < TW && p % S == 0) acc[((p + 2 * PERIOD) / S) % RING][jj] = b2;
                            continue;
                        }
                        uint32_t raw;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(raw) : "r"(sb_c + jj * pix_b));
                        const float2 xf = Half16<T>::unpack(raw);
                        const uint64_t x = f2_pack(xf.x, xf.y);
#pragma unroll
                        for (int ky = 0; ky < KS; ++ky) {
                            const int dd = p - ky + 2 * PERIOD;             // compile-time after unrolling
                            if (dd % S != 0) continue;
                            const int slot = (dd / S) % RING;
#pragma unroll
                            for (int kx = 0; kx < KS; ++kx) {
                                const int dj = jj - kx;
                                if (dj < 0 || (dj % S) != 0 || dj / S >= TW) continue;
                                const int j = dj / S;
                                if (ky == 0 && kx == 0) acc[slot][j] = fma2(x, wr[0], b2);
                                else acc[slot][j] = fma2(x, wr[ky * KS + kx], acc[slot][j]);
                            }
                        }
                    }
                } else if (p % S == 0) {                                    // padding row: only the ky = 0 initialisation
                    const int slot = (p / S) % RING;
#pragma unroll
                    for (int j = 0; j < TW; ++j) acc[slot][j] = b2;
                }
                ++iy;
                sb_c += rsb; if (sb_c == sb_c_end) sb_c -= NR * rsb;
                if ((p - (KS - 1) + 2 * PERIOD) % S == 0) {                 // an output row completes after this input row
                    const int slot = ((p - (KS - 1) + 2 * PERIOD) / S) % RING;
                    if (r >= KS - 1) {
#pragma unroll
                        for (int j = 0; j < TW; ++j) {
                            if (FULL || ox0 + j < OW) {
                                const float2 a = f2_unpack(acc[slot][j]);
                                const float y0 = fmaf(a.x, tanh_approx(a.x), a.x), y1 = fmaf(a.y, tanh_approx(a.y), a.y);
                                sums = add2(sums, f2_pack(y0, y1));
                                *reinterpret_cast<uint32_t*>(orow + j * C) = Half16<T>::pack(y0, y1);
                            }
                        }
                        orow += ro_step;
                    }
                }
            }
        }
    }
    cp_async_wait<0>();
    float* dst = partials + (((size_t)frame * segs + seg) * strips + strip) * C + c0;
    *reinterpret_cast<float2*>(dst) = f2_unpack(sums);
}

// DFD_MARCH_KERNEL_END

// rows per segment: a function of the layer shape only (batch-invariant partial sums)
static inline int march_rps(int OH) { return OH > 56 ? 56 : OH; }

int dw_march_slots(int OH, int OW) {
    const int rps = march_rps(OH);
    return ((OH + rps - 1) / rps) * ((OW + kMarchTW - 1) / kMarchTW);
}

// (k, stride, max regs, C, W, CB) instantiations with compile-time geometry: the twelve depthwise shapes of the
// 224x224 network with the channel block tools/sweep_dw.py measured fastest (march_cb() holds the same choice)
#ifndef DFD_MARCH_REG_5S1
#define DFD_MARCH_REG_5S1 168        // register cap of the 5x5 stride-1 instantiations (tools/build_variant.py sweeps it)
#endif
#define DFD_MARCH_SPEC_LIST \
    DFD_MARCH_SPEC(3, 1, 128, 32, 112, 16) DFD_MARCH_SPEC(3, 2, 128, 96, 112, 48) DFD_MARCH_SPEC(3, 1, 128, 144, 56, 16) \
    DFD_MARCH_SPEC(5, 2, 168, 144, 56, 16) DFD_MARCH_SPEC(5, 1, DFD_MARCH_REG_5S1, 240, 28, 16) DFD_MARCH_SPEC(3, 2, 128, 240, 28, 48) \
    DFD_MARCH_SPEC(3, 1, 128, 480, 14, 32) DFD_MARCH_SPEC(5, 1, DFD_MARCH_REG_5S1, 480, 14, 32) DFD_MARCH_SPEC(5, 1, DFD_MARCH_REG_5S1, 672, 14, 32) \
    DFD_MARCH_SPEC(5, 2, 168, 672, 14, 224) DFD_MARCH_SPEC(5, 1, DFD_MARCH_REG_5S1, 1152, 7, 64) DFD_MARCH_SPEC(3, 1, 128, 1152, 7, 64)

static int g_march_cb_override = 0;      // tuning aid (dfd_k_set_dw_channel_block): 0 = the table in march_cb()
void dw_march_set_cb(int cb) { g_march_cb_override = cb; }

// channel block.  The network's own shapes use the block tools/sweep_dw.py measured fastest on B200 (small blocks:
// more resident CTAs per SM, i.e. more independent cp.async rings in flight); any other shape takes the largest
// divisor of C (multiple of 8) whose CTA (strips * CB/2 threads) fits the thread limit.
static int march_cb(int C, int k, int stride, int W, int strips, int max_threads) {
    static const struct { int C, k, s, W, cb; } tuned[] = {
        {32, 3, 1, 112, 16},  {96, 3, 2, 112, 48},  {144, 3, 1, 56, 16},  {144, 5, 2, 56, 16},
        {240, 5, 1, 28, 16},  {240, 3, 2, 28, 48},  {480, 3, 1, 14, 32},  {480, 5, 1, 14, 32},
        {672, 5, 1, 14, 32},  {1152, 5, 1, 7, 64},  {1152, 3, 1, 7, 64},
    };
    for (const auto& t : tuned)
        if (t.C == C && t.k == k && t.s == stride && t.W == W && strips * (t.cb / 2) <= kMarchMaxThreads) return t.cb;
    int best = 0;
    for (int cb = 8; cb <= C; cb += 8)
        if (C % cb == 0 && strips * (cb / 2) <= max_threads) best = cb;
    return best;
}

bool dw_march_supported(int H, int W, int C, int k, int stride) {
    if ((k != 3 && k != 5) || (stride != 1 && stride != 2) || (C & 7)) return false;
    const int OW = (W + 2 * (k / 2) - k) / stride + 1;
    const int strips = (OW + kMarchTW - 1) / kMarchTW;
    return strips * 4 <= kMarchMaxThreads;
}

template <typename T>
static cudaError_t launch_march_t(const void* in, const float* w, const float* bias, void* out, float* partials,
                                  int64_t frames, int H, int W, int C, int k, int stride, cudaStream_t s) {
    constexpr int max_threads = 128;
    const int env_cb = g_march_cb_override;                                              // sweeps only (dfd_k_set_dw_channel_block)
    constexpr int NR = 6;
    const int pad = k / 2;
    const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
    const int strips = (OW + kMarchTW - 1) / kMarchTW;
    const int rps = march_rps(OH), segs = (OH + rps - 1) / rps;
    if (frames <= 0) return cudaSuccess;
    if (!dw_march_supported(H, W, C, k, stride)) return cudaErrorInvalidValue;
    int mt = max_threads < strips * 4 ? strips * 4 : max_threads;
    if (mt > kMarchMaxThreads) mt = kMarchMaxThreads;
    int CB = march_cb(C, k, stride, W, strips, mt);
    if (env_cb > 0 && (env_cb & 7) == 0 && C % env_cb == 0 && strips * (env_cb / 2) <= kMarchMaxThreads) CB = env_cb;
    const int threads = strips * (CB / 2);
    const int need = (strips * kMarchTW - 1) * stride + k;          // columns the last strip's window reaches
    const int pixw = need > W + 2 * pad ? need : W + 2 * pad;
    const int chunks = W * (CB / 8);
    const size_t smem = (size_t)NR * pixw * CB * 2;
    const int64_t grid = frames * segs * (C / CB);
    if (chunks > kMarchMaxK * threads || smem > 200 * 1024 || grid > 0x7fffffffLL) return cudaErrorInvalidValue;
#define DFD_MARCH_GO(kern) { \
        if (smem > 48 * 1024) { cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); if (e != cudaSuccess) return e; } \
        return launch_pdl(kern, dim3((unsigned)grid), dim3(threads), smem, s, (const T*)in, w, bias, (T*)out, partials, H, W, C, OH, OW, CB, strips, rps, segs, pixw); }
    // the network's own layer shapes at 224x224, fp16: compile-time geometry
#define DFD_MARCH_SPEC(KS, ST, MR, CC_, WW_, CB_) \
    if constexpr (std::is_same<T, __half>::value) if (k == KS && stride == ST && C == CC_ && W == WW_ && H == WW_ && CB == CB_) { \
        constexpr int kOW = (WW_ + 2 * (KS / 2) - KS) / ST + 1; \
        static_assert(kOW % kMarchTW == 0, "specialised shapes have whole strips"); \
        constexpr int kMK = (WW_ * (CB_ / 8) <= 2 * (kOW / kMarchTW) * (CB_ / 2)) ? 2 : 4; \
        auto kern = dwconv_march_kernel<__half, KS, ST, NR, MR, true, kMK, CC_, WW_, CB_>; \
        DFD_MARCH_GO(kern) }
    DFD_MARCH_SPEC_LIST
#undef DFD_MARCH_SPEC
#define DFD_MARCH(KS, ST, MR) if (k == KS && stride == ST) { \
        auto kern = dwconv_march_kernel<T, KS, ST, NR, MR, false, 4>; \
        if ((OW % kMarchTW) == 0) kern = chunks <= 2 * threads ? dwconv_march_kernel<T, KS, ST, NR, MR, true, 2> : dwconv_march_kernel<T, KS, ST, NR, MR, true, 4>; \
        DFD_MARCH_GO(kern) }
    DFD_MARCH(3, 1, 128) DFD_MARCH(5, 1, 168) DFD_MARCH(3, 2, 128) DFD_MARCH(5, 2, 168)
#undef DFD_MARCH
#undef DFD_MARCH_GO
    return cudaErrorInvalidValue;
}

int dw_num_partials(int OH, int OW, int C, int k, int stride) { (void)C; (void)k; (void)stride; return dw_march_slots(OH, OW); }

cudaError_t launch_dwconv(const void* in, const float* w, const float* bias, void* out, float* partials,
                          int64_t frames, int H, int W, int C, int k, int stride, int dtype, cudaStream_t s) {
    if (dtype == kDtypeFP16) return launch_march_t<__half>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
    return launch_march_t<__nv_bfloat16>(in, w, bias, out, partials, frames, H, W, C, k, stride, s);
}

}  // namespace dfd
