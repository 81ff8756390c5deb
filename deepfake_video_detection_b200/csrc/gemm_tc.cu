// K3 — pointwise (1x1) convolutions of EfficientNet-B0 as tcgen05 / TMEM GEMMs (sm_100a).
//
//   D[M,N] = act( (A .* gate)[M,K] · W[N,K]^T + bias[N] ) (+ R[M,N])
//
//   A   NHWC activations viewed as [M = frames*H*W, K = Cin], 16-bit, K contiguous (K-major)
//   W   conv weight [N = Cout, K = Cin] with BatchNorm folded, 16-bit, K-major (timm conv_pw / conv_pwl /
//       conv_head; reference call site pretrained_detector.py:116)
//   gate  squeeze-excite gate fp32 [frames][K], applied to the A operand while it is staged (project convs)
//   R   residual input of the block (stride-1, Cin == Cout blocks)
//   pool variant: conv_head + BN + SiLU + global average pool -> features fp32 [frames][N]
//
// Structure (one persistent CTA per SM, warp-specialised; roles are template parameters):
//   TMA issuer   ONE elected lane of a converged warp issues the copies of a pipeline stage with cp.async.bulk.tensor.2d: the A tile (64 K-elements
//                x 128 rows) and, when the weights are not resident, the W block (64 x N) — both land in the 128-byte
//                swizzled K-major UMMA layout, out-of-range K / M is zero-filled by the copy engine, completion is counted
//                in bytes on the stage mbarrier (`mbarrier.arrive.expect_tx`).  Up to 13 stages are in flight.
//                Weights up to 60 KB are loaded once per CTA (cp.async, no-swizzle layout) and stay resident.
//   transformers (gated project layers on small maps) wait for a landed stage, multiply the A tile in place by the
//                per-frame squeeze-excite gate (gate slice staged beside the tile by two loader warps), issue
//                `fence.proxy.async` and hand the stage to the MMA warp; the warps form groups that take alternate stages.
//   MMA issuer   one elected lane of a converged warp (`elect.sync`: operands stay in uniform registers) issues tcgen05.mma (M=128, N<=256, K=16) per 16-wide K step with descriptors that only
//                add to the 14-bit start-address field; accumulators in TMEM (ring of up to 8), tcgen05.commit releases
//                smem stages / signals the epilogue.
//   epilogue     8-16 warps in sets that take alternate tiles: tcgen05.ld (32 lanes x 16 columns), +bias, SiLU / exact GELU,
//                +residual (16-bit, or fp32 in place), pack, 32-byte stores; head variant: SiLU + average pool.
// Almost every EfficientNet layer is HBM-bound: the design goal there is "read A once, write D once, keep
// enough bytes in flight"; the ViT contractions are tensor-bound and reuse the same kernel with streamed weights.
#include "common.cuh"
#include "kernels.h"
#include "conv_map.h"
#include <cuda.h>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

namespace dfd {

constexpr int kBM = 128;            // rows per tile = TMEM lanes
constexpr int kKB = 64;             // K elements per pipeline stage (8 chunks of 8)
constexpr int kMaxStages = 16;
// Warp roles are template parameters: D-heavy layers (expand, head: SiLU on every output) get 16 epilogue
// and 4 loader warps; A-heavy gated project layers 8 epilogue, 4 loader and 8 transformer warps.  Epilogue
// warps come first so that (warp index % 4) is the TMEM lane quarter a warp may access.
constexpr uint32_t kAStageBytes = kBM * 128;                         // A tile of a stage: 128 rows x 64 elements, 128-byte swizzle (TMA)

struct GemmArgs {
    const void* A; const void* W; const float* bias; const float* gate; const void* R; void* D; float* feat;
    int64_t M; int K; int N; int HW;
    int NB, NBp, n_chunks;          // columns per work unit, padded to 16, units along N
    int rows_per_tile;              // 128, or frames_per_tile*HW for the pooled head
    int conv_h;                     // CONV variants (3x3 convolution over a zero-haloed map, see launch_gemm_tc_conv3x3): map height.
                                    //    The three conv_* fields sit in alignment holes, so the parameter layout (and the
                                    //    generated code) of the CONV == 0 kernels is exactly what was verified on the GPU.
    int64_t m_tiles;
    int tpf;                        // > 0: frame-aligned tiling (tpf tiles per frame, the last one partial) with per-frame weights
    int conv_w;                     // CONV variants: map width (the padded map is (conv_h + 2) x (conv_w + 2))
    int64_t w_frame_stride;         // elements between the weight matrices of consecutive frames (frame-aligned mode)
    int stages;
    int nacc, na;                   // TMEM accumulator buffers; epilogue group sets that take alternate tiles
    int nf_max, total_frames;       // gated: frames a tile can touch, frames in the tensor
    uint32_t g_stage_bytes;         // gated: bytes of the per-stage gate slice [nf_max][64] fp32
    int b_resident;                 // 1: whole W lives in shared memory for the CTA's lifetime; 2: only the CTA's own
                                    //    column chunk (grid is a multiple of n_chunks, so a CTA always works on the same chunk)
    int kchunks_pad;                // K/8 rounded up to even
    int dbg;                        // timing experiments (DFD_GEMM_DBG): 1 skip loads, 2 skip MMAs, 4 skip stores,
                                    //    8 skip the transformers' proxy fence, 16 per-thread (not per-warp) arrivals,
                                    //    32 per-role wait accounting, 64 transformers skip the tile, 128 transformers skip the stores
    int xg;                         // gated: transformer warp groups taking alternate stages
    uint32_t lbo_b, stage_bytes, b_stage_bytes, b_chunk_bytes, b_res_bytes, tmem_cols;
    float inv_hw;
    int conv_cpk;                   // CONV == 2: 64-element k-blocks per filter tap (= input channels / 64)
    // unit strides of the persistent loops, decomposed on the host so that no role divides per tile
    int64_t stride1, strideE;       // grid, na * grid
    int64_t d1_mt, dE_mt, d1_frame, dE_frame;
    int d1_nc, dE_nc, d1_t, dE_t;
};
static_assert(sizeof(GemmArgs) == 256, "GemmArgs layout is part of the verified kernels' parameter space");

// Walks the work units u = u0, u0 + stride, ... of one warp role: (M tile, N chunk) and, for frame-aligned tiling
// (per-frame weights), (frame, tile inside the frame).  Divisions happen once in init(); advance() only adds.
struct TileIter {
    int64_t u, mt, frame;
    int nc, t;
    __device__ __forceinline__ void init(const GemmArgs& p, int64_t u0) {
        u = u0; mt = u0 / p.n_chunks; nc = (int)(u0 - mt * p.n_chunks);
        frame = 0; t = 0;
        if (p.tpf > 0) { frame = mt / p.tpf; t = (int)(mt - frame * p.tpf); }
    }
    __device__ __forceinline__ void advance(const GemmArgs& p, int64_t stride, int64_t d_mt, int d_nc, int64_t d_frame, int d_t) {
        u += stride; mt += d_mt; nc += d_nc;
        int carry = 0;
        if (nc >= p.n_chunks) { nc -= p.n_chunks; ++mt; carry = 1; }
        if (p.tpf > 0) {
            t += d_t + carry; frame += d_frame;
            if (t >= p.tpf) { t -= p.tpf; ++frame; }
        }
    }
    __device__ __forceinline__ void next1(const GemmArgs& p) { advance(p, p.stride1, p.d1_mt, p.d1_nc, p.d1_frame, p.d1_t); }
    __device__ __forceinline__ void nextE(const GemmArgs& p) { advance(p, p.strideE, p.dE_mt, p.dE_nc, p.dE_frame, p.dE_t); }
    // first row and number of valid rows of the tile
    __device__ __forceinline__ int64_t m0(const GemmArgs& p) const {
        return p.tpf > 0 ? frame * p.HW + (int64_t)t * kBM : mt * p.rows_per_tile;
    }
    __device__ __forceinline__ int rows_valid(const GemmArgs& p, int64_t m0v) const {
        return p.tpf > 0 ? min(kBM, p.HW - t * kBM) : (int)min((int64_t)p.rows_per_tile, p.M - m0v);
    }
};

// CONV (3x3 stride-1 pad-1 convolutions of the resnet50 member without a gathered operand; the map between two such
// launches lives in a zero-haloed layout [frames][H+2][W+2][C] preceded by a guard of W+3 rows, so that filter tap (ky,kx)
// of padded pixel p is the plain row p + ky*(W+2) + kx of a 2-D tensor map and every tap tile is ONE ordinary TMA box):
//   1  pointwise GEMM whose output rows are scattered into that layout (interior pixels only; the halo stays zero)
//   2  implicit GEMM over it: k-block kb = (tap, 64-channel slice), A box at row m0 + ky*(W+2) + kx; M runs over the padded
//      pixels, halo rows are computed and dropped, interior rows are stored to the ordinary [frames*H*W][N] layout
// ACT: 0 none, 1 SiLU, 2 exact-erf GELU (ViT MLP), 3 ReLU applied AFTER the residual add (ResNet bottleneck: relu(bn(conv) + identity)).  F32OUT: fp32 D (and fp32 R when RES: the ViT residual stream, in place).
// TSTORE = 32 / 16 (expand layers whose column chunk is a multiple of 32 / 16): the epilogue stages 32 rows x TSTORE columns per warp
// in shared memory and writes them with ONE TMA store — a thread owns an accumulator ROW, so its direct 32-byte stores touch 32
// different lines per warp instruction; with the stores skipped the 192 -> 1152 layer at 7x7 ran in 54 instead of 91 us
// (profiles/r02_experimental.md).  The launcher selects 32-column boxes on 64-byte-aligned rows only (N = 96, 672, 1152): 16-column
// boxes write the same 32-byte row segments as the direct stores and measured exactly their time.
template <typename T, bool GATE, int ACT, bool RES, bool POOL, int kEpiWarps, int kProdWarps, int kXformWarps, bool F32OUT = false, int CONV = 0, int TSTORE = 0>
__global__ void __launch_bounds__((kEpiWarps + 1 + kProdWarps + kXformWarps) * 32, 1) gemm_tc_kernel(const GemmArgs p, const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmD) {
    static_assert(GATE == (kXformWarps > 0), "transformer warps exist exactly for gated layers");
    static_assert(TSTORE == 0 || ((TSTORE == 16 || TSTORE == 32) && !GATE && !RES && !POOL && !F32OUT && CONV == 0), "TMA-store epilogue: plain 16-bit outputs only");
    static_assert(CONV == 0 || (!GATE && !RES && !POOL && !F32OUT), "CONV variants: plain 16-bit epilogue only");
    constexpr int kProdThreads = kProdWarps * 32;
    constexpr int kXformThreads = kXformWarps * 32;
    constexpr int kGemmThreads = (kEpiWarps + 1 + kProdWarps + kXformWarps) * 32;
    constexpr int kMmaWarp = kEpiWarps;
    constexpr int kColGroups = kEpiWarps / 4;       // epilogue warps sharing a TMEM lane quarter split the columns
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // ---- shared memory carve-up ----------------------------------------------------------------
    // [resident W][pad to 1024][stages x (A tile 16 KB | streamed W block | gate slice), 1024-aligned][bias, pool, barriers]
    const uint32_t stage_bytes = p.stage_bytes;
    const uint32_t g_off = kAStageBytes + (p.b_resident ? 0u : p.b_stage_bytes);
    const uint32_t bres_base = smem_u32(smem_raw);
    const uint32_t smem_base = (bres_base + p.b_res_bytes + 1023u) & ~1023u;
    uint8_t* sp = smem_raw + (smem_base - bres_base) + (size_t)p.stages * stage_bytes;
    const uint32_t sm_out = smem_u32(sp);                         // TSTORE: one 2 KB staging slab per epilogue warp (1024-aligned)
    if (TSTORE != 0) sp += kEpiWarps * 2048;
    float* s_bias = reinterpret_cast<float*>(sp);                 sp += (size_t)((p.N + 3) & ~3) * 4;
    float* s_pool = reinterpret_cast<float*>(sp);                 if (POOL) sp += kColGroups * kBM * 17 * 4;
    sp = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sp) + 7) & ~uintptr_t(7));
    uint64_t* bars = reinterpret_cast<uint64_t*>(sp);             // full[S], empty[S], raw[S], tfull[8], tempty[8]
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 3 * kMaxStages + 16);

    const uint32_t bar_full = smem_u32(bars), bar_empty = smem_u32(bars + kMaxStages);
    const uint32_t bar_raw = smem_u32(bars + 2 * kMaxStages);
    const uint32_t bar_tfull = smem_u32(bars + 3 * kMaxStages), bar_tempty = smem_u32(bars + 3 * kMaxStages + 8);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    for (int i = threadIdx.x; i < p.N; i += kGemmThreads) s_bias[i] = p.bias[i];
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.stages; ++s) {
            // a landed stage = the TMA issuer's arrival + its transaction bytes
            mbar_init(bar_full + 8 * s, GATE ? (p.dbg & 16 ? kXformThreads : kXformWarps) / p.xg : 1);
            mbar_init(bar_empty + 8 * s, 1);
            mbar_init(bar_raw + 8 * s, 1 + 64);                 // TMA issuer + the 64 gate-slice copiers
        }
        for (int a = 0; a < p.nacc; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, (kColGroups / p.na) * (p.dbg & 16 ? 128 : 4)); }
        fence_barrier_init();
    }
    if (warp == kMmaWarp) tmem_alloc(smem_u32(s_tmem), p.tmem_cols);
    if (warp > kMmaWarp && warp <= kMmaWarp + kProdWarps && p.b_resident) {
        const int tp = threadIdx.x - (kMmaWarp + 1) * 32;
        const T* Wt = reinterpret_cast<const T*>(p.W);
        const int kch = p.K >> 3, per = p.NBp * p.kchunks_pad;
        const int c_lo = p.b_resident == 2 ? (int)(blockIdx.x % p.n_chunks) : 0;
        const int c_hi = p.b_resident == 2 ? c_lo + 1 : p.n_chunks;
        for (int c = c_lo; c < c_hi; ++c) {
            const int nbv = min(p.NB, p.N - c * p.NB);
            const uint32_t dst = bres_base + (p.b_resident == 2 ? 0 : c * p.b_chunk_bytes);
            for (int i = tp; i < per; i += kProdThreads) {
                const int r = i / p.kchunks_pad, q = i - r * p.kchunks_pad;
                const bool ok = r < nbv && q < kch;
                cp_async16(dst + q * p.lbo_b + r * 16, ok ? Wt + (size_t)(c * p.NB + r) * p.K + q * 8 : Wt, ok);
            }
        }
        cp_async_commit();
        cp_async_wait<0>();
        fence_proxy_async_smem();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *s_tmem;
    // programmatic dependent launch: the set-up above read only the bias and (resident case, never per-frame weights) W
    griddep_wait();

    const int num_kb = (p.K + kKB - 1) / kKB;
    const int64_t units = p.m_tiles * p.n_chunks;
    // DFD_GEMM_DBG & 32: per-role wait accounting (block 0 prints cycles spent in each barrier wait)
    const bool prof = (p.dbg & 32) != 0;
    long long w0 = 0, w1 = 0, t_begin = prof ? clock64() : 0;
#define DFD_TWAIT(acc, bar, par) { if (prof) { const long long c0_ = clock64(); mbar_wait(bar, par); acc += clock64() - c0_; } else mbar_wait(bar, par); }

    if (warp > kMmaWarp && warp <= kMmaWarp + kProdWarps) {
        // =================================== LOADERS =============================================
        // ONE thread issues the TMA copies of a stage — the A tile (box 64 x 128, zero fill outside the tensor) and,
        // for weights too large to stay resident, the W block (box 64 x NBp) — both into the 128-byte-swizzled UMMA
        // layout, completion counted in bytes on the stage barrier.  No other loader thread has work in the main loop.
        const int tp = threadIdx.x - (kMmaWarp + 1) * 32;
        if (tp < 32) {                                                     // the whole first loader warp, converged (elect.sync inside)
            if (tp == 0) { tma_prefetch_desc(&tmA); if (!p.b_resident) tma_prefetch_desc(&tmB); }
            const uint32_t tx_bytes = kAStageBytes + (p.b_resident ? 0u : p.b_stage_bytes);
            int stage = 0; uint32_t phase = 0;
            TileIter it; it.init(p, blockIdx.x);
            for (; it.u < units; it.next1(p)) {
                const int m0 = (int)it.m0(p);
                const int wrow = (int)it.frame * p.N + it.nc * p.NB;       // per-frame weights when tpf > 0
                ConvTapIter tap; tap.init(m0);                             // CONV == 2: (64-channel slice, row) of the k-block's A box
                for (int kb = 0; kb < num_kb; ++kb) {
                    DFD_TWAIT(w0, bar_empty + 8 * stage, phase ^ 1)
                    const uint32_t a_base = smem_base + stage * stage_bytes;
                    const uint32_t bar = (GATE ? bar_raw : bar_full) + 8 * stage;
                    mbar_arrive_expect_tx_elect(bar, (p.dbg & 1) ? 0u : tx_bytes);
                    if (!(p.dbg & 1)) {
                        if (CONV == 2) tma_load_2d_elect(a_base, &tmA, tap.ck * kKB, tap.row, bar);
                        else tma_load_2d_elect(a_base, &tmA, kb * kKB, m0, bar);
                        if (!p.b_resident) tma_load_2d_elect(a_base + kAStageBytes, &tmB, kb * kKB, wrow, bar);
                    }
                    if (CONV == 2) tap.next(p.conv_cpk, p.conv_w);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        } else if (GATE && tp >= 32 && tp < 96) {
            // gated layers: loader warps 1-2 stage the per-frame gate slice [nf_max][64] fp32 of every k-block next to
            // the A tile (cp.async; the landed-barrier counts their 64 completion arrivals besides the TMA bytes)
            const int i = tp - 32;
            const uint32_t fl = i >> 4, c = i & 15;
            int stage = 0; uint32_t phase = 0;
            TileIter it; it.init(p, blockIdx.x);
            for (; it.u < units; it.next1(p)) {
                const uint32_t f0 = (uint32_t)it.m0(p) / (uint32_t)p.HW;
                for (int kb = 0; kb < num_kb; ++kb) {
                    const int k0 = kb * kKB;
                    DFD_TWAIT(w0, bar_empty + 8 * stage, phase ^ 1)
                    const uint32_t g_base = smem_base + stage * stage_bytes + g_off;
                    for (uint32_t f = fl; f < (uint32_t)p.nf_max; f += 4) {
                        const bool ok = (f0 + f < (uint32_t)p.total_frames) && (k0 + (int)c * 4 < p.K);
                        cp_async16(g_base + f * 256 + c * 16, ok ? p.gate + (size_t)(f0 + f) * p.K + k0 + c * 4 : p.gate, ok);
                    }
                    cp_async_mbar_arrive_noinc(bar_raw + 8 * stage);
                    if (++stage == p.stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (GATE && warp > kMmaWarp + kProdWarps) {
        // =================================== TRANSFORMERS ========================================
        // The warps form p.xg groups that take alternate stages: a stage's lds -> multiply -> sts -> proxy fence
        // chain is latency-bound, so several stages are transformed concurrently.  One arrival per warp.
        const int tx = threadIdx.x - (kMmaWarp + 1 + kProdWarps) * 32;
        const int gthreads = kXformThreads / p.xg;
        const int grp = tx / gthreads, tg = tx - grp * gthreads;
        const int q = tg & 7, rb = tg >> 3, xstep = gthreads >> 3, xpasses = (kBM + xstep - 1) / xstep;      // (rows past rows_valid are skipped in the loop)
        int stage = 0; uint32_t phase = 0; int turn = 0;
        TileIter it; it.init(p, blockIdx.x);
        for (; it.u < units; it.next1(p)) {
            const int64_t m0 = it.m0(p);
            const int rows_valid = it.rows_valid(p, m0);
            const uint32_t f0 = (uint32_t)m0 / (uint32_t)p.HW;
            const uint32_t rem0 = (uint32_t)m0 - f0 * (uint32_t)p.HW + rb;      // row rb relative to frame f0
            for (int kb = 0; kb < num_kb; ++kb) {
                if (turn == grp) {
                    const int kc = min(8, (p.K - kb * kKB) >> 3);
                    DFD_TWAIT(w0, bar_raw + 8 * stage, phase)
                    if (q < kc && !(p.dbg & 64)) {
                        const uint32_t a_base = smem_base + stage * stage_bytes;
                        uint32_t fl = 0, rem = rem0;
                        while (rem >= (uint32_t)p.HW) { rem -= p.HW; ++fl; }
                        uint32_t fl_loaded = 0xffffffffu;               // the gate slice is re-read only when the frame changes
                        uint4 gA = make_uint4(0, 0, 0, 0), gB = gA;
#pragma unroll 4
                        for (int j = 0; j < xpasses; ++j) {
                            if (rb + xstep * j < rows_valid) {
                                if (fl != fl_loaded) {
                                    const uint32_t gaddr = a_base + g_off + fl * 256 + q * 32;
                                    gA = lds16(gaddr); gB = lds16(gaddr + 16); fl_loaded = fl;
                                }
                                const int r = rb + xstep * j;                       // 128-byte swizzle: chunk q of row r
                                const uint32_t addr = a_base + r * 128 + ((q ^ (r & 7)) << 4);
                                uint4 v = lds16(addr);
                                const float2 x0 = Half16<T>::unpack(v.x), x1 = Half16<T>::unpack(v.y);
                                const float2 x2 = Half16<T>::unpack(v.z), x3 = Half16<T>::unpack(v.w);
                                v.x = Half16<T>::pack(x0.x * __uint_as_float(gA.x), x0.y * __uint_as_float(gA.y));
                                v.y = Half16<T>::pack(x1.x * __uint_as_float(gA.z), x1.y * __uint_as_float(gA.w));
                                v.z = Half16<T>::pack(x2.x * __uint_as_float(gB.x), x2.y * __uint_as_float(gB.y));
                                v.w = Half16<T>::pack(x3.x * __uint_as_float(gB.z), x3.y * __uint_as_float(gB.w));
                                if (!(p.dbg & 128)) sts16(addr, v);
                                else asm volatile("" :: "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
                            }
                            rem += xstep;
                            while (rem >= (uint32_t)p.HW) { rem -= p.HW; ++fl; }
                        }
                    }
                    if (!(p.dbg & 8)) fence_proxy_async_smem();
                    if (p.dbg & 16) mbar_arrive(bar_full + 8 * stage);
                    else { __syncwarp(); if (lane == 0) mbar_arrive(bar_full + 8 * stage); }
                }
                if (++turn == p.xg) turn = 0;
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kMmaWarp) {
        // =================================== MMA ISSUER ==========================================
        const uint32_t idesc = umma_idesc(Half16<T>::kUmmaFormat, kBM, (uint32_t)p.NBp);
        const uint64_t a_d0 = umma_smem_desc_sw128(smem_base);
        const uint64_t b_d0 = p.b_resident ? umma_smem_desc(bres_base, p.lbo_b, 128) : umma_smem_desc_sw128(smem_base + kAStageBytes);
        const uint32_t a_hi = (uint32_t)(a_d0 >> 32), a_lo0 = (uint32_t)a_d0, b_hi = (uint32_t)(b_d0 >> 32);
        const uint32_t b_lo_res = (uint32_t)b_d0, b_lo_str = (uint32_t)b_d0;
        const uint32_t stage_step = stage_bytes >> 4, b_chunk_step = p.b_chunk_bytes >> 4, b_kb_step = (8 * p.lbo_b) >> 4;
        const uint32_t b_jstep = p.b_resident ? (2 * p.lbo_b) >> 4 : 2u;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0; uint32_t phase = 0; int acc = 0; uint32_t acc_phase = 0;
        TileIter it; it.init(p, blockIdx.x);
        for (; it.u < units; it.next1(p)) {
            const int nc = it.nc;
            DFD_TWAIT(w0, bar_tempty + 8 * acc, acc_phase ^ 1)
            tc_fence_after_sync();
            const uint32_t d_tmem = tmem_u + (uint32_t)(acc * p.NBp);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int kc = min(8, (p.K - kb * kKB) >> 3);
                const int steps = (kc + 1) >> 1;
                DFD_TWAIT(w1, bar_full + 8 * stage, phase)
                if (!(p.dbg & 512)) tc_fence_after_sync();
                {
                    // The whole warp runs this block converged; `elect.sync` inside the helpers picks the issuing lane, so the
                    // operands stay warp-uniform (behind `if (lane == 0)` every UTCHMMA operand went through an ELECT + R2UR loop:
                    // ~0.5 us per k-block in this warp whatever the tile did — measured with loads, MMAs and stores all skipped).
                    // descriptors differ only in the 14-bit start-address field: add to the low word
                    const uint32_t a_lo = a_lo0 + (uint32_t)stage * stage_step;
                    const uint32_t b_lo = p.b_resident ? b_lo_res + (uint32_t)(p.b_resident == 2 ? 0 : nc) * b_chunk_step + (uint32_t)kb * b_kb_step
                                                       : b_lo_str + (uint32_t)stage * stage_step;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        if (j < steps && !(p.dbg & 2))
                            umma_f16_elect(d_tmem, ((uint64_t)a_hi << 32) | (a_lo + 2u * j), ((uint64_t)b_hi << 32) | (b_lo + (uint32_t)j * b_jstep),
                                           idesc, (kb > 0 || j > 0) ? 1u : 0u);
                    }
                    if (p.dbg & 256) {                                  // timing experiment (with dbg & 2): plain arrives instead of commits
                        if (lane == 0) {
                            mbar_arrive(bar_empty + 8 * stage);
                            if (kb == num_kb - 1) mbar_arrive(bar_tfull + 8 * acc);
                        }
                    } else {
                        umma_commit_elect(bar_empty + 8 * stage);       // smem stage reusable once the MMAs retire
                        if (kb == num_kb - 1) umma_commit_elect(bar_tfull + 8 * acc);
                    }
                }
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
        }
    } else {
        // =================================== EPILOGUE ============================================
        // Epilogue groups (4 warps = 128 TMEM lanes each) are split into `na` sets that take alternate tiles, so
        // several tiles are drained concurrently; the kColGroups/na groups of one set split a tile's columns.
        const int q = warp & 3, half = warp >> 2;      // half = epilogue group of this warp
        const int gpt = kColGroups / p.na, set = half / gpt, sub = half - set * gpt;
        const int row = 32 * q + lane;
        T* D = reinterpret_cast<T*>(p.D);
        const T* R = reinterpret_cast<const T*>(p.R);
        int acc = set; uint32_t acc_phase = 0;         // set < na <= nacc, nacc is a multiple of na
        TileIter it; it.init(p, blockIdx.x + (int64_t)set * gridDim.x);
        for (; it.u < units; it.nextE(p)) {
            const int nc = it.nc;
            const int64_t m0 = it.m0(p);
            const int rows_valid = it.rows_valid(p, m0);
            const int n0 = nc * p.NB;
            const int nb_valid = min(p.NB, p.N - n0);
            bool valid = row < rows_valid;
            int64_t m = m0 + row;
            if (CONV == 1) {           // rows are interior pixels: scatter into the zero-haloed layout
                m = conv_pad_row((uint32_t)m, (uint32_t)p.conv_h, (uint32_t)p.conv_w);
            } else if (CONV == 2) {    // rows are padded pixels: only interior ones are outputs
                int64_t mo;
                const bool interior = conv_unpad_row((uint32_t)m, (uint32_t)p.conv_h, (uint32_t)p.conv_w, &mo);
                valid = valid && interior; m = mo;
            }
            DFD_TWAIT(w0, bar_tfull + 8 * acc, acc_phase)
            tc_fence_after_sync();
            const uint32_t t_row = tmem_base + ((uint32_t)(32 * q) << 16) + (uint32_t)(acc * p.NBp);
            const int slots = POOL ? rows_valid / p.HW : 0;
            if constexpr (TSTORE != 0) {
                // TSTORE columns at a time: TMEM -> + bias -> SiLU -> 16-bit -> the warp's slab [32 rows][2 TSTORE bytes] in the swizzle
                // of the output tensor map (64-byte rows: 16-byte chunk ^= row / 2 % 4; 32-byte rows: chunk ^= row / 4 % 2) -> one
                // TMA store of the 32 x TSTORE box (rows past M are clipped by the map)
                constexpr int NH = TSTORE / 16;
                const uint32_t slab = sm_out + (uint32_t)warp * 2048u;
                const uint32_t srow = slab + (uint32_t)lane * (2u * TSTORE);
                const uint32_t sw = TSTORE == 32 ? ((uint32_t)lane >> 1) & 3u : ((uint32_t)lane >> 2) & 1u;
                (void)valid; (void)m;
                for (int cg = sub; cg * TSTORE < p.NB; cg += gpt) {
                    uint32_t r[NH][16];
#pragma unroll
                    for (int h = 0; h < NH; ++h) tmem_ld16(t_row + cg * TSTORE + h * 16, r[h]);
                    tmem_ld_wait();
                    uint32_t pk[NH * 8];
#pragma unroll
                    for (int h = 0; h < NH; ++h)
#pragma unroll
                        for (int i = 0; i < 16; i += 2) {
                            uint64_t x = add2(f2_pack(__uint_as_float(r[h][i]), __uint_as_float(r[h][i + 1])),
                                              *reinterpret_cast<const uint64_t*>(&s_bias[n0 + cg * TSTORE + h * 16 + i]));
                            if (ACT == 1) x = neg_silu2(x);                   // -silu(x)
                            const float2 xf = f2_unpack(x);
                            pk[h * 8 + (i >> 1)] = ACT == 1 ? Half16<T>::pack(-xf.x, -xf.y) : Half16<T>::pack(xf.x, xf.y);
                        }
                    if (lane == 0) bulk_wait_group_read0();                     // the warp's previous store has read the slab
                    __syncwarp();
#pragma unroll
                    for (uint32_t j = 0; j < 2 * NH; ++j) sts16(srow + ((j ^ sw) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0 && !(p.dbg & 4)) {
                        tma_store_2d(&tmD, slab, n0 + cg * TSTORE, (int)m0 + 32 * q);
                        bulk_commit_group();
                    }
                }
            } else
            for (int c16 = sub; c16 * 16 < p.NBp; c16 += gpt) {
                uint32_t r[16];
                tmem_ld16(t_row + c16 * 16, r);
                tmem_ld_wait();
                const int ncol = nb_valid - c16 * 16;               // valid columns in this chunk (>=16, 8 or <=0)
                if (ncol <= 0) continue;
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; i += 2) {
                    const int col = min(n0 + c16 * 16 + i, p.N - 2);
                    uint64_t x = add2(f2_pack(__uint_as_float(r[i]), __uint_as_float(r[i + 1])),
                                      *reinterpret_cast<const uint64_t*>(&s_bias[col]));
                    if (ACT == 1) x = neg_silu2(x);                   // -silu(x); sign restored below
                    const float2 xf = f2_unpack(x);
                    if (ACT == 2) {
                        v[i] = 0.5f * xf.x * (1.0f + erff(xf.x * 0.70710678118654752f));
                        v[i + 1] = 0.5f * xf.y * (1.0f + erff(xf.y * 0.70710678118654752f));
                    } else {
                        v[i] = ACT == 1 ? -xf.x : xf.x; v[i + 1] = ACT == 1 ? -xf.y : xf.y;
                    }
                }
                if (POOL) {
                    // Batch-invariant average pool: stage the SiLU'd tile in shared memory, then add each frame's
                    // HW rows in an order that depends only on the row index inside the frame (4 fixed row
                    // quarters, combined as (p0+p1)+(p2+p3)), never on where the frame sits in the tile.
                    float* tile = s_pool + half * (kBM * 17);
                    if (valid) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) tile[row * 17 + i] = v[i];
                    }
                    asm volatile("bar.sync %0, 128;" :: "r"(1 + half) : "memory");
                    const int th = q * 32 + lane;
                    const int rq = (p.HW + 3) >> 2;
                    const int64_t frame0 = m0 / p.HW;
                    for (int item = th; item < slots * 64; item += 128) {
                        const int part = item & 3, col = (item >> 2) & 15, s = item >> 6;
                        const int r0 = part * rq, r1 = min(p.HW, r0 + rq);
                        float tot = 0.f;
                        for (int r = r0; r < r1; ++r) tot += tile[(s * p.HW + r) * 17 + col];
                        tot += __shfl_xor_sync(0xffffffffu, tot, 1);
                        tot += __shfl_xor_sync(0xffffffffu, tot, 2);
                        if (part == 0 && col < ncol) p.feat[(size_t)(frame0 + s) * p.N + n0 + c16 * 16 + col] = tot * p.inv_hw;
                    }
                    asm volatile("bar.sync %0, 128;" :: "r"(1 + half) : "memory");
                } else if (F32OUT) {
                    if (valid) {                                      // fp32 output (recurrent head): 64 bytes per chunk
                        float* dst = reinterpret_cast<float*>(p.D) + (size_t)m * p.N + n0 + c16 * 16;
                        if (RES) {                                    // fp32 residual (may alias D: same thread reads then writes)
                            const float* rs = reinterpret_cast<const float*>(p.R) + (size_t)m * p.N + n0 + c16 * 16;
                            const U32x8 r0 = ldg32_coherent(rs);
                            U32x8 r1 = r0;
                            if (ncol > 8) r1 = ldg32_coherent(rs + 8);
#pragma unroll
                            for (int i = 0; i < 8; ++i) { v[i] += __uint_as_float(r0.v[i]); if (ncol > 8) v[8 + i] += __uint_as_float(r1.v[i]); }
                        }
                        U32x8 o0, o1;
#pragma unroll
                        for (int i = 0; i < 8; ++i) { o0.v[i] = __float_as_uint(v[i]); o1.v[i] = __float_as_uint(v[8 + i]); }
                        stg32(dst, o0);
                        if (ncol > 8) stg32(dst + 8, o1);
                    }
                } else if (valid && !(p.dbg & 4)) {
                    T* dst = D + (size_t)m * p.N + n0 + c16 * 16;
                    const bool wide = ((p.N & 15) == 0);              // every 16-column chunk is then 32-byte aligned
                    if (RES) {
                        const T* rs = R + (size_t)m * p.N + n0 + c16 * 16;
                        uint32_t rr[8];
                        if (wide) {
                            const U32x8 t = ldg32(rs);
#pragma unroll
                            for (int i = 0; i < 8; ++i) rr[i] = t.v[i];
                        } else {
                            const uint4 t0 = ldg16_stream(rs);
                            const uint4 t1 = (ncol > 8) ? ldg16_stream(rs + 8) : make_uint4(0, 0, 0, 0);
                            rr[0] = t0.x; rr[1] = t0.y; rr[2] = t0.z; rr[3] = t0.w; rr[4] = t1.x; rr[5] = t1.y; rr[6] = t1.z; rr[7] = t1.w;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const float2 a = Half16<T>::unpack(rr[i]);
                            v[2 * i] += a.x; v[2 * i + 1] += a.y;
                        }
                    }
                    if (ACT == 3) {
#pragma unroll
                        for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                    }
                    U32x8 o;
#pragma unroll
                    for (int i = 0; i < 8; ++i) o.v[i] = Half16<T>::pack(v[2 * i], v[2 * i + 1]);
                    if (wide) {
                        stg32(dst, o);
                    } else {
                        stg16(dst, make_uint4(o.v[0], o.v[1], o.v[2], o.v[3]));
                        if (ncol > 8) stg16(dst + 8, make_uint4(o.v[4], o.v[5], o.v[6], o.v[7]));
                    }
                }
            }
            tc_fence_before_sync();
            if (p.dbg & 16) mbar_arrive(bar_tempty + 8 * acc);
            else { __syncwarp(); if (lane == 0) mbar_arrive(bar_tempty + 8 * acc); }     // one arrival per warp
            acc += p.na; if (acc >= p.nacc) { acc -= p.nacc; acc_phase ^= 1; }
        }
        if (TSTORE != 0 && lane == 0) bulk_wait_group0();                     // the slabs must outlive their stores
    }

    if (prof && blockIdx.x == 0 && lane == 0)
        printf("warp %2d (%s): total %lld cycles, wait0 %lld, wait1 %lld\n", warp,
               warp < kMmaWarp ? "epilogue" : warp == kMmaWarp ? "mma" : warp <= kMmaWarp + kProdWarps ? "loader" : "xform",
               clock64() - t_begin, w0, w1);
#undef DFD_TWAIT
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kMmaWarp) { tc_fence_after_sync(); tmem_dealloc(tmem_base, p.tmem_cols); }
}

// ---------------------------------------------------------------------------------------------------
static int g_num_sms = 0;

// Tensor map of a K-major operand (A, or streamed W): [M rows][K elements] 16-bit row-major, box = 64 elements x box_rows,
// 128-byte swizzle,
// zero fill outside the tensor (K tails, the last M tile).  cuTensorMapEncodeTiled comes from the driver through the
// runtime's entry-point query (no link against libcuda); maps are cached per thread by (pointer, M, K).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cudaError_t make_tmap(const void* A, int64_t M, int K, int box_rows, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    struct Key { const void* p; int64_t m; int k, b; bool operator==(const Key& o) const { return p == o.p && m == o.m && k == o.k && b == o.b; } };
    struct Hash { size_t operator()(const Key& k) const { return std::hash<const void*>()(k.p) ^ (size_t)k.m * 1315423911u ^ (size_t)k.k * 2654435761u ^ (size_t)k.b * 40503u; } };
    thread_local std::unordered_map<Key, CUtensorMap, Hash> cache;
    const Key key{A, M, K, box_rows};
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return cudaSuccess; }
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kKB, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    // 16-bit payload: the element type only matters for the (unused) NaN fill, UINT16 serves fp16 and bf16 alike
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(A), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
    if (cache.size() > 4096) cache.clear();
    cache.emplace(key, *out);
    return cudaSuccess;
}

// 16-bit output matrix [M][N] row-major: boxes of 32 (16) columns x 32 rows, 64-byte (32-byte) swizzle (the TSTORE epilogue's slabs)
static cudaError_t make_tmap_out(const void* D, int64_t M, int N, int box_cols, CUtensorMap* out) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t dims[2] = {(cuuint64_t)N, (cuuint64_t)M};
    const cuuint64_t strides[1] = {(cuuint64_t)N * 2};
    const cuuint32_t box[2] = {(cuuint32_t)box_cols, 32u};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(out, CU_TENSOR_MAP_DATA_TYPE_UINT16, 2, const_cast<void*>(D), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, box_cols == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

cudaError_t make_tmap_2d(const void* base, int64_t rows, int cols, int box_rows, void* out_tmap) {
    return make_tmap(base, rows, cols, box_rows, reinterpret_cast<CUtensorMap*>(out_tmap));
}

// fp32 matrix [rows][cols] row-major: boxes of 32 columns (128 bytes) x box_rows rows, 128-byte swizzle (TMA stores / reductions)
cudaError_t make_tmap_2d_f32(const float* base, int64_t rows, int cols, int box_rows, void* out_tmap) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
        if (e != cudaSuccess) return e;
        if (q != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    if ((cols & 3) || box_rows < 1 || box_rows > 256) return cudaErrorInvalidValue;
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 4};
    const cuuint32_t box[2] = {32u, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = encode(reinterpret_cast<CUtensorMap*>(out_tmap), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// columns per work unit: N split into equal chunks of <= 256 columns (multiple of 8)
static int gemm_n_chunks(int N) {
    int n_chunks = (N + 255) / 256;
    while ((N % n_chunks) != 0 || ((N / n_chunks) & 7)) ++n_chunks;
    return n_chunks;
}

static cudaError_t make_tmap_out(const void* D, int64_t M, int N, int box_cols, CUtensorMap* out);

template <typename KernelT>
static cudaError_t run(KernelT kernel, GemmArgs& a, int epi_warps, int prod_warps, int xform_warps, cudaStream_t s, int tstore = 0) {
    if (g_num_sms == 0) {
        int dev = 0; cudaError_t e = cudaGetDevice(&dev); if (e != cudaSuccess) return e;
        e = cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev); if (e != cudaSuccess) return e;
    }
    const int n_chunks = gemm_n_chunks(a.N);
    a.n_chunks = n_chunks;
    a.NB = a.N / n_chunks;
    a.NBp = (a.NB + 15) & ~15;
    a.lbo_b = (uint32_t)a.NBp * 16 + 16;
    a.kchunks_pad = ((a.K >> 3) + 1) & ~1;
    { static const int env_dbg = getenv("DFD_GEMM_DBG") ? atoi(getenv("DFD_GEMM_DBG")) : 0; a.dbg = env_dbg; }
#ifndef DFD_GEMM_XG
#define DFD_GEMM_XG 4
#endif
    a.xg = xform_warps > 0 ? DFD_GEMM_XG : 1;                      // transformer warp groups (take alternate stages)
    if (a.xg > xform_warps || (xform_warps % a.xg)) a.xg = 1;
    a.b_stage_bytes = (uint32_t)a.NBp * 128u;                      // streamed W block: NBp rows x 64 elements, 128-byte swizzle
    a.b_chunk_bytes = (uint32_t)a.kchunks_pad * a.lbo_b;
    {   // accumulator ring: as many buffers as fit in the 512 TMEM columns (<= 8); epilogue groups take alternate tiles
        const int groups = epi_warps / 4, fit = 512 / a.NBp;
        int na = groups; while (na > fit) na >>= 1;
        a.na = na;
        int nacc = (fit / na) * na; if (nacc > 8) nacc = 8 / na * na;
        a.nacc = nacc;
    }
    uint32_t cols = 32; while (cols < (uint32_t)(a.nacc * a.NBp)) cols <<= 1;
    a.tmem_cols = cols;
    const size_t fixed = (size_t)((a.N + 3) & ~3) * 4 + (a.feat ? (size_t)(epi_warps / 4) * kBM * 17 * 4 : 0) + 8 + (3 * kMaxStages + 16) * 8 + 16
                         + (tstore ? (size_t)epi_warps * 2048 : 0);
    a.nf_max = a.gate ? (kBM - 1) / a.HW + 2 : 0;
    a.total_frames = (int)((a.M + a.HW - 1) / a.HW);
    a.g_stage_bytes = (uint32_t)a.nf_max * 256u;
    const size_t budget = 227 * 1024;
    const size_t bres = (size_t)a.n_chunks * a.b_chunk_bytes;
    // Weights stay resident only up to 60 KB: beyond that the shared memory is worth more as pipeline stages (W blocks then
    // stream from L2 by TMA).  Measured: 672->112 gated 233 -> 178 us, 112->672 expand 88 -> 78 us per 2048 / 1024 frames.
    const size_t res_limit = 60 * 1024;      // (keeping the 73 KB of the 64 -> 64 implicit 3x3 layer resident changed nothing: 213 vs 209 us)
#ifndef DFD_GEMM_CHUNK_RESLIM
#define DFD_GEMM_CHUNK_RESLIM (60 * 1024)      // limit for keeping ONE column chunk per CTA (tools/build_variant.py sweeps it)
#endif
    a.b_resident = 0;
    if (a.tpf == 0) {
        if (bres <= res_limit) a.b_resident = 1;
        else if (a.n_chunks > 1 && a.n_chunks <= g_num_sms && a.b_chunk_bytes <= (size_t)DFD_GEMM_CHUNK_RESLIM) a.b_resident = 2;
    }
    a.b_res_bytes = a.b_resident == 1 ? (uint32_t)bres : (a.b_resident == 2 ? a.b_chunk_bytes : 0u);
    const size_t stage_bytes = ((size_t)kAStageBytes + (a.b_resident ? 0 : a.b_stage_bytes) + a.g_stage_bytes + 1023) & ~size_t(1023);
    a.stage_bytes = (uint32_t)stage_bytes;
    const size_t bres_pad = ((size_t)a.b_res_bytes + 1023) & ~size_t(1023);      // + worst-case alignment of the stage ring
    int stages = (int)((budget - fixed - bres_pad - 1024) / stage_bytes);
    if (stages > kMaxStages) stages = kMaxStages;
#ifdef DFD_GEMM_STAGE_CAP
    if (stages > DFD_GEMM_STAGE_CAP) stages = DFD_GEMM_STAGE_CAP;
#endif
    if (stages < 3) return cudaErrorInvalidValue;
    a.stages = stages;
    while (a.xg > 1 && (a.xg > stages || a.xg > 4)) a.xg >>= 1;      // groups take alternate stages: never more groups than stages
    const size_t smem = bres_pad + 1024 + (size_t)stages * stage_bytes + fixed;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t units = a.m_tiles * a.n_chunks;
    unsigned grid = (unsigned)(units < g_num_sms ? units : g_num_sms);
    if (a.b_resident == 2) grid = (unsigned)((g_num_sms / a.n_chunks) * a.n_chunks);     // every CTA keeps one column chunk
    if ((int64_t)grid > units) grid = (unsigned)units;
    {
        auto split = [&](int64_t stride, int64_t& d_mt, int& d_nc, int64_t& d_frame, int& d_t) {
            d_mt = stride / a.n_chunks; d_nc = (int)(stride % a.n_chunks);
            d_frame = a.tpf > 0 ? d_mt / a.tpf : 0; d_t = a.tpf > 0 ? (int)(d_mt % a.tpf) : 0;
        };
        a.stride1 = grid; a.strideE = (int64_t)a.na * grid;
        split(a.stride1, a.d1_mt, a.d1_nc, a.d1_frame, a.d1_t);
        split(a.strideE, a.dE_mt, a.dE_nc, a.dE_frame, a.dE_t);
    }
    CUtensorMap tmA, tmB;
    if (a.conv_cpk > 0)     // CONV == 2: A is the zero-haloed map, guard rows before and after: [M + 2 (W+3)][channels]
        e = make_tmap(a.A, a.M + 2 * (int64_t)(a.conv_w + 3), a.conv_cpk * kKB, kBM, &tmA);
    else
        e = make_tmap(a.A, a.M, a.K, kBM, &tmA);
    if (e != cudaSuccess) return e;
    if (!a.b_resident) {
        if (a.tpf > 0 && a.w_frame_stride != (int64_t)a.N * a.K) return cudaErrorInvalidValue;
        e = make_tmap(a.W, a.tpf > 0 ? (a.M / a.HW) * a.N : (int64_t)a.N, a.K, a.NBp, &tmB);
        if (e != cudaSuccess) return e;
    } else {
        tmB = tmA;
    }
    CUtensorMap tmD = tmA;
    if (tstore) {
        if (a.tpf != 0 || (a.NB % tstore)) return cudaErrorInvalidValue;
        e = make_tmap_out(a.D, a.M, a.N, tstore, &tmD);
        if (e != cudaSuccess) return e;
    }
    return launch_pdl(kernel, dim3(grid), dim3((epi_warps + 1 + prod_warps + xform_warps) * 32), smem, s, a, tmA, tmB, tmD);
}

template <typename T>
static cudaError_t launch_t(GemmArgs& a, int act, cudaStream_t s) {
    const bool gate = a.gate != nullptr, res = a.R != nullptr;
    // <T, GATE, ACT, RES, POOL, epilogue warps, loader warps, transformer warps>
    if (a.feat) return run(gemm_tc_kernel<T, false, 1, false, true, 16, 4, 0>, a, 16, 4, 0, s);
    // gated layers: the transformers' lds -> multiply -> sts -> fence chain per stage is latency-bound (they were busy 80 % of the
    // kernel with 8 warps while the MMA warp waited 60 % for transformed stages): 12 warps in 4 groups of 3 are 6-11 % faster on
    // every gated layer (1152 -> 192 @7x7: 109.5 -> 101.3 us; 672 -> 112 @14x14: 179.2 -> 159.7 us); 16 warps spill (64 registers)
#ifndef DFD_GEMM_XFORM_WARPS
#define DFD_GEMM_XFORM_WARPS 12     // transformer warps / epilogue warps of the gated layers (tools/build_variant.py sweeps them)
#endif
#ifndef DFD_GEMM_GATED_EPI
#define DFD_GEMM_GATED_EPI 8
#endif
    constexpr int XW = DFD_GEMM_XFORM_WARPS, GE = DFD_GEMM_GATED_EPI;
    if (gate && res && !act) return run(gemm_tc_kernel<T, true, 0, true, false, GE, 4, XW>, a, GE, 4, XW, s);
    if (gate && !res && !act) return run(gemm_tc_kernel<T, true, 0, false, false, GE, 4, XW>, a, GE, 4, XW, s);
    if (!gate && !res && act == 1) {
#ifndef DFD_GEMM_TSTORE
#define DFD_GEMM_TSTORE 1      // 0: direct stores everywhere (A/B builds)
#endif
        const int nb = a.N / gemm_n_chunks(a.N);
        if (DFD_GEMM_TSTORE && a.tpf == 0 && (nb & 31) == 0)
            return run(gemm_tc_kernel<T, false, 1, false, false, 16, 4, 0, false, 0, 32>, a, 16, 4, 0, s, 32);
        // (16-column boxes for N = 240 / 480 measured exactly the time of the direct 32-byte stores — 183.3 / 107.5 us: it is the 32-byte
        // row segment, not the instruction that writes it, that is slow — so those layers keep the direct stores)
        return run(gemm_tc_kernel<T, false, 1, false, false, 16, 4, 0>, a, 16, 4, 0, s);
    }
    if (!gate && !res && act == 2) return run(gemm_tc_kernel<T, false, 2, false, false, 16, 4, 0>, a, 16, 4, 0, s);
    if (!gate && !res && act == 3) return run(gemm_tc_kernel<T, false, 3, false, false, 8, 4, 0>, a, 8, 4, 0, s);
    if (!gate && res && act == 3) return run(gemm_tc_kernel<T, false, 3, true, false, 8, 4, 0>, a, 8, 4, 0, s);
    if (!gate && !res && !act) return run(gemm_tc_kernel<T, false, 0, false, false, 8, 4, 0>, a, 8, 4, 0, s);
    if (!gate && res && !act) return run(gemm_tc_kernel<T, false, 0, true, false, 8, 4, 0>, a, 8, 4, 0, s);
    return cudaErrorInvalidValue;
}

cudaError_t launch_gemm_tc(const void* A, const void* W, const float* bias, const float* gate, const void* R,
                           void* D, int64_t M, int K, int N, int HW, int act, int dtype, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if ((K & 7) || (N & 7) || K < 8 || N < 8 || HW <= 0) return cudaErrorInvalidValue;
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = bias; a.gate = gate; a.R = R; a.D = D; a.feat = nullptr;
    a.M = M; a.K = K; a.N = N; a.HW = HW;
    a.rows_per_tile = kBM; a.m_tiles = (M + kBM - 1) / kBM; a.inv_hw = 0.f;
    if (dtype == kDtypeFP16) return launch_t<__half>(a, act, s);
    return launch_t<__nv_bfloat16>(a, act, s);
}

// ---- 3x3 stride-1 pad-1 convolution without a gathered operand (resnet50 bottleneck conv2; DFD_RESNET_IMPLICIT=1) ------
// The producer (conv1, a pointwise GEMM) scatters its rows into the zero-haloed layout described at the kernel template;
// the consumer reads nine shifted boxes of it per 64-channel slice.  Element counts of that buffer:
int64_t conv3x3_padded_rows(int64_t frames, int H, int Wd) { return frames * (int64_t)(H + 2) * (Wd + 2) + 2 * (int64_t)(Wd + 3); }

// D_pad[guard + padded(f,y,x)][N] = relu(A[(f,y,x)][K] * W[N,K]^T + bias); halo and guard rows are never written (the
// caller zeroes the buffer once per geometry).
cudaError_t launch_gemm_tc_padout(const void* A, const void* W, const float* bias, void* Dpad, int64_t frames, int H, int Wd,
                                  int K, int N, int dtype, cudaStream_t s) {
    if (frames <= 0) return cudaSuccess;
    if ((K & 7) || (N & 7) || K < 8 || N < 8 || H <= 0 || Wd <= 0 || frames * (int64_t)(H + 2) * (Wd + 2) > (1ll << 30)) return cudaErrorInvalidValue;
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = bias; a.D = Dpad;
    a.M = frames * H * Wd; a.K = K; a.N = N; a.HW = H * Wd; a.conv_h = H; a.conv_w = Wd;
    a.rows_per_tile = kBM; a.m_tiles = (a.M + kBM - 1) / kBM;
    if (dtype == kDtypeFP16) return run(gemm_tc_kernel<__half, false, 3, false, false, 8, 4, 0, false, 1>, a, 8, 4, 0, s);
    return run(gemm_tc_kernel<__nv_bfloat16, false, 3, false, false, 8, 4, 0, false, 1>, a, 8, 4, 0, s);
}

// D[(f,y,x)][N] = relu(sum over taps (ky,kx) and channels c of Apad[padded(f, y+ky-1, x+kx-1)][c] * W[N][(ky*3+kx)*C + c] + bias)
cudaError_t launch_gemm_tc_conv3x3(const void* Apad, const void* W, const float* bias, void* D, int64_t frames, int H, int Wd,
                                   int C, int N, int dtype, cudaStream_t s) {
    if (frames <= 0) return cudaSuccess;
    if ((C % kKB) || (N & 7) || N < 8 || H <= 0 || Wd <= 0 || frames * (int64_t)(H + 2) * (Wd + 2) > (1ll << 30)) return cudaErrorInvalidValue;
    GemmArgs a{};
    a.A = Apad; a.W = W; a.bias = bias; a.D = D;
    a.M = frames * (int64_t)(H + 2) * (Wd + 2); a.K = 9 * C; a.N = N; a.HW = H * Wd; a.conv_h = H; a.conv_w = Wd; a.conv_cpk = C / kKB;
    a.rows_per_tile = kBM; a.m_tiles = (a.M + kBM - 1) / kBM;
    if (dtype == kDtypeFP16) return run(gemm_tc_kernel<__half, false, 3, false, false, 8, 4, 0, false, 2>, a, 8, 4, 0, s);
    return run(gemm_tc_kernel<__nv_bfloat16, false, 3, false, false, 8, 4, 0, false, 2>, a, 8, 4, 0, s);
}

// Project conv with the SE gate folded into per-frame weights Wf[frame][N][K] (scale_weights kernel): tiles are
// frame-aligned so every tile uses one weight matrix and the A operand needs no transformation.
cudaError_t launch_gemm_tc_framew(const void* A, const void* Wf, const float* bias, const void* R, void* D,
                                  int64_t M, int K, int N, int HW, int dtype, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if ((K & 7) || (N & 7) || K < 8 || N < 8 || HW <= 0 || (M % HW) != 0) return cudaErrorInvalidValue;
    GemmArgs a{};
    a.A = A; a.W = Wf; a.bias = bias; a.gate = nullptr; a.R = R; a.D = D; a.feat = nullptr;
    a.M = M; a.K = K; a.N = N; a.HW = HW;
    a.tpf = (HW + kBM - 1) / kBM; a.w_frame_stride = (int64_t)N * K;
    a.rows_per_tile = kBM; a.m_tiles = (M / HW) * a.tpf; a.inv_hw = 0.f;
    if (dtype == kDtypeFP16) return launch_t<__half>(a, 0, s);
    return launch_t<__nv_bfloat16>(a, 0, s);
}

template <typename T>
__global__ void scale_weights_kernel(const T* __restrict__ W, const float* __restrict__ gate, T* __restrict__ Wf, int N, int K, int Kg, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = 8 consecutive k (16 bytes in, 16 bytes out)
    if (i >= total) return;
    const int k8 = (int)(i % (K / 8));
    const int64_t fn = i / (K / 8);
    const int n = (int)(fn % N);
    const int64_t f = fn / N;
    const uint4 w = __ldg(reinterpret_cast<const uint4*>(W) + (size_t)n * (K / 8) + k8);
    griddep_wait();                                                             // the gate comes from the SE kernel before this one
    const float4* gp = reinterpret_cast<const float4*>(gate + (size_t)f * Kg + (8 * k8) % Kg);      // Kg % 8 == 0: the group does not wrap
    const float4 g0 = __ldg(gp), g1 = __ldg(gp + 1);
    const float2 x0 = Half16<T>::unpack(w.x), x1 = Half16<T>::unpack(w.y), x2 = Half16<T>::unpack(w.z), x3 = Half16<T>::unpack(w.w);
    uint4 o;
    o.x = Half16<T>::pack(x0.x * g0.x, x0.y * g0.y); o.y = Half16<T>::pack(x1.x * g0.z, x1.y * g0.w);
    o.z = Half16<T>::pack(x2.x * g1.x, x2.y * g1.y); o.w = Half16<T>::pack(x3.x * g1.z, x3.y * g1.w);
    reinterpret_cast<uint4*>(Wf)[i] = o;
}
// Wf[f][n][k] = W[n][k] * gate[f][k % Kg]   (Kg < K: pixel-packed block-diagonal weights, the gate repeats per pixel)
cudaError_t launch_scale_weights(const void* W, const float* gate, void* Wf, int64_t frames, int N, int K, int Kg, int dtype, cudaStream_t s) {
    const int64_t total = frames * N * (K / 8);
    if (total <= 0) return cudaSuccess;
    if (Kg <= 0 || (Kg & 7) || (K & 7) || K % Kg) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)((total + 255) / 256);
    if (dtype == kDtypeFP16) return launch_pdl(scale_weights_kernel<__half>, dim3(grid), dim3(256), 0, s, (const __half*)W, gate, (__half*)Wf, N, K, Kg, total);
    return launch_pdl(scale_weights_kernel<__nv_bfloat16>, dim3(grid), dim3(256), 0, s, (const __nv_bfloat16*)W, gate, (__nv_bfloat16*)Wf, N, K, Kg, total);
}

cudaError_t launch_gemm_tc_f32out(const void* A, const void* W, const float* bias, const float* R, float* D,
                                  int64_t M, int K, int N, int dtype, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
    if ((K & 7) || (N & 7) || K < 8 || N < 8) return cudaErrorInvalidValue;
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = bias; a.gate = nullptr; a.R = R; a.D = D; a.feat = nullptr;
    a.M = M; a.K = K; a.N = N; a.HW = 1;
    a.rows_per_tile = kBM; a.m_tiles = (M + kBM - 1) / kBM; a.inv_hw = 0.f;
    // 16 epilogue warps: with 12 k-blocks per tile (ViT attention projection, K = 768) the fp32 residual read + write of an
    // 8-warp epilogue is as long as the main loop (measured on B200: ViT-B/16 at batch 512 28.73 -> 28.57 ms)
    if (R) {
        if (dtype == kDtypeFP16) return run(gemm_tc_kernel<__half, false, 0, true, false, 16, 4, 0, true>, a, 16, 4, 0, s);
        return run(gemm_tc_kernel<__nv_bfloat16, false, 0, true, false, 16, 4, 0, true>, a, 16, 4, 0, s);
    }
    if (dtype == kDtypeFP16) return run(gemm_tc_kernel<__half, false, 0, false, false, 16, 4, 0, true>, a, 16, 4, 0, s);
    return run(gemm_tc_kernel<__nv_bfloat16, false, 0, false, false, 16, 4, 0, true>, a, 16, 4, 0, s);
}

cudaError_t launch_gemm_tc_pool(const void* A, const void* W, const float* bias, float* feat,
                                int64_t M, int K, int N, int HW, int dtype, cudaStream_t s) {
    if (M <= 0) return cudaSuccess;
#ifndef DFD_HEAD_TRANSPOSED
#define DFD_HEAD_TRANSPOSED 1      // 0: always the row-major POOL epilogue below (A/B builds)
#endif
    if (DFD_HEAD_TRANSPOSED && head_pool_tc_supported(M, K, N, HW)) return launch_head_pool_tc(A, W, bias, feat, M, K, N, HW, dtype, s);
    if ((K & 7) || (N & 7) || HW <= 0 || HW > kBM || (M % HW) != 0 || kBM / HW > 4) return cudaErrorInvalidValue;
    GemmArgs a{};
    a.A = A; a.W = W; a.bias = bias; a.gate = nullptr; a.R = nullptr; a.D = nullptr; a.feat = feat;
    a.M = M; a.K = K; a.N = N; a.HW = HW;
    a.rows_per_tile = (kBM / HW) * HW; a.m_tiles = (M + a.rows_per_tile - 1) / a.rows_per_tile; a.inv_hw = 1.0f / (float)HW;
    if (dtype == kDtypeFP16) return launch_t<__half>(a, 1, s);
    return launch_t<__nv_bfloat16>(a, 1, s);
}

}  // namespace dfd
