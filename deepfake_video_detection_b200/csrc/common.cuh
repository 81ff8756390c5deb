// Shared device helpers for the sm_100a kernels: 16-bit storage traits, SiLU, vector ld/st,
// mbarrier / tcgen05 / TMEM PTX wrappers.  Everything here is inline device code.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cstdint>
#include <cstdio>
#include <utility>

namespace dfd {

// ------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of the scoring step is launched with the programmatic-stream-serialisation
// attribute (`launch_pdl`), so the 68 dependent launches of a step overlap their launch latency and constant-only prologues
// (barrier init, TMEM allocation, weight staging, shared-memory zeroing) with the tail of the kernel before them — in a plain
// stream and, as programmatic edges, in a captured CUDA graph.  The contract every kernel keeps:
//   * `griddep_wait()` is executed by every CTA BEFORE its first read of anything another kernel produced and BEFORE its first
//     global write (the buffers of the step are reused, so a write may race with the previous kernel's reads otherwise); it returns
//     when the preceding kernel has completed and its writes are visible.  Completion is transitive: the preceding kernel itself
//     waited for its predecessor.
//   * no kernel triggers its dependents early (`griddepcontrol.launch_dependents`): the next kernel's CTAs become resident as this
//     kernel's CTAs exit.  An explicit trigger at the top of every kernel measured the same at 2048 frames and SLOWER for small
//     batches (8 frames: 0.95 vs 0.84 ms in a stream, 0.93 vs 0.84 ms under graph replay) — profiles/r02_experimental.md.
// DFD_PDL=0 builds plain launches (the instructions are no-ops then): `tools/build_variant.py` uses it for A/B timing.
// ------------------------------------------------------------------------------------------
#ifndef DFD_PDL
#define DFD_PDL 1
#endif
__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = DFD_PDL ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

// ------------------------------------------------------------------------------------------
// 16-bit storage types.  Activations and GEMM operands are stored as fp16 (default) or bf16;
// all accumulation is fp32.  dtype codes match include/dfd_b200.h (DFD_DTYPE_*).
// ------------------------------------------------------------------------------------------
enum : int { kDtypeBF16 = 0, kDtypeFP16 = 1 };

template <typename T> struct Half16;
template <> struct Half16<__half> {
    using T2 = __half2;
    static constexpr int kCode = kDtypeFP16;
    static constexpr uint32_t kUmmaFormat = 0;   // F16F32Format::F16
    __device__ __forceinline__ static float2 unpack(uint32_t v) {
        return __half22float2(*reinterpret_cast<const __half2*>(&v));
    }
    __device__ __forceinline__ static uint32_t pack(float a, float b) {
        __half2 h = __floats2half2_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static float to_float(__half v) { return __half2float(v); }
    __device__ __forceinline__ static __half from_float(float v) { return __float2half_rn(v); }
    // d = a*b + c with a, b 16-bit and fp32 accumulate (single FHFMA, no unpack)
    __device__ __forceinline__ static float fma_mixed(uint16_t a, uint16_t b, float c) {
        float d;
        asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
        return d;
    }
};
template <> struct Half16<__nv_bfloat16> {
    using T2 = __nv_bfloat162;
    static constexpr int kCode = kDtypeBF16;
    static constexpr uint32_t kUmmaFormat = 1;   // F16F32Format::BF16
    __device__ __forceinline__ static float2 unpack(uint32_t v) {
        return make_float2(__uint_as_float(v << 16), __uint_as_float(v & 0xffff0000u));
    }
    __device__ __forceinline__ static uint32_t pack(float a, float b) {
        __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    __device__ __forceinline__ static float to_float(__nv_bfloat16 v) { return __bfloat162float(v); }
    __device__ __forceinline__ static __nv_bfloat16 from_float(float v) { return __float2bfloat16_rn(v); }
    __device__ __forceinline__ static float fma_mixed(uint16_t a, uint16_t b, float c) {
        float d;
        asm("fma.rn.f32.bf16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
        return d;
    }
};

// ImageNet statistics of app.py:1774-1775 and the reference's exact fp32 prep arithmetic
// (`u8/255` then `(x-mean)/std`, IEEE divides).
__device__ __forceinline__ float imagenet_mean(int c) { return c == 0 ? 0.485f : (c == 1 ? 0.456f : 0.406f); }
__device__ __forceinline__ float imagenet_std(int c) { return c == 0 ? 0.229f : (c == 1 ? 0.224f : 0.225f); }
__device__ __forceinline__ float prep_value(int c, int v) {
    const float x = __fdiv_rn((float)v, 255.0f);
    return __fdiv_rn(__fsub_rn(x, imagenet_mean(c)), imagenet_std(c));
}

// x * sigmoid(x), fp32.  ex2.approx (2 ulp) + rcp.approx (1 ulp): well below 16-bit storage error.
__device__ __forceinline__ float silu_f(float x) {
    return __fdividef(x, 1.0f + __expf(-x));
}
__device__ __forceinline__ float sigmoid_f(float x) {
    return __fdividef(1.0f, 1.0f + __expf(-x));
}

// Packed fp32x2 arithmetic (sm_100: the full fp32 rate needs the .f32x2 forms) and a SiLU for register
// pairs that spends ONE MUFU per element (ex2) instead of two: the reciprocal of 1+e^-x is a bit-trick seed
// refined by three Newton steps on the FMA pipe (relative error ~1e-7).  Used where MUFU, not FMA, is the
// scarce pipe (GEMM epilogues).  The result is returned NEGATED (-x*sigmoid(x)); callers flip the sign bits
// of the packed 16-bit pair for free.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) { float2 v = make_float2(a, b); return *reinterpret_cast<uint64_t*>(&v); }
__device__ __forceinline__ float2 f2_unpack(uint64_t v) { return *reinterpret_cast<float2*>(&v); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) {
    uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d;
}
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#ifndef DFD_SILU_NR
#define DFD_SILU_NR 0
#endif
#ifndef DFD_SILU_TANH
#define DFD_SILU_TANH 1
#endif
// x*sigmoid(x) = h + h*tanh(h), h = x/2: ONE MUFU (tanh.approx.f32, rel. error 2^-11) and two FMA-pipe ops.
__device__ __forceinline__ float tanh_approx(float x) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(x)); return t; }
__device__ __forceinline__ float silu_tanh(float x) {
    const float h = 0.5f * x;
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
    return fmaf(h, t, h);
}
__device__ __forceinline__ uint64_t neg_silu2(uint64_t x) {
#if DFD_SILU_TANH
    {
        const float2 xf = f2_unpack(x);
        return f2_pack(-silu_tanh(xf.x), -silu_tanh(xf.y));
    }
#endif
    const float2 t = f2_unpack(mul2(x, f2_pack(-1.4426950408889634f, -1.4426950408889634f)));
#if !DFD_SILU_NR
    {   // two-MUFU form: ex2 + rcp.approx (fewer issue slots, twice the XU work)
        float r0, r1;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(1.0f + ex2_approx(fminf(t.x, 126.f))));
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(1.0f + ex2_approx(fminf(t.y, 126.f))));
        return mul2(x, f2_pack(-r0, -r1));
    }
#endif
    const uint64_t e = f2_pack(ex2_approx(fminf(t.x, 126.f)), ex2_approx(fminf(t.y, 126.f)));
    const uint64_t d = add2(e, f2_pack(1.f, 1.f));                       // 1 + e^-x  in [1, 2^126]
    const float2 df = f2_unpack(d);
    // s = -(1/d): seed -(magic - bits(d)), then s <- s * (2 + d*s)  (Newton on the negated reciprocal)
    uint64_t s = f2_pack(__int_as_float((0x7EF311C7 - __float_as_int(df.x)) | 0x80000000),
                         __int_as_float((0x7EF311C7 - __float_as_int(df.y)) | 0x80000000));
    const uint64_t two = f2_pack(2.f, 2.f);
    s = mul2(s, fma2(d, s, two));
    s = mul2(s, fma2(d, s, two));
    s = mul2(s, fma2(d, s, two));
    return mul2(x, s);                                                   // = -x * sigmoid(x)
}

// ------------------------------------------------------------------------------------------
// vector global access
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg16(const void* p) {      // read-only 128-bit
    return __ldg(reinterpret_cast<const uint4*>(p));
}
__device__ __forceinline__ uint4 ldg16_stream(const void* p) {   // streaming: do not allocate in L1
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg16(void* p, const uint4& v) {
    *reinterpret_cast<uint4*>(p) = v;
}
struct alignas(32) U32x8 { uint32_t v[8]; };
__device__ __forceinline__ void stg32(void* p, const U32x8& r) {   // 256-bit store (sm_100+)
    asm volatile("st.global.v8.b32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};"
                 :: "r"(r.v[0]), "r"(r.v[1]), "r"(r.v[2]), "r"(r.v[3]),
                    "r"(r.v[4]), "r"(r.v[5]), "r"(r.v[6]), "r"(r.v[7]), "l"(p) : "memory");
}
__device__ __forceinline__ U32x8 ldg32(const void* p) {            // 256-bit read-only load (sm_100+)
    U32x8 r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]),
                   "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p));
    return r;
}

__device__ __forceinline__ U32x8 ldg32_coherent(const void* p) {   // 256-bit load of data this kernel also writes (in-place residual)
    U32x8 r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]),
                   "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7]) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ uint4 ldg16_coherent(const void* p) {   // 128-bit load of data this kernel also writes
    uint4 r;
    asm volatile("ld.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p) : "memory");
    return r;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void sts16(uint32_t addr, const uint4& v) {
    asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__device__ __forceinline__ uint4 lds16(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
// 16-byte asynchronous global -> shared copy (LDGSTS), L1 bypass; `valid == false` writes 16 zero bytes
// without touching global memory (src must still be a mapped address).
__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* gptr, bool valid) {
    const uint32_t n = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(saddr), "l"(gptr), "r"(n) : "memory");
}
// Make the mbarrier track this thread's prior cp.async copies: its pending count is incremented now and
// decremented when they have all landed (pair it with a normal arrive; CUTLASS cpasync_barrier_arrive).
__device__ __forceinline__ void cp_async_mbar_arrive(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
}
// Same, but the completion counts as one of the barrier's EXPECTED arrivals (no pending-count increment): one
// mbarrier operation per thread and stage instead of three (increment, completion, plain arrive).
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint32_t bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

// ------------------------------------------------------------------------------------------
// mbarrier (shared::cta)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" :: "r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#ifndef DFD_MBAR_TEST_WAIT
#define DFD_MBAR_TEST_WAIT 0
#endif
#if DFD_MBAR_TEST_WAIT
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#else
    asm volatile("{\n\t.reg .pred p;\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch fails with an error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        if (mbar_try_wait(bar, parity)) return;
    }
    printf("dfd: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
// make generic-proxy shared-memory writes visible to the async proxy (UMMA operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {   // one full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // same warp as alloc
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync()  { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem] · B[smem]^T, 16-bit inputs, fp32 accumulate, issued by ONE thread.
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on an mbarrier when all previously issued tcgen05.mma of this thread have completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" :: "r"(bar) : "memory");
}
// TMEM → registers: this thread's lane, 16 consecutive fp32 columns
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): core matrix = 8 rows x 16 B,
// rows 16 B apart; 8-row groups `sbo` bytes apart; the two 16-byte K chunks of one MMA `lbo` bytes apart.
// Bit layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor), version field = 1 on sm_100.
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)((lbo >> 4) & 0x3fffu) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3fffu) << 32;
    d |= (uint64_t)1 << 46;
    return d;   // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// Same for the 128-byte-swizzled K-major layout TMA produces (box = 64 elements x rows, CU_TENSOR_MAP_SWIZZLE_128B):
// rows 128 B apart, 8-row groups 1024 B apart, 16-byte chunk index XORed with (row & 7).  A 16-element K step
// advances the start address by 32 bytes inside the swizzle atom.  The tile base must be 1024-byte aligned.
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);
    d |= (uint64_t)1 << 16;                        // LBO: unused for swizzled K-major
    d |= (uint64_t)(1024 >> 4) << 32;              // SBO
    d |= (uint64_t)1 << 46;                        // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
    return d;
}
// TMA: 2-D tiled bulk copy global -> shared, completion counted in bytes on an mbarrier (issued by ONE thread)
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar), "r"(bytes) : "memory");
}
// whole-warp forms (`elect.sync` picks the issuing lane; operands stay warp-uniform): the producer warp of a pipeline
__device__ __forceinline__ void tma_load_2d_elect(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
                 :: "r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx_elect(uint32_t bar, uint32_t bytes) {
    asm volatile("{\n\t.reg .pred q;\n\t.reg .b64 st;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" :: "l"(tmap) : "memory");
}
// Instruction descriptor for kind::f16: fp32 accumulate, K-major A and B, M x N tile.
__device__ __host__ constexpr uint32_t umma_idesc(uint32_t ab_format, uint32_t m, uint32_t n) {
    return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ------------------------------------------------------------------------------------------
// CTA pairs (thread-block clusters of 2): tcgen05 cta_group::2, cluster-scope mbarrier traffic, TMA stores
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
// every thread of every CTA in the cluster
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory location in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t saddr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" :: "r"(cluster_addr) : "memory");
}
// wait on a local barrier whose arrivals may come from the peer CTA
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\t"
                     "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
                     "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
    }
    printf("dfd: cluster mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {   // the same warp of BOTH CTAs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_dst), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(taddr), "r"(ncols) : "memory");
}
// D[tmem of both CTAs] (+)= A[128 rows in each CTA's smem] · B[N/2 rows in each CTA's smem]^T; issued by ONE thread of the leader CTA
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on the barrier at the same shared-memory offset in BOTH CTAs when the pair's MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
// The same, executed by a WHOLE converged warp: `elect.sync` picks the issuing lane inside the instruction's predicate, so control
// flow stays warp-uniform and the compiler keeps descriptors / addresses in uniform registers.  (Behind `if (lane == 0)` every
// operand of every UTCHMMA goes through an ELECT + 5 x R2UR + branch sequence: ~130 dependent instructions = ~620 cycles per 64-wide
// k-block in the issuing warp, more than the 512 cycles the four MMAs of a 256 x 256 x 64 block take.)
__device__ __forceinline__ void umma_f16_pair_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_pair_elect(uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t}"
                 :: "r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_f16_elect(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit_elect(uint32_t bar) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" :: "r"(bar) : "memory");
}
// TMA load issued by either CTA of a pair: data lands in the issuing CTA's shared memory, bytes are counted on `bar_cluster`
// (a shared::cluster address: the leader's stage barrier)
__device__ __forceinline__ void tma_load_2d_pair(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar_cluster) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 :: "r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair_elect(uint32_t smem_dst, const void* tmap, int c0, int c1, uint32_t bar_cluster) {
    asm volatile("{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
                 "@q cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
                 :: "r"(smem_dst), "l"(tmap), "r"(c0), "r"(c1), "r"(bar_cluster) : "memory");
}
// TMA store shared -> global (rows / columns outside the tensor are dropped), bulk-group completion
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tmap), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
// TMA reduction shared -> global: global[box] += shared[box] (fp32 add performed at the L2; rows / columns outside the tensor are dropped)
__device__ __forceinline__ void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tmap), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory"); }

// exact GELU x * Phi(x) without erff: erfc(t / sqrt 2) = 2^-q(t) with a degree-6 polynomial q fitted on [0, 6] (beyond 6 the
// term is below 1e-8), one MUFU (ex2).  Max abs error 5e-7, max error relative to max(|gelu|, 1e-2) 6e-6 (fit and check:
// tools/fit_gelu.py) — far below the 16-bit rounding of the output.
__device__ __forceinline__ float gelu_erfc_poly(float x) {
    const float ax = fabsf(x), t = fminf(ax, 6.0f);
    float q = -3.3095075e-05f;
    q = fmaf(q, t, 0.00076925056f); q = fmaf(q, t, -0.008080854f); q = fmaf(q, t, 0.053412355f);
    q = fmaf(q, t, 0.45877084f); q = fmaf(q, t, 1.1512016f); q = fmaf(q, t, -6.878746e-06f);
    return fmaf(-0.5f * ax, ex2_approx(-q), fmaxf(x, 0.f));
}
// the same for a register pair on the packed fp32x2 pipe (the polynomial is 6 of the 11 instructions per element)
__device__ __forceinline__ uint64_t gelu_erfc_poly2(uint64_t x) {
    const float2 xf = f2_unpack(x);
    const uint64_t t = f2_pack(fminf(fabsf(xf.x), 6.0f), fminf(fabsf(xf.y), 6.0f));
    uint64_t q = fma2(f2_pack(-3.3095075e-05f, -3.3095075e-05f), t, f2_pack(0.00076925056f, 0.00076925056f));
    q = fma2(q, t, f2_pack(-0.008080854f, -0.008080854f)); q = fma2(q, t, f2_pack(0.053412355f, 0.053412355f));
    q = fma2(q, t, f2_pack(0.45877084f, 0.45877084f)); q = fma2(q, t, f2_pack(1.1512016f, 1.1512016f));
    q = fma2(q, t, f2_pack(-6.878746e-06f, -6.878746e-06f));
    const float2 qf = f2_unpack(q);
    return fma2(f2_pack(-0.5f * fabsf(xf.x), -0.5f * fabsf(xf.y)), f2_pack(ex2_approx(-qf.x), ex2_approx(-qf.y)),
                f2_pack(fmaxf(xf.x, 0.f), fmaxf(xf.y, 0.f)));
}

}  // namespace dfd
