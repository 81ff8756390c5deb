// K6 — ViT-B/16 frame encoder (BASELINE config 5): reference src/models.py:88-107 `ViTFeatureExtractor`, i.e.
// timm `vit_base_patch16_224(num_classes=0)` (un-vendored dependency, requirements.txt:12): x (B,3,224,224) fp32 ->
// CLS feature (B,768) fp32.
//
// Every dense contraction (patch embedding, qkv, attention projection, the two MLP layers: 96 % of the 35 GFLOP per
// image) runs on the CTA-pair tcgen05/TMEM GEMM of gemm_pair.cu with 16-bit operands and fp32 accumulation:
//   patchify (fp32 NCHW -> 16-bit [B*196, 768], column = c*256 + ky*16 + kx, the Conv2d weight order)
//   -> GEMM(+bias) -> assemble tokens: [cls | patches] + pos_embed -> residual stream X fp32 [B*197, 768]
//   12 x { X += fc2 output of the previous block ; LN1(X) -> 16-bit ; qkv GEMM ; attention (softmax(QK^T/8)V per image and
//          head on tcgen05, vit_attn_tc.cu) ; proj GEMM -> 16-bit ; X += proj output ; LN2(X) ; fc1 GEMM + exact GELU ; fc2 GEMM }
//   X[cls] += last fc2 output ; final LayerNorm on the CLS rows only -> fp32 features.
// The residual stream stays fp32 end to end; the branch outputs are rounded to 16 bits once (GEMM output) and added to it
// inside the LayerNorm kernel that needs the sum anyway (one read-modify-write of X per branch instead of a second one in a
// GEMM epilogue).  LayerNorm statistics and the softmax are fp32.  Every reduction has a fixed order: results do not
// depend on the batch.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dfd_b200.h"
#include "common.cuh"
#include "kernels.h"

#ifndef DFD_VIT_RES_IN_GEMM
#define DFD_VIT_RES_IN_GEMM 1      // 1: proj / fc2 add their output to the fp32 residual stream in the GEMM epilogue (HBM traffic in the
#endif                             //    shadow of a tensor-bound kernel); 0: 16-bit branch outputs, added inside the next LayerNorm kernel
                                   //    (measured on B200 at batch 512: 22.42 vs 22.67 ms per forward)
namespace {
constexpr int kDim = 768, kDepth = 12, kPatch = 16, kImg = 224, kGridP = 14;
constexpr int kPatches = kGridP * kGridP, kTokens = kPatches + 1, kMlp = 3072, kPatchK = 3 * kPatch * kPatch;
}

struct dfd_vit_weights {
    int dtype;
    void* patch_w; float* patch_b;            // [768][768] 16-bit, [768]
    float *cls, *pos;                          // [768], [197][768]
    struct Block {
        float *ln1_w, *ln1_b, *ln2_w, *ln2_b;
        void *qkv_w, *proj_w, *fc1_w, *fc2_w;  // 16-bit [N][K] (K-major, as nn.Linear stores them)
        float *qkv_b, *proj_b, *fc1_b, *fc2_b;
    } blk[kDepth];
    float *norm_w, *norm_b;
    void* arena;
};

namespace dfd {

// x fp32 (B,3,224,224) -> A [B*196][768] 16-bit: row = image*196 + py*14 + px, column = c*256 + ky*16 + kx
template <typename T>
__global__ void vit_patchify_kernel(const float* __restrict__ x, T* __restrict__ a, int64_t total8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = 8 consecutive kx
    if (i >= total8) return;
    const int kx0 = (int)(i & 1) * 8;
    const int ky = (int)((i >> 1) & 15);
    const int c = (int)((i >> 5) % 3);
    const int64_t row = i / 96;
    const int px = (int)(row % kGridP), py = (int)((row / kGridP) % kGridP);
    const int64_t b = row / kPatches;
    const float* src = x + (((size_t)b * 3 + c) * kImg + (py * kPatch + ky)) * kImg + px * kPatch + kx0;
    const float4 v0 = __ldg(reinterpret_cast<const float4*>(src)), v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
    uint4 o;
    o.x = Half16<T>::pack(v0.x, v0.y); o.y = Half16<T>::pack(v0.z, v0.w);
    o.z = Half16<T>::pack(v1.x, v1.y); o.w = Half16<T>::pack(v1.z, v1.w);
    *reinterpret_cast<uint4*>(a + (size_t)row * kPatchK + c * 256 + ky * 16 + kx0) = o;
}

// X[b][0] = cls + pos[0];  X[b][1+i] = P[b*196+i] + pos[1+i]   (P already carries the conv bias)
template <typename T>
__global__ void vit_assemble_kernel(const T* __restrict__ p, const float* __restrict__ cls, const float* __restrict__ pos,
                                    float* __restrict__ x, int64_t total8) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;          // one thread = 8 channels
    if (i >= total8) return;
    const int c8 = (int)(i % (kDim / 8)) * 8;
    const int64_t row = i / (kDim / 8);
    const int tok = (int)(row % kTokens);
    const int64_t b = row / kTokens;
    float v[8];
    if (tok == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = cls[c8 + j];
    } else {
        const uint4 r = *reinterpret_cast<const uint4*>(p + ((size_t)b * kPatches + tok - 1) * kDim + c8);
        const float2 a0 = Half16<T>::unpack(r.x), a1 = Half16<T>::unpack(r.y), a2 = Half16<T>::unpack(r.z), a3 = Half16<T>::unpack(r.w);
        v[0] = a0.x; v[1] = a0.y; v[2] = a1.x; v[3] = a1.y; v[4] = a2.x; v[5] = a2.y; v[6] = a3.x; v[7] = a3.y;
    }
    const float* ps = pos + (size_t)tok * kDim + c8;
    float* dst = x + (size_t)row * kDim + c8;
    *reinterpret_cast<float4*>(dst) = make_float4(v[0] + ps[0], v[1] + ps[1], v[2] + ps[2], v[3] + ps[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(v[4] + ps[4], v[5] + ps[5], v[6] + ps[6], v[7] + ps[7]);
}

// LayerNorm over 768 channels, eps 1e-6, one warp per row: fp32 statistics (mean, then centred variance).
// Row r of the residual stream lives at x + r * in_stride floats.  With `delta` (the 16-bit output of the previous branch,
// row r at delta + r * in_stride elements) the row becomes x + delta first and is written back (the residual add).
// OUT = 16-bit GEMM operand or fp32 (final norm).
template <typename OUT, typename DT>
__global__ void vit_layernorm_kernel(float* __restrict__ x, int64_t in_stride, const DT* __restrict__ delta, const float* __restrict__ w,
                                     const float* __restrict__ b, OUT* __restrict__ y, int64_t rows) {
    const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (r >= rows) return;
    const int lane = threadIdx.x & 31;
    float4* src = reinterpret_cast<float4*>(x + (size_t)r * in_stride);
    float4 v[6];
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) v[j] = src[lane + 32 * j];
    if (delta) {
        const uint2* dp = reinterpret_cast<const uint2*>(delta + (size_t)r * in_stride);
#pragma unroll
        for (int j = 0; j < 6; ++j) {
            const uint2 d = __ldg(dp + lane + 32 * j);
            const float2 d0 = Half16<DT>::unpack(d.x), d1 = Half16<DT>::unpack(d.y);
            v[j].x += d0.x; v[j].y += d0.y; v[j].z += d1.x; v[j].w += d1.y;
            src[lane + 32 * j] = v[j];
        }
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * (1.0f / kDim);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const float a = v[j].x - mean, c = v[j].y - mean, d = v[j].z - mean, e = v[j].w - mean;
        q += (a * a + c * c) + (d * d + e * e);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * (1.0f / kDim) + 1e-6f);
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        const int c0 = (lane + 32 * j) * 4;
        const float4 g = *reinterpret_cast<const float4*>(w + c0), be = *reinterpret_cast<const float4*>(b + c0);
        const float o0 = (v[j].x - mean) * rstd * g.x + be.x, o1 = (v[j].y - mean) * rstd * g.y + be.y;
        const float o2 = (v[j].z - mean) * rstd * g.z + be.z, o3 = (v[j].w - mean) * rstd * g.w + be.w;
        if constexpr (sizeof(OUT) == 4) {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(y) + (size_t)r * kDim + c0) = make_float4(o0, o1, o2, o3);
        } else {
            uint2 o; o.x = Half16<OUT>::pack(o0, o1); o.y = Half16<OUT>::pack(o2, o3);
            *reinterpret_cast<uint2*>(y + (size_t)r * kDim + c0) = o;
        }
    }
}

// ---- attention: csrc/vit_attn_tc.cu (tcgen05 / TMEM; the mma.sync + ldmatrix kernel it replaced measured 152 TFLOP/s against
// 270 TFLOP/s at batch 512 on B200 and was removed) ----------------------------------------------------------------------

// ---- DeepfakeModel head (src/models.py:199-291): SimpleGCN over the frame graph + mean pool + classifier ------------
//   g = relu(fc2(relu(fc1(A_norm @ H))));  logits = Linear(64 -> C)(relu(Linear(128 -> 64)(mean_n g)))      (:186-197, :283-291)
// One CTA per video, fp32.  fc1(A @ H) is evaluated as A @ (H W1^T) + b1 (same function, N x 256 intermediate instead
// of N x 768).  Every sum runs in a fixed order.  Weights are packed transposed so that threads read consecutive words.
constexpr int kGcnF = 768, kGcnHid = 256, kGcnOut = 128, kGcnC1 = 64, kGcnMaxNodes = 64, kGcnChunk = 8;
struct GcnW { const float *w1t, *b1, *w2t, *b2, *c1t, *cb1, *c2, *cb2; int classes; };

__global__ void __launch_bounds__(256) gcn_head_kernel(GcnW w, const float* __restrict__ feats, const float* __restrict__ adj,
                                                       int N, float* __restrict__ logits) {
    extern __shared__ float gs[];
    float* sT = gs;                          // [N][256]  H W1^T, later g2 [N][128]
    float* sG = sT + N * kGcnHid;            // [N][256]  relu(A T + b1)
    float* sA = sG + N * kGcnHid;            // [N][N]
    float* sH = sA + N * N;                  // [8][768] chunk of the node features; later pooled[128] + c1[64]
    const int b = blockIdx.x, t = threadIdx.x;
    const float* H = feats + (size_t)b * N * kGcnF;
    for (int i = t; i < N * N; i += 256) sA[i] = adj[(size_t)b * N * N + i];
    for (int n0 = 0; n0 < N; n0 += kGcnChunk) {
        const int nn = min(kGcnChunk, N - n0);
        __syncthreads();
        for (int i = t; i < nn * kGcnF; i += 256) sH[i] = H[(size_t)n0 * kGcnF + i];
        __syncthreads();
        float acc[kGcnChunk];
#pragma unroll
        for (int n = 0; n < kGcnChunk; ++n) acc[n] = 0.f;
        for (int k = 0; k < kGcnF; ++k) {
            const float wv = __ldg(w.w1t + (size_t)k * kGcnHid + t);
#pragma unroll
            for (int n = 0; n < kGcnChunk; ++n) acc[n] = fmaf(sH[n * kGcnF + k], wv, acc[n]);     // rows >= nn hold stale data: unused
        }
#pragma unroll
        for (int n = 0; n < kGcnChunk; ++n) if (n < nn) sT[(n0 + n) * kGcnHid + t] = acc[n];
    }
    __syncthreads();
    {
        const float bj = __ldg(w.b1 + t);
        for (int n = 0; n < N; ++n) {
            float a = 0.f;
            for (int m = 0; m < N; ++m) a = fmaf(sA[n * N + m], sT[m * kGcnHid + t], a);
            sG[n * kGcnHid + t] = fmaxf(a + bj, 0.f);
        }
    }
    __syncthreads();
    {
        const int o = t & (kGcnOut - 1), half = t >> 7;
        const float bo = __ldg(w.b2 + o);
        for (int n = half; n < N; n += 2) {
            float a = 0.f;
            for (int j = 0; j < kGcnHid; ++j) a = fmaf(sG[n * kGcnHid + j], __ldg(w.w2t + (size_t)j * kGcnOut + o), a);
            sT[n * kGcnOut + o] = fmaxf(a + bo, 0.f);
        }
    }
    __syncthreads();
    float* pooled = sH; float* c1 = sH + kGcnOut;
    if (t < kGcnOut) {
        float a = 0.f;
        for (int n = 0; n < N; ++n) a += sT[n * kGcnOut + t];
        pooled[t] = a / (float)N;
    }
    __syncthreads();
    if (t < kGcnC1) {
        float a = __ldg(w.cb1 + t);
        for (int o = 0; o < kGcnOut; ++o) a = fmaf(pooled[o], __ldg(w.c1t + (size_t)o * kGcnC1 + t), a);
        c1[t] = fmaxf(a, 0.f);
    }
    __syncthreads();
    if (t < w.classes) {
        float a = __ldg(w.cb2 + t);
        for (int i = 0; i < kGcnC1; ++i) a = fmaf(c1[i], __ldg(w.c2 + (size_t)t * kGcnC1 + i), a);
        logits[(size_t)b * w.classes + t] = a;
    }
}

}  // namespace dfd

struct dfd_gcn_weights { dfd::GcnW w; void* arena; };

namespace {
thread_local std::string g_vit_err;
int vfail(int code, const std::string& m) { g_vit_err = m; return code; }
uint16_t vh16(float v, int dtype) {
    if (dtype == DFD_DTYPE_FP16) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
    __nv_bfloat16 h = __float2bfloat16_rn(v); uint16_t u; memcpy(&u, &h, 2); return u;
}
size_t vup(size_t b) { return (b + 255) & ~size_t(255); }
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* dfd_vit_last_error(void) { return g_vit_err.c_str(); }

int dfd_vit_pack_weights(int n, const char* const* names, const float* const* data, const int64_t* numel,
                         int dtype, dfd_vit_weights_t** out) {
    if (!names || !data || !numel || !out || n <= 0) return vfail(DFD_EINVAL, "dfd_vit_pack_weights: null argument");
    if (dtype != DFD_DTYPE_BF16 && dtype != DFD_DTYPE_FP16) return vfail(DFD_EINVAL, "dfd_vit_pack_weights: unknown dtype");
    std::unordered_map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n; ++i) {
        if (!names[i]) continue;
        std::string k = names[i];
        if (k.rfind("vit.", 0) == 0) k = k.substr(4);             // the reference keeps the timm model under `.vit` (models.py:93)
        t[k] = {data[i], numel[i]};
    }
    std::string missing;
    auto get = [&](const std::string& k, int64_t ne) -> const float* {
        auto it = t.find(k);
        if (it == t.end() || it->second.second != ne) { if (missing.empty()) missing = k; return nullptr; }
        return it->second.first;
    };
    std::vector<uint8_t> host;
    auto alloc = [&](size_t nb) { size_t o = (host.size() + 255) & ~size_t(255); host.resize(o + nb, 0); return o; };
    auto put16 = [&](const std::string& k, int64_t ne) {
        const size_t off = alloc((size_t)ne * 2);
        const float* p = get(k, ne);
        if (p) { uint16_t* d = reinterpret_cast<uint16_t*>(host.data() + off); for (int64_t i = 0; i < ne; ++i) d[i] = vh16(p[i], dtype); }
        return off;
    };
    auto put32 = [&](const std::string& k, int64_t ne) {
        const size_t off = alloc((size_t)ne * 4);
        const float* p = get(k, ne);
        if (p) memcpy(host.data() + off, p, (size_t)ne * 4);
        return off;
    };
    size_t o_patch_w = put16("patch_embed.proj.weight", (int64_t)kDim * kPatchK), o_patch_b = put32("patch_embed.proj.bias", kDim);
    size_t o_cls = put32("cls_token", kDim), o_pos = put32("pos_embed", (int64_t)kTokens * kDim);
    size_t ob[kDepth][12];
    for (int i = 0; i < kDepth; ++i) {
        const std::string p = "blocks." + std::to_string(i) + ".";
        ob[i][0] = put32(p + "norm1.weight", kDim); ob[i][1] = put32(p + "norm1.bias", kDim);
        ob[i][2] = put32(p + "norm2.weight", kDim); ob[i][3] = put32(p + "norm2.bias", kDim);
        ob[i][4] = put16(p + "attn.qkv.weight", (int64_t)3 * kDim * kDim); ob[i][5] = put32(p + "attn.qkv.bias", 3 * kDim);
        ob[i][6] = put16(p + "attn.proj.weight", (int64_t)kDim * kDim);    ob[i][7] = put32(p + "attn.proj.bias", kDim);
        ob[i][8] = put16(p + "mlp.fc1.weight", (int64_t)kMlp * kDim);      ob[i][9] = put32(p + "mlp.fc1.bias", kMlp);
        ob[i][10] = put16(p + "mlp.fc2.weight", (int64_t)kDim * kMlp);     ob[i][11] = put32(p + "mlp.fc2.bias", kDim);
    }
    size_t o_nw = put32("norm.weight", kDim), o_nb = put32("norm.bias", kDim);
    if (!missing.empty()) return vfail(DFD_EKEY, "dfd_vit_pack_weights: state_dict tensor " + missing + " absent or wrong size");
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size()) != cudaSuccess) return vfail(DFD_ECUDA, "cudaMalloc(vit weights) failed");
    if (cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return vfail(DFD_ECUDA, "cudaMemcpy(vit weights) failed"); }
    auto* W = new dfd_vit_weights();
    uint8_t* d = reinterpret_cast<uint8_t*>(dev);
    auto F = [&](size_t off) { return reinterpret_cast<float*>(d + off); };
    W->dtype = dtype; W->arena = dev;
    W->patch_w = d + o_patch_w; W->patch_b = F(o_patch_b); W->cls = F(o_cls); W->pos = F(o_pos);
    for (int i = 0; i < kDepth; ++i) {
        auto& b = W->blk[i];
        b.ln1_w = F(ob[i][0]); b.ln1_b = F(ob[i][1]); b.ln2_w = F(ob[i][2]); b.ln2_b = F(ob[i][3]);
        b.qkv_w = d + ob[i][4]; b.qkv_b = F(ob[i][5]); b.proj_w = d + ob[i][6]; b.proj_b = F(ob[i][7]);
        b.fc1_w = d + ob[i][8]; b.fc1_b = F(ob[i][9]); b.fc2_w = d + ob[i][10]; b.fc2_b = F(ob[i][11]);
    }
    W->norm_w = F(o_nw); W->norm_b = F(o_nb);
    *out = W;
    return DFD_OK;
}

void dfd_vit_free_weights(dfd_vit_weights_t* w) { if (w) { if (w->arena) cudaFree(w->arena); delete w; } }

int dfd_vit_workspace_bytes(int64_t images, size_t* bytes) {
    if (!bytes || images <= 0) return vfail(DFD_EINVAL, "dfd_vit_workspace_bytes: bad argument");
    const size_t M = (size_t)images * kTokens;
    *bytes = vup(M * kDim * 4) + (DFD_VIT_RES_IN_GEMM ? 1 : 2) * vup(M * kDim * 2) + vup(M * kMlp * 2) + 1024;
    return DFD_OK;
}

// kernel-level entry (include/dfd_b200_kernels.h): the attention of one ViT block (the kernel the encoder runs)
int dfd_k_vit_attention(const void* d_qkv, void* d_o, int64_t images, int dtype, void* stream) {
    if (!d_qkv || !d_o || images <= 0) return vfail(DFD_EINVAL, "dfd_k_vit_attention: bad argument");
    if (dtype != DFD_DTYPE_FP16 && dtype != DFD_DTYPE_BF16) return vfail(DFD_EINVAL, "dfd_k_vit_attention: unknown dtype");
    cudaError_t e = dfd::launch_vit_attention_tc(d_qkv, d_o, images, dtype, (cudaStream_t)stream);
    if (e != cudaSuccess) return vfail(DFD_ECUDA, std::string("vit attention: ") + cudaGetErrorString(e));
    return DFD_OK;
}

int dfd_vit_features(const dfd_vit_weights_t* w, const float* d_in, int64_t images, float* d_features,
                     void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!w || !d_in || !d_features || !d_workspace) return vfail(DFD_EINVAL, "dfd_vit_features: null pointer");
    size_t need = 0;
    int rc = dfd_vit_workspace_bytes(images, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return vfail(DFD_ENOMEM, "dfd_vit_features: workspace too small");
    if (images * kTokens * (int64_t)kMlp > 0x7fffffffffLL) return vfail(DFD_EINVAL, "dfd_vit_features: batch too large");
    cudaStream_t s = (cudaStream_t)stream;
    const int64_t M = images * kTokens, MP = images * kPatches;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    float* X = reinterpret_cast<float*>(ws);   ws += vup((size_t)M * kDim * 4);     // fp32 residual stream
    void* H16 = ws;                            ws += vup((size_t)M * kDim * 2);     // LN output / attention output / patch GEMM output
#if !DFD_VIT_RES_IN_GEMM
    void* P16 = ws;                            ws += vup((size_t)M * kDim * 2);     // branch output (proj / fc2) waiting to be added to X
#endif
    void* BIG = ws;                                                                  // patches / qkv / MLP hidden
    const int dt = w->dtype;
    const bool f16 = dt == DFD_DTYPE_FP16;
    cudaError_t e;
    dfd::reset_launches();
#define VIT_CK(call, what) do { e = (call); dfd::note_launch(what); if (e != cudaSuccess) return vfail(DFD_ECUDA, std::string(what) + ": " + cudaGetErrorString(e)); } while (0)
    // y = LN(x (+= delta)) over `rows` rows `stride` elements apart
    auto ln = [&](float* x, int64_t stride, const void* delta, const float* g, const float* b, void* out, bool out_f32, int64_t rows) {
        const unsigned grid = (unsigned)((rows + 7) / 8);
        if (f16) {
            if (out_f32) dfd::vit_layernorm_kernel<float, __half><<<grid, 256, 0, s>>>(x, stride, (const __half*)delta, g, b, (float*)out, rows);
            else dfd::vit_layernorm_kernel<__half, __half><<<grid, 256, 0, s>>>(x, stride, (const __half*)delta, g, b, (__half*)out, rows);
        } else {
            if (out_f32) dfd::vit_layernorm_kernel<float, __nv_bfloat16><<<grid, 256, 0, s>>>(x, stride, (const __nv_bfloat16*)delta, g, b, (float*)out, rows);
            else dfd::vit_layernorm_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, s>>>(x, stride, (const __nv_bfloat16*)delta, g, b, (__nv_bfloat16*)out, rows);
        }
        return cudaGetLastError();
    };
    {
        const int64_t n8 = MP * (kPatchK / 8);
        const unsigned grid = (unsigned)((n8 + 255) / 256);
        if (f16) dfd::vit_patchify_kernel<__half><<<grid, 256, 0, s>>>(d_in, (__half*)BIG, n8);
        else dfd::vit_patchify_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(d_in, (__nv_bfloat16*)BIG, n8);
        VIT_CK(cudaGetLastError(), "vit patchify");
        VIT_CK(dfd::launch_gemm_pair(BIG, w->patch_w, w->patch_b, H16, MP, kPatchK, kDim, 0, dt, s), "vit patch-embed gemm");
        const int64_t a8 = M * (kDim / 8);
        const unsigned agrid = (unsigned)((a8 + 255) / 256);
        if (f16) dfd::vit_assemble_kernel<__half><<<agrid, 256, 0, s>>>((const __half*)H16, w->cls, w->pos, X, a8);
        else dfd::vit_assemble_kernel<__nv_bfloat16><<<agrid, 256, 0, s>>>((const __nv_bfloat16*)H16, w->cls, w->pos, X, a8);
        VIT_CK(cudaGetLastError(), "vit assemble");
    }
    for (int i = 0; i < kDepth; ++i) {
        const auto& b = w->blk[i];
#if DFD_VIT_RES_IN_GEMM
        VIT_CK(ln(X, kDim, nullptr, b.ln1_w, b.ln1_b, H16, false, M), "vit norm1");
        VIT_CK(dfd::launch_gemm_pair(H16, b.qkv_w, b.qkv_b, BIG, M, kDim, 3 * kDim, 0, dt, s), "vit qkv gemm");
        VIT_CK(dfd::launch_vit_attention_tc(BIG, H16, images, dt, s), "vit attention");
        VIT_CK(dfd::launch_gemm_pair_residual(H16, b.proj_w, b.proj_b, X, M, kDim, kDim, dt, s), "vit proj gemm");
        VIT_CK(ln(X, kDim, nullptr, b.ln2_w, b.ln2_b, H16, false, M), "vit norm2");
        VIT_CK(dfd::launch_gemm_pair(H16, b.fc1_w, b.fc1_b, BIG, M, kDim, kMlp, 2, dt, s), "vit fc1 gemm");
        VIT_CK(dfd::launch_gemm_pair_residual(BIG, b.fc2_w, b.fc2_b, X, M, kMlp, kDim, dt, s), "vit fc2 gemm");
#else
        VIT_CK(ln(X, kDim, i ? P16 : nullptr, b.ln1_w, b.ln1_b, H16, false, M), "vit norm1");            // X += fc2 output of block i-1
        VIT_CK(dfd::launch_gemm_pair(H16, b.qkv_w, b.qkv_b, BIG, M, kDim, 3 * kDim, 0, dt, s), "vit qkv gemm");
        VIT_CK(dfd::launch_vit_attention_tc(BIG, H16, images, dt, s), "vit attention");
        VIT_CK(dfd::launch_gemm_pair(H16, b.proj_w, b.proj_b, P16, M, kDim, kDim, 0, dt, s), "vit proj gemm");
        VIT_CK(ln(X, kDim, P16, b.ln2_w, b.ln2_b, H16, false, M), "vit norm2");                          // X += attention branch
        VIT_CK(dfd::launch_gemm_pair(H16, b.fc1_w, b.fc1_b, BIG, M, kDim, kMlp, 2, dt, s), "vit fc1 gemm");
        VIT_CK(dfd::launch_gemm_pair(BIG, b.fc2_w, b.fc2_b, P16, M, kMlp, kDim, 0, dt, s), "vit fc2 gemm");
#endif
    }
#if DFD_VIT_RES_IN_GEMM
    VIT_CK(ln(X, (int64_t)kTokens * kDim, nullptr, w->norm_w, w->norm_b, d_features, true, images), "vit final norm");   // CLS rows only
#else
    VIT_CK(ln(X, (int64_t)kTokens * kDim, P16, w->norm_w, w->norm_b, d_features, true, images), "vit final norm");   // CLS rows only
#endif
#undef VIT_CK
    return DFD_OK;
}

int dfd_gcn_pack_weights(int n, const char* const* names, const float* const* data, const int64_t* numel,
                         int num_classes, dfd_gcn_weights_t** out) {
    using namespace dfd;
    if (!names || !data || !numel || !out || n <= 0 || num_classes < 1 || num_classes > 64) return vfail(DFD_EINVAL, "dfd_gcn_pack_weights: bad argument");
    std::unordered_map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n; ++i) if (names[i]) t[names[i]] = {data[i], numel[i]};
    std::string missing;
    auto get = [&](const std::string& k, int64_t ne) -> const float* {
        auto it = t.find(k);
        if (it == t.end() || it->second.second != ne) { if (missing.empty()) missing = k; return nullptr; }
        return it->second.first;
    };
    std::vector<float> host;
    auto put = [&](const float* p, int rows, int cols, bool transpose) {          // -> offset (floats), 64-float aligned
        const size_t off = (host.size() + 63) & ~size_t(63);
        host.resize(off + (size_t)rows * cols, 0.f);
        if (p) for (int r = 0; r < rows; ++r) for (int c = 0; c < cols; ++c)
            host[off + (transpose ? (size_t)c * rows + r : (size_t)r * cols + c)] = p[(size_t)r * cols + c];
        return off;
    };
    const size_t o_w1 = put(get("gcn.fc1.weight", (int64_t)kGcnHid * kGcnF), kGcnHid, kGcnF, true), o_b1 = put(get("gcn.fc1.bias", kGcnHid), 1, kGcnHid, false);
    const size_t o_w2 = put(get("gcn.fc2.weight", (int64_t)kGcnOut * kGcnHid), kGcnOut, kGcnHid, true), o_b2 = put(get("gcn.fc2.bias", kGcnOut), 1, kGcnOut, false);
    const size_t o_c1 = put(get("classifier.0.weight", (int64_t)kGcnC1 * kGcnOut), kGcnC1, kGcnOut, true), o_cb1 = put(get("classifier.0.bias", kGcnC1), 1, kGcnC1, false);
    const size_t o_c2 = put(get("classifier.3.weight", (int64_t)num_classes * kGcnC1), num_classes, kGcnC1, false), o_cb2 = put(get("classifier.3.bias", num_classes), 1, num_classes, false);
    if (!missing.empty()) return vfail(DFD_EKEY, "dfd_gcn_pack_weights: state_dict tensor " + missing + " absent or wrong size (vit_out 768, gcn 256/128, classifier 64 are the built sizes)");
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size() * 4) != cudaSuccess) return vfail(DFD_ECUDA, "cudaMalloc(gcn weights) failed");
    if (cudaMemcpy(dev, host.data(), host.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return vfail(DFD_ECUDA, "cudaMemcpy(gcn weights) failed"); }
    const float* d = reinterpret_cast<const float*>(dev);
    auto* W = new dfd_gcn_weights();
    W->arena = dev;
    W->w = GcnW{d + o_w1, d + o_b1, d + o_w2, d + o_b2, d + o_c1, d + o_cb1, d + o_c2, d + o_cb2, num_classes};
    *out = W;
    return DFD_OK;
}

void dfd_gcn_free_weights(dfd_gcn_weights_t* w) { if (w) { if (w->arena) cudaFree(w->arena); delete w; } }

int dfd_gcn_head(const dfd_gcn_weights_t* w, const float* d_feats, const float* d_adj, int64_t videos, int nodes,
                 float* d_logits, void* stream) {
    using namespace dfd;
    if (!w || !d_feats || !d_adj || !d_logits) return vfail(DFD_EINVAL, "dfd_gcn_head: null pointer");
    if (videos <= 0 || nodes < 1 || nodes > kGcnMaxNodes) return vfail(DFD_EINVAL, "dfd_gcn_head: 1 <= nodes <= 64 and videos >= 1 required");
    const size_t smem = ((size_t)nodes * 2 * kGcnHid + (size_t)nodes * nodes + (size_t)kGcnChunk * kGcnF) * 4;
    cudaError_t e = cudaFuncSetAttribute(gcn_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return vfail(DFD_ECUDA, std::string("gcn head smem: ") + cudaGetErrorString(e));
    gcn_head_kernel<<<(unsigned)videos, 256, smem, (cudaStream_t)stream>>>(w->w, d_feats, d_adj, nodes, d_logits);
    e = cudaGetLastError();
    if (e != cudaSuccess) return vfail(DFD_ECUDA, std::string("gcn head: ") + cudaGetErrorString(e));
    return DFD_OK;
}

#pragma GCC visibility pop
}  // extern "C"
