// K7 — `resnet50` ensemble member: reference src/pretrained_detector.py:38-41 builds the trunk as
// nn.Sequential(*list(torchvision.models.resnet50().children())[:-1]) (conv1 7x7 s2, bn1, relu, maxpool 3x3 s2, layer1..4 of
// Bottleneck blocks, avgpool), :103-143 pools the 2048-d frame features over time and classifies; EnsembleDetector
// (:146-218) combines the members.
//
// Every convolution is a tcgen05/TMEM GEMM (gemm_tc.cu) over NHWC 16-bit activations with BatchNorm folded into the weights:
//   1x1 stride-1 convs read the activation matrix [F*H*W, C] directly;
//   3x3 convs, the stride-2 1x1 downsample convs and the 7x7 stem go through an explicit gather (im2col) into a
//   [rows, KH*KW*C] operand — simple and correct first; the TMA im2col-free form (4-D tensor maps, zero-filled halos) is
//   the obvious next step;
//   ReLU is fused into the GEMM epilogue, for conv3 AFTER the residual add (relu(bn3(conv3) + identity), torchvision
//   Bottleneck.forward).
// Max-pool, global average pool and the attention pool + head (generic feature width) are small kernels of their own.
// fp32 accumulation everywhere, fixed-order reductions (batch-invariant results).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dfd_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace {
constexpr int kRnFeat = 2048, kRnImg = 224, kRnStemK = 152;       // 7*7*3 = 147 padded to a multiple of 8
struct RnLayer { int blocks, width, stride; };
constexpr RnLayer kRnLayers[4] = {{3, 64, 1}, {4, 128, 2}, {6, 256, 2}, {3, 512, 2}};
constexpr int kRnMaxChunk = 256;                                   // frames per trunk pass (workspace = 11.9 MB per frame)
}

struct RnConv { void* w; float* b; int cin, cout, k, stride; };
struct RnBlock { RnConv c1, c2, c3, ds; bool has_ds; };
struct dfd_resnet_weights {
    int dtype;
    RnConv stem;
    std::vector<RnBlock> blocks;
    float *att_w1, *att_b1, *att_w2, *att_b2, *fc1_w, *fc1_b, *fc2_w, *fc2_b;    // [64][2048], [64], [64], [1], [256][2048], [256], [2][256], [2]
    void* arena;
};

namespace dfd {

// stem gather: x fp32 (F,3,224,224) -> A [F*112*112][152] 16-bit, column = (ky*7 + kx)*3 + c (conv 7x7 s2 p3), zero padded
template <typename T>
__global__ void rn_stem_im2col_kernel(const float* __restrict__ x, T* __restrict__ a, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           // one thread = 8 consecutive columns
    if (i >= total) return;
    constexpr int kChunks = kRnStemK / 8, O = kRnImg / 2;
    const int c8 = (int)(i % kChunks);
    const int64_t row = i / kChunks;
    const int ox = (int)(row % O), oy = (int)((row / O) % O);
    const int64_t f = row / (O * O);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int k = c8 * 8 + j;
        v[j] = 0.f;
        if (k < 147) {
            const int tap = k / 3, c = k - tap * 3, ky = tap / 7, kx = tap - ky * 7;
            const int iy = 2 * oy - 3 + ky, ix = 2 * ox - 3 + kx;
            if (iy >= 0 && iy < kRnImg && ix >= 0 && ix < kRnImg) v[j] = __ldg(x + (((size_t)f * 3 + c) * kRnImg + iy) * kRnImg + ix);
        }
    }
    uint4 o;
    o.x = Half16<T>::pack(v[0], v[1]); o.y = Half16<T>::pack(v[2], v[3]); o.z = Half16<T>::pack(v[4], v[5]); o.w = Half16<T>::pack(v[6], v[7]);
    *reinterpret_cast<uint4*>(a + (size_t)row * kRnStemK + c8 * 8) = o;
}

// generic gather for NHWC 16-bit maps: in [F][H][W][C] -> A [F*OH*OW][k*k*C], column = (ky*k + kx)*C + c, pad k/2
template <typename T>
__global__ void rn_im2col_kernel(const T* __restrict__ in, T* __restrict__ a, int H, int W, int C, int OH, int OW, int k, int stride, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           // one thread = 8 channels of one tap
    if (i >= total) return;
    const int cpt = C >> 3;
    const int c8 = (int)(i % cpt);
    const int64_t t1 = i / cpt;
    const int tap = (int)(t1 % (k * k));
    const int64_t row = t1 / (k * k);
    const int ox = (int)(row % OW), oy = (int)((row / OW) % OH);
    const int64_t f = row / ((int64_t)OW * OH);
    const int ky = tap / k, kx = tap - ky * k, pad = k / 2;
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = *reinterpret_cast<const uint4*>(in + (((size_t)f * H + iy) * W + ix) * C + c8 * 8);
    *reinterpret_cast<uint4*>(a + ((size_t)row * (k * k) + tap) * C + c8 * 8) = v;
}

// DFD_RESNET_SMALL_KERNELS_BEGIN   (tools/host_emul/ runs the three kernels below, unchanged, on CPU threads)
// max-pool 3x3 s2 p1 on NHWC (padding never wins: -inf)
template <typename T>
__global__ void rn_maxpool_kernel(const T* __restrict__ in, T* __restrict__ out, int H, int W, int C, int OH, int OW, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           // one thread = 8 channels of one output pixel
    if (i >= total) return;
    const int cpt = C >> 3;
    const int c8 = (int)(i % cpt);
    const int64_t px = i / cpt;
    const int ox = (int)(px % OW), oy = (int)((px / OW) % OH);
    const int64_t f = px / ((int64_t)OW * OH);
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) m[j] = -INFINITY;
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy - 1 + ky;
        if (iy < 0 || iy >= H) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int ix = 2 * ox - 1 + kx;
            if (ix < 0 || ix >= W) continue;
            const uint4 v = *reinterpret_cast<const uint4*>(in + (((size_t)f * H + iy) * W + ix) * C + c8 * 8);
            const float2 a0 = Half16<T>::unpack(v.x), a1 = Half16<T>::unpack(v.y), a2 = Half16<T>::unpack(v.z), a3 = Half16<T>::unpack(v.w);
            m[0] = fmaxf(m[0], a0.x); m[1] = fmaxf(m[1], a0.y); m[2] = fmaxf(m[2], a1.x); m[3] = fmaxf(m[3], a1.y);
            m[4] = fmaxf(m[4], a2.x); m[5] = fmaxf(m[5], a2.y); m[6] = fmaxf(m[6], a3.x); m[7] = fmaxf(m[7], a3.y);
        }
    }
    uint4 o;
    o.x = Half16<T>::pack(m[0], m[1]); o.y = Half16<T>::pack(m[2], m[3]); o.z = Half16<T>::pack(m[4], m[5]); o.w = Half16<T>::pack(m[6], m[7]);
    *reinterpret_cast<uint4*>(out + (size_t)px * C + c8 * 8) = o;
}

// global average pool: in [F][HW][C] 16-bit -> feat fp32 [F][C], rows added in index order
template <typename T>
__global__ void rn_avgpool_kernel(const T* __restrict__ in, float* __restrict__ feat, int HW, int C, int64_t total) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;           // one thread = 2 channels of one frame
    if (i >= total) return;
    const int c2 = (int)(i % (C >> 1));
    const int64_t f = i / (C >> 1);
    const T* p = in + (size_t)f * HW * C + 2 * c2;
    float s0 = 0.f, s1 = 0.f;
    for (int r = 0; r < HW; ++r) {
        const float2 v = Half16<T>::unpack(*reinterpret_cast<const uint32_t*>(p + (size_t)r * C));
        s0 += v.x; s1 += v.y;
    }
    const float inv = 1.0f / (float)HW;
    *reinterpret_cast<float2*>(feat + (size_t)f * C + 2 * c2) = make_float2(s0 * inv, s1 * inv);
}

// DFD_RESNET_SMALL_KERNELS_END

}  // namespace dfd

namespace {
thread_local std::string g_rn_err;
int nfail(int code, const std::string& m) { g_rn_err = m; return code; }
uint16_t nh16(float v, int dtype) {
    if (dtype == DFD_DTYPE_FP16) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
    __nv_bfloat16 h = __float2bfloat16_rn(v); uint16_t u; memcpy(&u, &h, 2); return u;
}
size_t nup(size_t b) { return (b + 255) & ~size_t(255); }
// per-frame element counts of the trunk buffers (16-bit elements)
constexpr size_t kRnAct = (size_t)56 * 56 * 256;                  // largest activation (= 112*112*64)
constexpr size_t kRnCol = (size_t)112 * 112 * kRnStemK;           // largest gathered operand (stem; layer1 conv2 = 3136*576 is smaller)
size_t rn_frame_bytes() { return 5 * nup(kRnAct * 2) + nup(kRnCol * 2); }
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* dfd_resnet_last_error(void) { return g_rn_err.c_str(); }

int dfd_resnet50_pack_weights(int n, const char* const* names, const float* const* data, const int64_t* numel,
                              int dtype, dfd_resnet_weights_t** out) {
    if (!names || !data || !numel || !out || n <= 0) return nfail(DFD_EINVAL, "dfd_resnet50_pack_weights: null argument");
    if (dtype != DFD_DTYPE_BF16 && dtype != DFD_DTYPE_FP16) return nfail(DFD_EINVAL, "dfd_resnet50_pack_weights: unknown dtype");
    std::unordered_map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n; ++i) if (names[i]) t[names[i]] = {data[i], numel[i]};
    std::string missing;
    auto get = [&](const std::string& k, int64_t ne) -> const float* {
        auto it = t.find(k);
        if (it == t.end() || it->second.second != ne || !it->second.first) { if (missing.empty()) missing = k; return nullptr; }
        return it->second.first;
    };
    std::vector<uint8_t> host;
    auto alloc = [&](size_t nb) { size_t o = (host.size() + 255) & ~size_t(255); host.resize(o + nb, 0); return o; };
    struct COff { size_t w, b; int cin, cout, k, stride, kpad; };
    // conv [N][C][k][k] (no bias) + BN(eps 1e-5) -> 16-bit [N][kpad] with column (ky*k + kx)*C + c, fp32 bias
    auto pack_conv = [&](const std::string& conv, const std::string& bn, int cin, int cout, int k, int stride, int kpad) {
        COff o{alloc((size_t)cout * kpad * 2), alloc((size_t)cout * 4), cin, cout, k, stride, kpad};
        const float* w = get(conv + ".weight", (int64_t)cout * cin * k * k);
        const float* g = get(bn + ".weight", cout); const float* be = get(bn + ".bias", cout);
        const float* mu = get(bn + ".running_mean", cout); const float* va = get(bn + ".running_var", cout);
        if (!w || !g || !be || !mu || !va) return o;
        uint16_t* wd = reinterpret_cast<uint16_t*>(host.data() + o.w);
        float* bd = reinterpret_cast<float*>(host.data() + o.b);
        for (int nn = 0; nn < cout; ++nn) {
            const float sc = g[nn] / sqrtf(va[nn] + 1e-5f);
            bd[nn] = be[nn] - mu[nn] * sc;
            for (int c = 0; c < cin; ++c)
                for (int ky = 0; ky < k; ++ky)
                    for (int kx = 0; kx < k; ++kx)
                        wd[(size_t)nn * kpad + (size_t)(ky * k + kx) * cin + c] = nh16(w[(((size_t)nn * cin + c) * k + ky) * k + kx] * sc, dtype);
        }
        return o;
    };
    const COff stem = pack_conv("backbone.0", "backbone.1", 3, 64, 7, 2, kRnStemK);
    struct BOff { COff c1, c2, c3, ds; bool has_ds; };
    std::vector<BOff> boffs;
    int cin = 64;
    for (int l = 0; l < 4; ++l)
        for (int b = 0; b < kRnLayers[l].blocks; ++b) {
            const int width = kRnLayers[l].width, s = b == 0 ? kRnLayers[l].stride : 1, cout = width * 4;
            const std::string p = "backbone." + std::to_string(4 + l) + "." + std::to_string(b) + ".";
            BOff bo{};
            bo.c1 = pack_conv(p + "conv1", p + "bn1", cin, width, 1, 1, cin);
            bo.c2 = pack_conv(p + "conv2", p + "bn2", width, width, 3, s, 9 * width);
            bo.c3 = pack_conv(p + "conv3", p + "bn3", width, cout, 1, 1, width);
            bo.has_ds = b == 0;                                   // torchvision: first block of every layer (channel change and/or stride)
            if (bo.has_ds) bo.ds = pack_conv(p + "downsample.0", p + "downsample.1", cin, cout, 1, s, cin);
            boffs.push_back(bo);
            cin = cout;
        }
    size_t hoff[8];
    {
        const char* keys[8] = {"temporal_attention.0.weight", "temporal_attention.0.bias", "temporal_attention.2.weight",
                               "temporal_attention.2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"};
        const int64_t ne[8] = {64 * kRnFeat, 64, 64, 1, 256 * kRnFeat, 256, 2 * 256, 2};
        for (int i = 0; i < 8; ++i) {
            hoff[i] = alloc((size_t)ne[i] * 4);
            const float* p = get(keys[i], ne[i]);
            if (p) memcpy(host.data() + hoff[i], p, (size_t)ne[i] * 4);
        }
    }
    if (!missing.empty()) return nfail(DFD_EKEY, "dfd_resnet50_pack_weights: state_dict tensor " + missing + " absent or wrong size");
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size()) != cudaSuccess) return nfail(DFD_ECUDA, "cudaMalloc(resnet weights) failed");
    if (cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return nfail(DFD_ECUDA, "cudaMemcpy(resnet weights) failed"); }
    uint8_t* d = reinterpret_cast<uint8_t*>(dev);
    auto conv = [&](const COff& o) { return RnConv{d + o.w, reinterpret_cast<float*>(d + o.b), o.cin, o.cout, o.k, o.stride}; };
    auto* W = new dfd_resnet_weights();
    W->dtype = dtype; W->arena = dev; W->stem = conv(stem);
    for (const BOff& bo : boffs) {
        RnBlock b{conv(bo.c1), conv(bo.c2), conv(bo.c3), bo.has_ds ? conv(bo.ds) : RnConv{}, bo.has_ds};
        W->blocks.push_back(b);
    }
    auto F = [&](int i) { return reinterpret_cast<float*>(d + hoff[i]); };
    W->att_w1 = F(0); W->att_b1 = F(1); W->att_w2 = F(2); W->att_b2 = F(3); W->fc1_w = F(4); W->fc1_b = F(5); W->fc2_w = F(6); W->fc2_b = F(7);
    *out = W;
    return DFD_OK;
}

void dfd_resnet50_free_weights(dfd_resnet_weights_t* w) { if (w) { if (w->arena) cudaFree(w->arena); delete w; } }

int dfd_resnet50_workspace_bytes(int64_t frames, size_t* bytes) {
    if (!bytes || frames <= 0) return nfail(DFD_EINVAL, "dfd_resnet50_workspace_bytes: bad argument");
    const size_t chunk = (size_t)(frames < kRnMaxChunk ? frames : kRnMaxChunk);
    *bytes = rn_frame_bytes() * chunk + nup((size_t)frames * kRnFeat * 4) + 1024;
    return DFD_OK;
}

// trunk on one chunk of frames: x fp32 (n,3,224,224) -> feat fp32 (n,2048)
static int rn_trunk_chunk(const dfd_resnet_weights* w, const float* x, int64_t n, float* feat, uint8_t* ws, cudaStream_t s) {
    using namespace dfd;
    const int dt = w->dtype;
    const bool f16 = dt == DFD_DTYPE_FP16;
    uint8_t* buf[5];
    for (int i = 0; i < 5; ++i) { buf[i] = ws; ws += nup(kRnAct * 2) * (size_t)n; }
    uint8_t* col = ws;
    cudaError_t e;
#define RN_CK(call, what) do { e = (call); dfd::note_launch(what); if (e != cudaSuccess) return nfail(DFD_ECUDA, std::string(what) + ": " + cudaGetErrorString(e)); } while (0)
    auto im2col = [&](const void* in, void* a, int H, int W, int C, int OH, int OW, int k, int stride) {
        const int64_t total = n * OH * OW * (int64_t)(k * k) * (C / 8);
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (f16) rn_im2col_kernel<__half><<<grid, 256, 0, s>>>((const __half*)in, (__half*)a, H, W, C, OH, OW, k, stride, total);
        else rn_im2col_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)a, H, W, C, OH, OW, k, stride, total);
        return cudaGetLastError();
    };
    // stem: gather -> GEMM (+bias, ReLU) -> max-pool
    {
        const int64_t rows = n * 112 * 112, total = rows * (kRnStemK / 8);
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (f16) rn_stem_im2col_kernel<__half><<<grid, 256, 0, s>>>(x, (__half*)col, total);
        else rn_stem_im2col_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(x, (__nv_bfloat16*)col, total);
        RN_CK(cudaGetLastError(), "resnet stem gather");
        RN_CK(launch_gemm_tc(col, w->stem.w, w->stem.b, nullptr, nullptr, buf[0], rows, kRnStemK, 64, 1, 3, dt, s), "resnet stem gemm");
        const int64_t ptotal = n * 56 * 56 * (64 / 8);
        const unsigned pgrid = (unsigned)((ptotal + 255) / 256);
        if (f16) rn_maxpool_kernel<__half><<<pgrid, 256, 0, s>>>((const __half*)buf[0], (__half*)buf[1], 112, 112, 64, 56, 56, ptotal);
        else rn_maxpool_kernel<__nv_bfloat16><<<pgrid, 256, 0, s>>>((const __nv_bfloat16*)buf[0], (__nv_bfloat16*)buf[1], 112, 112, 64, 56, 56, ptotal);
        RN_CK(cudaGetLastError(), "resnet maxpool");
    }
    // The 13 stride-1 3x3 convs run as implicit GEMMs — conv1 scatters its rows into a zero-haloed map kept in `col`, conv2 reads nine shifted TMA
    // boxes of it (gemm_tc.cu, CONV variants) — instead of gathering a 9x larger operand.  The halo is zeroed once per
    // (H, C) geometry: inside a layer only interior rows are ever rewritten.
    // (measured on B200, 256 frames: 14.19 -> 11.50 ms against gathering those operands; the three stride-2 3x3 convs,
    // the strided 1x1 shortcuts and the 7x7 stem still gather).
    int pad_h = 0, pad_c = 0;
    int cur = 1, H = 56;                                           // buf[cur] holds the block input [n][H][H][cin]
    for (const RnBlock& b : w->blocks) {
        const int s2 = b.c2.stride, OH = (H + 2 - 3) / s2 + 1;
        uint8_t* y = buf[cur];
        uint8_t* o1 = buf[(cur + 1) % 5]; uint8_t* o2 = buf[(cur + 2) % 5]; uint8_t* idn = buf[(cur + 3) % 5]; uint8_t* nxt = buf[(cur + 4) % 5];
        const int64_t rows_in = n * H * H, rows_out = n * OH * OH;
        if (s2 == 1 && (b.c2.cin % 64) == 0) {
            if (pad_h != H || pad_c != b.c2.cin) {
                RN_CK(cudaMemsetAsync(col, 0, (size_t)conv3x3_padded_rows(n, H, H) * b.c2.cin * 2, s), "resnet halo memset");
                pad_h = H; pad_c = b.c2.cin;
            }
            RN_CK(launch_gemm_tc_padout(y, b.c1.w, b.c1.b, col, n, H, H, b.c1.cin, b.c1.cout, dt, s), "resnet conv1 (haloed output)");
            RN_CK(launch_gemm_tc_conv3x3(col, b.c2.w, b.c2.b, o2, n, H, H, b.c2.cin, b.c2.cout, dt, s), "resnet conv2 (implicit)");
        } else {
            RN_CK(launch_gemm_tc(y, b.c1.w, b.c1.b, nullptr, nullptr, o1, rows_in, b.c1.cin, b.c1.cout, 1, 3, dt, s), "resnet conv1");
            RN_CK(im2col(o1, col, H, H, b.c2.cin, OH, OH, 3, s2), "resnet conv2 gather");
            RN_CK(launch_gemm_tc(col, b.c2.w, b.c2.b, nullptr, nullptr, o2, rows_out, 9 * b.c2.cin, b.c2.cout, 1, 3, dt, s), "resnet conv2");
            pad_h = 0;                                             // the gather overwrote the haloed map
        }
        const void* res = y;
        if (b.has_ds) {
            const void* a = y;
            if (b.ds.stride != 1) { RN_CK(im2col(y, col, H, H, b.ds.cin, OH, OH, 1, b.ds.stride), "resnet downsample gather"); a = col; }
            RN_CK(launch_gemm_tc(a, b.ds.w, b.ds.b, nullptr, nullptr, idn, rows_out, b.ds.cin, b.ds.cout, 1, 0, dt, s), "resnet downsample");
            res = idn;
        }
        RN_CK(launch_gemm_tc(o2, b.c3.w, b.c3.b, nullptr, res, nxt, rows_out, b.c3.cin, b.c3.cout, 1, 3, dt, s), "resnet conv3");
        cur = (cur + 4) % 5; H = OH;
    }
    {
        const int64_t total = n * (kRnFeat / 2);
        const unsigned grid = (unsigned)((total + 255) / 256);
        if (f16) rn_avgpool_kernel<__half><<<grid, 256, 0, s>>>((const __half*)buf[cur], feat, H * H, kRnFeat, total);
        else rn_avgpool_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)buf[cur], feat, H * H, kRnFeat, total);
        RN_CK(cudaGetLastError(), "resnet avgpool");
    }
#undef RN_CK
    return DFD_OK;
}

int dfd_resnet50_score_videos(const dfd_resnet_weights_t* w, const float* d_in, const int32_t* d_offsets, int64_t videos, int64_t frames,
                              int max_frames_per_video, int use_attention, float* d_logits, float* d_frame_scores, float* d_features_out,
                              void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!w || !d_in || !d_offsets || !d_logits || !d_workspace) return nfail(DFD_EINVAL, "dfd_resnet50_score_videos: null pointer");
    if (videos <= 0 || frames <= 0 || max_frames_per_video <= 0 || max_frames_per_video > 1024) return nfail(DFD_EINVAL, "dfd_resnet50_score_videos: bad counts (1..1024 frames per video)");
    size_t need = 0;
    int rc = dfd_resnet50_workspace_bytes(frames, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return nfail(DFD_ENOMEM, "dfd_resnet50_score_videos: workspace too small");
    cudaStream_t s = (cudaStream_t)stream;
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    const size_t chunk = (size_t)(frames < kRnMaxChunk ? frames : kRnMaxChunk);
    float* feat = d_features_out ? d_features_out : reinterpret_cast<float*>(ws + rn_frame_bytes() * chunk);
    dfd::reset_launches();
    for (int64_t f0 = 0; f0 < frames; f0 += kRnMaxChunk) {
        const int64_t nfr = frames - f0 < kRnMaxChunk ? frames - f0 : kRnMaxChunk;
        rc = rn_trunk_chunk(w, d_in + (size_t)f0 * 3 * kRnImg * kRnImg, nfr, feat + (size_t)f0 * kRnFeat, ws, s);
        if (rc) return rc;
    }
    // the same attention pool + head kernel as the efficientnet_b0 member (poolhead.cu), instantiated for 2048 features
    const dfd::HeadWeights hw{w->att_w1, w->att_b1, w->att_w2, w->att_b2, w->fc1_w, w->fc1_b, w->fc2_w, w->fc2_b};
    cudaError_t e = dfd::launch_pool_head(hw, feat, d_offsets, videos, frames, kRnFeat, use_attention, d_logits, d_frame_scores, s);
    dfd::note_launch("resnet pool head");
    if (e != cudaSuccess) return nfail(DFD_ECUDA, std::string("resnet pool head: ") + cudaGetErrorString(e));
    return DFD_OK;
}

#pragma GCC visibility pop
}  // extern "C"
