// C ABI (include/dfd_b200.h, include/dfd_b200_kernels.h): weight packing (BN fold, repack), the
// EfficientNet-B0 layer schedule, workspace planning, error reporting.  Host code only; the kernels are in
// preprocess.cu / stem.cu / dwconv.cu / se.cu / gemm_tc.cu / poolhead.cu.
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
#include "../../include/dfd_b200_kernels.h"
#include "kernels.h"
#include "conv_map.h"

namespace {

thread_local std::string g_err;
thread_local int g_launches = 0;

// ---- optional per-launch profiling (dfd_profile_*): CUDA events around every launch on the launch stream
enum KClass { KC_PREPROCESS = 0, KC_STEM, KC_DWCONV, KC_SE, KC_GEMM_EXPAND, KC_GEMM_PROJECT, KC_GEMM_HEAD_POOL, KC_POOL_HEAD, KC_MBCONV_FUSED, KC_COUNT };
const char* const kClassNames[KC_COUNT] = {"preprocess", "stem", "dwconv_se_squeeze", "se_gate", "gemm_expand", "gemm_project",
                                           "gemm_head_pool", "attn_pool_head", "expand_dwconv_fused"};
struct ProfRec { int cls; cudaEvent_t a, b; double bytes, flops; };
thread_local bool g_prof_on = false;
thread_local std::vector<ProfRec> g_prof;
thread_local int g_cls = 0;
thread_local double g_bytes = 0, g_flops = 0;
thread_local cudaStream_t g_prof_stream = nullptr;
// declare what the NEXT launch is (class, algorithmic bytes = activations read once + written once, flops)
inline void prof_next(int cls, double bytes, double flops, cudaStream_t s) { g_cls = cls; g_bytes = bytes; g_flops = flops; g_prof_stream = s; }
struct ProfScope {
    ProfRec r{}; bool on;
    ProfScope() : on(g_prof_on) {
        if (on) { cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, g_prof_stream); }
    }
    ~ProfScope() {
        if (on) { cudaEventRecord(r.b, g_prof_stream); r.cls = g_cls; r.bytes = g_bytes; r.flops = g_flops; g_prof.push_back(r); }
    }
};

int fail(int code, const std::string& msg) { g_err = msg; return code; }
}  // namespace
namespace dfd {
void reset_launches() { g_launches = 0; }
void note_launch(const char* what) { if (!strstr(what, "smem") && !strstr(what, "memset")) ++g_launches; }
}  // namespace dfd
namespace {
int cuda_fail(cudaError_t e, const char* what) {
    g_err = std::string(what) + ": " + cudaGetErrorString(e);
    return DFD_ECUDA;
}
#define DFD_CUDA(call, what)                                   \
    do {                                                       \
        cudaError_t _e = (call);                               \
        if (_e != cudaSuccess) return cuda_fail(_e, what);     \
    } while (0)
#define DFD_LAUNCH(call, what)                                 \
    do {                                                       \
        cudaError_t _e;                                        \
        { ProfScope _ps; _e = (call); }                        \
        ++g_launches;                                          \
        if (_e != cudaSuccess) return cuda_fail(_e, what);     \
    } while (0)

// timm efficientnet_b0 stage spec: (repeats, kernel, stride, expand, out_channels)  — SURVEY.md §8c
const int kStages[7][5] = {{1, 3, 1, 1, 16}, {2, 3, 2, 6, 24}, {2, 5, 2, 6, 40}, {3, 3, 2, 6, 80},
                           {3, 5, 1, 6, 112}, {4, 5, 2, 6, 192}, {1, 3, 1, 6, 320}};
constexpr int kNumBlocks = 16;
constexpr float kBnEps = 1e-5f;

struct BlockW {
    int cin, mid, cout, k, stride, rd;
    bool has_expand, has_skip;
    void* exp_w; float* exp_b;                         // [mid][cin] 16-bit, [mid]
    float* dw_w; float* dw_b;                          // [k*k][mid], [mid]
    float *se_w1, *se_b1, *se_w2t, *se_b2;             // [rd][mid], [rd], [rd][mid], [mid]
    void* proj_w; float* proj_b;                       // [cout][mid] 16-bit, [cout]
    // Pixel packing for narrow layers (K < 64): r consecutive pixels of the NHWC map are ONE GEMM row of r*K channels
    // against the block-diagonal weight diag(W, .., W) [r*N][r*K]; the output rows [M/r][r*N] are the same bytes as
    // [M][N].  Full 128-byte TMA rows and r times fewer tiles for free (the tensor pipe has the headroom).
    int exp_pack, proj_pack;                           // r (1 = not packed)
    void* exp_wp; float* exp_bp; void* proj_wp; float* proj_bp;
};

// r = pixels per GEMM row for a pointwise layer with K input channels (r*K <= 64)
int pack_factor(int K) { return K <= 16 ? 4 : (K <= 32 ? 2 : 1); }

// frames per pass of the trunk (bounds the workspace; results do not depend on it: tests/test_gpu_path.py)
int64_t chunk_frames() {
    const char* e = getenv("DFD_CHUNK_FRAMES");
    long v = e ? atol(e) : 0;
    return v > 0 ? v : 2048;
}

}  // namespace

struct dfd_weights {
    int dtype;
    float *stem_w, *stem_b;                            // [27][32], [32]
    void* stem_w16;                                    // [32][96] 16-bit hi/lo split for the tcgen05 stem
    void* stem_wrow; float* stem_b4;                   // row-variant stem: fp16 [2][32][32], fp32 [4][32] (pack_stem_row)
    BlockW blocks[kNumBlocks];
    void* head_w; float* head_b;                       // [1280][320] 16-bit, [1280]
    dfd::HeadWeights hw;
    void* arena;
    size_t arena_bytes;
};

namespace {

// ---------------------------------------------------------------------------------------------------
// packing
// ---------------------------------------------------------------------------------------------------
struct HostArena {
    std::vector<uint8_t> bytes;
    size_t alloc(size_t n) {                           // 256-byte aligned offsets
        size_t off = (bytes.size() + 255) & ~size_t(255);
        bytes.resize(off + n, 0);
        return off;
    }
};

struct Tensors {
    std::unordered_map<std::string, std::pair<const float*, int64_t>> map;
    std::string missing;
    const float* get(const std::string& key, int64_t numel) {
        auto it = map.find(key);
        if (it == map.end() || it->second.second != numel || it->second.first == nullptr) {
            if (missing.empty()) missing = key + (it == map.end() ? " (absent)" : " (wrong element count)");
            return nullptr;
        }
        return it->second.first;
    }
};

uint16_t to_h16(float v, int dtype) {
    if (dtype == DFD_DTYPE_FP16) {
        if (v > 65504.f) v = 65504.f;
        if (v < -65504.f) v = -65504.f;
        __half h = __float2half_rn(v);
        uint16_t u; memcpy(&u, &h, 2); return u;
    }
    __nv_bfloat16 h = __float2bfloat16_rn(v);
    uint16_t u; memcpy(&u, &h, 2); return u;
}

// BN fold factors: scale = gamma / sqrt(var + eps), shift = beta - mean * scale   (fp32)
bool bn_fold(Tensors& t, const std::string& p, int c, std::vector<float>& scale, std::vector<float>& shift) {
    const float* g = t.get(p + ".weight", c);
    const float* b = t.get(p + ".bias", c);
    const float* m = t.get(p + ".running_mean", c);
    const float* v = t.get(p + ".running_var", c);
    if (!g || !b || !m || !v) return false;
    scale.resize(c); shift.resize(c);
    for (int i = 0; i < c; ++i) {
        scale[i] = g[i] / sqrtf(v[i] + kBnEps);
        shift[i] = b[i] - m[i] * scale[i];
    }
    return true;
}

// pointwise conv [N][K] + BN -> 16-bit [N][K] (K-major) + fp32 bias
bool pack_pw(Tensors& t, const std::string& conv, const std::string& bn, int N, int K, int dtype,
             HostArena& a, size_t& w_off, size_t& b_off) {
    const float* w = t.get(conv + ".weight", (int64_t)N * K);
    std::vector<float> sc, sh;
    if (!w || !bn_fold(t, bn, N, sc, sh)) return false;
    w_off = a.alloc((size_t)N * K * 2);
    b_off = a.alloc((size_t)N * 4);
    uint16_t* wd = reinterpret_cast<uint16_t*>(a.bytes.data() + w_off);
    for (int n = 0; n < N; ++n)
        for (int k = 0; k < K; ++k) wd[(size_t)n * K + k] = to_h16(w[(size_t)n * K + k] * sc[n], dtype);
    memcpy(a.bytes.data() + b_off, sh.data(), (size_t)N * 4);
    return true;
}

// block-diagonal copy of a packed pointwise layer: [r*N][r*K] 16-bit (+ bias repeated r times)
void pack_pw_diag(HostArena& a, size_t w_off, size_t b_off, int N, int K, int r, size_t& wp_off, size_t& bp_off) {
    wp_off = a.alloc((size_t)r * N * r * K * 2);
    bp_off = a.alloc((size_t)r * N * 4);
    const uint16_t* w = reinterpret_cast<const uint16_t*>(a.bytes.data() + w_off);
    const float* b = reinterpret_cast<const float*>(a.bytes.data() + b_off);
    uint16_t* wd = reinterpret_cast<uint16_t*>(a.bytes.data() + wp_off);     // zero-initialised by alloc
    float* bd = reinterpret_cast<float*>(a.bytes.data() + bp_off);
    for (int i = 0; i < r; ++i)
        for (int n = 0; n < N; ++n) {
            bd[i * N + n] = b[n];
            for (int k = 0; k < K; ++k) wd[((size_t)i * N + n) * (r * K) + i * K + k] = w[(size_t)n * K + k];
        }
}

// stem weights [27][32] fp32 (tap-major) -> 16-bit [32][96] = [w_hi | w_hi | w_lo], taps padded 27 -> 32
size_t pack_stem16(HostArena& a, const float* p27x32, int dtype) {
    size_t off = a.alloc(32 * 96 * 2);
    uint16_t* d = reinterpret_cast<uint16_t*>(a.bytes.data() + off);
    for (int o = 0; o < 32; ++o)
        for (int i = 0; i < 27; ++i) {
            const float w = p27x32[i * 32 + o];
            const uint16_t hi = to_h16(w, dtype);
            float hf;
            if (dtype == DFD_DTYPE_FP16) { __half h; memcpy(&h, &hi, 2); hf = __half2float(h); }
            else { __nv_bfloat16 h; memcpy(&h, &hi, 2); hf = __bfloat162float(h); }
            d[o * 96 + i] = hi; d[o * 96 + 32 + i] = hi; d[o * 96 + 64 + i] = to_h16(w - hf, dtype);
        }
    return off;
}

// Row-variant stem operands (stem_tc.cu): fp16 [hi|lo][32 oc][32 k], k = ky*10 + kx*3 + c, holding
// w' = 256 * w / (255 * std_c) split in two fp16 terms, and the four bias vectors [top*2 + left][32] that absorb
// -sum_inb w * mean_c / std_c over the taps that are inside the image (tensor prep of app.py:1772-1780 folded in).
void pack_stem_row(HostArena& a, const float* p27x32, const float* bias32, size_t& w_off, size_t& b4_off) {
    const double mean[3] = {(double)0.485f, (double)0.456f, (double)0.406f}, stdv[3] = {(double)0.229f, (double)0.224f, (double)0.225f};
    w_off = a.alloc(2 * 32 * 32 * 2);
    b4_off = a.alloc(4 * 32 * 4);
    uint16_t* wd = reinterpret_cast<uint16_t*>(a.bytes.data() + w_off);
    float* bd = reinterpret_cast<float*>(a.bytes.data() + b4_off);
    for (int o = 0; o < 32; ++o) {
        for (int ky = 0; ky < 3; ++ky)
            for (int kx = 0; kx < 3; ++kx)
                for (int c = 0; c < 3; ++c) {
                    const double w = 256.0 * (double)p27x32[((ky * 3 + kx) * 3 + c) * 32 + o] / (255.0 * stdv[c]);
                    const __half hi = __float2half_rn((float)w);
                    const __half lo = __float2half_rn((float)(w - (double)__half2float(hi)));
                    const int k = ky * 10 + kx * 3 + c;
                    memcpy(&wd[(size_t)o * 32 + k], &hi, 2);
                    memcpy(&wd[(size_t)(32 + o) * 32 + k], &lo, 2);
                }
        for (int cs = 0; cs < 4; ++cs) {
            double b = (double)bias32[o];
            for (int ky = (cs & 2) ? 1 : 0; ky < 3; ++ky)
                for (int kx = (cs & 1) ? 1 : 0; kx < 3; ++kx)
                    for (int c = 0; c < 3; ++c) b -= (double)p27x32[((ky * 3 + kx) * 3 + c) * 32 + o] * mean[c] / stdv[c];
            bd[cs * 32 + o] = (float)b;
        }
    }
}

size_t pack_f32(HostArena& a, const float* src, size_t n) {
    size_t off = a.alloc(n * 4);
    memcpy(a.bytes.data() + off, src, n * 4);
    return off;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int dfd_abi_version(void) { return DFD_ABI_VERSION; }
const char* dfd_last_error(void) { return g_err.c_str(); }
int dfd_last_launch_count(void) { return g_launches; }
int dfd_weights_dtype(const dfd_weights_t* w) { return w ? w->dtype : DFD_EINVAL; }

int dfd_pack_weights(int n_tensors, const char* const* names, const float* const* data, const int64_t* numel,
                     int dtype, dfd_weights_t** out) {
    if (!names || !data || !numel || !out || n_tensors <= 0) return fail(DFD_EINVAL, "dfd_pack_weights: null argument");
    if (dtype != DFD_DTYPE_BF16 && dtype != DFD_DTYPE_FP16) return fail(DFD_EINVAL, "dfd_pack_weights: unknown dtype");
    Tensors t;
    for (int i = 0; i < n_tensors; ++i)
        if (names[i]) t.map[names[i]] = {data[i], numel[i]};

    HostArena a;
    struct Off { size_t exp_w, exp_b, dw_w, dw_b, se_w1, se_b1, se_w2t, se_b2, proj_w, proj_b, exp_wp, exp_bp, proj_wp, proj_bp; } off[kNumBlocks];
    dfd_weights W{};
    W.dtype = dtype;

    // stem: [32][3][3][3] + BN -> fp32 [(ky*3+kx)*3+c][32]
    size_t stem_w_off = 0, stem_b_off = 0, stem_w16_off = 0, stem_wrow_off = 0, stem_b4_off = 0;
    {
        const float* w = t.get("backbone.0.weight", 32 * 27);
        std::vector<float> sc, sh;
        if (w && bn_fold(t, "backbone.1", 32, sc, sh)) {
            std::vector<float> p(27 * 32);
            for (int o = 0; o < 32; ++o)
                for (int c = 0; c < 3; ++c)
                    for (int ky = 0; ky < 3; ++ky)
                        for (int kx = 0; kx < 3; ++kx)
                            p[((ky * 3 + kx) * 3 + c) * 32 + o] = w[((o * 3 + c) * 3 + ky) * 3 + kx] * sc[o];
            stem_w_off = pack_f32(a, p.data(), p.size());
            stem_w16_off = pack_stem16(a, p.data(), dtype);
            stem_b_off = pack_f32(a, sh.data(), 32);
            pack_stem_row(a, p.data(), sh.data(), stem_wrow_off, stem_b4_off);
        }
    }
    int bi = 0, cin = 32;
    for (int s = 0; s < 7; ++s) {
        for (int b = 0; b < kStages[s][0]; ++b, ++bi) {
            BlockW& B = W.blocks[bi];
            B.cin = cin; B.k = kStages[s][1]; B.stride = (b == 0) ? kStages[s][2] : 1;
            B.mid = cin * kStages[s][3]; B.cout = kStages[s][4];
            B.has_expand = kStages[s][3] != 1;
            B.has_skip = (B.stride == 1 && B.cin == B.cout);
            B.rd = std::max(1, (int)lrintf(cin * 0.25f));
            const std::string p = "backbone.2." + std::to_string(s) + "." + std::to_string(b);
            const std::string bn_dw = p + (B.has_expand ? ".bn2" : ".bn1");
            const std::string bn_pr = p + (B.has_expand ? ".bn3" : ".bn2");
            const std::string conv_pr = p + (B.has_expand ? ".conv_pwl" : ".conv_pw");
            Off& o = off[bi];
            memset(&o, 0, sizeof(o));
            B.exp_pack = 1;      // expand layers are epilogue-bound (SiLU on 6x the channels): packing measured 10-18% SLOWER there
                                 // (16->96 at 112x112: 535 us unpacked, 596 us r=2, 631 us r=4 per 1024 frames), so only project layers pack
            B.proj_pack = pack_factor(B.mid);
            if (B.has_expand && pack_pw(t, p + ".conv_pw", p + ".bn1", B.mid, B.cin, dtype, a, o.exp_w, o.exp_b) && B.exp_pack > 1)
                pack_pw_diag(a, o.exp_w, o.exp_b, B.mid, B.cin, B.exp_pack, o.exp_wp, o.exp_bp);
            {   // depthwise [mid][1][k][k] + BN -> fp32 [k*k][mid]
                const int kk = B.k * B.k;
                const float* w = t.get(p + ".conv_dw.weight", (int64_t)B.mid * kk);
                std::vector<float> sc, sh;
                if (w && bn_fold(t, bn_dw, B.mid, sc, sh)) {
                    std::vector<float> pw((size_t)kk * B.mid);
                    for (int c = 0; c < B.mid; ++c)
                        for (int i = 0; i < kk; ++i) pw[(size_t)i * B.mid + c] = w[(size_t)c * kk + i] * sc[c];
                    o.dw_w = pack_f32(a, pw.data(), pw.size());
                    o.dw_b = pack_f32(a, sh.data(), B.mid);
                }
            }
            {   // squeeze-excite: conv_reduce [rd][mid], conv_expand [mid][rd] -> transposed [rd][mid]
                const float* w1 = t.get(p + ".se.conv_reduce.weight", (int64_t)B.rd * B.mid);
                const float* b1 = t.get(p + ".se.conv_reduce.bias", B.rd);
                const float* w2 = t.get(p + ".se.conv_expand.weight", (int64_t)B.mid * B.rd);
                const float* b2 = t.get(p + ".se.conv_expand.bias", B.mid);
                if (w1 && b1 && w2 && b2) {
                    std::vector<float> w2t((size_t)B.rd * B.mid);
                    for (int c = 0; c < B.mid; ++c)
                        for (int j = 0; j < B.rd; ++j) w2t[(size_t)j * B.mid + c] = w2[(size_t)c * B.rd + j];
                    o.se_w1 = pack_f32(a, w1, (size_t)B.rd * B.mid);
                    o.se_b1 = pack_f32(a, b1, B.rd);
                    o.se_w2t = pack_f32(a, w2t.data(), w2t.size());
                    o.se_b2 = pack_f32(a, b2, B.mid);
                }
            }
            if (pack_pw(t, conv_pr, bn_pr, B.cout, B.mid, dtype, a, o.proj_w, o.proj_b) && B.proj_pack > 1)
                pack_pw_diag(a, o.proj_w, o.proj_b, B.cout, B.mid, B.proj_pack, o.proj_wp, o.proj_bp);
            cin = B.cout;
        }
    }
    size_t head_w_off = 0, head_b_off = 0;
    pack_pw(t, "backbone.3", "backbone.4", 1280, 320, dtype, a, head_w_off, head_b_off);
    size_t hoff[8] = {0};
    {
        const char* keys[8] = {"temporal_attention.0.weight", "temporal_attention.0.bias", "temporal_attention.2.weight",
                               "temporal_attention.2.bias", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"};
        const int64_t n[8] = {64 * 1280, 64, 64, 1, 256 * 1280, 256, 2 * 256, 2};
        for (int i = 0; i < 8; ++i) {
            const float* p = t.get(keys[i], n[i]);
            if (p) hoff[i] = pack_f32(a, p, (size_t)n[i]);
        }
    }
    if (!t.missing.empty()) return fail(DFD_EKEY, "dfd_pack_weights: state_dict tensor " + t.missing);

    void* dev = nullptr;
    DFD_CUDA(cudaMalloc(&dev, a.bytes.size()), "cudaMalloc(weights)");
    cudaError_t e = cudaMemcpy(dev, a.bytes.data(), a.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dev); return cuda_fail(e, "cudaMemcpy(weights)"); }
    uint8_t* d = reinterpret_cast<uint8_t*>(dev);
    auto F = [&](size_t o) { return reinterpret_cast<float*>(d + o); };
    W.arena = dev; W.arena_bytes = a.bytes.size();
    W.stem_w = F(stem_w_off); W.stem_b = F(stem_b_off); W.stem_w16 = d + stem_w16_off;
    W.stem_wrow = d + stem_wrow_off; W.stem_b4 = F(stem_b4_off);
    for (int i = 0; i < kNumBlocks; ++i) {
        BlockW& B = W.blocks[i]; const Off& o = off[i];
        B.exp_w = B.has_expand ? d + o.exp_w : nullptr; B.exp_b = B.has_expand ? F(o.exp_b) : nullptr;
        B.dw_w = F(o.dw_w); B.dw_b = F(o.dw_b);
        B.se_w1 = F(o.se_w1); B.se_b1 = F(o.se_b1); B.se_w2t = F(o.se_w2t); B.se_b2 = F(o.se_b2);
        B.proj_w = d + o.proj_w; B.proj_b = F(o.proj_b);
        B.exp_wp = B.exp_pack > 1 ? d + o.exp_wp : nullptr; B.exp_bp = B.exp_pack > 1 ? F(o.exp_bp) : nullptr;
        B.proj_wp = B.proj_pack > 1 ? d + o.proj_wp : nullptr; B.proj_bp = B.proj_pack > 1 ? F(o.proj_bp) : nullptr;
    }
    W.head_w = d + head_w_off; W.head_b = F(head_b_off);
    W.hw.att_w1 = F(hoff[0]); W.hw.att_b1 = F(hoff[1]); W.hw.att_w2 = F(hoff[2]); W.hw.att_b2 = F(hoff[3]);
    W.hw.fc1_w = F(hoff[4]); W.hw.fc1_b = F(hoff[5]); W.hw.fc2_w = F(hoff[6]); W.hw.fc2_b = F(hoff[7]);
    *out = new dfd_weights(W);
    return DFD_OK;
}

void dfd_free_weights(dfd_weights_t* w) {
    if (!w) return;
    if (w->arena) cudaFree(w->arena);
    delete w;
}

int dfd_preprocess_u8hwc_to_nchw(const uint8_t* d_in, void* d_out, int64_t frames, int H, int W, int dtype, void* stream) {
    if (!d_in || !d_out) return fail(DFD_EINVAL, "dfd_preprocess: null pointer");
    if (frames < 0 || H <= 0 || W <= 0 || ((int64_t)H * W) % 16 != 0) return fail(DFD_EINVAL, "dfd_preprocess: H*W must be a positive multiple of 16");
    if (dtype != DFD_DTYPE_BF16 && dtype != DFD_DTYPE_FP16) return fail(DFD_EINVAL, "dfd_preprocess: unknown dtype");
    g_launches = 0;
    prof_next(KC_PREPROCESS, (double)frames * H * W * 3 * 3, 0, (cudaStream_t)stream);
    DFD_LAUNCH(dfd::launch_preprocess(d_in, d_out, frames, H, W, dtype, (cudaStream_t)stream), "preprocess kernel");
    return DFD_OK;
}

#pragma GCC visibility pop
}  // extern "C"

// ---------------------------------------------------------------------------------------------------
// workspace plan + layer schedule
// ---------------------------------------------------------------------------------------------------
namespace {

// project layers on maps with at least this many pixels fold the SE gate into per-frame weights (gemm_tc.cu)
constexpr int kFrameWeightsMinHW = 784;

struct Plan {
    size_t io_elems, e_elems, d_elems, part_floats, gate_floats, wf_elems;     // per frame
    size_t per_frame_bytes() const {
        auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
        return 2 * up(io_elems * 2) + up(e_elems * 2) + up(d_elems * 2) + up(part_floats * 4) + up(gate_floats * 4) + up(wf_elems * 2);
    }
};

bool shape_ok(int H, int W) {
    return H > 0 && W > 0 && (H % 32) == 0 && (W % 32) == 0 && (H / 32) * (W / 32) <= 128 && 128 / ((H / 32) * (W / 32)) <= 4;
}

Plan make_plan(int H, int W) {
    Plan p{};
    int h = H / 2, w = W / 2, cin = 32;
    p.io_elems = (size_t)h * w * 32;
    for (int s = 0; s < 7; ++s)
        for (int b = 0; b < kStages[s][0]; ++b) {
            const int k = kStages[s][1], st = (b == 0) ? kStages[s][2] : 1, mid = cin * kStages[s][3], cout = kStages[s][4];
            const int oh = (h + 2 * (k / 2) - k) / st + 1, ow = (w + 2 * (k / 2) - k) / st + 1;
            if (kStages[s][3] != 1) p.e_elems = std::max(p.e_elems, (size_t)h * w * mid);
            p.d_elems = std::max(p.d_elems, (size_t)oh * ow * mid);
            p.part_floats = std::max(p.part_floats, (size_t)dfd::dw_num_partials(oh, ow, mid, k, st) * mid);
            p.gate_floats = std::max(p.gate_floats, (size_t)mid);
            if (oh * ow >= kFrameWeightsMinHW) { const size_t r = pack_factor(mid); p.wf_elems = std::max(p.wf_elems, r * r * mid * cout); }
            p.io_elems = std::max(p.io_elems, (size_t)oh * ow * cout);
            h = oh; w = ow; cin = cout;
        }
    return p;
}

size_t in_frame_bytes(int in_kind, int H, int W) {
    const size_t px = (size_t)H * W * 3;
    return in_kind == DFD_IN_U8_HWC ? px : (in_kind == DFD_IN_F32_NCHW ? px * 4 : px * 2);
}

// r > 1: pixel-packed call (Wt / bias are the block-diagonal copies): M/r rows of r*K channels -> r*N outputs
int run_gemm(const void* A, const void* Wt, const float* bias, const float* gate, const void* R, void* D,
             int64_t M, int K, int N, int HW, int act, int dtype, cudaStream_t s, int r = 1) {
    prof_next(gate ? KC_GEMM_PROJECT : KC_GEMM_EXPAND, (double)M * (K + N + (R ? N : 0)) * 2, 2.0 * M * K * N, s);
    M /= r; K *= r; N *= r; HW /= r;
    DFD_LAUNCH(dfd::launch_gemm_tc(A, Wt, bias, gate, R, D, M, K, N, HW, act, dtype, s), "gemm (tcgen05)");
    return DFD_OK;
}

int run_trunk_chunk(const dfd_weights* w, const void* in, int in_kind, int64_t frames, int H, int W,
                    float* feat, uint8_t* ws, const Plan& plan, cudaStream_t s) {
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    uint8_t* io[2];
    io[0] = ws;                                  ws += up(plan.io_elems * 2) * frames;
    io[1] = ws;                                  ws += up(plan.io_elems * 2) * frames;
    uint8_t* bufE = ws;                          ws += up(plan.e_elems * 2) * frames;
    uint8_t* bufD = ws;                          ws += up(plan.d_elems * 2) * frames;
    float* part = reinterpret_cast<float*>(ws);  ws += up(plan.part_floats * 4) * frames;
    float* gate = reinterpret_cast<float*>(ws);  ws += up(plan.gate_floats * 4) * frames;
    uint8_t* bufWf = ws;
    const int dt = w->dtype;

    int h = H / 2, wd = W / 2, cur = 0;
    prof_next(KC_STEM, (double)frames * (in_frame_bytes(in_kind, H, W) + (double)h * wd * 32 * 2), 2.0 * frames * h * wd * 27 * 32, s);
    if (in_kind == DFD_IN_U8_HWC)      // uint8 crops: tensor prep folded into the tcgen05 stem; float inputs (the module's forward): CUDA-core stem
        DFD_LAUNCH(dfd::launch_stem_tc(reinterpret_cast<const uint8_t*>(in), w->stem_w16, w->stem_b, w->stem_wrow, w->stem_b4, io[cur], frames, H, W, dt, s), "stem kernel (tcgen05)");
    else
        DFD_LAUNCH(dfd::launch_stem(in, in_kind, w->stem_w, w->stem_b, io[cur], frames, H, W, dt, s), "stem kernel");
    for (int i = 0; i < kNumBlocks; ++i) {
        const BlockW& B = w->blocks[i];
        const void* x = io[cur];
        const void* e = x;
        // the early, HBM-bound blocks of the 224x224 network run expand 1x1 + depthwise as ONE kernel (mbconv_fused.cu): the
        // expanded tensor never goes to HBM; same rounding points, same partial-sum layout
        const bool fused = B.has_expand && dfd::mbconv_fused_supported(h, wd, B.cin, B.mid, B.k, B.stride);
        if (B.has_expand && !fused) {
            const int r = (B.exp_pack > 1 && (h * wd) % B.exp_pack == 0) ? B.exp_pack : 1;
            int rc = run_gemm(x, r > 1 ? B.exp_wp : B.exp_w, r > 1 ? B.exp_bp : B.exp_b, nullptr, nullptr, bufE,
                              frames * h * wd, B.cin, B.mid, h * wd, 1, dt, s, r);
            if (rc) return rc;
            e = bufE;
        }
        const int pad = B.k / 2;
        const int oh = (h + 2 * pad - B.k) / B.stride + 1, ow = (wd + 2 * pad - B.k) / B.stride + 1;
        if (fused) {
            prof_next(KC_MBCONV_FUSED, (double)frames * ((double)h * wd * B.cin + (double)oh * ow * B.mid) * 2,
                      2.0 * frames * ((double)h * wd * B.cin * B.mid + (double)oh * ow * B.mid * B.k * B.k), s);
            DFD_LAUNCH(dfd::launch_mbconv_fused(x, B.exp_w, B.exp_b, B.dw_w, B.dw_b, bufD, part, frames, h, wd, B.cin, B.mid, B.k, B.stride, dt, s),
                       "fused expand + depthwise kernel");
        } else {
            prof_next(KC_DWCONV, (double)frames * B.mid * ((double)h * wd + (double)oh * ow) * 2, 2.0 * frames * oh * ow * B.mid * B.k * B.k, s);
            DFD_LAUNCH(dfd::launch_dwconv(e, B.dw_w, B.dw_b, bufD, part, frames, h, wd, B.mid, B.k, B.stride, dt, s), "depthwise kernel");
        }
        const int nparts = dfd::dw_num_partials(oh, ow, B.mid, B.k, B.stride);
        prof_next(KC_SE, (double)frames * B.mid * (nparts + 1) * 4, 4.0 * frames * B.mid * B.rd, s);
        DFD_LAUNCH(dfd::launch_se(part, nparts, 1.0f / (float)(oh * ow), B.se_w1, B.se_b1, B.se_w2t, B.se_b2, gate,
                                  frames, B.mid, B.rd, s), "squeeze-excite kernel");
        if (oh * ow >= kFrameWeightsMinHW) {
            // big maps: SE gate folded into per-frame weights, ungated GEMM on frame-aligned tiles
            prof_next(KC_GEMM_PROJECT, (double)frames * B.cout * B.mid * 2 * 2, 0, s);
            const int r = (B.proj_pack > 1 && (oh * ow) % B.proj_pack == 0) ? B.proj_pack : 1;
            DFD_LAUNCH(dfd::launch_scale_weights(r > 1 ? B.proj_wp : B.proj_w, gate, bufWf, frames, B.cout * r, B.mid * r, B.mid, dt, s), "per-frame weight scaling");
            prof_next(KC_GEMM_PROJECT, (double)frames * oh * ow * (B.mid + B.cout + (B.has_skip ? B.cout : 0)) * 2, 2.0 * frames * oh * ow * B.mid * B.cout, s);
            DFD_LAUNCH(dfd::launch_gemm_tc_framew(bufD, bufWf, r > 1 ? B.proj_bp : B.proj_b, B.has_skip ? x : nullptr, io[cur ^ 1],
                                                  frames * oh * ow / r, B.mid * r, B.cout * r, oh * ow / r, dt, s), "gemm (tcgen05, per-frame weights)");
        } else {
            int rc = run_gemm(bufD, B.proj_w, B.proj_b, gate, B.has_skip ? x : nullptr, io[cur ^ 1],
                              frames * oh * ow, B.mid, B.cout, oh * ow, 0, dt, s);
            if (rc) return rc;
        }
        cur ^= 1; h = oh; wd = ow;
    }
    prof_next(KC_GEMM_HEAD_POOL, (double)frames * h * wd * 320 * 2 + (double)frames * 1280 * 4, 2.0 * frames * h * wd * 320 * 1280, s);
    DFD_LAUNCH(dfd::launch_gemm_tc_pool(io[cur], w->head_w, w->head_b, feat, frames * h * wd, 320, 1280, h * wd, dt, s), "head (tcgen05)");
    return DFD_OK;
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int dfd_workspace_bytes(int64_t frames, int H, int W, size_t* bytes) {
    if (!bytes || frames < 0) return fail(DFD_EINVAL, "dfd_workspace_bytes: bad argument");
    if (!shape_ok(H, W)) return fail(DFD_EINVAL, "unsupported crop size: H and W must be multiples of 32 with 32 <= (H/32)*(W/32) <= 128 (224x224 is the supported size)");
    const Plan p = make_plan(H, W);
    const int64_t chunk = std::min<int64_t>(std::max<int64_t>(frames, 1), chunk_frames());
    *bytes = p.per_frame_bytes() * (size_t)chunk + 256;
    return DFD_OK;
}

int dfd_effnet_b0_features(const dfd_weights_t* w, const void* d_in, int in_kind, int64_t frames, int H, int W,
                           float* d_features, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!w || !d_in || !d_features || !d_workspace) return fail(DFD_EINVAL, "dfd_effnet_b0_features: null pointer");
    if (in_kind < 0 || in_kind > 2) return fail(DFD_EINVAL, "dfd_effnet_b0_features: unknown input layout");
    if (frames < 0) return fail(DFD_EINVAL, "dfd_effnet_b0_features: negative frame count");
    size_t need = 0;
    int rc = dfd_workspace_bytes(frames, H, W, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return fail(DFD_ENOMEM, "dfd_effnet_b0_features: workspace too small (" + std::to_string(workspace_bytes) + " < " + std::to_string(need) + ")");
    g_launches = 0;
    const Plan p = make_plan(H, W);
    const int64_t chunk = chunk_frames();
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    const size_t fb = in_frame_bytes(in_kind, H, W);
    for (int64_t f0 = 0; f0 < frames; f0 += chunk) {
        const int64_t n = std::min(chunk, frames - f0);
        rc = run_trunk_chunk(w, reinterpret_cast<const uint8_t*>(d_in) + (size_t)f0 * fb, in_kind, n, H, W,
                             d_features + (size_t)f0 * DFD_FEATURE_DIM, ws, p, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return DFD_OK;
}

int dfd_attn_pool_head(const dfd_weights_t* w, const float* d_features, const int32_t* d_offsets, int64_t videos,
                       int64_t frames, int use_attention, float* d_logits, float* d_frame_scores, void* stream) {
    if (!w || !d_features || !d_offsets || !d_logits) return fail(DFD_EINVAL, "dfd_attn_pool_head: null pointer");
    if (videos < 0 || frames < 0 || frames > INT32_MAX) return fail(DFD_EINVAL, "dfd_attn_pool_head: bad video / frame count");
    g_launches = 0;
    prof_next(KC_POOL_HEAD, (double)frames * 1280 * 4 * 2, 2.0 * frames * 1280 * 64 + 2.0 * videos * 1280 * 256, (cudaStream_t)stream);
    DFD_LAUNCH(dfd::launch_pool_head(w->hw, d_features, d_offsets, videos, frames, DFD_FEATURE_DIM, use_attention, d_logits, d_frame_scores,
                                     (cudaStream_t)stream), "pool+head kernel");
    return DFD_OK;
}

int dfd_score_workspace_bytes(int64_t frames, int H, int W, size_t* bytes) {
    size_t b = 0;
    int rc = dfd_workspace_bytes(frames, H, W, &b);
    if (rc) return rc;
    *bytes = b + (size_t)std::max<int64_t>(frames, 1) * DFD_FEATURE_DIM * 4 + 256;
    return DFD_OK;
}

int dfd_score_videos(const dfd_weights_t* w, const void* d_in, int in_kind, const int32_t* d_offsets, int64_t videos,
                     int64_t frames, int H, int W, int use_attention, float* d_logits, float* d_frame_scores,
                     float* d_features_out, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!w || !d_in || !d_offsets || !d_logits || !d_workspace) return fail(DFD_EINVAL, "dfd_score_videos: null pointer");
    size_t need = 0, trunk = 0;
    int rc = dfd_score_workspace_bytes(frames, H, W, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return fail(DFD_ENOMEM, "dfd_score_videos: workspace too small (" + std::to_string(workspace_bytes) + " < " + std::to_string(need) + ")");
    dfd_workspace_bytes(frames, H, W, &trunk);
    uint8_t* ws = reinterpret_cast<uint8_t*>(d_workspace);
    float* feat = d_features_out;
    if (!feat) feat = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(ws + trunk) + 255) & ~uintptr_t(255));
    rc = dfd_effnet_b0_features(w, d_in, in_kind, frames, H, W, feat, d_workspace, trunk, stream);
    if (rc) return rc;
    const int n = g_launches;
    rc = dfd_attn_pool_head(w, feat, d_offsets, videos, frames, use_attention, d_logits, d_frame_scores, stream);
    g_launches += n;
    return rc;
}

// ---- profiling ---------------------------------------------------------------------------------------
int dfd_profile_enable(int on) {
    for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
    g_prof.clear();
    g_prof_on = on != 0;
    return DFD_OK;
}
int dfd_profile_collect(dfd_profile_entry* out, int max_entries, int* n_out) {
    if (!out || !n_out || max_entries < KC_COUNT) return fail(DFD_EINVAL, "dfd_profile_collect: need room for all kernel classes");
    for (int c = 0; c < KC_COUNT; ++c) {
        memset(&out[c], 0, sizeof(out[c]));
        strncpy(out[c].name, kClassNames[c], sizeof(out[c].name) - 1);
    }
    for (auto& r : g_prof) {
        DFD_CUDA(cudaEventSynchronize(r.b), "cudaEventSynchronize(profile)");
        float ms = 0.f;
        DFD_CUDA(cudaEventElapsedTime(&ms, r.a, r.b), "cudaEventElapsedTime(profile)");
        out[r.cls].launches += 1; out[r.cls].ms += ms; out[r.cls].bytes += r.bytes; out[r.cls].flops += r.flops;
    }
    *n_out = KC_COUNT;
    return DFD_OK;
}

// ---- kernel-level entry points (include/dfd_b200_kernels.h) ------------------------------------------
int dfd_k_stem(const void* d_in, int in_kind, const float* d_w, const float* d_bias, void* d_out, int64_t frames,
               int H, int W, int dtype, void* stream) {
    g_launches = 0;
    DFD_LAUNCH(dfd::launch_stem(d_in, in_kind, d_w, d_bias, d_out, frames, H, W, dtype, (cudaStream_t)stream), "stem kernel");
    return DFD_OK;
}
int dfd_k_stem_tc(const uint8_t* d_in, const float* h_w27x32, const float* d_bias, void* d_out, int64_t frames,
                  int H, int W, int dtype, void* stream) {
    if (!d_in || !h_w27x32 || !d_bias || !d_out) return fail(DFD_EINVAL, "dfd_k_stem_tc: null pointer");
    float hb[32];
    DFD_CUDA(cudaMemcpy(hb, d_bias, sizeof(hb), cudaMemcpyDeviceToHost), "cudaMemcpy(stem bias)");
    HostArena a;
    const size_t off = pack_stem16(a, h_w27x32, dtype);
    size_t wrow_off = 0, b4_off = 0;
    pack_stem_row(a, h_w27x32, hb, wrow_off, b4_off);
    uint8_t* dw = nullptr;
    DFD_CUDA(cudaMalloc(&dw, a.bytes.size()), "cudaMalloc(stem operands)");
    cudaError_t e = cudaMemcpy(dw, a.bytes.data(), a.bytes.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dw); return cuda_fail(e, "cudaMemcpy(stem operands)"); }
    g_launches = 0;
    { ProfScope ps; e = dfd::launch_stem_tc(d_in, dw + off, d_bias, dw + wrow_off, reinterpret_cast<const float*>(dw + b4_off), d_out,
                                            frames, H, W, dtype, (cudaStream_t)stream); }
    ++g_launches;
    cudaError_t e2 = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(dw);
    if (e != cudaSuccess) return cuda_fail(e, "stem kernel (tcgen05)");
    if (e2 != cudaSuccess) return cuda_fail(e2, "stem kernel (tcgen05) sync");
    return DFD_OK;
}
int dfd_k_pack_stem_row(const float* h_w27x32, const float* h_bias32, uint16_t* h_wrow, float* h_bias4) {
    if (!h_w27x32 || !h_bias32 || !h_wrow || !h_bias4) return fail(DFD_EINVAL, "dfd_k_pack_stem_row: null pointer");
    HostArena a;
    size_t w_off = 0, b4_off = 0;
    pack_stem_row(a, h_w27x32, h_bias32, w_off, b4_off);
    memcpy(h_wrow, a.bytes.data() + w_off, 2 * 32 * 32 * 2);
    memcpy(h_bias4, a.bytes.data() + b4_off, 4 * 32 * 4);
    return DFD_OK;
}
void dfd_k_set_dw_channel_block(int cb) { dfd::dw_march_set_cb(cb); }
int dfd_k_dw_num_partials(int OH, int OW, int C, int k, int stride) { return dfd::dw_num_partials(OH, OW, C, k, stride); }
int dfd_k_dwconv(const void* d_in, const float* d_w, const float* d_bias, void* d_out, float* d_partials,
                 int64_t frames, int H, int W, int C, int k, int stride, int dtype, void* stream) {
    g_launches = 0;
    DFD_LAUNCH(dfd::launch_dwconv(d_in, d_w, d_bias, d_out, d_partials, frames, H, W, C, k, stride, dtype, (cudaStream_t)stream), "depthwise kernel");
    return DFD_OK;
}
int dfd_k_se(const float* d_partials, int nparts, float inv_hw, const float* d_w1, const float* d_b1, const float* d_w2t,
             const float* d_b2, float* d_gate, int64_t frames, int C, int rd, void* stream) {
    g_launches = 0;
    DFD_LAUNCH(dfd::launch_se(d_partials, nparts, inv_hw, d_w1, d_b1, d_w2t, d_b2, d_gate, frames, C, rd, (cudaStream_t)stream), "squeeze-excite kernel");
    return DFD_OK;
}
int dfd_k_gemm(const void* d_A, const void* d_W, const float* d_bias, const float* d_gate, const void* d_R, void* d_D,
               int64_t M, int K, int N, int HW, int act, int dtype, int impl, void* stream) {
    g_launches = 0;
    if (impl == 2) {                 // gated project conv via per-frame weights (needs gate, act == 0, M % HW == 0)
        if (!d_gate || act || HW <= 0 || (M % HW)) return fail(DFD_EINVAL, "dfd_k_gemm impl 2: needs a gate, act == 0 and M % HW == 0");
        void* wf = nullptr;
        DFD_CUDA(cudaMalloc(&wf, (size_t)(M / HW) * N * K * 2), "cudaMalloc(per-frame weights)");
        cudaError_t e = dfd::launch_scale_weights(d_W, d_gate, wf, M / HW, N, K, K, dtype, (cudaStream_t)stream);
        if (e == cudaSuccess) e = dfd::launch_gemm_tc_framew(d_A, wf, d_bias, d_R, d_D, M, K, N, HW, dtype, (cudaStream_t)stream);
        g_launches = 2;
        cudaError_t e2 = cudaStreamSynchronize((cudaStream_t)stream);
        cudaFree(wf);
        if (e != cudaSuccess) return cuda_fail(e, "gemm (tcgen05, per-frame weights)");
        if (e2 != cudaSuccess) return cuda_fail(e2, "gemm (tcgen05, per-frame weights) sync");
        return DFD_OK;
    }
    if (impl == 3) {                 // CTA-pair GEMM (gemm_pair.cu, the ViT contractions): no gate, act 0 or 2 (GELU), N % 256 == 0;
                                     // d_R == d_D: fp32 residual stream updated in place (X += A W^T + bias, act 0)
        if (d_gate || (d_R && (d_R != d_D || act)) || (act != 0 && act != 2) || !dfd::gemm_pair_supported(M, K, N))
            return fail(DFD_EINVAL, "dfd_k_gemm impl 3: no gate, act 0 or 2, N a multiple of 256, residual only in place (d_R == d_D, fp32, act 0)");
        if (d_R) DFD_LAUNCH(dfd::launch_gemm_pair_residual(d_A, d_W, d_bias, (float*)d_D, M, K, N, dtype, (cudaStream_t)stream), "gemm (tcgen05, CTA pairs, fp32 residual)");
        else DFD_LAUNCH(dfd::launch_gemm_pair(d_A, d_W, d_bias, d_D, M, K, N, act, dtype, (cudaStream_t)stream), "gemm (tcgen05, CTA pairs)");
        return DFD_OK;
    }
    if (impl != 0) return fail(DFD_EINVAL, "dfd_k_gemm: impl must be 0 (tcgen05 GEMM), 2 (per-frame weights) or 3 (CTA pairs)");
    DFD_LAUNCH(dfd::launch_gemm_tc(d_A, d_W, d_bias, d_gate, d_R, d_D, M, K, N, HW, act, dtype, (cudaStream_t)stream), "gemm (tcgen05)");
    return DFD_OK;
}
int dfd_k_gemm_pool(const void* d_A, const void* d_W, const float* d_bias, float* d_feat, int64_t M, int K, int N,
                    int HW, int dtype, int impl, void* stream) {
    g_launches = 0;
    if (impl != 0) return fail(DFD_EINVAL, "dfd_k_gemm_pool: impl must be 0 (tcgen05 GEMM)");
    DFD_LAUNCH(dfd::launch_gemm_tc_pool(d_A, d_W, d_bias, d_feat, M, K, N, HW, dtype, (cudaStream_t)stream), "head (tcgen05)");
    return DFD_OK;
}

// relu(conv3x3(relu(conv1x1(x)))) of a resnet bottleneck through the haloed-map path (gemm_tc.cu CONV variants): one memset,
// one scattering pointwise GEMM, one implicit 3x3 GEMM (the path resnet.cu runs for its 13 stride-1 3x3 convolutions).
int dfd_k_conv1x1_conv3x3(const void* d_in, const void* d_w1, const float* d_b1, const void* d_w2, const float* d_b2, void* d_out,
                          int64_t frames, int H, int W, int K, int C, int N, int dtype, void* d_pad, size_t pad_bytes, void* stream) {
    g_launches = 0;
    if (!d_in || !d_w1 || !d_b1 || !d_w2 || !d_b2 || !d_out || !d_pad) return fail(DFD_EINVAL, "dfd_k_conv1x1_conv3x3: null pointer");
    if (frames <= 0 || H <= 0 || W <= 0 || (C % 64) || (K & 7) || (N & 7)) return fail(DFD_EINVAL, "dfd_k_conv1x1_conv3x3: C must be a multiple of 64, K and N of 8");
    const size_t need = (size_t)dfd::conv3x3_padded_rows(frames, H, W) * C * 2;
    if (pad_bytes < need) return fail(DFD_ENOMEM, "dfd_k_conv1x1_conv3x3: haloed map needs " + std::to_string(need) + " bytes");
    DFD_CUDA(cudaMemsetAsync(d_pad, 0, need, (cudaStream_t)stream), "cudaMemsetAsync(haloed map)");
    DFD_LAUNCH(dfd::launch_gemm_tc_padout(d_in, d_w1, d_b1, d_pad, frames, H, W, K, C, dtype, (cudaStream_t)stream), "conv1x1 -> haloed map");
    DFD_LAUNCH(dfd::launch_gemm_tc_conv3x3(d_pad, d_w2, d_b2, d_out, frames, H, W, C, N, dtype, (cudaStream_t)stream), "implicit conv3x3");
    return DFD_OK;
}

// expand 1x1 + BN + SiLU fused into the row-marching depthwise kernel (mbconv_fused.cu): the early blocks of the trunk
int dfd_k_mbconv_fused(const void* d_x, const void* d_we, const float* d_be, const float* d_w, const float* d_bias, void* d_out,
                       float* d_partials, int64_t frames, int H, int W, int cin, int mid, int k, int stride, int dtype, void* stream) {
    g_launches = 0;
    if (!d_x || !d_we || !d_be || !d_w || !d_bias || !d_out || !d_partials) return fail(DFD_EINVAL, "dfd_k_mbconv_fused: null pointer");
    if (!dfd::mbconv_fused_supported(H, W, cin, mid, k, stride)) return fail(DFD_EINVAL, "dfd_k_mbconv_fused: unsupported block shape");
    DFD_LAUNCH(dfd::launch_mbconv_fused(d_x, d_we, d_be, d_w, d_bias, d_out, d_partials, frames, H, W, cin, mid, k, stride, dtype, (cudaStream_t)stream),
               "fused expand + depthwise kernel");
    return DFD_OK;
}
int dfd_k_mbconv_fused_supported(int H, int W, int cin, int mid, int k, int stride) { return dfd::mbconv_fused_supported(H, W, cin, mid, k, stride) ? 1 : 0; }

// HOST-ONLY: the row maps of the haloed layout, computed by the functions the kernels use (csrc/conv_map.h)
int64_t dfd_k_conv3x3_maps(int frames, int H, int W, int cpk, int64_t* h_pad_row, int64_t* h_out_row, int32_t* h_tap_row, int32_t* h_tap_col) {
    if (frames <= 0 || H <= 0 || W <= 0 || cpk <= 0) return fail(DFD_EINVAL, "dfd_k_conv3x3_maps: bad geometry");
    if (h_pad_row)
        for (int64_t m = 0; m < (int64_t)frames * H * W; ++m) h_pad_row[m] = dfd::conv_pad_row((uint32_t)m, (uint32_t)H, (uint32_t)W);
    if (h_out_row)
        for (int64_t p = 0; p < (int64_t)frames * (H + 2) * (W + 2); ++p) {
            int64_t o;
            h_out_row[p] = dfd::conv_unpad_row((uint32_t)p, (uint32_t)H, (uint32_t)W, &o) ? o : -1;
        }
    if (h_tap_row && h_tap_col) {
        dfd::ConvTapIter it; it.init(0);
        for (int kb = 0; kb < 9 * cpk; ++kb) { h_tap_row[kb] = it.row; h_tap_col[kb] = it.ck * 64; it.next(cpk, W); }
    }
    return dfd::conv3x3_padded_rows(frames, H, W);
}

#pragma GCC visibility pop
}  // extern "C"
