// K5 — LogicRNNLSTM temporal head (reference src/RNNModel.py:5-41 LogicCell, :43-147 LogicRNNLSTM), used on
// EfficientNet features in BASELINE config 3 (wiring: src/evaluate.py:143-192).
//
// Per time step and layer the reference runs six Linear(in+H -> H) on cat(x, h) and one Linear(H -> H) on h
// (RNNModel.py:21-31).  Here they are ONE tcgen05 GEMM per (step, layer): the seven weight matrices are stacked
// to W[7H, K] (the `not` gate gets zero columns over the x part; for layers >= 1 the reference feeds
// cat(h_temp, h_temp), so the two column halves are pre-added and K = H), operands 16-bit, fp32 accumulate, fp32
// gate pre-activations out (gemm_tc.cu, F32OUT).  A fused cell kernel does the gate math in fp32
// (RNNModel.py:24-39) and writes h straight into the next GEMM's A operand, so nothing is re-laid-out.
// The attention pool over time + classifier + sigmoid (:128-133) is one CTA per sequence.
// Quirks kept: a single (h, c) pair is threaded through all layers (:103-115); the length mask zeroes outputs
// but the recurrence runs over padded steps (:120-125); softmax runs over all T (:128).  The reference's
// sort-by-length without un-sorting (:92-95) is done by the Python wrapper with the same torch call.
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/dfd_b200.h"
#include "common.cuh"
#include "kernels.h"

struct dfd_rnn_weights {
    int dtype, input_size, hidden, layers;
    std::vector<void*> w;          // per layer: [7H][K_l] 16-bit
    std::vector<float*> b;         // per layer: [7H]
    float *att_w1t, *att_b1, *att_w2, *att_b2;     // [H][H] transposed (k-major), [H], [H], [1]
    float *cls_w1t, *cls_b1, *cls_w2, *cls_b2;     // [H][H] transposed, [H], [H], [1]
    void* arena;
};

namespace dfd {

// x fp32 [B][T][IN] -> A0[t][b][0:IN] (16-bit, row stride K0); zero the h part of step 0 and the cell state
template <typename T>
__global__ void rnn_pack_x_kernel(const float* __restrict__ x, T* __restrict__ a0, float* __restrict__ c,
                                  int B, int Tn, int IN, int H, int K0) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t nx = (int64_t)B * Tn * IN;
    if (i < nx) {
        const int k = (int)(i % IN);
        const int64_t bt = i / IN;
        const int t = (int)(bt % Tn), b = (int)(bt / Tn);
        a0[((size_t)t * B + b) * K0 + k] = Half16<T>::from_float(x[i]);
    } else if (i < nx + (int64_t)B * H) {
        const int64_t j = i - nx;
        const int b = (int)(j / H), k = (int)(j % H);
        a0[(size_t)b * K0 + IN + k] = Half16<T>::from_float(0.f);
        c[j] = 0.f;
    }
}

// DFD_RNN_KERNELS_BEGIN   (tools/host_emul/ runs the two kernels below, unchanged, on CPU threads)
// LogicCell gate math (RNNModel.py:24-39).  g fp32 [B][7H] = (and, or, forget, input, cell, output, not) pre-activations.
template <typename T>
__global__ void rnn_cell_kernel(const float* __restrict__ g, float* __restrict__ c, T* __restrict__ h16, int h16_stride,
                                float* __restrict__ out, int out_stride, const int* __restrict__ lengths, int t,
                                int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, k = i - b * H;
    const float* gr = g + (size_t)b * 7 * H + k;
    const float and_o = 1.f / (1.f + expf(-gr[0]));
    const float or_o = 1.f / (1.f + expf(-gr[H]));
    const float forget = 1.f / (1.f + expf(-gr[2 * H]));
    const float input_g = 1.f / (1.f + expf(-gr[3 * H]));
    const float cell_t = tanhf(gr[4 * H]);
    const float output = 1.f / (1.f + expf(-gr[5 * H]));
    const float not_o = tanhf(gr[6 * H]);
    const float cell_new = forget * c[i] + input_g * cell_t;
    const float cell_logic = and_o * cell_new + or_o * not_o;
    const float h_new = output * tanhf(cell_logic);
    c[i] = cell_logic;
    if (h16) h16[(size_t)b * h16_stride + k] = Half16<T>::from_float(h_new);
    if (out) {
        const float m = (lengths == nullptr || t < lengths[b]) ? 1.f : 0.f;          // :120-125
        out[(size_t)b * out_stride + k] = h_new * m;
    }
}

// attention over time + classifier + sigmoid (RNNModel.py:128-133); one CTA per sequence, thread j = hidden unit j
__global__ void rnn_head_kernel(const float* __restrict__ outs, int Tn, int H,
                                const float* __restrict__ w1t, const float* __restrict__ b1, const float* __restrict__ w2,
                                const float* __restrict__ b2, const float* __restrict__ c1t, const float* __restrict__ cb1,
                                const float* __restrict__ c2, const float* __restrict__ cb2, float* __restrict__ prob) {
    extern __shared__ float sm[];
    float* s_o = sm;                 // [T][H]
    float* s_red = sm + Tn * H;      // [blockDim/32]
    float* s_a = s_red + 32;         // [T]
    float* s_ctx = s_a + Tn;         // [H]
    const int b = blockIdx.x, j = threadIdx.x, warp = j >> 5, lane = j & 31, nw = blockDim.x >> 5;
    const float* o = outs + (size_t)b * Tn * H;
    for (int i = j; i < Tn * H; i += blockDim.x) s_o[i] = o[i];
    __syncthreads();
    for (int t = 0; t < Tn; ++t) {                                  // score_t = w2 . tanh(W1 o_t + b1) + b2
        float part = 0.f;
        for (int jj = j; jj < H; jj += blockDim.x) {
            float acc = b1[jj];
            for (int k = 0; k < H; ++k) acc = fmaf(s_o[t * H + k], __ldg(w1t + (size_t)k * H + jj), acc);
            part += tanhf(acc) * w2[jj];
        }
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0) s_red[warp] = part;
        __syncthreads();
        if (j == 0) { float s = b2[0]; for (int w = 0; w < nw; ++w) s += s_red[w]; s_a[t] = s; }
        __syncthreads();
    }
    if (j == 0) {                                                   // softmax over T (dim=1)
        float mx = -INFINITY; for (int t = 0; t < Tn; ++t) mx = fmaxf(mx, s_a[t]);
        float sum = 0.f; for (int t = 0; t < Tn; ++t) { s_a[t] = expf(s_a[t] - mx); sum += s_a[t]; }
        for (int t = 0; t < Tn; ++t) s_a[t] /= sum;
    }
    __syncthreads();
    for (int k = j; k < H; k += blockDim.x) {                       // context = sum_t a_t * o_t
        float acc = 0.f;
        for (int t = 0; t < Tn; ++t) acc += s_a[t] * s_o[t * H + k];
        s_ctx[k] = acc;
    }
    __syncthreads();
    float part = 0.f;                                               // classifier: Linear -> ReLU -> Linear(H,1)
    for (int jj = j; jj < H; jj += blockDim.x) {
        float acc = cb1[jj];
        for (int k = 0; k < H; ++k) acc = fmaf(s_ctx[k], __ldg(c1t + (size_t)k * H + jj), acc);
        part += fmaxf(acc, 0.f) * c2[jj];
    }
    for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
    if (lane == 0) s_red[warp] = part;
    __syncthreads();
    if (j == 0) { float s = cb2[0]; for (int w = 0; w < nw; ++w) s += s_red[w]; prob[b] = 1.f / (1.f + expf(-s)); }
}

// DFD_RNN_KERNELS_END

}  // namespace dfd

namespace {
thread_local std::string g_rnn_err;
int rfail(int code, const std::string& m) { g_rnn_err = m; return code; }
uint16_t h16(float v, int dtype) {
    if (dtype == DFD_DTYPE_FP16) { __half h = __float2half_rn(v); uint16_t u; memcpy(&u, &h, 2); return u; }
    __nv_bfloat16 h = __float2bfloat16_rn(v); uint16_t u; memcpy(&u, &h, 2); return u;
}
}  // namespace

extern "C" {
#pragma GCC visibility push(default)

const char* dfd_rnn_last_error(void) { return g_rnn_err.c_str(); }

int dfd_rnn_pack_weights(int n, const char* const* names, const float* const* data, const int64_t* numel,
                         int input_size, int hidden, int layers, int dtype, dfd_rnn_weights_t** out) {
    if (!names || !data || !numel || !out || n <= 0) return rfail(DFD_EINVAL, "dfd_rnn_pack_weights: null argument");
    if (input_size % 8 || hidden % 32 || hidden > 1024 || layers < 1) return rfail(DFD_EINVAL, "dfd_rnn_pack_weights: input_size % 8 == 0 and hidden % 32 == 0 (<= 1024) required");
    if (dtype != DFD_DTYPE_BF16 && dtype != DFD_DTYPE_FP16) return rfail(DFD_EINVAL, "dfd_rnn_pack_weights: unknown dtype");
    std::unordered_map<std::string, std::pair<const float*, int64_t>> t;
    for (int i = 0; i < n; ++i) if (names[i]) t[names[i]] = {data[i], numel[i]};
    std::string missing;
    auto get = [&](const std::string& k, int64_t ne) -> const float* {
        auto it = t.find(k);
        if (it == t.end() || it->second.second != ne) { if (missing.empty()) missing = k; return nullptr; }
        return it->second.first;
    };
    const int H = hidden;
    std::vector<uint8_t> host;
    auto alloc = [&](size_t nb) { size_t o = (host.size() + 255) & ~size_t(255); host.resize(o + nb, 0); return o; };
    std::vector<size_t> w_off(layers), b_off(layers);
    const char* gates[6] = {"and_gate", "or_gate", "forget_gate", "input_gate", "cell_gate", "output_gate"};   // RNNModel.py:11-19
    for (int l = 0; l < layers; ++l) {
        const int in_l = l == 0 ? input_size : H, Kl = l == 0 ? input_size + H : H;
        const std::string p = "logic_cells." + std::to_string(l) + ".";
        w_off[l] = alloc((size_t)7 * H * Kl * 2); b_off[l] = alloc((size_t)7 * H * 4);
        for (int gi = 0; gi < 7; ++gi) {
            const bool is_not = gi == 6;
            const std::string key = p + (is_not ? "not_gate" : gates[gi]);
            const float* w = get(key + ".weight", (int64_t)H * (is_not ? H : in_l + H));
            const float* b = get(key + ".bias", H);
            if (!w || !b) continue;
            uint16_t* wd = reinterpret_cast<uint16_t*>(host.data() + w_off[l]) + (size_t)gi * H * Kl;
            float* bd = reinterpret_cast<float*>(host.data() + b_off[l]) + (size_t)gi * H;
            for (int j = 0; j < H; ++j) {
                bd[j] = b[j];
                for (int k = 0; k < Kl; ++k) {
                    float v;
                    if (l == 0) v = is_not ? (k < input_size ? 0.f : w[(size_t)j * H + (k - input_size)]) : w[(size_t)j * Kl + k];
                    else v = is_not ? w[(size_t)j * H + k] : w[(size_t)j * 2 * H + k] + w[(size_t)j * 2 * H + H + k];   // cat(h, h)
                    wd[(size_t)j * Kl + k] = h16(v, dtype);
                }
            }
        }
    }
    size_t hoff[8] = {0};
    {
        const char* keys[8] = {"attention.0.weight", "attention.0.bias", "attention.2.weight", "attention.2.bias",
                               "classifier.0.weight", "classifier.0.bias", "classifier.3.weight", "classifier.3.bias"};
        const int64_t ne[8] = {(int64_t)H * H, H, H, 1, (int64_t)H * H, H, H, 1};
        for (int i = 0; i < 8; ++i) {
            const float* p = get(keys[i], ne[i]);
            hoff[i] = alloc((size_t)ne[i] * 4);
            if (!p) continue;
            float* d = reinterpret_cast<float*>(host.data() + hoff[i]);
            if (i == 0 || i == 4) { for (int j = 0; j < H; ++j) for (int k = 0; k < H; ++k) d[(size_t)k * H + j] = p[(size_t)j * H + k]; }
            else memcpy(d, p, (size_t)ne[i] * 4);
        }
    }
    if (!missing.empty()) return rfail(DFD_EKEY, "dfd_rnn_pack_weights: state_dict tensor " + missing + " absent or wrong size");
    void* dev = nullptr;
    if (cudaMalloc(&dev, host.size()) != cudaSuccess) return rfail(DFD_ECUDA, "cudaMalloc(rnn weights) failed");
    if (cudaMemcpy(dev, host.data(), host.size(), cudaMemcpyHostToDevice) != cudaSuccess) { cudaFree(dev); return rfail(DFD_ECUDA, "cudaMemcpy(rnn weights) failed"); }
    auto* W = new dfd_rnn_weights();
    W->dtype = dtype; W->input_size = input_size; W->hidden = H; W->layers = layers; W->arena = dev;
    uint8_t* d = reinterpret_cast<uint8_t*>(dev);
    for (int l = 0; l < layers; ++l) { W->w.push_back(d + w_off[l]); W->b.push_back(reinterpret_cast<float*>(d + b_off[l])); }
    auto F = [&](int i) { return reinterpret_cast<float*>(d + hoff[i]); };
    W->att_w1t = F(0); W->att_b1 = F(1); W->att_w2 = F(2); W->att_b2 = F(3);
    W->cls_w1t = F(4); W->cls_b1 = F(5); W->cls_w2 = F(6); W->cls_b2 = F(7);
    *out = W;
    return DFD_OK;
}

void dfd_rnn_free_weights(dfd_rnn_weights_t* w) { if (w) { if (w->arena) cudaFree(w->arena); delete w; } }

int dfd_rnn_workspace_bytes(const dfd_rnn_weights_t* w, int64_t batch, int T, size_t* bytes) {
    if (!w || !bytes || batch <= 0 || T <= 0) return rfail(DFD_EINVAL, "dfd_rnn_workspace_bytes: bad argument");
    const size_t H = w->hidden, K0 = w->input_size + H, B = (size_t)batch;
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    *bytes = up((size_t)(T + 1) * B * K0 * 2) + up(B * H * 2) + up(B * 7 * H * 4) + up(B * H * 4) + up(B * T * H * 4) + 1024;
    return DFD_OK;
}

int dfd_rnn_forward(const dfd_rnn_weights_t* w, const float* d_x, const int32_t* d_lengths, int64_t batch, int T,
                    float* d_prob, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (!w || !d_x || !d_prob || !d_workspace) return rfail(DFD_EINVAL, "dfd_rnn_forward: null pointer");
    size_t need = 0;
    int rc = dfd_rnn_workspace_bytes(w, batch, T, &need);
    if (rc) return rc;
    if (workspace_bytes < need) return rfail(DFD_ENOMEM, "dfd_rnn_forward: workspace too small");
    const int H = w->hidden, IN = w->input_size, K0 = IN + H, B = (int)batch;
    if ((size_t)T * H * 4 + 4096 > 200 * 1024) return rfail(DFD_EINVAL, "dfd_rnn_forward: sequence too long for the head kernel (T*H*4 <= 196 KB)");
    cudaStream_t s = (cudaStream_t)stream;
    auto up = [](size_t b) { return (b + 255) & ~size_t(255); };
    uint8_t* ws = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(d_workspace) + 255) & ~uintptr_t(255));
    uint8_t* a0 = ws;                                   ws += up((size_t)(T + 1) * B * K0 * 2);     // A operand of layer 0, one slab per step
    uint8_t* a1 = ws;                                   ws += up((size_t)B * H * 2);                // A operand of layers >= 1
    float* g = reinterpret_cast<float*>(ws);            ws += up((size_t)B * 7 * H * 4);
    float* c = reinterpret_cast<float*>(ws);            ws += up((size_t)B * H * 4);
    float* outs = reinterpret_cast<float*>(ws);
    const bool f16 = w->dtype == DFD_DTYPE_FP16;
    cudaError_t e;
    dfd::reset_launches();
#define RNN_CK(call, what) do { e = (call); dfd::note_launch(what); if (e != cudaSuccess) return rfail(DFD_ECUDA, std::string(what) + ": " + cudaGetErrorString(e)); } while (0)
    {
        const int64_t n = (int64_t)B * T * IN + (int64_t)B * H;
        const unsigned grid = (unsigned)((n + 255) / 256);
        if (f16) dfd::rnn_pack_x_kernel<__half><<<grid, 256, 0, s>>>(d_x, (__half*)a0, c, B, T, IN, H, K0);
        else dfd::rnn_pack_x_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>(d_x, (__nv_bfloat16*)a0, c, B, T, IN, H, K0);
        RNN_CK(cudaGetLastError(), "rnn_pack_x");
    }
    const unsigned cgrid = (unsigned)(((size_t)B * H + 255) / 256);
    for (int t = 0; t < T; ++t) {
        uint8_t* a0_t = a0 + (size_t)t * B * K0 * 2;
        uint8_t* a0_next = a0 + (size_t)(t + 1) * B * K0 * 2 + (size_t)IN * 2;       // h slot of step t+1
        for (int l = 0; l < w->layers; ++l) {
            const bool last = l == w->layers - 1;
            RNN_CK(dfd::launch_gemm_tc_f32out(l == 0 ? a0_t : a1, w->w[l], w->b[l], nullptr, g, B, l == 0 ? K0 : H, 7 * H, w->dtype, s), "rnn gate gemm");
            void* h16 = last ? (void*)a0_next : (void*)a1;
            const int stride = last ? K0 : H;
            float* o = last ? outs + (size_t)t * H : nullptr;
            if (f16) dfd::rnn_cell_kernel<__half><<<cgrid, 256, 0, s>>>(g, c, (__half*)h16, stride, o, T * H, d_lengths, t, B, H);
            else dfd::rnn_cell_kernel<__nv_bfloat16><<<cgrid, 256, 0, s>>>(g, c, (__nv_bfloat16*)h16, stride, o, T * H, d_lengths, t, B, H);
            RNN_CK(cudaGetLastError(), "rnn_cell");
        }
    }
    {
        const size_t smem = ((size_t)T * H + 32 + T + H) * 4;
        RNN_CK(cudaFuncSetAttribute(dfd::rnn_head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "rnn head smem");
        dfd::rnn_head_kernel<<<(unsigned)B, 256, smem, s>>>(outs, T, H, w->att_w1t, w->att_b1, w->att_w2, w->att_b2,
                                                          w->cls_w1t, w->cls_b1, w->cls_w2, w->cls_b2, d_prob);
        RNN_CK(cudaGetLastError(), "rnn_head");
    }
#undef RNN_CK
    return DFD_OK;
}

#pragma GCC visibility pop
}  // extern "C"
