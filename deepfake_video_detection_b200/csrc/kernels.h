// Host-side launchers of the sm_100a kernels (internal; the public surface is include/dfd_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>

namespace dfd {

// kernel-launch accounting shared by the engines (api.cu owns the thread-local counter behind dfd_last_launch_count)
void reset_launches();
void note_launch(const char* what);      // counts one launch unless `what` names a non-kernel call ("smem" attribute, "memset")

// K1 (preprocess.cu): uint8 HWC -> 16-bit NCHW, ImageNet normalisation (app.py:1772-1780, 2084-2085)
cudaError_t launch_preprocess(const uint8_t* in, void* out, int64_t frames, int H, int W, int dtype, cudaStream_t s);

// stem (stem.cu): conv3x3 s2 p1 3->32 + folded BN + SiLU -> NHWC.  w: fp32 [27][32] (tap-major: (ky*3+kx)*3+c), bias fp32 [32]
cudaError_t launch_stem(const void* in, int in_kind, const float* w, const float* bias, void* out,
                        int64_t frames, int H, int W, int dtype, cudaStream_t s);

// stem as tcgen05 implicit GEMM for uint8 crops (stem_tc.cu).  w16: 16-bit [32][96] = [w_hi | w_hi | w_lo] per output
// channel, each block 27 taps ((ky*3+kx)*3+c) zero-padded to 32.
// Row variant (output width <= 128, used when wrow / bias4 are given): raw uint8 bytes are the fp16 A operand, fetched
// with one bulk copy per output row; wrow fp16 [2][32][32] + bias4 fp32 [4][32] fold the tensor prep (api.cu pack_stem_row).
cudaError_t launch_stem_tc(const uint8_t* in, const void* w16, const float* bias, const void* wrow, const float* bias4, void* out,
                           int64_t frames, int H, int W, int dtype, cudaStream_t s);

// K2 (dwconv_march.cu): depthwise kxk (k in {3,5}, stride in {1,2}, pad k/2) + folded BN + SiLU, NHWC, plus the
// squeeze-excite spatial sums as per-block partials.  w: fp32 [k*k][C], bias fp32 [C].
// partials: fp32 [frames][dw_num_partials][C] (sums of SiLU outputs over groups of output tiles).
int dw_num_partials(int OH, int OW, int C, int k, int stride);
cudaError_t launch_dwconv(const void* in, const float* w, const float* bias, void* out, float* partials,
                          int64_t frames, int H, int W, int C, int k, int stride, int dtype, cudaStream_t s);

bool dw_march_supported(int H, int W, int C, int k, int stride);      // geometry the kernel can launch (maps up to 448 columns)
int dw_march_slots(int OH, int OW);
void dw_march_set_cb(int cb);      // tuning aid: force the channel block of the next launches (0 = default)

// K2' (mbconv_fused.cu): expand 1x1 + BN + SiLU fused into the row-marching depthwise kernel for the early HBM-bound
// blocks of the 224x224 network.  x [frames][H][W][cin] 16-bit, we [mid][cin] 16-bit + be fp32 [mid]; the rest as launch_dwconv
// (partials: dw_march_slots rows).
bool mbconv_fused_supported(int H, int W, int cin, int mid, int k, int stride);
cudaError_t launch_mbconv_fused(const void* x, const void* we, const float* be, const float* w, const float* bias, void* out,
                                float* partials, int64_t frames, int H, int W, int cin, int mid, int k, int stride, int dtype,
                                cudaStream_t s);

// K2 tail (se.cu): mean -> FC(C->rd)+bias -> SiLU -> FC(rd->C)+bias -> sigmoid.  gate fp32 [frames][C].
// w1 fp32 [rd][C], w2t fp32 [rd][C] (conv_expand transposed), b1 [rd], b2 [C].
cudaError_t launch_se(const float* partials, int nparts, float inv_hw, const float* w1, const float* b1,
                      const float* w2t, const float* b2, float* gate, int64_t frames, int C, int rd, cudaStream_t s);

// K3 (gemm_tc.cu): pointwise conv as tcgen05/TMEM GEMM.  D[M,N] = act((A .* gate)[M,K] * W[N,K]^T + bias) (+ R)
// A, W, R, D 16-bit; bias fp32 [N]; gate fp32 [M/HW][K] or null; R [M,N] or null; act: 0 none, 1 SiLU, 2 exact GELU, 3 ReLU after the residual add.
cudaError_t launch_gemm_tc(const void* A, const void* W, const float* bias, const float* gate, const void* R,
                           void* D, int64_t M, int K, int N, int HW, int act, int dtype, cudaStream_t s);
// gated project conv for big maps (HW >= 784): fold the SE gate into per-frame weights, then an ungated GEMM on
// frame-aligned tiles.  Wf: 16-bit [frames][N][K] scratch.
// Kg = gate row length (K, or K / r for pixel-packed block-diagonal weights: the gate repeats per packed pixel)
cudaError_t launch_scale_weights(const void* W, const float* gate, void* Wf, int64_t frames, int N, int K, int Kg, int dtype, cudaStream_t s);
cudaError_t launch_gemm_tc_framew(const void* A, const void* Wf, const float* bias, const void* R, void* D,
                                  int64_t M, int K, int N, int HW, int dtype, cudaStream_t s);
// D[M,N] (fp32) = A[M,K] * W[N,K]^T + bias (+ R fp32, may alias D) — gate pre-activations of the recurrent head
// (rnn.cu) and the fp32 residual stream of the ViT encoder (vit.cu)
cudaError_t launch_gemm_tc_f32out(const void* A, const void* W, const float* bias, const float* R, float* D,
                                  int64_t M, int K, int N, int dtype, cudaStream_t s);
// K3b (gemm_pair.cu): D[M,N] = act(A[M,K] * W[N,K]^T + bias) on CTA pairs (tcgen05 cta_group::2, 256 x 256 tiles, TMA-store
// epilogue); act 0 none | 2 exact GELU; N a multiple of 256.  The ViT-B/16 contractions.
bool gemm_pair_supported(int64_t M, int K, int N);
cudaError_t launch_gemm_pair(const void* A, const void* W, const float* bias, void* D, int64_t M, int K, int N, int act, int dtype, cudaStream_t s);
// X[M,N] (fp32, updated in place) += A[M,K] * W[N,K]^T + bias — the fp32 residual stream of the ViT encoder
cudaError_t launch_gemm_pair_residual(const void* A, const void* W, const float* bias, float* X, int64_t M, int K, int N, int dtype, cudaStream_t s);
// conv_head + BN + SiLU + global average pool (pretrained_detector.py:116 tail): feat fp32 [M/HW][N]
cudaError_t launch_gemm_tc_pool(const void* A, const void* W, const float* bias, float* feat,
                                int64_t M, int K, int N, int HW, int dtype, cudaStream_t s);
// K3h (head_pool_tc.cu): the same op as a transposed GEMM (channels on the TMEM lanes, a frame's pixels along a thread's row: the pool is a
// serial sum in the thread) for maps of up to 64 pixels, K <= 320, N a multiple of 128; launch_gemm_tc_pool dispatches to it
bool head_pool_tc_supported(int64_t M, int K, int N, int HW);
cudaError_t launch_head_pool_tc(const void* A, const void* W, const float* bias, float* feat, int64_t M, int K, int N, int HW, int dtype, cudaStream_t s);
// 3x3 stride-1 pad-1 convolution + bias + ReLU as an implicit GEMM (resnet50 conv2): the
// producing pointwise conv scatters its rows into a zero-haloed map of conv3x3_padded_rows(frames,H,W) x N elements (zeroed
// by the caller once per geometry), the 3x3 conv reads nine shifted TMA boxes of it.  Weights [N][(ky*3+kx)*C + c].
int64_t conv3x3_padded_rows(int64_t frames, int H, int W);
cudaError_t launch_gemm_tc_padout(const void* A, const void* W, const float* bias, void* Dpad, int64_t frames, int H, int Wd,
                                  int K, int N, int dtype, cudaStream_t s);
cudaError_t launch_gemm_tc_conv3x3(const void* Apad, const void* W, const float* bias, void* D, int64_t frames, int H, int Wd,
                                   int C, int N, int dtype, cudaStream_t s);
// 2-D tensor map (CUtensorMap written to *out_tmap) over a row-major 16-bit matrix [rows][cols]: boxes of 64 columns x box_rows
// rows, 128-byte swizzle, zero fill outside the tensor (gemm_tc.cu; cached per (pointer, shape))
cudaError_t make_tmap_2d(const void* base, int64_t rows, int cols, int box_rows, void* out_tmap);
// the same over an fp32 matrix: boxes of 32 columns x box_rows rows (gemm_pair.cu adds its tiles into the residual stream through it)
cudaError_t make_tmap_2d_f32(const float* base, int64_t rows, int cols, int box_rows, void* out_tmap);
// ViT attention on tcgen05 (vit_attn_tc.cu): qkv [images*197][2304] 16-bit -> o [images*197][768], softmax(Q K^T / 8) V per (image, head)
cudaError_t launch_vit_attention_tc(const void* qkv, void* o, int64_t images, int dtype, cudaStream_t s);

// K4 (poolhead.cu): temporal attention pool + fc1/ReLU/fc2 per video (pretrained_detector.py:123-141)
struct HeadWeights {
    const float *att_w1, *att_b1, *att_w2, *att_b2;   // [64][D], [64], [64], [1]
    const float *fc1_w, *fc1_b, *fc2_w, *fc2_b;       // [256][D], [256], [2][256], [2]
};
// feature_dim D = 1280 (efficientnet_b0) or 2048 (resnet50); a video that is empty, longer than 1024 frames or whose offsets leave
// the feature matrix gets NaN logits
cudaError_t launch_pool_head(const HeadWeights& hw, const float* feat, const int32_t* offsets, int64_t videos, int64_t frames,
                             int feature_dim, int use_attention, float* logits, float* frame_scores, cudaStream_t s);

}  // namespace dfd
