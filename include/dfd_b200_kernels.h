/* dfd_b200 — kernel-level C entry points (one launch each).  Same conventions as dfd_b200.h.
 * These exist so that every kernel can be parity-tested and profiled on its own through the C ABI;
 * a binding for the reference only needs dfd_b200.h.  Layouts:
 *   activations   NHWC, 16-bit (DFD_DTYPE_*), i.e. a [frames*H*W, C] row-major matrix
 *   BatchNorm     already folded into weight/bias by the caller (dfd_pack_weights does this)
 * Stream ordering: the EfficientNet kernels are launched with programmatic dependent launch.  Each of them reads its WEIGHT and
 * BIAS operands (d_w, d_bias, d_W, SE matrices ...) in a prologue that may run before the kernel enqueued ahead of it on `stream`
 * has finished; activations, gates, partial sums and offsets are read (and every output is written) only after that kernel has
 * completed.  So weights must be complete in device memory when the call is enqueued (a cudaMemcpy, or an earlier
 * synchronisation) — never the product of the kernel immediately ahead on the same stream.  dfd_pack_weights guarantees this
 * for the handle-based entry points of dfd_b200.h.
 */
#ifndef DFD_B200_KERNELS_H
#define DFD_B200_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* timm conv_stem + bn1 (backbone.0/.1): conv3x3 s2 p1 3->32 + bias + SiLU.  d_w fp32 [(ky*3+kx)*3+c][32]. */
int dfd_k_stem(const void* d_in, int in_kind, const float* d_w, const float* d_bias, void* d_out,
               int64_t frames, int H, int W, int dtype, void* stream);

/* the same stem as a tcgen05/TMEM implicit GEMM for uint8 crops (hi/lo split operands, ~fp32 accuracy).
 * h_w27x32: HOST fp32 weights in the layout above (split + uploaded inside; synchronous; test/profiling aid). */
int dfd_k_stem_tc(const uint8_t* d_in, const float* h_w27x32, const float* d_bias, void* d_out,
                  int64_t frames, int H, int W, int dtype, void* stream);

/* timm conv_dw + bn: depthwise kxk (k 3|5, stride 1|2, pad k/2) + bias + SiLU, plus the squeeze-excite
 * spatial sums as d_partials fp32 [frames][dfd_k_dw_num_partials(OH,OW,C,k,stride)][C].  d_w fp32 [k*k][C]. */
int dfd_k_dw_num_partials(int OH, int OW, int C, int k, int stride);
/* tuning aid for tools/sweep_dw.py: force the channel block (channels per CTA) of the following depthwise launches; 0 = built-in choice */
void dfd_k_set_dw_channel_block(int cb);
int dfd_k_dwconv(const void* d_in, const float* d_w, const float* d_bias, void* d_out, float* d_partials,
                 int64_t frames, int H, int W, int C, int k, int stride, int dtype, void* stream);

/* timm SqueezeExcite: gate = sigmoid(W2 * SiLU(W1 * mean + b1) + b2), fp32 [frames][C].
 * d_w1 [rd][C], d_w2t [rd][C] (conv_expand transposed). */
int dfd_k_se(const float* d_partials, int nparts, float inv_hw, const float* d_w1, const float* d_b1,
             const float* d_w2t, const float* d_b2, float* d_gate, int64_t frames, int C, int rd, void* stream);

/* pointwise conv: D[M,N] = act((A .* gate)[M,K] * W[N,K]^T + bias) (+ R).  gate fp32 [M/HW][K] or NULL,
 * R [M,N] or NULL, act 0|1 (SiLU).  impl 0 = tcgen05/TMEM kernel (gate applied to the A operand),
 * 2 = tcgen05 kernel with the gate folded into per-frame
 * weights on frame-aligned tiles (what the engine uses for maps of >= 784 pixels; synchronous in this test entry),
 * 3 = CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 tiles, TMA-store epilogue: the ViT-B/16 contractions; no gate, no
 * 16-bit residual, act 0 | 2 (exact GELU), N a multiple of 256; with d_R == d_D the output is an fp32 matrix updated in place,
 * X += A W^T + bias: the encoder's residual stream). */
int dfd_k_gemm(const void* d_A, const void* d_W, const float* d_bias, const float* d_gate, const void* d_R,
               void* d_D, int64_t M, int K, int N, int HW, int act, int dtype, int impl, void* stream);

/* conv_head + bn2 + SiLU + global average pool: feat fp32 [M/HW][N]. */
int dfd_k_gemm_pool(const void* d_A, const void* d_W, const float* d_bias, float* d_feat, int64_t M, int K,
                    int N, int HW, int dtype, int impl, void* stream);

/* torchvision Bottleneck conv1 + bn1 + relu ->
 * conv2 (3x3, stride 1, pad 1) + bn2 + relu (reference trunk: src/pretrained_detector.py:38-41) without a gathered operand.
 * d_in [frames*H*W][K], d_w1 [C][K], d_w2 [N][(ky*3+kx)*C + c] (BN folded), d_out [frames*H*W][N], all 16-bit; biases fp32.
 * d_pad: scratch for the zero-haloed intermediate map, dfd_k_conv3x3_maps(frames,H,W,..) * C * 2 bytes. */
int dfd_k_conv1x1_conv3x3(const void* d_in, const void* d_w1, const float* d_b1, const void* d_w2, const float* d_b2, void* d_out,
                          int64_t frames, int H, int W, int K, int C, int N, int dtype, void* d_pad, size_t pad_bytes, void* stream);

/* timm InvertedResidual conv_pw + bn1 + SiLU ->
 * conv_dw + bn2 + SiLU (+ squeeze-excite sums) as ONE kernel for the early blocks of the 224x224 network
 * ((cin, mid, map, k, stride) = (16,96,112,3,2), (24,144,56,3,1), (24,144,56,5,2); dfd_k_mbconv_fused_supported tells):
 * the expanded tensor stays in shared memory.  d_x [frames*H*W][cin], d_we [mid][cin] 16-bit, d_be fp32 [mid]; d_w, d_bias,
 * d_out, d_partials as dfd_k_dwconv. */
int dfd_k_mbconv_fused_supported(int H, int W, int cin, int mid, int k, int stride);
int dfd_k_mbconv_fused(const void* d_x, const void* d_we, const float* d_be, const float* d_w, const float* d_bias, void* d_out,
                       float* d_partials, int64_t frames, int H, int W, int cin, int mid, int k, int stride, int dtype, void* stream);

/* timm vit_base_patch16_224 attention of one block (reference: src/models.py:88-107): d_qkv [images*197][2304] 16-bit in timm's
 * column order (which*768 + head*64 + d) -> d_o [images*197][768] = softmax(Q K^T / 8) V per (image, head); the tcgen05 / TMEM
 * kernel the encoder runs (csrc/vit_attn_tc.cu). */
int dfd_k_vit_attention(const void* d_qkv, void* d_o, int64_t images, int dtype, void* stream);

/* HOST-ONLY (no GPU needed): row maps of that zero-haloed layout, computed by the very functions the kernels use
 * (csrc/conv_map.h).  h_pad_row [frames*H*W]: physical row of every interior pixel; h_out_row [frames*(H+2)*(W+2)]: output
 * row of every padded pixel, -1 for halo pixels; h_tap_row / h_tap_col [9*cpk]: row offset and channel column of the A box
 * of k-block kb = tap*cpk + slice.  Any pointer may be NULL.  Returns the number of rows of the layout (guards included). */
int64_t dfd_k_conv3x3_maps(int frames, int H, int W, int cpk, int64_t* h_pad_row, int64_t* h_out_row, int32_t* h_tap_row, int32_t* h_tap_col);

/* HOST-ONLY (no GPU needed): the operands dfd_pack_weights builds for the row-variant stem (stem_tc.cu) from BN-folded
 * weights h_w27x32 fp32 [(ky*3+kx)*3+c][32] and bias h_bias32: h_wrow fp16 bits [hi|lo][32 oc][32 k] with
 * k = ky*10 + kx*3 + c holding 256*w/(255*std_c) split in two fp16 terms, h_bias4 fp32 [top*2+left][32] = bias minus
 * sum over the in-bounds taps of w*mean_c/std_c (the tensor prep of app.py:1772-1780 folded into the conv). */
int dfd_k_pack_stem_row(const float* h_w27x32, const float* h_bias32, uint16_t* h_wrow, float* h_bias4);

/* HOST-ONLY (no GPU needed): the fixed-point bicubic coefficient table the resize kernels use for one axis
 * (Pillow's precompute_coeffs + normalize_coeffs_8bpc).  h_bounds int32 [out_size][2] = (first source index, taps),
 * h_coeffs int32 [out_size][*ksize] (row stride = *ksize, unused taps 0); h_coeffs may be NULL to query *ksize only.
 * Returns the number of coefficients per row, or a negative DFD_E* code. */
int dfd_k_resize_coeffs(int in_size, int out_size, int32_t* h_bounds, int32_t* h_coeffs, int* ksize);

#ifdef __cplusplus
}
#endif
#endif /* DFD_B200_KERNELS_H */
