/* dfd_b200 — C ABI of the B200 (sm_100a) EfficientNet-B0 frame-scoring path.
 *
 * The reference (SaiPranav1506/DeepFake-Video-Detection) is pure Python and has no FFI/plugin boundary
 * (SURVEY.md §8b); its boundary is the nn.Module contract of
 *     src/pretrained_detector.py:15-143   PretrainedBackboneDetector (ctor :21-29, forward :103-143)
 *     app.py:1772-1780                    imagenet_normalize
 *     app.py:2084-2094                    tensor prep + model call + softmax
 * This header is what a binding for that path would bind: plain pointers and sizes, no torch types.
 * Every pointer named `d_*` is DEVICE memory on the current CUDA device; `stream` is a cudaStream_t
 * passed as void* (0 = legacy default stream).  All entry points are asynchronous on `stream` unless
 * stated, re-entrant (no hidden mutable state besides the immutable packed weights), and return
 * 0 on success or a negative DFD_E* code; dfd_last_error() gives the thread-local message.
 * No exception crosses this boundary.
 */
#ifndef DFD_B200_H
#define DFD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DFD_ABI_VERSION 1

/* 16-bit storage type of activations and GEMM operands (accumulation is always fp32). */
#define DFD_DTYPE_BF16 0
#define DFD_DTYPE_FP16 1

/* layout of the frames handed to dfd_effnet_b0_features / dfd_score_videos */
#define DFD_IN_U8_HWC 0      /* uint8 (F,H,W,3) RGB crops: prep of app.py:2084-2085 is fused into the stem   */
#define DFD_IN_F32_NCHW 1    /* float (F,3,H,W), already normalised: what forward() receives (:103-116)      */
#define DFD_IN_H16_NCHW 2    /* fp16/bf16 (F,3,H,W), output of dfd_preprocess_u8hwc_to_nchw                   */

#define DFD_OK 0
#define DFD_EINVAL (-1)      /* bad argument (null pointer, unsupported shape, unknown dtype ...)             */
#define DFD_ECUDA (-2)       /* a CUDA runtime call or kernel launch failed                                  */
#define DFD_ENOMEM (-3)      /* workspace too small / allocation failed                                      */
#define DFD_EKEY (-4)        /* a required state_dict tensor is missing or has the wrong element count       */

#define DFD_FEATURE_DIM 1280 /* pretrained_detector.py:49                                                     */
#define DFD_NUM_CLASSES 2

typedef struct dfd_weights dfd_weights_t;   /* opaque: BN-folded, repacked weights resident in HBM */

int dfd_abi_version(void);
const char* dfd_last_error(void);

/* ---- weights -------------------------------------------------------------------------------------
 * Replaces `model.load_state_dict(state_dict)` + `model.to(DEVICE)` (app.py:1718,1760; agent_system.py:89-91).
 * `names[i]` are reference state_dict keys (e.g. "backbone.2.1.0.conv_dw.weight"),
 * `data[i]` HOST fp32 pointers with `numel[i]` elements (`num_batches_tracked` entries may be omitted).
 * BatchNorm is folded in fp32 (w' = w*g/sqrt(var+1e-5), b' = beta - mean*g/sqrt(var+1e-5)), then cast
 * to `dtype` and repacked K-major for the tensor-core GEMMs.  Synchronous.  Free with dfd_free_weights. */
int dfd_pack_weights(int n_tensors, const char* const* names, const float* const* data,
                     const int64_t* numel, int dtype, dfd_weights_t** out);
void dfd_free_weights(dfd_weights_t* w);
int dfd_weights_dtype(const dfd_weights_t* w);

/* ---- K1: tensor prep (app.py:2084-2085 + imagenet_normalize app.py:1772-1780) ---------------------
 * uint8 (F,H,W,3) -> `dtype` (F,3,H,W): y = ((u8/255) - mean[c]) / std[c], evaluated in fp32 exactly as
 * the reference does, then rounded once to the 16-bit type.  H*W must be a multiple of 16. */
int dfd_preprocess_u8hwc_to_nchw(const uint8_t* d_in, void* d_out, int64_t frames, int H, int W,
                                 int dtype, void* stream);

/* ---- trunk: `self.backbone(x_flat)` (pretrained_detector.py:116) ----------------------------------
 * frames (layout `in_kind`) -> pooled features fp32 (F,1280).  H = W = 224 is the supported crop size
 * (any H, W multiple of 32 with (H/32)*(W/32) <= 128 works).  `d_workspace` must hold at least
 * dfd_workspace_bytes(frames, H, W) bytes and must not be shared between concurrent calls. */
int dfd_workspace_bytes(int64_t frames, int H, int W, size_t* bytes);
int dfd_effnet_b0_features(const dfd_weights_t* w, const void* d_in, int in_kind, int64_t frames,
                           int H, int W, float* d_features, void* d_workspace, size_t workspace_bytes,
                           void* stream);

/* ---- pool + head: pretrained_detector.py:123-141 --------------------------------------------------
 * `d_offsets` int32 (V+1): video v owns frames [offsets[v], offsets[v+1]) of `d_features` (ragged T,
 * each video is scored exactly as one reference B=1 call; an empty video is an error).
 * use_attention != 0: sigmoid-MLP scores -> softmax over T -> weighted feature sum (:123-131);
 * use_attention == 0: feature mean, frame_scores = 1/T (:132-135).  Then fc1/ReLU/fc2 (:138-141).
 * Outputs: d_logits fp32 (V,2), d_frame_scores fp32 (F,) in frame order (may be NULL). */
int dfd_attn_pool_head(const dfd_weights_t* w, const float* d_features, const int32_t* d_offsets,
                       int64_t videos, int64_t frames, int use_attention, float* d_logits,
                       float* d_frame_scores, void* stream);

/* ---- whole path: app.py:2084-2089 for a batch of videos -------------------------------------------
 * = dfd_effnet_b0_features + dfd_attn_pool_head.  `d_features_out` (F,1280) may be NULL, in which case
 * the features live in the workspace (dfd_score_workspace_bytes accounts for them). */
int dfd_score_workspace_bytes(int64_t frames, int H, int W, size_t* bytes);
int dfd_score_videos(const dfd_weights_t* w, const void* d_in, int in_kind, const int32_t* d_offsets,
                     int64_t videos, int64_t frames, int H, int W, int use_attention, float* d_logits,
                     float* d_frame_scores, float* d_features_out, void* d_workspace,
                     size_t workspace_bytes, void* stream);

/* number of kernel launches the last dfd_effnet_b0_features / dfd_score_videos call on this thread made */
int dfd_last_launch_count(void);

/* ---- temporal RNN head (BASELINE config 3): src/RNNModel.py:43-147 LogicRNNLSTM ------------------------
 * Weights by reference state_dict key ("logic_cells.{l}.{and,or,not,forget,input,cell,output}_gate.{weight,bias}",
 * "attention.{0,2}.*", "classifier.{0,3}.*"), HOST fp32.  The seven gate Linears of a LogicCell (:11-19) are
 * stacked into one [7H, K] 16-bit GEMM operand per layer.
 * dfd_rnn_forward: d_x fp32 (B,T,input_size) [already sorted by length when d_lengths is given, as the
 * reference does at :92-95], d_lengths int32 (B,) or NULL -> d_prob fp32 (B,) = sigmoid(classifier(context)). */
typedef struct dfd_rnn_weights dfd_rnn_weights_t;
const char* dfd_rnn_last_error(void);
int dfd_rnn_pack_weights(int n_tensors, const char* const* names, const float* const* data, const int64_t* numel,
                         int input_size, int hidden, int layers, int dtype, dfd_rnn_weights_t** out);
void dfd_rnn_free_weights(dfd_rnn_weights_t* w);
int dfd_rnn_workspace_bytes(const dfd_rnn_weights_t* w, int64_t batch, int T, size_t* bytes);
int dfd_rnn_forward(const dfd_rnn_weights_t* w, const float* d_x, const int32_t* d_lengths, int64_t batch, int T,
                    float* d_prob, void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- ViT-B/16 frame encoder (BASELINE config 5): src/models.py:88-107 ViTFeatureExtractor ------------------
 * = timm `vit_base_patch16_224(num_classes=0)` (models.py:93): 224x224 images -> CLS feature after the final norm.
 * Weights by reference state_dict key, with or without the `vit.` prefix ("vit.patch_embed.proj.weight",
 * "vit.cls_token", "vit.pos_embed", "vit.blocks.{i}.{norm1,norm2}.{weight,bias}",
 * "vit.blocks.{i}.attn.{qkv,proj}.{weight,bias}", "vit.blocks.{i}.mlp.{fc1,fc2}.{weight,bias}", "vit.norm.*"), HOST
 * fp32; Linear / conv weights are cast to `dtype` (GEMM operands), everything else stays fp32.
 * dfd_vit_features: d_in fp32 (B,3,224,224), already normalised (what forward() receives, models.py:105-107)
 * -> d_features fp32 (B,768).  Workspace: dfd_vit_workspace_bytes(B) bytes, not shared between concurrent calls. */
typedef struct dfd_vit_weights dfd_vit_weights_t;
const char* dfd_vit_last_error(void);
int dfd_vit_pack_weights(int n_tensors, const char* const* names, const float* const* data, const int64_t* numel,
                         int dtype, dfd_vit_weights_t** out);
void dfd_vit_free_weights(dfd_vit_weights_t* w);
int dfd_vit_workspace_bytes(int64_t images, size_t* bytes);
int dfd_vit_features(const dfd_vit_weights_t* w, const float* d_in, int64_t images, float* d_features,
                     void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- `resnet50` ensemble member: src/pretrained_detector.py:38-41 (torchvision trunk), :103-143 ------
 * Weights by reference state_dict key ("backbone.0.weight" = conv1, "backbone.1.*" = bn1, "backbone.{4..7}.{b}.conv{1,2,3}.weight",
 * ".bn{1,2,3}.*", ".downsample.{0,1}.*", "temporal_attention.{0,2}.*", "fc1.*", "fc2.*"), HOST fp32; BatchNorm is folded in fp32.
 * dfd_resnet50_score_videos: d_in fp32 (frames,3,224,224) already normalised (what forward() receives), d_offsets int32
 * (videos+1), max_frames_per_video >= the longest video -> d_logits fp32 (videos,2), d_frame_scores fp32 (frames) or NULL,
 * d_features_out fp32 (frames,2048) or NULL.  Workspace: dfd_resnet50_workspace_bytes(frames). */
typedef struct dfd_resnet_weights dfd_resnet_weights_t;
const char* dfd_resnet_last_error(void);
int dfd_resnet50_pack_weights(int n_tensors, const char* const* names, const float* const* data, const int64_t* numel,
                              int dtype, dfd_resnet_weights_t** out);
void dfd_resnet50_free_weights(dfd_resnet_weights_t* w);
int dfd_resnet50_workspace_bytes(int64_t frames, size_t* bytes);
int dfd_resnet50_score_videos(const dfd_resnet_weights_t* w, const float* d_in, const int32_t* d_offsets, int64_t videos, int64_t frames,
                              int max_frames_per_video, int use_attention, float* d_logits, float* d_frame_scores, float* d_features_out,
                              void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- DeepfakeModel head: src/models.py:177-197 SimpleGCN + :283-291 mean pool + classifier ------
 * logits = classifier(mean_n relu(fc2(relu(fc1(A_norm @ H))))) on the ViT frame features H of each video.
 * Weights by reference state_dict key ("gcn.fc1.*", "gcn.fc2.*", "classifier.0.*", "classifier.3.*"), HOST fp32; built for
 * the reference's default sizes (vit_out 768, gcn 256 / 128, classifier 64).  d_feats fp32 (videos*nodes, 768) = output
 * of dfd_vit_features, d_adj fp32 (videos, nodes, nodes) = normalize_adjacency (src/utils.py:95-104), nodes <= 64
 * (app.py:2230 uses 16) -> d_logits fp32 (videos, num_classes).  Errors: dfd_vit_last_error(). */
typedef struct dfd_gcn_weights dfd_gcn_weights_t;
int dfd_gcn_pack_weights(int n_tensors, const char* const* names, const float* const* data, const int64_t* numel,
                         int num_classes, dfd_gcn_weights_t** out);
void dfd_gcn_free_weights(dfd_gcn_weights_t* w);
int dfd_gcn_head(const dfd_gcn_weights_t* w, const float* d_feats, const float* d_adj, int64_t videos, int nodes,
                 float* d_logits, void* stream);

/* ---- K0: crop + resize in front of the path: app.py:1964-1978, src/data_prepare.py:54-56 ----------
 * `pil.crop((x1, y1, x2, y2)).resize((S, S))` — Pillow's BICUBIC (antialiased, separable, 8-bit intermediate, 22-bit
 * fixed-point coefficients) reproduced bit-exactly, straight from full uint8 HWC video frames resident on the device.
 * `h_boxes` is a HOST array: frame_offset = byte offset of the frame inside d_frames, frame_w / frame_h its extent in
 * pixels, (x1, y1, x2, y2) the box already clamped to the frame as the reference does at app.py:1968-1973 (an empty or
 * out-of-frame box is DFD_EINVAL — the reference skips those).  d_out: uint8 (n, S, S, 3), ready for dfd_score_videos
 * (DFD_IN_U8_HWC).  The coefficient tables are built on the host and uploaded on `stream` before the call returns. */
typedef struct dfd_crop_box {
    int64_t frame_offset;
    int32_t frame_w, frame_h;
    int32_t x1, y1, x2, y2;
} dfd_crop_box;
const char* dfd_resize_last_error(void);
int dfd_crop_resize_workspace_bytes(const dfd_crop_box* h_boxes, int64_t n, int out_size, size_t* bytes);
int dfd_crop_resize_u8(const uint8_t* d_frames, const dfd_crop_box* h_boxes, int64_t n, int out_size, uint8_t* d_out,
                       void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- measurement aid --------------------------------------------------------------------------------
 * dfd_profile_enable(1): from now on every kernel launched by this thread through this library is
 * bracketed by CUDA events recorded on the launch stream.  dfd_profile_collect synchronises on them and
 * returns one entry per kernel class: launches, summed device time, algorithmic bytes (activations read
 * once + written once, weights excluded — SURVEY.md §8d) and flops.  dfd_profile_enable(0) stops and
 * discards.  Not for use inside a timed region. */
typedef struct dfd_profile_entry {
    char name[32];
    int launches;
    double ms, bytes, flops;
} dfd_profile_entry;
int dfd_profile_enable(int on);
int dfd_profile_collect(dfd_profile_entry* out, int max_entries, int* n_out);

#ifdef __cplusplus
}
#endif
#endif /* DFD_B200_H */
