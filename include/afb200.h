/*
 * afb200.h — C ABI of libafb200.so, the B200-native (sm_100a) AltFreezing
 * clip-classification hot path.
 *
 * The reference (Mariachiar/Spatiotemporal-Deepfake-Detection-for-Live-Video-Calls)
 * is pure Python and has no FFI of its own; every entry point below names the
 * reference interface it stands in for (paths relative to the reference root).
 * Conventions: plain pointers and sizes only, no C++/torch types; every call
 * returns an af_status (0 = ok, negative = error) and never throws; the text of
 * the last error on the calling thread is af_last_error().  Device pointers are
 * owned by the caller (PyTorch's allocator in the Python host); the engine owns
 * its weights and workspace.  Calls on one handle are not re-entrant (the
 * reference's services are single-threaded singletons, altfreezing/TEST2.py:64-70).
 * `stream` is a cudaStream_t passed as void* (0 = the legacy default stream); all
 * *_dev entry points are asynchronous on it.
 */
#ifndef AFB200_H
#define AFB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFB200_VERSION 102

typedef struct af_engine* af_handle;

typedef enum {
  AF_OK = 0,
  AF_ERR_INVALID = -1,      /* bad argument / unsupported shape                     */
  AF_ERR_CUDA = -2,         /* a CUDA runtime/driver call failed (see af_last_error) */
  AF_ERR_UNSUPPORTED = -3,  /* no sm_100 device                                      */
  AF_ERR_NOMEM = -4
} af_status;

typedef enum { AF_F32 = 0, AF_BF16 = 1, AF_F16 = 2, AF_U8 = 3 } af_dtype;

/* Arithmetic of the trunk. FP32: fp32 activations/weights/accumulate (parity gate
 * 1e-3 on logits). BF16: bf16 activations/weights, fp32 accumulate on tcgen05
 * tensor cores (parity gate 2e-2 and same decision at logit 0). */
/* AF_PREC_TF32: fp32 storage and fp32 accumulation like AF_PREC_FP32, trunk convolutions on the tensor cores with TF32
 * operands (what cuDNN does for the reference's fp32 model under torch's default allow_tf32). */
typedef enum { AF_PREC_FP32 = 0, AF_PREC_BF16 = 1, AF_PREC_TF32 = 2 } af_precision;

/* One Conv3d with its eval-mode BatchNorm3d already folded by the host
 * (W' = W*gamma/sqrt(var+eps), b' = beta - mean*gamma/sqrt(var+eps)).
 * Replaces nn.Conv3d + nn.BatchNorm3d pairs built at
 * altfreezing/slowfast/models/stem_helper.py:156-171 and resnet_helper.py:255-309,411-423. */
typedef struct {
  const float* weight; /* host, [cout][cin][kt][kh][kw] fp32 (PyTorch Conv3d layout) */
  const float* bias;   /* host, [cout] fp32                                          */
  int32_t cin, cout;
  int32_t kt, kh, kw;
  int32_t st, sh, sw;
  int32_t pt, ph, pw;
} af_conv_desc;

/* One bottleneck ResBlock: relu(shortcut + c(b(a(x)))) — resnet_helper.py:438-444,311-326.
 * Indices into af_weights.convs; branch1 = -1 for identity shortcuts. */
typedef struct {
  int32_t branch1, a, b, c;
  int32_t temporal_pool_before; /* 1: MaxPool3d [2,1,1] on the block input
                                   (pathway0_pool, video_model_builder.py:474-480,566-568) */
  int32_t spatial_pool;         /* 1: MaxPool3d((1,2,2)) behind the BatchNorm of conv b and of branch1 — what the
                                   FTCN-TT plugin puts where the I3D has a spatial stride
                                   (model/classifier/i3d_temporal_var_fix_dropout_tt_cfg.py:222-267); 0 for i3d_ori */
} af_block_desc;

/* One pre-norm transformer layer of the FTCN-TT head (model/classifier/time_transformer.py:75-88):
 * x += to_out(attention(LayerNorm(x)));  x += fc2(GELU(fc1(LayerNorm(x)))).  Host pointers, fp32, nn.Linear layout. */
typedef struct {
  const float *ln1_w, *ln1_b;
  const float* qkv_w;              /* [3*heads*dim_head][dim], no bias (time_transformer.py:37) */
  const float *out_w, *out_b;      /* [dim][heads*dim_head], [dim] */
  const float *ln2_w, *ln2_b;
  const float *fc1_w, *fc1_b;      /* [mlp_dim][dim], [mlp_dim] */
  const float *fc2_w, *fc2_b;      /* [dim][mlp_dim], [dim] */
} af_tt_layer;

/* TransformerHead + TimeTransformer (i3d_temporal_var_fix_dropout_tt_cfg.py:126-196, time_transformer.py:219-279):
 * per-frame spatial means -> `tokens` vectors of `dim` channels, cls token + learned positions, `depth` layers,
 * LayerNorm + Linear(dim->1) on the cls token. */
typedef struct {
  int32_t dim, tokens, heads, dim_head, mlp_dim, depth;
  const af_tt_layer* layers;
  const float* cls_token;      /* [dim] */
  const float* pos_embedding;  /* [tokens+1][dim] */
  const float *norm_w, *norm_b;
  const float* fc_w;           /* [dim] */
  float fc_b;
} af_tt_head;

/* The whole network: stem conv (+BN+ReLU) + MaxPool3d [1,3,3]/[1,2,2]/[0,1,1]
 * (stem_helper.py:173-178), the blocks, global average pool + Linear(F->1)
 * (head_helper.py:74-95). */
typedef struct {
  int32_t n_convs;
  const af_conv_desc* convs;
  int32_t stem;            /* index of the stem conv */
  int32_t n_blocks;
  const af_block_desc* blocks;
  const float* fc_weight;  /* host, [feature_dim] */
  float fc_bias;
  int32_t feature_dim;     /* 2048 */
  int32_t clip_t, clip_s;  /* 32, 224 (cfg.clip_size, cfg.imsize; setting/i3d_ori.yaml:20,60) */
  /* FTCN-TT plugin only (zero / NULL for i3d_ori): */
  int32_t stem_pool2;        /* 1: MaxPool3d((1,2,2)) between the stem's BatchNorm and its ReLU; the stem conv is then
                                k[5,1,1] s[1,1,1] (temporal_only_conv on s1.pathway0_stem) */
  const af_tt_head* tt_head; /* non-NULL: transformer head on the per-frame means of the last stage instead of
                                average-pool + Linear; fc_weight may then be NULL and feature_dim = tt_head->dim */
} af_weights;

/* Per-frame source description for the crop kernel: one decoded RGB (or BGR) u8
 * frame in device memory and that frame's enlarged face box ("big box",
 * altfreezing/test_tools/utils.py:13-24), exclusive lower-right corner. */
typedef struct {
  const uint8_t* data; /* device pointer to pixel (0,0), 3 interleaved channels */
  int64_t pitch;       /* bytes per row */
  int32_t height, width;
  int32_t box[4];      /* x1, y1, x2, y2 in frame coordinates */
} af_frame_desc;

/* Per-clip geometry computed on the host by the similarity estimator
 * (altfreezing/test_tools/warp_for_xray.py:556-560) and the union-box logic of
 * FasterCropAlignXRay.__call__ (faster_crop_align_xray.py:22-50). */
typedef struct {
  double tfm[6];        /* forward 2x3 map canvas -> SxS crop, row-major */
  int32_t left_top[2];  /* canvas origin in frame coordinates (x, y)     */
  int32_t canvas_wh[2]; /* canvas size (w, h) = union of the clip's big boxes */
} af_clip_geom;

const char* af_last_error(void);
int32_t af_version(void);
/* Number of the engine's own CUDA kernel launches since creation (bench "gpu_launches"). */
int64_t af_launch_count(af_handle h);

/* Build an engine on `device`: uploads and re-lays-out the folded weights, allocates
 * the activation workspace for up to max_batch clips per call.
 * Replaces `PluginLoader.get_classifier("i3d_ori")().to(device).eval()` + `.load(ckpt)`
 * (altfreezing/demo.py:403-404; model/_base.py:20-23,39-104) for the device side. */
af_status af_create(af_handle* out, int32_t device, const af_weights* w, int32_t max_batch,
                    int32_t precision);
af_status af_destroy(af_handle h);

/* Process-wide diagnostic knobs (tests only).  "block_n" = 64 | 128 | 256 forces the output-tile width of the
 * generic tcgen05 conv kernel wherever Cout allows (0 = the planner's own choice), so that every tile shape the
 * production schedule can pick is testable on small inputs. */
af_status af_set_global_option(const char* name, int64_t value);

/* Counters for the benchmark's roofline line.  With option "profile_events" = 1 every
 * conv launch is bracketed by CUDA events on its stream (no synchronisation); the getters
 * below synchronise the device and aggregate them since the last "reset_stats" option call:
 *   "conv_umma_ms" / "conv_umma_launches" / "conv_umma_flops"   tcgen05 conv kernel
 *   "conv_simt_ms" / "conv_simt_launches" / "conv_simt_flops"   CUDA-core conv kernel
 *   "conv_bytes"                                                algorithmic activation+weight bytes of all convs
 * The tcgen05 launches are also split by which roofline bounds them — algorithmic FLOP/byte of the launch against the
 * ridge set with option "ridge_x1000" (1000 x peak FLOP/s / peak byte/s; default 208):
 *   "conv_tensor_bound_ms" / "_flops" / "_launches"             launches above the ridge
 *   "conv_hbm_bound_ms" / "_bytes" / "_flops" / "_launches"     launches below it */
af_status af_get_stat(af_handle h, const char* name, double* value);

/* Tuning knobs (chunk sizes of the batch schedule); name/value pairs, optional. */
af_status af_set_option(af_handle h, const char* name, int64_t value);

/* clf(x)["final_output"]: x is a normalised clip tensor [B,3,T,S,S] of `dtype` in device
 * memory with arbitrary element strides (B,C,T,H,W) — contiguous NCTHW, a permuted NTHWC
 * view (demo.py:317) or channels_last_3d (TEST2.py:155).  Writes fp32 logits [B] (no
 * activation, head_helper.py:90-94) and, if non-null, the pooled features [B,feature_dim]
 * (the input of head.projection that altfreezing/feature.py:106-114 hooks).
 * Replaces ModelBase.forward -> I3D8x8.forward -> ResNet.forward
 * (model/_base.py:25-26, model/classifier/i3d_ori.py:92-104, video_model_builder.py:561-578). */
af_status af_forward(af_handle h, const void* clip_dev, int32_t dtype, const int64_t strides[5],
                     int32_t batch, float* logits_dev, float* features_dev, void* stream);

/* af_forward plus per-frame features: frame_features_dev [B, T/2, feature_dim] = spatial mean of the last stage
 * per output frame (their temporal mean is the pooled feature of af_forward).  This is the `backbone(x) -> [B,T',D]`
 * contract of AltFreezingRGBEncoder (dualrun/model/dual_rgb.py:26-44; the reference ships no adapter, SURVEY.md
 * §8a row A13). */
af_status af_forward_frames(af_handle h, const void* clip_dev, int32_t dtype, const int64_t strides[5],
                            int32_t batch, float* logits_dev, float* frame_features_dev, void* stream);

/* ClassifierSvc.infer_scores (altfreezing/TEST2.py:151-204, test/af_realtime.py:75-96) on
 * device buffers: u8 aligned clips [B,T,S,S,3] RGB -> (x-255*mean)/(255*std) -> trunk ->
 * logits [B] (and sigmoid scores [B] if scores_dev != NULL). mean/std are the three
 * per-channel constants already multiplied by 255 (demo.py:84-87). */
af_status af_infer_u8(af_handle h, const uint8_t* clips_dev, int32_t batch, const float mean255[3],
                      const float std255[3], float* logits_dev, float* scores_dev,
                      float* features_dev, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies the u8 clips to the device,
 * runs af_infer_u8, copies logits/scores back and synchronises the stream. This is the
 * end-to-end entry the benchmark's `e2e` number is measured through. */
af_status af_infer_u8_host(af_handle h, const uint8_t* clips_host, int32_t batch,
                           const float mean255[3], const float std255[3], float* logits_host,
                           float* scores_host, void* stream);

/* Pipelined form of af_infer_u8_host for callers that score batch after batch (the flush loops of
 * altfreezing/TEST2.py:393-439 and test/af_realtime.py:318-360, batch_eval-style offline scoring): af_submit_u8_host
 * starts the upload of `clips_host` on an internal copy stream, queues pack + trunk + the read-back of logits and
 * scores behind it on `stream`, and returns at once with a ticket (0 or 1); af_wait blocks until that submission's
 * results are on the host and copies them out.  Two submissions may be outstanding, so the upload of batch i+1
 * overlaps the compute of batch i.  `clips_host` must stay valid until af_wait returns (pinned memory for a truly
 * asynchronous copy); results of a ticket must be collected before the slot is submitted to again. */
af_status af_submit_u8_host(af_handle h, const uint8_t* clips_host, int32_t batch, const float mean255[3],
                            const float std255[3], void* stream, int32_t* ticket);
af_status af_wait(af_handle h, int32_t ticket, float* logits_host, float* scores_host);

/* FasterCropAlignXRay.process_single for a batch of clips, bit-exact with
 * cv2.warpAffine(INTER_LINEAR, BORDER_CONSTANT 0) (faster_crop_align_xray.py:77-88):
 * gathers straight from the decoded frames (frames[b*T+t]), masks to each frame's big
 * box, writes u8 [B,T,S,S,3].  `frames_dev` and `geom_dev` are device arrays.
 * bgr != 0 swaps channels 0 and 2 while reading (decoded BGR -> RGB, demo.py:245-269). */
af_status af_crop_u8(const af_frame_desc* frames_dev, const af_clip_geom* geom_dev, int32_t batch,
                     int32_t frames_per_clip, int32_t size, int32_t bgr, uint8_t* out_dev,
                     void* stream);

/* Device frame ring feed (stands in for the per-track deques of decoded frames the streaming
 * callers keep on the host, test/af_realtime.py:456-479, TEST2.py:354-391): queue `n` host->device copies, item i
 * being rows [row0[i], row1[i]) of the decoded frame frames_host[i] (row pitch `pitch` bytes, pinned memory for
 * truly asynchronous copies) into the same rows of ring slot slots[i] (slot k starts at ring_dev + k*slot_stride).
 * A caller uploads each NEW frame once - only the rows its face box covers - and every overlapping window then
 * gathers from the ring through af_crop_infer / af_crop_u8 descriptors. */
af_status af_ring_put_rows(uint8_t* ring_dev, int64_t slot_stride, int64_t pitch, int32_t n, const int32_t* slots,
                           const uint8_t* const* frames_host, const int32_t* row0, const int32_t* row1, void* stream);

/* The same feed restricted to the face box: item i copies the pixels [x0,x1) x [y0,y1) of boxes_xyxy[4i..4i+3] (the
 * enlarged crop box get_crop_box returns, altfreezing/test_tools/utils.py:13-24; byte columns widened to 16-byte
 * multiples) with one strided host->device copy.  K1 never reads a pixel outside a frame's box (the reference pastes the
 * crop onto a zero canvas, faster_crop_align_xray.py:60-75), so nothing else of the frame has to reach the device:
 * about a third of the bytes of whole rows for a 720p call frame. */
af_status af_ring_put_boxes(uint8_t* ring_dev, int64_t slot_stride, int64_t pitch, int32_t n, const int32_t* slots,
                            const uint8_t* const* frames_host, const int32_t* boxes_xyxy, void* stream);

/* K1 on its own (SURVEY.md 8b `crop_pack`): the same warp followed by the callers' pack step
 *   x = (float(u8) - 255*mean_c) / (255*std_c), NTHWC -> NCTHW   (demo.py:84-87,317-319; TEST2.py:147-158)
 * written straight into a caller tensor viewed as [B,3,T,S,S] with ELEMENT strides `out_strides`
 * (contiguous NCTHW, channels_last_3d, a slice of a larger batch ...), fp32 or bf16 — the aligned u8 clip
 * never exists in memory.  Equals af_crop_u8 followed by that pack step bit for bit (IEEE fp32 division). */
af_status af_crop_pack(const af_frame_desc* frames_dev, const af_clip_geom* geom_dev, int32_t batch,
                       int32_t frames_per_clip, int32_t size, int32_t bgr, const float mean255[3],
                       const float std255[3], void* clip_out_dev, int32_t out_dtype /* AF_F32 | AF_BF16 */,
                       const int64_t out_strides[5], void* stream);

/* Fused fast path: the same warp, normalised and written directly into the engine's
 * internal clip layout, followed by the trunk.  Replaces crop_align_func + the three
 * torch pack lines + classifier(images_t) of the hot loop (demo.py:309-328). */
af_status af_crop_infer(af_handle h, const af_frame_desc* frames_dev, const af_clip_geom* geom_dev,
                        int32_t batch, int32_t bgr, const float mean255[3], const float std255[3],
                        float* logits_dev, float* scores_dev, float* features_dev, void* stream);

/* Test/diagnostic entry: one conv (+bias, +residual, +ReLU) on NDHWC device tensors of the
 * engine's precision (fp32 or bf16), through the same kernels the trunk uses.
 * impl: 0 = engine's choice, 1 = force the SIMT kernel, 2 = force the tcgen05 kernel. */
af_status af_conv_ndhwc(const void* x_dev, const af_conv_desc* conv_host, const void* residual_dev,
                        void* y_dev, int32_t batch, int32_t t, int32_t hgt, int32_t wid,
                        int32_t relu, int32_t precision, int32_t impl, void* stream);

/* Test/diagnostic entry for the fused projection shortcut of a block's first ResBlock
 * (resnet_helper.py:411-423,438-441: relu(branch1_bn(branch1(x)) + branch2(x))): the pointwise
 * `conv` over x_dev [B,t,hgt,wid,cin] plus the pointwise, spatially strided `shortcut` over
 * x2_dev [B,t,hgt2,wid2,cin2], accumulated in one tcgen05 GEMM (bf16 NDHWC in/out, fp32 accumulate). */
af_status af_conv_shortcut_ndhwc(const void* x_dev, const af_conv_desc* conv_host, const void* x2_dev,
                                 const af_conv_desc* shortcut_host, void* y_dev, int32_t batch, int32_t t,
                                 int32_t hgt, int32_t wid, int32_t hgt2, int32_t wid2, int32_t relu,
                                 void* stream);

/* Test/diagnostic entry: the fused stem kernel (K3: conv k[5,7,7] s[1,2,2] p[2,3,3] + folded BN + ReLU +
 * MaxPool3d k[1,3,3] s[1,2,2] p[0,1,1]; stem_helper.py:156-178) on its own, exactly as the bf16 trunk launches it.
 * clip_dev: bf16 [B,T,S,S,4] (channel 3 ignored); y_dev: bf16 [B,T,S/4,S/4,64].  per_frame_kernel = 0 runs the
 * temporal-sweep kernel when T % 4 == 0 (the production path), 1 forces the per-frame row-halo kernel. */
af_status af_stem_pool_ndhwc4(const void* clip_dev, const af_conv_desc* stem_host, void* y_dev, int32_t batch,
                              int32_t t, int32_t s, int32_t per_frame_kernel, void* stream);

/* Test/diagnostic entry: the fused tail of an s2 bottleneck block (conv_bc_fused.cu),
 *   y = relu(c(relu(b(x))) + residual)          (identity-shortcut blocks), or
 *   y = relu(c(relu(b(x))) + shortcut(x2))      (first block of the stage: projection shortcut as a second K block)
 * with b = 1x3x3 64->64, c = 1x1x1 64->256, shortcut = 1x1x1 stride 1 64->256, all with folded BN
 * (resnet_helper.py:311-326,411-423,438-444) on bf16 NDHWC device tensors: x, x2 [B,T,H,W,64], residual / y
 * [B,T,H,W,256].  Give EITHER residual_dev OR (x2_dev, shortcut_host); the other(s) NULL. */
af_status af_conv_bc_fused_ndhwc(const void* x_dev, const af_conv_desc* conv_b_host, const af_conv_desc* conv_c_host,
                                 const void* residual_dev, const void* x2_dev, const af_conv_desc* shortcut_host,
                                 void* y_dev, int32_t batch, int32_t t, int32_t hgt, int32_t wid, void* stream);

/* The same fused tail for the LAST block of s2, with the next stage's MaxPool3d k = s = [2,1,1]
 * (video_model_builder.py:474-480,566-568: pathway0_pool before s3) taken in the epilogue:
 *   y[b, j] = max(f(x)[b, 2j], f(x)[b, 2j+1]),  f = relu(c(relu(b(x))) + residual)
 * x [B,T,H,W,64], residual [B,T,H,W,256], y [B,T/2,H,W,256] (bf16 NDHWC device tensors), T even. */
af_status af_conv_bc_fused_tpool_ndhwc(const void* x_dev, const af_conv_desc* conv_b_host, const af_conv_desc* conv_c_host,
                                       const void* residual_dev, void* y_dev, int32_t batch, int32_t t, int32_t hgt,
                                       int32_t wid, void* stream);

/* Copy intermediate activations of the LAST af_forward/af_infer call out for stage
 * parity tests: which = 1..5 (s1..s5 outputs) as fp32 NCTHW [B,C,T,H,W] into out_dev.
 * Requires option "keep_stages" = 1 (costs extra memory). */
af_status af_get_stage(af_handle h, int32_t which, float* out_dev, int64_t capacity_elems,
                       int32_t dims_out[5], void* stream);

#ifdef __cplusplus
}
#endif
#endif /* AFB200_H */
