// Shared declarations for the afb200 CUDA sources (internal; the public ABI is include/afb200.h).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

namespace afb {

typedef __nv_bfloat16 bf16;

// A convolution expressed on channels-last (NDHWC) activations.
// Input may be a strided view (the padded clip buffer); output / residual are dense [M, Cout].
struct ConvProblem {
  const void* x;
  const void* w;       // SIMT: float [taps][cin][cout];  UMMA: bf16 [taps][cout][cin]
  const float* bias;   // [cout]
  const float* bias_host = nullptr;   // host copy of `bias` (kernels that take the bias through their launch parameters)
  const void* res;     // nullptr or dense [M, cout]
  void* y;             // dense [M, cout]
  int B, Ti, Hi, Wi, Cin;
  long long xsB, xsT, xsH, xsW;   // element strides of the input view (channel stride 1)
  int To, Ho, Wo, Cout;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  int relu;
  long long M;         // B*To*Ho*Wo
  int pool_t = 0;      // conv_umma pointwise only: fuse MaxPool3d k=s=[2,1,1] over frame pairs; y is then
                       // [B, To/2, Ho, Wo, Cout] (needs relu, To even, Ho*Wo % 64 == 0)
  // conv_umma only: a second operand source accumulated into the same output tile — the block's projection
  // shortcut (pointwise conv of stride [1,sh2,sw2] over the block input, resnet_helper.py:411-423) fused into
  // its `c` conv: y = relu(conv(x) + conv2(x2) + bias), bias = both folded-BN biases summed by the caller.
  const void* x2 = nullptr;     // dense NDHWC [B, T2, H2, W2, Cin2]
  const void* w2 = nullptr;     // bf16 [Cout][Cin2]
  int Cin2 = 0, T2 = 0, H2 = 0, W2 = 0, sh2 = 1, sw2 = 1;
  int pool_hw = 0;     // conv_rows only: fuse MaxPool3d k[1,3,3] s[1,2,2] p[0,1,1]; y is then the
                       // ZERO-INITIALISED pooled tensor [B*To, Ho/2, Wo/2, Cout] (needs relu)
};

void set_error(const char* fmt, ...);
extern thread_local long long g_launches;
// Upper bound on the CTAs (= SMs, all tcgen05 kernels are persistent with 1 CTA/SM) a conv launch may occupy;
// 0 = whole device.  Lets two half-batches run side by side on disjoint SM sets (engine option "sm_limit").
extern thread_local int g_sm_limit;
inline int limit_grid(int want, int sms) { const int cap = (g_sm_limit > 0 && g_sm_limit < sms) ? g_sm_limit : sms; return want < cap ? want : cap; }   // kernels launched by this thread (copied into engines)

// Per-op timing to stderr when AFB200_TRACE=1 (synchronises; bring-up/profiling aid only).
struct OpTrace {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaStream_t s;
  static bool enabled();
  explicit OpTrace(cudaStream_t st);
  void done(const char* what, double flops, double bytes);
};

#define AFB_CUDA(call)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (call);                                                             \
    if (_e != cudaSuccess) {                                                             \
      afb::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__,                \
                     cudaGetErrorString(_e));                                            \
      return AF_ERR_CUDA;                                                                \
    }                                                                                    \
  } while (0)

// conv_simt.cu
int conv_simt_launch(const ConvProblem& p, bool is_bf16, cudaStream_t s);
// conv_umma.cu
bool conv_umma_supported(const ConvProblem& p);
int conv_umma_launch(const ConvProblem& p, cudaStream_t s);
int conv_umma_init();
void conv_umma_timeline_dump(int n);
void conv_umma_force_block_n(int bn);
void conv_umma_set_pair_mode(int mode);
// conv_rows.cu (row-halo tcgen05 kernel for Cout=64 layers with vertical taps)
bool conv_rows_supported(const ConvProblem& p);
int conv_rows_launch(const ConvProblem& p, cudaStream_t s);
int conv_rows_init();
// conv_tsweep.cu (temporal-sweep tcgen05 kernel for 3x1x1 convs with Cout = 64)
bool conv_tsweep_supported(const ConvProblem& p);
int conv_tsweep_launch(const ConvProblem& p, cudaStream_t s);
int conv_tsweep_init();
// conv_tf32.cu: fp32 NDHWC activations, fp32 weights [taps][Cout][Cin], tcgen05.mma kind::tf32
int conv_tf32_init();
bool conv_tf32_supported(const ConvProblem& p);
int conv_tf32_launch(const ConvProblem& p, cudaStream_t s);
int conv_tf32_stem_launch(const void* clip_phys, int B, int T, int S, const void* w35, const float* bias, void* y, cudaStream_t s);
int conv_stem_direct_launch(const void* clip_phys, int B, int T, int S, const void* w35, const float* bias_host, void* y,
                            int pool, cudaStream_t s, int force_per_frame = 0);
// conv_bc_fused.cu (s2 bottleneck tail: 1x3x3 64->64 + ReLU, then 1x1x1 64->256 + residual + ReLU, one kernel)
bool conv_bc_fused_supported(const ConvProblem& b, const ConvProblem& c);
int conv_bc_fused_launch(const ConvProblem& b, const ConvProblem& c, cudaStream_t s);
int conv_bc_fused_init();
int ftcn_stem_umma_launch(const void* clip_phys, int B, int T, int S, const void* w2, const float* bias_host, void* y,
                          cudaStream_t s);
// pool_head.cu
int maxpool_spatial_launch(const void* x, void* y, int B, int T, int H, int W, int C, bool is_bf16,
                           cudaStream_t s);   // k[1,3,3] s[1,2,2] p[0,1,1]
int maxpool_temporal_launch(const void* x, void* y, int B, int T, int H, int W, int C, bool is_bf16,
                            cudaStream_t s);  // k[2,1,1] s[2,1,1]
int head_pool_slices(int P);      // partial-sum slices head_launch uses for P positions (sizes features_ws)
int head_launch(const void* x, int B, int P, int C, bool is_bf16, const float* fc_w, float fc_b,
                float* features_ws, float* features_out, float* logits, float* scores,
                cudaStream_t s);
// mean over P positions: x [N, P, C] -> out fp32 [N, C]
int spatial_mean_launch(const void* x, int N, int P, int C, bool is_bf16, float* out, cudaStream_t s);
int ndhwc_to_ncthw_f32_launch(const void* x, float* y, int B, int T, int H, int W, int C,
                              bool is_bf16, cudaStream_t s);
// crop_pack.cu
struct ClipLayout {     // engine-internal normalised clip: padded NDHWC4
  void* base;           // points at logical element (b=0,t=0,y=0,x=0,c=0)
  long long sB, sT, sH, sW;   // element strides
  int T, S;
  bool is_bf16;
  long long sC = 0;     // 0: 4 packed channels per pixel (the engine's NDHWC4); else 3 channels sC elements apart
};
int stem_unfold_launch(const ClipLayout& clip, int clip0, int B, void* U, cudaStream_t s);
int pack_clip_launch(const void* src, int dtype, const long long strides[5], int B,
                     const ClipLayout& dst, cudaStream_t s);
int pack_u8_launch(const uint8_t* src, int B, const float mean[3], const float stdv[3],
                   const ClipLayout& dst, cudaStream_t s);
struct FrameDesc { const uint8_t* data; long long pitch; int height, width; int box[4]; };
struct ClipGeom { double tfm[6]; int left_top[2]; int canvas_wh[2]; };
int crop_launch(const FrameDesc* frames, const ClipGeom* geom, int B, int T, int S, int bgr,
                uint8_t* out_u8, const ClipLayout* dst, const float mean[3], const float stdv[3],
                cudaStream_t s);

// ftcn.cu (FTCN-TT plugin: fused temporal-only stem, 2x2 max-pool, transformer head; all pointers device memory)
struct TTLayerDev {
  const float *ln1_w, *ln1_b, *qkv_w, *out_w, *out_b, *ln2_w, *ln2_b, *fc1_w, *fc1_b, *fc2_w, *fc2_b;
};
struct TTHeadDev {
  int dim = 0, tokens = 0, heads = 0, dim_head = 0, mlp_dim = 0;
  std::vector<TTLayerDev> layers;
  const float *cls_token = nullptr, *pos_embedding = nullptr, *norm_w = nullptr, *norm_b = nullptr, *fc_w = nullptr;
  float fc_b = 0.f;
};
int ftcn_stem_launch(const ClipLayout& clip, int clip0, int B, const float* w_tap_c_cout, const float* bias, void* y,
                     cudaStream_t s);
int maxpool_hw2_launch(const void* x, void* y, long long BT, int H, int W, int C, bool is_bf16, cudaStream_t s);
long long tt_head_workspace_floats(const TTHeadDev& h, int B);
// tokens fp32 [B, h.tokens, h.dim] -> logits / sigmoid scores [B], optional normalised cls features [B, h.dim]
int tt_head_launch(const TTHeadDev& h, const float* tokens, int B, float* ws, float* features_out, float* logits,
                   float* scores, cudaStream_t s);

}  // namespace afb
