// C-ABI entry points and the engine (weights, workspace, batch schedule) of libafb200.so.
// Public contract: include/afb200.h.  The schedule restates ResNet.forward
// (altfreezing/slowfast/models/video_model_builder.py:561-578): stem -> s2 -> temporal
// max-pool -> s3 -> s4 -> s5 -> head, with every BatchNorm folded into its conv and the
// residual add + ReLU fused into the `c` conv's epilogue (resnet_helper.py:438-444).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "../../include/afb200.h"
#include "common.cuh"

namespace afb {

static thread_local char g_err[1024] = "";
thread_local long long g_launches = 0;
thread_local int g_sm_limit = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool OpTrace::enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("AFB200_TRACE"); on = (e && e[0] == '1') ? 1 : 0; }
  return on == 1;
}
OpTrace::OpTrace(cudaStream_t st) : s(st) {
  if (!enabled()) return;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s);
}
void OpTrace::done(const char* what, double flops, double bytes) {
  if (!enabled()) return;
  cudaEventRecord(e1, s);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  fprintf(stderr, "[afb200] %-46s %8.3f ms  %8.1f TFLOP/s  %8.1f GB/s\n", what, ms, flops / (ms * 1e9), bytes / (ms * 1e6));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
}

static inline float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

struct ConvLayer {
  int cin, cout, cin_p;
  int kt, kh, kw, st, sh, sw, pt, ph, pw;
  float* w_simt = nullptr;  // [taps][cin_p][cout] fp32 (bf16-rounded values in bf16 precision)
  bf16* w_umma = nullptr;   // [taps][cout][cin_p] bf16 (bf16 precision only)
  float* w_tf32 = nullptr;  // [taps][cout][cin_p] fp32 (tf32 precision only: the tensor-core kernel's B operand)
  float* bias = nullptr;    // [cout]
  std::vector<float> bias_h;  // host copy of bias
};

static void free_layer(ConvLayer& L) {
  if (L.w_simt) cudaFree(L.w_simt);
  if (L.w_umma) cudaFree(L.w_umma);
  if (L.w_tf32) cudaFree(L.w_tf32);
  if (L.bias) cudaFree(L.bias);
  L.w_simt = nullptr; L.w_umma = nullptr; L.w_tf32 = nullptr; L.bias = nullptr;
}

// fp32 -> nearest value with 10 mantissa bits (round half away from zero, like cvt.rna.tf32.f32): the tensor core
// truncates its TF32 operands, rounding here keeps the weights' error unbiased
static float tf32_round(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  if ((u & 0x7F800000u) == 0x7F800000u) return v;      // inf / nan
  u = (u + 0x1000u) & ~0x1FFFu;
  memcpy(&v, &u, 4);
  return v;
}

// Re-lay-out one folded conv for the kernels and upload it.
static int upload_layer(const af_conv_desc& d, bool is_bf16, ConvLayer& L, bool tf32 = false) {
  L.cin = d.cin; L.cout = d.cout; L.cin_p = (d.cin + 3) / 4 * 4;
  L.kt = d.kt; L.kh = d.kh; L.kw = d.kw; L.st = d.st; L.sh = d.sh; L.sw = d.sw;
  L.pt = d.pt; L.ph = d.ph; L.pw = d.pw;
  const int taps = d.kt * d.kh * d.kw;
  const size_t n = (size_t)taps * L.cin_p * d.cout;
  std::vector<float> ws(n, 0.f);
  std::vector<bf16> wu;
  if (is_bf16) wu.assign(n, __float2bfloat16_rn(0.f));
  for (int co = 0; co < d.cout; ++co)
    for (int ci = 0; ci < d.cin; ++ci)
      for (int tp = 0; tp < taps; ++tp) {
        float v = d.weight[((size_t)co * d.cin + ci) * taps + tp];
        if (is_bf16) {
          v = bf16_round(v);
          wu[((size_t)tp * d.cout + co) * L.cin_p + ci] = __float2bfloat16_rn(v);
        }
        ws[((size_t)tp * L.cin_p + ci) * d.cout + co] = v;
      }
  AFB_CUDA(cudaMalloc(&L.w_simt, n * sizeof(float)));
  AFB_CUDA(cudaMemcpy(L.w_simt, ws.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  if (is_bf16) {
    AFB_CUDA(cudaMalloc(&L.w_umma, n * sizeof(bf16)));
    AFB_CUDA(cudaMemcpy(L.w_umma, wu.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
  }
  if (tf32 && !is_bf16 && L.cin_p % 32 == 0) {      // K-major per tap for the tensor-core kernel (conv_tf32.cu)
    std::vector<float> wt(n, 0.f);
    for (int co = 0; co < d.cout; ++co)
      for (int ci = 0; ci < d.cin; ++ci)
        for (int tp = 0; tp < taps; ++tp)
          wt[((size_t)tp * d.cout + co) * L.cin_p + ci] = tf32_round(d.weight[((size_t)co * d.cin + ci) * taps + tp]);
    AFB_CUDA(cudaMalloc(&L.w_tf32, n * sizeof(float)));
    AFB_CUDA(cudaMemcpy(L.w_tf32, wt.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  }
  AFB_CUDA(cudaMalloc(&L.bias, d.cout * sizeof(float)));
  AFB_CUDA(cudaMemcpy(L.bias, d.bias, d.cout * sizeof(float), cudaMemcpyHostToDevice));
  L.bias_h.assign(d.bias, d.bias + d.cout);
  return AF_OK;
}

// bf16 engine: the stem conv re-expressed on the x/row-pair-unfolded clip (crop_pack.cu,
// stem_unfold_kernel): 64 "channels" k = dyy*28 + dx*4 + c, taps (dt, p) with input row
// dy = 2p + dyy.  W'[dt][p][cout][k] = W[cout][c][dt][2p+dyy][dx].
static int upload_stem_unfolded(const af_conv_desc& d, ConvLayer& L) {
  if (d.kt != 5 || d.kh != 7 || d.kw != 7 || d.cin != 3 || d.st != 1 || d.sh != 2 || d.sw != 2 || d.pt != 2 ||
      d.ph != 3 || d.pw != 3)
    return AF_ERR_INVALID;
  L.cin = 64; L.cin_p = 64; L.cout = d.cout;
  L.kt = 5; L.kh = 4; L.kw = 1; L.st = 1; L.sh = 1; L.sw = 1; L.pt = 2; L.ph = 0; L.pw = 0;
  const int taps = 20;
  const size_t n = (size_t)taps * d.cout * 64;
  std::vector<bf16> wu(n, __float2bfloat16_rn(0.f));
  std::vector<float> ws(n, 0.f);
  for (int co = 0; co < d.cout; ++co)
    for (int c = 0; c < 3; ++c)
      for (int dt = 0; dt < 5; ++dt)
        for (int dy = 0; dy < 7; ++dy)
          for (int dx = 0; dx < 7; ++dx) {
            const float v = bf16_round(d.weight[((((size_t)co * 3 + c) * 5 + dt) * 7 + dy) * 7 + dx]);
            const int pr = dy / 2, dyy = dy % 2, k = dyy * 28 + dx * 4 + c, tap = dt * 4 + pr;
            wu[((size_t)tap * d.cout + co) * 64 + k] = __float2bfloat16_rn(v);
            ws[((size_t)tap * 64 + k) * d.cout + co] = v;
          }
  AFB_CUDA(cudaMalloc(&L.w_umma, n * sizeof(bf16)));
  AFB_CUDA(cudaMemcpy(L.w_umma, wu.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
  AFB_CUDA(cudaMalloc(&L.w_simt, n * sizeof(float)));
  AFB_CUDA(cudaMemcpy(L.w_simt, ws.data(), n * sizeof(float), cudaMemcpyHostToDevice));
  AFB_CUDA(cudaMalloc(&L.bias, d.cout * sizeof(float)));
  AFB_CUDA(cudaMemcpy(L.bias, d.bias, d.cout * sizeof(float), cudaMemcpyHostToDevice));
  L.bias_h.assign(d.bias, d.bias + d.cout);
  return AF_OK;
}

// Direct stem (conv_rows.cu: conv_stem_direct_launch): W35[dt*7+dy][cout][dx*4+c] = W[cout][c][dt][dy][dx].
static int upload_stem_direct(const af_conv_desc& d, bf16** out) {
  const size_t n = (size_t)35 * d.cout * 32;
  std::vector<bf16> w(n, __float2bfloat16_rn(0.f));
  for (int co = 0; co < d.cout; ++co)
    for (int c = 0; c < 3; ++c)
      for (int dt = 0; dt < 5; ++dt)
        for (int dy = 0; dy < 7; ++dy)
          for (int dx = 0; dx < 7; ++dx)
            w[((size_t)(dt * 7 + dy) * d.cout + co) * 32 + dx * 4 + c] =
                __float2bfloat16_rn(d.weight[((((size_t)co * 3 + c) * 5 + dt) * 7 + dy) * 7 + dx]);
  AFB_CUDA(cudaMalloc(out, n * sizeof(bf16)));
  AFB_CUDA(cudaMemcpy(*out, w.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
  return AF_OK;
}

// FTCN-TT stem on the tensor cores (conv_rows.cu: ftcn_stem_umma_kernel): the k[5,1,1] conv as a GEMM over horizontal
// pixel pairs.  W2[chunk = tap dt (5 = zero)][n][e]: n < 64 -> output channel n of the pair's EVEN pixel, weights on the
// pair's elements e = 0..2; n >= 64 -> channel n-64 of the ODD pixel, weights on e = 4..6.
static int upload_ftcn_stem_w2(const af_conv_desc& d, bf16** out) {
  const size_t n = (size_t)6 * 128 * 8;
  std::vector<bf16> w(n, __float2bfloat16_rn(0.f));
  for (int co = 0; co < 64; ++co)
    for (int c = 0; c < 3; ++c)
      for (int dt = 0; dt < 5; ++dt) {
        const bf16 v = __float2bfloat16_rn(d.weight[((size_t)co * 3 + c) * 5 + dt]);
        w[((size_t)dt * 128 + co) * 8 + c] = v;
        w[((size_t)dt * 128 + 64 + co) * 8 + 4 + c] = v;
      }
  AFB_CUDA(cudaMalloc(out, n * sizeof(bf16)));
  AFB_CUDA(cudaMemcpy(*out, w.data(), n * sizeof(bf16), cudaMemcpyHostToDevice));
  return AF_OK;
}

struct Dims { int T, H, W, C; long long elems() const { return (long long)T * H * W * C; } };

static Dims conv_out(const ConvLayer& L, Dims in) {
  Dims o;
  o.T = (in.T + 2 * L.pt - L.kt) / L.st + 1;
  o.H = (in.H + 2 * L.ph - L.kh) / L.sh + 1;
  o.W = (in.W + 2 * L.pw - L.kw) / L.sw + 1;
  o.C = L.cout;
  return o;
}

}  // namespace afb

using namespace afb;

struct af_engine {
  int device = 0;
  bool is_bf16 = false;
  bool tf32 = false;        // fp32 storage, trunk convs on the tensor cores with TF32 operands (conv_tf32.cu)
  int T = 32, S = 224, max_batch = 1;
  std::vector<ConvLayer> convs;
  ConvLayer stem_u;          // unfolded stem (bf16 engine), valid if has_stem_u
  bool has_stem_u = false;
  bf16* ftcn_w2 = nullptr;   // FTCN-TT tensor-core stem weights [6][128][8] bf16 (upload_ftcn_stem_w2)
  bf16* stem_w35 = nullptr;  // direct stem weights [35 (dt,dy)][cout][32 (dx*4+c)] bf16
  float* stem_w35_tf32 = nullptr;   // the same layout in fp32 (tf32 precision: conv_tf32_stem_launch)
  int stem_direct = 0;       // 1 usable, 0 not tried / disabled, -1 tensor-map encode refused
  int stem = 0;
  std::vector<af_block_desc> blocks;
  std::vector<float*> fused_bias;   // per block: bias(c) + bias(branch1) for the fused projection shortcut (or null)
  std::vector<std::vector<float>> fused_bias_h;   // host copies of fused_bias
  float* fc_w = nullptr;
  float fc_b = 0.f;
  int feat_dim = 0;
  // FTCN-TT plugin: temporal-only stem with a 2x2 max-pool behind its BN, transformer head
  bool stem_pool2 = false;
  bool has_tt = false;
  TTHeadDev tt;
  std::vector<float*> tt_arrays;     // device copies the TTHeadDev pointers refer to
  float* tt_ws = nullptr;            // transformer activations for max_batch clips
  float* tok_ws = nullptr;           // per-frame mean features [max_batch * tokens, dim]
  int cb_front = 32, cb_back = 32;  // clips per chunk: stem..s2 / s3..head (tuned on B200)
  int conv_impl = 0;       // 0 auto, 1 force SIMT, 2 force UMMA where supported
  bool keep_stages = false;
  int sm_limit = 0;              // CTAs per conv launch (0 = all SMs)
  float* frame_feat_out = nullptr;   // set by af_forward_frames for the duration of one call
  bool pooled_already = false;   // the previous block's `c` conv already applied the temporal max-pool
  long long launches = 0;

  // clip buffer (padded NDHWC4) for max_batch clips
  void* clip_raw = nullptr;
  ClipLayout clip;
  // workspace
  size_t esz = 4;
  int split = -1;            // first block of the "back" part (-1: none)
  long long front_max = 0, back_max = 0, group_elems = 0;
  void* fbuf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void* bbuf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  void* gbuf = nullptr;      // group input of the back part [cb_back, ...]
  float* feat_ws = nullptr;  // [max_batch * HEAD slices, feat_dim] pooled-feature partial sums (pool_head.cu)
  int feat_slices = 16;
  uint8_t* u8_stage = nullptr;
  cudaStream_t copy_stream = nullptr;          // H2D of host clips, overlapped with the trunk
  std::vector<cudaEvent_t> copy_events;
  float* out_stage = nullptr;  // [2*max_batch] logits, scores (device staging for *_host calls)
  // pipelined host API (af_submit_u8_host / af_wait): two slots, each with its own staging and events
  struct HostSlot {
    uint8_t* u8_dev = nullptr;      // [max_batch] u8 clips
    float* out_dev = nullptr;       // [2*max_batch] logits, scores
    float* out_pinned = nullptr;    // same, pinned host memory
    cudaEvent_t copied = nullptr, done = nullptr;
    int batch = 0;                  // > 0 while a submission is outstanding
  } host_slot[2];
  int next_slot = 0;
  // event-based conv timing (option profile_events)
  bool profile_events = false;
  struct EvRec { cudaEvent_t e0, e1; int kind; double flops, bytes; };
  std::vector<EvRec> ev_recs;
  std::vector<cudaEvent_t> ev_pool;
  // kind 0: tcgen05 launch whose algorithmic intensity is above the roofline ridge (tensor-bound), 1: CUDA-core conv,
  // 2: tcgen05 launch below the ridge (HBM-bound)
  // 3: the feeder of a front chunk (crop / pack kernel = K1)
  double stat_ms[4] = {0, 0, 0, 0}, stat_flops[4] = {0, 0, 0, 0}, stat_launches[4] = {0, 0, 0, 0}, stat_kbytes[4] = {0, 0, 0, 0};
  double stat_bytes = 0;
  double ridge_flop_per_byte = 208.0;   // measured sustained bf16 TFLOP/s / measured HBM TB/s (option "ridge_x1000")
  // kept stages (fp32 NCTHW) for parity tests
  float* stage_buf[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  Dims stage_dims[5];
  int stage_batch = 0;
};

namespace afb {

// Brackets one conv launch with CUDA events when option profile_events is on (no synchronisation).
struct ProfRec {
  af_engine* e;
  cudaStream_t s;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  bool on;
  ProfRec(af_engine* e_, cudaStream_t s_) : e(e_), s(s_), on(e_ && e_->profile_events) {
    if (!on) return;
    for (cudaEvent_t* pe : {&e0, &e1}) {
      if (e->ev_pool.empty()) { cudaEventCreate(pe); } else { *pe = e->ev_pool.back(); e->ev_pool.pop_back(); }
    }
    cudaEventRecord(e0, s);
  }
  void done(int kind, double flops, double bytes) {
    if (!on) return;
    cudaEventRecord(e1, s);
    e->ev_recs.push_back({e0, e1, kind, flops, bytes});
  }
};

// A projection shortcut fused into the conv it is added to (bf16 tcgen05 path only): conv `L` over `x`,
// both folded biases pre-summed in `bias`.
struct FusedShortcut { const ConvLayer* L; const void* x; Dims in; const float* bias; const float* bias_host; };

static ConvProblem make_problem(const ConvLayer& L, const void* x, Dims in, long long xsB, long long xsT, long long xsH,
                                long long xsW, int B, const void* res, void* y, bool relu, int pool_hw, int pool_t,
                                const FusedShortcut* sc) {
  ConvProblem p;
  p.x = x; p.bias = sc ? sc->bias : L.bias; p.res = res; p.y = y;
  p.bias_host = sc ? sc->bias_host : (L.bias_h.empty() ? nullptr : L.bias_h.data());
  if (sc) {
    p.x2 = sc->x; p.w2 = sc->L->w_umma; p.Cin2 = sc->L->cin_p;
    p.T2 = sc->in.T; p.H2 = sc->in.H; p.W2 = sc->in.W; p.sh2 = sc->L->sh; p.sw2 = sc->L->sw;
  }
  p.B = B; p.Ti = in.T; p.Hi = in.H; p.Wi = in.W; p.Cin = L.cin_p;
  p.xsB = xsB; p.xsT = xsT; p.xsH = xsH; p.xsW = xsW;
  Dims o = conv_out(L, in);
  p.To = o.T; p.Ho = o.H; p.Wo = o.W; p.Cout = L.cout;
  p.kt = L.kt; p.kh = L.kh; p.kw = L.kw; p.st = L.st; p.sh = L.sh; p.sw = L.sw;
  p.pt = L.pt; p.ph = L.ph; p.pw = L.pw;
  p.relu = relu ? 1 : 0;
  p.M = (long long)B * o.T * o.H * o.W;
  p.pool_hw = pool_hw;
  p.pool_t = pool_t;
  p.w = L.w_umma;
  return p;
}

static ConvProblem dense_problem(const ConvLayer& L, const void* x, Dims in, int B, const void* res, void* y, bool relu) {
  const long long sW = in.C, sH = (long long)in.W * in.C, sT = sH * in.H, sB = sT * in.T;
  return make_problem(L, x, in, sB, sT, sH, sW, B, res, y, relu, 0, 0, nullptr);
}

// The `b` -> `c` tail of an s2-shaped bottleneck block in one kernel (conv_bc_fused.cu) when the shapes allow it.
// Returns 1 if it ran, 0 if the caller should run the two convs separately, < 0 on error.
static int try_fused_bc(af_engine* e, const ConvLayer& Lb, const ConvLayer& Lc, const void* xb, Dims db_in, int B,
                        const void* res, void* y, cudaStream_t s, const FusedShortcut* sc = nullptr, int pool_t = 0) {
  static const bool off = getenv("AFB200_NO_FUSED_BC") != nullptr;
  static const bool off_sc = getenv("AFB200_NO_FUSED_BC_SHORTCUT") != nullptr;
  static const bool off_tp = getenv("AFB200_NO_FUSED_BC_TPOOL") != nullptr;
  if (off || (sc && off_sc) || (pool_t && off_tp) || !Lb.w_umma || !Lc.w_umma) return 0;
  const ConvProblem pb = dense_problem(Lb, xb, db_in, B, nullptr, nullptr, true);
  const Dims dmid = conv_out(Lb, db_in);
  const long long sW = dmid.C, sH = (long long)dmid.W * dmid.C, sT = sH * dmid.H, sB = sT * dmid.T;
  const ConvProblem pc = make_problem(Lc, nullptr, dmid, sB, sT, sH, sW, B, sc ? nullptr : res, y, true, 0, pool_t, sc);
  if (!conv_bc_fused_supported(pb, pc)) return 0;
  OpTrace tr(s);
  ProfRec prec(e, s);
  int rc = conv_bc_fused_launch(pb, pc, s);
  if (rc) return rc;
  const double k2 = sc ? (double)sc->L->cin_p : 0.0;
  const double flops = 2.0 * (double)pb.M * (Lb.cout * 9.0 * Lb.cin_p + (double)Lc.cout * (Lc.cin_p + k2));
  const double bytes = ((double)pb.M * (Lb.cin_p + k2 + (sc ? 1.0 : pool_t ? 1.5 : 2.0) * Lc.cout) + 9.0 * Lb.cin_p * Lb.cout +
                        (Lc.cin_p + k2) * Lc.cout) * 2.0;
  prec.done(flops >= (e ? e->ridge_flop_per_byte : 208.0) * bytes ? 0 : 2, flops, bytes);
  if (OpTrace::enabled()) {
    char nm[128];
    snprintf(nm, sizeof(nm), "conv fused b k1x3x3 + c k1x1x1 M=%lld N=%d->%d %s", pb.M, Lb.cout, Lc.cout,
             sc ? "+shortcut" : pool_t ? "+res +tpool" : "+res");
    tr.done(nm, flops, bytes);
  }
  return 1;
}

static int run_conv(af_engine* e, const ConvLayer& L, const void* x, Dims in, long long xsB, long long xsT,
                    long long xsH, long long xsW, int B, const void* res, void* y, bool relu, cudaStream_t s,
                    int impl_override = -1, int pool_hw = 0, int pool_t = 0, const FusedShortcut* sc = nullptr) {
  ConvProblem p = make_problem(L, x, in, xsB, xsT, xsH, xsW, B, res, y, relu, pool_hw, pool_t, sc);
  Dims o = conv_out(L, in);
  const bool is_bf16 = e ? e->is_bf16 : (L.w_umma != nullptr);
  const int impl = impl_override >= 0 ? impl_override : (e ? e->conv_impl : 0);
  OpTrace tr(s);
  int rc;
  const char* which = "simt";
  const double Kd = (double)L.kt * L.kh * L.kw * L.cin_p;
  const double K2d = sc ? (double)sc->L->cin_p : 0.0;
  const double conv_flops = 2.0 * (double)p.M * L.cout * (Kd + K2d);
  const double conv_bytes = ((double)B * in.elems() + (sc ? (double)B * sc->in.elems() : 0.0) +
                             (double)p.M * L.cout * ((res ? 1.0 : 0.0) + (pool_t ? 0.5 : pool_hw ? 0.25 : 1.0)) +
                             (Kd + K2d) * L.cout) * (is_bf16 ? 2.0 : 4.0);
  ProfRec prec(e, s);
  p.w = L.w_umma;
  if (pool_hw && !(is_bf16 && (impl == 0 || impl == 3) && conv_rows_supported(p))) {
    set_error("fused max-pool needs the row-halo tcgen05 kernel");
    return AF_ERR_INVALID;
  }
  if (pool_t && !(is_bf16 && impl != 1 && impl != 3 && conv_umma_supported(p))) {
    set_error("fused temporal max-pool needs the pointwise tcgen05 kernel");
    return AF_ERR_INVALID;
  }
  if (sc && !(is_bf16 && impl != 1 && impl != 3 && impl != 6 && conv_umma_supported(p))) {
    set_error("fused projection shortcut needs the pointwise tcgen05 kernel");
    return AF_ERR_INVALID;
  }
  if (sc) {
    rc = conv_umma_launch(p, s);
    which = "umma";
  } else if (!pool_t && is_bf16 && (impl == 0 || impl == 3) && conv_rows_supported(p)) {
    rc = conv_rows_launch(p, s);
    which = "urows";
  } else if (!pool_t && is_bf16 && (impl == 0 || impl == 6) && conv_tsweep_supported(p)) {
    rc = conv_tsweep_launch(p, s);
    which = "usweep";
  } else if (is_bf16 && impl != 1 && impl != 3 && impl != 6 && conv_umma_supported(p)) {
    rc = conv_umma_launch(p, s);
    which = "umma";
  } else if (is_bf16 && impl >= 2) {
    set_error("tcgen05 conv kernel (impl %d) does not take this shape", impl);
    return AF_ERR_INVALID;
  } else if (!is_bf16 && L.w_tf32 && (e ? e->tf32 : true) && impl != 1 && !pool_hw && !pool_t && !sc && conv_tf32_supported(p)) {
    p.w = L.w_tf32;
    rc = conv_tf32_launch(p, s);
    which = "utf32";
  } else {
    p.w = L.w_simt;
    rc = conv_simt_launch(p, is_bf16, s);
  }
  prec.done(which[0] != 'u' ? 1 : (conv_flops >= (e ? e->ridge_flop_per_byte : 208.0) * conv_bytes ? 0 : 2), conv_flops, conv_bytes);
  if (OpTrace::enabled()) {
    char nm[128];
    snprintf(nm, sizeof(nm), "conv %s k%dx%dx%d s%d M=%lld N=%d K=%d%s%s", which, L.kt, L.kh, L.kw, L.sh, p.M, L.cout,
             (int)(Kd + K2d), res ? " +res" : "", sc ? " +shortcut" : "");
    tr.done(nm, conv_flops, conv_bytes);
  }
  return rc;
}

static int dense_conv(af_engine* e, int idx, const void* x, Dims in, int B, const void* res, void* y, bool relu,
                      cudaStream_t s) {
  const long long sW = in.C, sH = (long long)in.W * in.C, sT = sH * in.H, sB = sT * in.T;
  return run_conv(e, e->convs[idx], x, in, sB, sT, sH, sW, B, res, y, relu, s);
}

static int keep_stage(af_engine* e, int which, const void* x, Dims d, int b0, int B, cudaStream_t s) {
  if (!e->keep_stages) return AF_OK;
  float*& buf = e->stage_buf[which];
  if (!buf) AFB_CUDA(cudaMalloc(&buf, (size_t)e->max_batch * d.elems() * sizeof(float)));
  e->stage_dims[which] = d;
  return ndhwc_to_ncthw_f32_launch(x, buf + (long long)b0 * d.elems(), B, d.T, d.H, d.W, d.C, e->is_bf16, s);
}

static bool is_stage_end(const af_engine* e, int bi) {
  return bi + 1 == (int)e->blocks.size() || e->blocks[bi + 1].branch1 >= 0;
}

// Run blocks [b_begin, b_end) on `B` clips whose input is `x` (dense NDHWC, dims `d`), using
// the 5 scratch buffers `buf`.  Returns the output buffer/dims through x/d.
static int run_blocks(af_engine* e, int b_begin, int b_end, const void*& x, Dims& d, int B, void* const buf[5],
                      int clip0, int& stage_no, cudaStream_t s) {
  for (int bi = b_begin; bi < b_end; ++bi) {
    const af_block_desc& blk = e->blocks[bi];
    // scratch buffers that do not alias the block input
    std::vector<void*> freeb;
    for (int i = 0; i < 5; ++i)
      if (buf[i] != x) freeb.push_back(buf[i]);
    if (blk.temporal_pool_before && e->pooled_already) {
      e->pooled_already = false;                 // fused into the producer's epilogue
    } else if (blk.temporal_pool_before) {
      void* pooled = freeb.back();
      freeb.pop_back();
      OpTrace tr(s);
      int rc = maxpool_temporal_launch(x, pooled, B, d.T, d.H, d.W, d.C, e->is_bf16, s);
      if (rc) return rc;
      tr.done("temporal maxpool 2x1x1", 0.0, (double)B * d.elems() * 1.5 * e->esz);
      for (int i = 0; i < 5; ++i)
        if (buf[i] == x) freeb.push_back(buf[i]);
      x = pooled; d.T /= 2;
    }
    if (freeb.size() < 4) { set_error("internal: scratch exhausted"); return AF_ERR_INVALID; }
    void *ya = freeb[0], *yb = freeb[1], *ysc = freeb[2], *yout = freeb[3];
    const void* shortcut = x;
    // Projection shortcuts are not run as convs of their own on the bf16 engine: their channel blocks are
    // accumulated into the block's `c` conv (one GEMM over K = Cin_c + Cin_x), which saves writing and
    // re-reading the widest tensor of the stage.
    static const bool no_scfuse = getenv("AFB200_NO_FUSED_SHORTCUT") != nullptr;
    FusedShortcut fsc = {nullptr, nullptr, d, nullptr, nullptr};
    if (blk.branch1 >= 0) {
      const ConvLayer& L1 = e->convs[blk.branch1];
      const ConvLayer& Lc1 = e->convs[blk.c];
      const bool fuse_sc = e->is_bf16 && e->conv_impl == 0 && !no_scfuse && !blk.spatial_pool && bi < (int)e->fused_bias.size() &&
                           e->fused_bias[bi] && L1.kt == 1 && L1.kh == 1 && L1.kw == 1 && L1.st == 1 &&
                           L1.cin_p % 64 == 0 && Lc1.kt == 1 && Lc1.kh == 1 && Lc1.kw == 1 && Lc1.st == 1 &&
                           Lc1.sh == 1 && Lc1.sw == 1;
      if (fuse_sc) {
        fsc.L = &L1; fsc.x = x; fsc.in = d; fsc.bias = e->fused_bias[bi]; fsc.bias_host = e->fused_bias_h[bi].data();
        shortcut = nullptr;
      } else {
        int rc = dense_conv(e, blk.branch1, x, d, B, nullptr, ysc, false, s);
        if (rc) return rc;
        shortcut = ysc;
      }
    }
    int rc = dense_conv(e, blk.a, x, d, B, nullptr, ya, true, s);
    if (rc) return rc;
    Dims da = conv_out(e->convs[blk.a], d);
    Dims db = conv_out(e->convs[blk.b], da);
    // fuse the next block's temporal max-pool into this `c` conv's epilogue when possible
    static const bool no_tfuse = getenv("AFB200_NO_FUSED_TPOOL") != nullptr;
    const ConvLayer& Lc = e->convs[blk.c];
    const bool fuse_t = e->is_bf16 && e->conv_impl == 0 && !e->keep_stages && !no_tfuse && !fsc.L && !blk.spatial_pool &&
                        bi + 1 < (int)e->blocks.size() &&
                        e->blocks[bi + 1].temporal_pool_before && (db.T % 2 == 0) && ((db.H * db.W) % 64 == 0) &&
                        Lc.kt == 1 && Lc.kh == 1 && Lc.kw == 1 && Lc.sh == 1 && Lc.sw == 1 && Lc.st == 1;
    // blocks of s2: `b` and `c` (+residual / +projection shortcut / +the next stage's temporal max-pool) as ONE kernel
    // (conv_bc_fused.cu)
    if (e->is_bf16 && e->conv_impl == 0 && !blk.spatial_pool) {
      rc = try_fused_bc(e, e->convs[blk.b], Lc, ya, da, B, shortcut, yout, s, fsc.L ? &fsc : nullptr, fuse_t ? 1 : 0);
      if (rc < 0) return rc;
      if (rc == 1) {
        d = conv_out(Lc, db);
        if (fuse_t) { d.T /= 2; e->pooled_already = true; }
        x = yout;
        if (is_stage_end(e, bi)) {
          rc = keep_stage(e, stage_no, x, d, clip0, B, s);
          if (rc) return rc;
          ++stage_no;
        }
        continue;
      }
    }
    rc = dense_conv(e, blk.b, ya, da, B, nullptr, yb, true, s);
    if (rc) return rc;
    if (blk.spatial_pool) {
      // FTCN-TT: MaxPool3d((1,2,2)) behind b_bn and branch1_bn (ReLU and max commute, so the conv epilogue's
      // ReLU stays where it is).  `ya` is dead after conv b and `yb` after its pooling, so they take the pooled maps.
      OpTrace tr(s);
      rc = maxpool_hw2_launch(yb, ya, (long long)B * db.T, db.H, db.W, db.C, e->is_bf16, s);
      if (rc) return rc;
      tr.done("maxpool 1x2x2 (b)", 0.0, (double)B * db.elems() * 1.25 * e->esz);
      db.H /= 2; db.W /= 2;
      if (shortcut == ysc) {
        const Dims dsc = conv_out(e->convs[blk.branch1], d);
        OpTrace tr2(s);
        rc = maxpool_hw2_launch(ysc, yb, (long long)B * dsc.T, dsc.H, dsc.W, dsc.C, e->is_bf16, s);
        if (rc) return rc;
        tr2.done("maxpool 1x2x2 (branch1)", 0.0, (double)B * dsc.elems() * 1.25 * e->esz);
        shortcut = yb;
      }
      void* t0 = ya; ya = yb; yb = t0;             // from here on `yb` names the pooled b output
    }
    {
      const long long sW = db.C, sH = (long long)db.W * db.C, sT = sH * db.H, sB = sT * db.T;
      rc = run_conv(e, Lc, yb, db, sB, sT, sH, sW, B, shortcut, yout, true, s, -1, 0, fuse_t ? 1 : 0,
                    fsc.L ? &fsc : nullptr);
    }
    if (rc) return rc;
    d = conv_out(e->convs[blk.c], db);
    if (fuse_t) { d.T /= 2; e->pooled_already = true; }
    x = yout;
    if (is_stage_end(e, bi)) {
      rc = keep_stage(e, stage_no, x, d, clip0, B, s);
      if (rc) return rc;
      ++stage_no;
    }
  }
  return AF_OK;
}

// Fills the engine's clip buffer for clips [f0, f0+fB) right before the trunk consumes them, so the
// packed chunk is still L2-resident when the stem reads it and host copies overlap earlier chunks.
typedef std::function<int(int f0, int fB, cudaStream_t s)> Feeder;

static ClipLayout clip_at(const af_engine* e, int f0) {
  ClipLayout c = e->clip;
  c.base = (char*)e->clip.base + (long long)f0 * e->clip.sB * (long long)e->esz;
  return c;
}

// (f0, fB) of every front chunk in the order run_trunk visits them
static std::vector<std::pair<int, int>> front_chunks(const af_engine* e, int B) {
  std::vector<std::pair<int, int>> v;
  for (int g0 = 0; g0 < B; g0 += e->cb_back) {
    const int gB = (B - g0) < e->cb_back ? (B - g0) : e->cb_back;
    for (int f0 = g0; f0 < g0 + gB; f0 += e->cb_front)
      v.push_back({f0, (g0 + gB - f0) < e->cb_front ? (g0 + gB - f0) : e->cb_front});
  }
  return v;
}

// The trunk on clips [0,B); `feed` packs each front chunk into e->clip just in time.
static int run_trunk(af_engine* e, int B, const Feeder& feed, float* logits, float* scores, float* features,
                     cudaStream_t s) {
  const ConvLayer& stem = e->convs[e->stem];
  const Dims din = {e->T, e->S, e->S, stem.cin_p};
  Dims dpre = conv_out(stem, din);
  if (e->stem_pool2) { dpre.H /= 2; dpre.W /= 2; }      // FTCN-TT: MaxPool3d((1,2,2)) behind the stem's BN
  const Dims dpool = {dpre.T, (dpre.H + 2 - 3) / 2 + 1, (dpre.W + 2 - 3) / 2 + 1, dpre.C};
  const int nblk = (int)e->blocks.size();
  const int split = e->split < 0 ? nblk : e->split;
  e->stage_batch = B;
  e->pooled_already = false;
  g_sm_limit = e->sm_limit;

  for (int g0 = 0; g0 < B; g0 += e->cb_back) {
    const int gB = (B - g0) < e->cb_back ? (B - g0) : e->cb_back;
    Dims dg = {0, 0, 0, 0};
    const void* xg = e->gbuf;
    for (int f0 = g0; f0 < g0 + gB; f0 += e->cb_front) {
      const int fB = (g0 + gB - f0) < e->cb_front ? (g0 + gB - f0) : e->cb_front;
      const char* xin = (const char*)e->clip.base + (long long)f0 * e->clip.sB * e->esz;
      int rc;
      {
        ProfRec prec(e, s);
        rc = feed(f0, fB, s);
        prec.done(3, 0.0, 0.0);
      }
      if (rc) return rc;
      bool fused_pool = false;
      bool stem_done = false;
      if (e->stem_pool2) {
        // FTCN-TT stem: conv k[5,1,1] + BN + max-pool 2x2 + ReLU + max-pool 3x3/2 in one kernel
        OpTrace tr(s);
        if (e->ftcn_w2 && e->conv_impl == 0) {         // tensor-core form (bf16 engine)
          AFB_CUDA(cudaMemsetAsync(e->fbuf[1], 0, (size_t)fB * dpool.elems() * e->esz, s));
          ProfRec prec(e, s);
          const char* phys = (const char*)e->clip_raw + (long long)f0 * e->clip.sB * (long long)e->esz;
          rc = ftcn_stem_umma_launch(phys, fB, e->T, e->S, e->ftcn_w2, stem.bias_h.data(), e->fbuf[1], s);
          prec.done(2, 2.0 * (double)fB * e->T * e->S * e->S * 64 * 15,
                    (double)fB * ((double)(e->T + 4) * (e->S + 6) * (e->S + 8) * 4 + dpool.elems()) * 2.0);
        } else {
          rc = ftcn_stem_launch(e->clip, f0, fB, stem.w_simt, stem.bias, e->fbuf[1], s);
        }
        if (rc) return rc;
        tr.done("ftcn stem k5x1x1 +pool2 +pool3/2", 2.0 * (double)fB * e->T * e->S * e->S * 64 * 15,
                (double)fB * ((double)e->T * e->S * e->S * 4 + dpool.elems()) * e->esz);
        stem_done = true;
        fused_pool = true;
      } else if (e->stem_direct == 1 && e->conv_impl == 0 && (dpre.H % 2 == 0) && (dpre.W % 2 == 0)) {
        // stem conv + BN + ReLU + max-pool in ONE kernel, reading the padded clip directly
        OpTrace tr(s);
        AFB_CUDA(cudaMemsetAsync(e->fbuf[1], 0, (size_t)fB * dpool.elems() * e->esz, s));
        ProfRec prec(e, s);
        const char* phys = (const char*)e->clip_raw + (long long)f0 * e->clip.sB * (long long)e->esz;
        rc = conv_stem_direct_launch(phys, fB, e->T, e->S, e->stem_w35, e->stem_u.bias_h.data(), e->fbuf[1], 1, s);
        if (rc == AF_ERR_CUDA && strstr(af_last_error(), "cuTensorMapEncodeTiled")) {
          e->stem_direct = -1;                    // driver refused the overlapping-window map: use the unfolded path
        } else {
          if (rc) return rc;
          stem_done = true;
          fused_pool = true;
          // algorithmic stem FLOPs (K = 3*5*7*7 = 735), not the 1120 issued
          prec.done(0, 2.0 * (double)fB * dpre.elems() * 735.0,
                    (double)fB * ((double)(e->T + 4) * (e->S + 6) * (e->S + 8) * 4 + dpool.elems()) * 2.0);
          tr.done("stem direct k5x7x7 s2 +pool (urows)", 2.0 * (double)fB * dpre.elems() * 1120.0,
                  (double)fB * ((double)(e->T + 4) * (e->S + 6) * (e->S + 8) * 4 + dpool.elems()) * 2.0);
        }
      }
      if (stem_done) {
      } else if (e->has_stem_u && e->conv_impl != 1) {
        OpTrace tr(s);
        rc = stem_unfold_launch(e->clip, f0, fB, e->fbuf[2], s);
        if (rc) return rc;
        const Dims du = {e->T, e->S / 2 + 3, e->S / 2, 64};
        tr.done("stem unfold", 0.0, (double)fB * (du.elems() + (double)e->T * e->S * e->S * 4) * 2.0);
        const long long sW = 64, sH = (long long)du.W * 64, sT = sH * du.H, sB = sT * du.T;
        static const bool no_fuse = getenv("AFB200_NO_FUSED_POOL") != nullptr;
        fused_pool = e->conv_impl == 0 && !no_fuse && (dpre.H % 2 == 0) && (dpre.W % 2 == 0);
        if (fused_pool) {
          AFB_CUDA(cudaMemsetAsync(e->fbuf[1], 0, (size_t)fB * dpool.elems() * e->esz, s));
          rc = run_conv(e, e->stem_u, e->fbuf[2], du, sB, sT, sH, sW, fB, nullptr, e->fbuf[1], true, s, -1, 1);
        } else {
          rc = run_conv(e, e->stem_u, e->fbuf[2], du, sB, sT, sH, sW, fB, nullptr, e->fbuf[0], true, s);
        }
      } else if (e->stem_w35_tf32 && e->conv_impl != 1) {
        // tf32 precision: the stem on the tensor cores straight from the padded fp32 clip (conv_tf32.cu, stem form)
        OpTrace tr(s);
        ProfRec prec(e, s);
        const char* phys = (const char*)e->clip_raw + (long long)f0 * e->clip.sB * (long long)e->esz;
        rc = conv_tf32_stem_launch(phys, fB, e->T, e->S, e->stem_w35_tf32, stem.bias, e->fbuf[0], s);
        prec.done(0, 2.0 * (double)fB * dpre.elems() * 735.0,
                  (double)fB * ((double)(e->T + 4) * (e->S + 6) * (e->S + 8) * 4 + dpre.elems()) * 4.0);
        tr.done("stem tf32 k5x7x7 s2", 2.0 * (double)fB * dpre.elems() * 1120.0,
                (double)fB * ((double)(e->T + 4) * (e->S + 6) * (e->S + 8) * 4 + dpre.elems()) * 4.0);
      } else {
        rc = run_conv(e, stem, xin, din, e->clip.sB, e->clip.sT, e->clip.sH, e->clip.sW, fB, nullptr, e->fbuf[0], true, s);
      }
      if (rc) return rc;
      if (!fused_pool) {
        OpTrace tr(s);
        rc = maxpool_spatial_launch(e->fbuf[0], e->fbuf[1], fB, dpre.T, dpre.H, dpre.W, dpre.C, e->is_bf16, s);
        if (rc) return rc;
        tr.done("stem maxpool 1x3x3", 0.0, (double)fB * (dpre.elems() + dpool.elems()) * e->esz);
      }
      rc = keep_stage(e, 0, e->fbuf[1], dpool, f0, fB, s);
      if (rc) return rc;
      const void* x = e->fbuf[1];
      Dims d = dpool;
      int stage_no = 1;
      rc = run_blocks(e, 0, split, x, d, fB, e->fbuf, f0, stage_no, s);
      if (rc) return rc;
      // hand the chunk to the back part: through the group buffer when several front chunks feed one back
      // group, directly (no copy) when the chunk IS the group
      dg = d;
      if (fB == gB) {
        xg = x;
      } else {
        AFB_CUDA(cudaMemcpyAsync((char*)e->gbuf + (long long)(f0 - g0) * d.elems() * e->esz, x,
                                 (size_t)fB * d.elems() * e->esz, cudaMemcpyDeviceToDevice, s));
      }
    }
    const void* x = xg;
    Dims d = dg;
    int stage_no = 1;
    for (int bi = 0; bi < split; ++bi) stage_no += is_stage_end(e, bi) ? 1 : 0;
    int rc = run_blocks(e, split, nblk, x, d, gB, e->bbuf, g0, stage_no, s);
    if (rc) return rc;
    OpTrace trh(s);
    if (e->has_tt) {
      // TransformerHead: AvgPool3d((1,H,W)) per frame -> tokens [gB, T', C] -> TimeTransformer
      if (d.T != e->tt.tokens || d.C != e->tt.dim) { set_error("tt head: trunk gives %d x %d, head takes %d x %d", d.T, d.C, e->tt.tokens, e->tt.dim); return AF_ERR_INVALID; }
      rc = spatial_mean_launch(x, gB * d.T, d.H * d.W, d.C, e->is_bf16, e->tok_ws, s);
      if (!rc)
        rc = tt_head_launch(e->tt, e->tok_ws, gB, e->tt_ws, features ? features + (long long)g0 * e->feat_dim : nullptr,
                            logits ? logits + g0 : nullptr, scores ? scores + g0 : nullptr, s);
      if (rc) return rc;
      trh.done("head frame means + transformer", 0.0, (double)gB * d.elems() * e->esz);
    } else {
      rc = head_launch(x, gB, d.T * d.H * d.W, d.C, e->is_bf16, e->fc_w, e->fc_b,
                       e->feat_ws + (long long)g0 * e->feat_slices * e->feat_dim,
                       features ? features + (long long)g0 * e->feat_dim : nullptr, logits ? logits + g0 : nullptr,
                       scores ? scores + g0 : nullptr, s);
      if (rc) return rc;
      trh.done("head avgpool+fc", 0.0, (double)gB * d.elems() * e->esz);
    }
    if (e->frame_feat_out && e->has_tt) {         // the transformer's tokens ARE the per-frame means
      AFB_CUDA(cudaMemcpyAsync(e->frame_feat_out + (long long)g0 * d.T * e->feat_dim, e->tok_ws,
                               (size_t)gB * d.T * d.C * sizeof(float), cudaMemcpyDeviceToDevice, s));
    } else if (e->frame_feat_out) {     // per-frame spatial means [gB*T', C]: the same pooling kernel over H*W positions
      rc = spatial_mean_launch(x, gB * d.T, d.H * d.W, d.C, e->is_bf16, e->frame_feat_out + (long long)g0 * d.T * e->feat_dim, s);
      if (rc) return rc;
    }
  }
  return AF_OK;
}

static int plan_workspace(af_engine* e) {
  // walk the network once for one clip to size the scratch buffers
  const ConvLayer& stem = e->convs[e->stem];
  Dims d = conv_out(stem, Dims{e->T, e->S, e->S, stem.cin_p});
  if (e->stem_pool2) { d.H /= 2; d.W /= 2; }
  long long fmax = e->stem_pool2 ? 0 : d.elems();      // the FTCN-TT stem kernel never materialises its conv output
  if (e->has_stem_u) {
    const long long u = (long long)e->T * (e->S / 2 + 3) * (e->S / 2) * 64;
    if (u > fmax) fmax = u;
  }
  d = Dims{d.T, (d.H + 2 - 3) / 2 + 1, (d.W + 2 - 3) / 2 + 1, d.C};
  long long bmax = 0;
  e->split = -1;
  for (size_t bi = 0; bi < e->blocks.size(); ++bi) {
    const af_block_desc& blk = e->blocks[bi];
    long long& mx = (e->split >= 0 || blk.temporal_pool_before) ? bmax : fmax;
    if (blk.temporal_pool_before && e->split < 0) {
      e->split = (int)bi;
      e->group_elems = d.elems();
    }
    if (blk.temporal_pool_before) { d.T /= 2; if (d.elems() > mx) mx = d.elems(); }
    Dims da = conv_out(e->convs[blk.a], d);
    Dims db = conv_out(e->convs[blk.b], da);
    long long sc_elems = 0;
    if (blk.branch1 >= 0) sc_elems = conv_out(e->convs[blk.branch1], d).elems();
    for (long long v : {d.elems(), da.elems(), db.elems(), sc_elems})
      if (v > mx) mx = v;
    if (blk.spatial_pool) { db.H /= 2; db.W /= 2; }
    Dims dc = conv_out(e->convs[blk.c], db);
    if (dc.elems() > mx) mx = dc.elems();
    if (dc.C != e->convs[blk.c].cout) { set_error("internal dims"); return AF_ERR_INVALID; }
    d = dc;
  }
  if (e->split < 0) e->group_elems = d.elems();
  if (d.C != e->feat_dim) { set_error("af_create: trunk ends with %d channels, head expects %d", d.C, e->feat_dim); return AF_ERR_INVALID; }
  e->front_max = fmax;
  e->back_max = bmax > 0 ? bmax : 1;
  return AF_OK;
}

static int alloc_workspace(af_engine* e) {
  for (int i = 0; i < 5; ++i) {
    AFB_CUDA(cudaMalloc(&e->fbuf[i], (size_t)e->cb_front * e->front_max * e->esz));
    AFB_CUDA(cudaMalloc(&e->bbuf[i], (size_t)e->cb_back * e->back_max * e->esz));
  }
  AFB_CUDA(cudaMalloc(&e->gbuf, (size_t)e->cb_back * e->group_elems * e->esz));
  return AF_OK;
}

static void free_workspace(af_engine* e) {
  for (int i = 0; i < 5; ++i) {
    if (e->fbuf[i]) cudaFree(e->fbuf[i]);
    if (e->bbuf[i]) cudaFree(e->bbuf[i]);
    e->fbuf[i] = e->bbuf[i] = nullptr;
  }
  if (e->gbuf) cudaFree(e->gbuf);
  e->gbuf = nullptr;
}

}  // namespace afb

extern "C" {

const char* af_last_error(void) { return afb::g_err; }
int32_t af_version(void) { return AFB200_VERSION; }
int64_t af_launch_count(af_handle h) { return h ? h->launches : 0; }

af_status af_destroy(af_handle h) {
  if (!h) return AF_OK;
  cudaSetDevice(h->device);
  for (auto& L : h->convs) free_layer(L);
  free_layer(h->stem_u);
  if (h->stem_w35) cudaFree(h->stem_w35);
  if (h->stem_w35_tf32) cudaFree(h->stem_w35_tf32);
  if (h->ftcn_w2) cudaFree(h->ftcn_w2);
  for (float* fb : h->fused_bias)
    if (fb) cudaFree(fb);
  free_workspace(h);
  if (h->fc_w) cudaFree(h->fc_w);
  for (float* p : h->tt_arrays) cudaFree(p);
  if (h->tt_ws) cudaFree(h->tt_ws);
  if (h->tok_ws) cudaFree(h->tok_ws);
  if (h->clip_raw) cudaFree(h->clip_raw);
  if (h->feat_ws) cudaFree(h->feat_ws);
  if (h->u8_stage) cudaFree(h->u8_stage);
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (auto ev : h->copy_events) cudaEventDestroy(ev);
  if (h->out_stage) cudaFree(h->out_stage);
  for (auto& hs : h->host_slot) {
    if (hs.u8_dev) cudaFree(hs.u8_dev);
    if (hs.out_dev) cudaFree(hs.out_dev);
    if (hs.out_pinned) cudaFreeHost(hs.out_pinned);
    if (hs.copied) cudaEventDestroy(hs.copied);
    if (hs.done) cudaEventDestroy(hs.done);
  }
  for (int i = 0; i < 5; ++i)
    if (h->stage_buf[i]) cudaFree(h->stage_buf[i]);
  for (auto& r : h->ev_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
  for (auto ev : h->ev_pool) cudaEventDestroy(ev);
  delete h;
  return AF_OK;
}

// Device copy of the FTCN-TT transformer head's parameters.
static int upload_tt_head(af_engine* e, const af_tt_head& h) {
  if (h.dim <= 0 || h.tokens <= 0 || h.heads <= 0 || h.dim_head <= 0 || h.mlp_dim <= 0 || h.depth <= 0 || !h.layers ||
      !h.cls_token || !h.pos_embedding || !h.norm_w || !h.norm_b || !h.fc_w) {
    set_error("af_create: incomplete transformer head");
    return AF_ERR_INVALID;
  }
  auto up = [&](const float* host, size_t n, const float** dev) -> int {
    if (!host) { set_error("af_create: transformer head has a null array"); return AF_ERR_INVALID; }
    float* d = nullptr;
    AFB_CUDA(cudaMalloc(&d, n * sizeof(float)));
    e->tt_arrays.push_back(d);
    AFB_CUDA(cudaMemcpy(d, host, n * sizeof(float), cudaMemcpyHostToDevice));
    *dev = d;
    return AF_OK;
  };
  TTHeadDev& t = e->tt;
  t.dim = h.dim; t.tokens = h.tokens; t.heads = h.heads; t.dim_head = h.dim_head; t.mlp_dim = h.mlp_dim; t.fc_b = h.fc_b;
  const size_t D = h.dim, inner = (size_t)h.heads * h.dim_head, mlp = h.mlp_dim;
  int rc = up(h.cls_token, D, &t.cls_token);
  if (!rc) rc = up(h.pos_embedding, (size_t)(h.tokens + 1) * D, &t.pos_embedding);
  if (!rc) rc = up(h.norm_w, D, &t.norm_w);
  if (!rc) rc = up(h.norm_b, D, &t.norm_b);
  if (!rc) rc = up(h.fc_w, D, &t.fc_w);
  for (int i = 0; i < h.depth && !rc; ++i) {
    const af_tt_layer& L = h.layers[i];
    TTLayerDev dl = {};
    rc = up(L.ln1_w, D, &dl.ln1_w);
    if (!rc) rc = up(L.ln1_b, D, &dl.ln1_b);
    if (!rc) rc = up(L.qkv_w, 3 * inner * D, &dl.qkv_w);
    if (!rc) rc = up(L.out_w, D * inner, &dl.out_w);
    if (!rc) rc = up(L.out_b, D, &dl.out_b);
    if (!rc) rc = up(L.ln2_w, D, &dl.ln2_w);
    if (!rc) rc = up(L.ln2_b, D, &dl.ln2_b);
    if (!rc) rc = up(L.fc1_w, mlp * D, &dl.fc1_w);
    if (!rc) rc = up(L.fc1_b, mlp, &dl.fc1_b);
    if (!rc) rc = up(L.fc2_w, D * mlp, &dl.fc2_w);
    if (!rc) rc = up(L.fc2_b, D, &dl.fc2_b);
    if (!rc) t.layers.push_back(dl);
  }
  if (rc) return rc;
  e->has_tt = true;
  return AF_OK;
}

static af_status create_impl(af_engine* e, const af_weights* w) {
  AFB_CUDA(cudaSetDevice(e->device));
  cudaDeviceProp prop;
  AFB_CUDA(cudaGetDeviceProperties(&prop, e->device));
  if (prop.major != 10) {
    set_error("af_create: device %d is sm_%d%d; this library is built for sm_100a only", e->device, prop.major, prop.minor);
    return AF_ERR_UNSUPPORTED;
  }
  if (e->is_bf16) {
    int rc = conv_umma_init();
    if (!rc) rc = conv_rows_init();
    if (!rc) rc = conv_tsweep_init();
    if (!rc) rc = conv_bc_fused_init();
    if (rc) return (af_status)rc;
  }
  if (e->tf32) {
    int rc = conv_tf32_init();
    if (rc) return (af_status)rc;
  }
  e->convs.resize(w->n_convs);
  for (int i = 0; i < w->n_convs; ++i) {
    int rc = upload_layer(w->convs[i], e->is_bf16, e->convs[i], e->tf32);
    if (rc) return (af_status)rc;
  }
  e->stem_pool2 = w->stem_pool2 != 0;
  if (e->stem_pool2) {
    const af_conv_desc& sc = w->convs[w->stem];
    if (sc.cin != 3 || sc.cout != 64 || sc.kt != 5 || sc.kh != 1 || sc.kw != 1 || sc.st != 1 || sc.sh != 1 || sc.sw != 1 ||
        sc.pt != 2 || sc.ph != 0 || sc.pw != 0 || e->S % 4 != 0) {
      set_error("af_create: stem_pool2 needs the FTCN-TT stem (3->64, k[5,1,1], s1, p[2,0,0]) and clip_s %% 4 == 0");
      return AF_ERR_INVALID;
    }
    if (e->is_bf16 && e->S % 32 == 0 && getenv("AFB200_NO_FTCN_UMMA") == nullptr) {
      int rc = upload_ftcn_stem_w2(sc, &e->ftcn_w2);
      if (rc) return (af_status)rc;
    }
  }
  for (int bi = 0; bi < w->n_blocks; ++bi)
    if (w->blocks[bi].spatial_pool && w->blocks[bi].branch1 < 0) {
      set_error("af_create: block %d has spatial_pool but no projection shortcut", bi);
      return AF_ERR_INVALID;
    }
  if (e->is_bf16 && (e->S % 2 == 0) && !e->stem_pool2) {
    int rc = upload_stem_unfolded(w->convs[w->stem], e->stem_u);
    if (rc == AF_OK) e->has_stem_u = true;
    else if (rc != AF_ERR_INVALID) return (af_status)rc;
    if (e->has_stem_u && e->stem_u.cout == 64 && getenv("AFB200_NO_STEM_DIRECT") == nullptr) {
      rc = upload_stem_direct(w->convs[w->stem], &e->stem_w35);
      if (rc) return (af_status)rc;
      e->stem_direct = 1;
    }
  }
  if (e->tf32 && !e->stem_pool2 && (e->S % 32 == 0) && getenv("AFB200_NO_TF32_STEM") == nullptr) {
    const af_conv_desc& d = w->convs[w->stem];
    if (d.kt == 5 && d.kh == 7 && d.kw == 7 && d.cin == 3 && d.cout == 64 && d.st == 1 && d.sh == 2 && d.sw == 2 && d.pt == 2 &&
        d.ph == 3 && d.pw == 3) {
      std::vector<float> w35((size_t)35 * 64 * 32, 0.f);
      for (int co = 0; co < 64; ++co)
        for (int c = 0; c < 3; ++c)
          for (int dt = 0; dt < 5; ++dt)
            for (int dy = 0; dy < 7; ++dy)
              for (int dx = 0; dx < 7; ++dx)
                w35[((size_t)(dt * 7 + dy) * 64 + co) * 32 + dx * 4 + c] =
                    tf32_round(d.weight[((((size_t)co * 3 + c) * 5 + dt) * 7 + dy) * 7 + dx]);
      AFB_CUDA(cudaMalloc(&e->stem_w35_tf32, w35.size() * sizeof(float)));
      AFB_CUDA(cudaMemcpy(e->stem_w35_tf32, w35.data(), w35.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
  }
  e->blocks.assign(w->blocks, w->blocks + w->n_blocks);
  e->fused_bias.assign(w->n_blocks, nullptr);
  e->fused_bias_h.assign(w->n_blocks, std::vector<float>());
  for (int bi = 0; bi < w->n_blocks && e->is_bf16; ++bi) {
    const af_block_desc& blk = w->blocks[bi];
    if (blk.branch1 < 0 || w->convs[blk.branch1].cout != w->convs[blk.c].cout) continue;
    const int n = w->convs[blk.c].cout;
    std::vector<float> sum(n);
    for (int i = 0; i < n; ++i) sum[i] = w->convs[blk.c].bias[i] + w->convs[blk.branch1].bias[i];
    AFB_CUDA(cudaMalloc(&e->fused_bias[bi], n * sizeof(float)));
    AFB_CUDA(cudaMemcpy(e->fused_bias[bi], sum.data(), n * sizeof(float), cudaMemcpyHostToDevice));
    e->fused_bias_h[bi] = sum;
  }
  if (w->tt_head) {
    int rc = upload_tt_head(e, *w->tt_head);
    if (rc) return (af_status)rc;
  } else {
    AFB_CUDA(cudaMalloc(&e->fc_w, w->feature_dim * sizeof(float)));
    AFB_CUDA(cudaMemcpy(e->fc_w, w->fc_weight, w->feature_dim * sizeof(float), cudaMemcpyHostToDevice));
  }
  // padded clip buffer: T+4 frames, S+6 rows, S+8 columns, 4 channels; pads stay zero forever.  The logical origin is at
  // padded (frame 2, row 3, column 3) - (frame 2, row 4, column 4) for the FTCN-TT variant, whose stem reads
  // 16-byte-aligned pixel pairs of even-aligned row pairs
  const long long Tp = e->T + 4, Hp = e->S + 6, Wp = e->S + 8;
  const size_t clip_bytes = (size_t)e->max_batch * Tp * Hp * Wp * 4 * e->esz;
  AFB_CUDA(cudaMalloc(&e->clip_raw, clip_bytes));
  AFB_CUDA(cudaMemset(e->clip_raw, 0, clip_bytes));
  e->clip.sW = 4; e->clip.sH = Wp * 4; e->clip.sT = Hp * Wp * 4; e->clip.sB = Tp * Hp * Wp * 4;
  e->clip.T = e->T; e->clip.S = e->S; e->clip.is_bf16 = e->is_bf16;
  e->clip.base = (char*)e->clip_raw + (2 * e->clip.sT + (e->stem_pool2 ? 4 : 3) * (e->clip.sH + e->clip.sW)) * (long long)e->esz;
  int rc = plan_workspace(e);
  if (rc) return (af_status)rc;
  if (e->cb_back > e->max_batch) e->cb_back = e->max_batch;
  if (e->cb_front > e->cb_back) e->cb_front = e->cb_back;
  rc = alloc_workspace(e);
  if (rc) return (af_status)rc;
  AFB_CUDA(cudaMalloc(&e->feat_ws, (size_t)e->max_batch * e->feat_slices * e->feat_dim * sizeof(float)));
  AFB_CUDA(cudaMalloc(&e->out_stage, (size_t)2 * e->max_batch * sizeof(float)));
  if (e->has_tt) {
    AFB_CUDA(cudaMalloc(&e->tt_ws, (size_t)tt_head_workspace_floats(e->tt, e->max_batch) * sizeof(float)));
    AFB_CUDA(cudaMalloc(&e->tok_ws, (size_t)e->max_batch * e->tt.tokens * e->tt.dim * sizeof(float)));
  }
  return AF_OK;
}

af_status af_create(af_handle* out, int32_t device, const af_weights* w, int32_t max_batch, int32_t precision) {
  if (!out || !w || max_batch <= 0 || w->n_convs <= 0 || w->n_blocks <= 0 || !w->convs || !w->blocks) {
    set_error("af_create: invalid arguments");
    return AF_ERR_INVALID;
  }
  if (precision != AF_PREC_FP32 && precision != AF_PREC_BF16 && precision != AF_PREC_TF32) {
    set_error("af_create: unknown precision %d", precision);
    return AF_ERR_INVALID;
  }
  for (int i = 0; i < w->n_convs; ++i) {
    const af_conv_desc& c = w->convs[i];
    if (!c.weight || !c.bias || c.cout % 64 != 0 || c.cin <= 0) {
      set_error("af_create: conv %d unsupported (cin=%d cout=%d)", i, c.cin, c.cout);
      return AF_ERR_INVALID;
    }
  }
  af_engine* e = new af_engine();
  e->device = device;
  e->is_bf16 = precision == AF_PREC_BF16;
  e->tf32 = precision == AF_PREC_TF32;
  e->esz = e->is_bf16 ? 2 : 4;
  e->T = w->clip_t; e->S = w->clip_s; e->max_batch = max_batch;
  e->stem = w->stem; e->fc_b = w->fc_bias; e->feat_dim = w->feature_dim;
  if (!w->tt_head && !w->fc_weight) { set_error("af_create: fc_weight is NULL and there is no transformer head"); delete e; return AF_ERR_INVALID; }
  af_status rc = create_impl(e, w);
  if (rc != AF_OK) { af_destroy(e); return rc; }
  *out = e;
  return AF_OK;
}

af_status af_set_option(af_handle h, const char* name, int64_t value) {
  if (!h || !name) { set_error("af_set_option: null"); return AF_ERR_INVALID; }
  std::string n(name);
  if (n == "keep_stages") { h->keep_stages = value != 0; return AF_OK; }
  if (n == "conv_impl") { h->conv_impl = (int)value; return AF_OK; }
  if (n == "sm_limit") { h->sm_limit = (int)value; return AF_OK; }
  if (n == "profile_events") { h->profile_events = value != 0; return AF_OK; }
  if (n == "dump_timeline") { conv_umma_timeline_dump((int)value); return AF_OK; }
  if (n == "ridge_x1000") { h->ridge_flop_per_byte = (double)value / 1000.0; return AF_OK; }
  if (n == "reset_stats") {
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (auto& r : h->ev_recs) { h->ev_pool.push_back(r.e0); h->ev_pool.push_back(r.e1); }
    h->ev_recs.clear();
    for (int k = 0; k < 4; ++k) h->stat_ms[k] = h->stat_flops[k] = h->stat_launches[k] = h->stat_kbytes[k] = 0;
    h->stat_bytes = 0;
    return AF_OK;
  }
  if (n == "chunk_front" || n == "chunk_back") {
    if (value <= 0) { set_error("af_set_option: %s must be positive", name); return AF_ERR_INVALID; }
    AFB_CUDA(cudaSetDevice(h->device));
    AFB_CUDA(cudaDeviceSynchronize());
    free_workspace(h);
    if (n == "chunk_front") h->cb_front = (int)value; else h->cb_back = (int)value;
    if (h->cb_back > h->max_batch) h->cb_back = h->max_batch;
    if (h->cb_front > h->cb_back) h->cb_front = h->cb_back;
    return (af_status)alloc_workspace(h);
  }
  set_error("af_set_option: unknown option '%s'", name);
  return AF_ERR_INVALID;
}

af_status af_set_global_option(const char* name, int64_t value) {
  if (!name) { set_error("af_set_global_option: null"); return AF_ERR_INVALID; }
  std::string n(name);
  if (n == "block_n") {
    if (value != 0 && value != 64 && value != 128 && value != 256) { set_error("af_set_global_option: block_n must be 0, 64, 128 or 256"); return AF_ERR_INVALID; }
    conv_umma_force_block_n((int)value);
    return AF_OK;
  }
  if (n == "pair") {
    if (value < -1 || value > 1) { set_error("af_set_global_option: pair must be -1, 0 or 1"); return AF_ERR_INVALID; }
    conv_umma_set_pair_mode((int)value);
    return AF_OK;
  }
  set_error("af_set_global_option: unknown option '%s'", name);
  return AF_ERR_INVALID;
}

af_status af_get_stat(af_handle h, const char* name, double* value) {
  if (!h || !name || !value) { set_error("af_get_stat: null"); return AF_ERR_INVALID; }
  AFB_CUDA(cudaSetDevice(h->device));
  AFB_CUDA(cudaDeviceSynchronize());
  for (auto& r : h->ev_recs) {          // fold finished event pairs into the running sums
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      h->stat_ms[r.kind] += ms; h->stat_flops[r.kind] += r.flops; h->stat_launches[r.kind] += 1; h->stat_kbytes[r.kind] += r.bytes;
      h->stat_bytes += r.bytes;
    }
    h->ev_pool.push_back(r.e0); h->ev_pool.push_back(r.e1);
  }
  h->ev_recs.clear();
  std::string n(name);
  if (n == "conv_umma_ms") *value = h->stat_ms[0] + h->stat_ms[2];
  else if (n == "conv_umma_flops") *value = h->stat_flops[0] + h->stat_flops[2];
  else if (n == "conv_umma_launches") *value = h->stat_launches[0] + h->stat_launches[2];
  else if (n == "conv_tensor_bound_ms") *value = h->stat_ms[0];
  else if (n == "conv_tensor_bound_flops") *value = h->stat_flops[0];
  else if (n == "conv_tensor_bound_launches") *value = h->stat_launches[0];
  else if (n == "conv_hbm_bound_ms") *value = h->stat_ms[2];
  else if (n == "conv_hbm_bound_bytes") *value = h->stat_kbytes[2];
  else if (n == "conv_hbm_bound_flops") *value = h->stat_flops[2];
  else if (n == "conv_hbm_bound_launches") *value = h->stat_launches[2];
  else if (n == "conv_simt_ms") *value = h->stat_ms[1];
  else if (n == "conv_simt_flops") *value = h->stat_flops[1];
  else if (n == "conv_simt_launches") *value = h->stat_launches[1];
  else if (n == "conv_bytes") *value = h->stat_bytes;
  else if (n == "feed_ms") *value = h->stat_ms[3];
  else if (n == "feed_launches") *value = h->stat_launches[3];
  else { set_error("af_get_stat: unknown stat '%s'", name); return AF_ERR_INVALID; }
  return AF_OK;
}

static af_status check_batch(af_handle h, int32_t batch, const char* who) {
  if (!h) { set_error("%s: null handle", who); return AF_ERR_INVALID; }
  if (batch <= 0 || batch > h->max_batch) {
    set_error("%s: batch %d outside 1..max_batch=%d", who, batch, h->max_batch);
    return AF_ERR_INVALID;
  }
  return AF_OK;
}

af_status af_forward(af_handle h, const void* clip_dev, int32_t dtype, const int64_t strides[5], int32_t batch,
                     float* logits_dev, float* features_dev, void* stream) {
  af_status rc = check_batch(h, batch, "af_forward");
  if (rc) return rc;
  if (!clip_dev || !strides || !logits_dev) { set_error("af_forward: null pointer"); return AF_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  AFB_CUDA(cudaSetDevice(h->device));
  const long long before = g_launches;
  long long q[5] = {strides[0], strides[1], strides[2], strides[3], strides[4]};
  const size_t elt = dtype == AF_F32 ? 4 : 2;
  Feeder feed = [&](int f0, int fB, cudaStream_t st) {
    return pack_clip_launch((const char*)clip_dev + (long long)f0 * q[0] * (long long)elt, dtype, q, fB, clip_at(h, f0), st);
  };
  int r = (dtype == AF_F32 || dtype == AF_BF16 || dtype == AF_F16) ? AF_OK : AF_ERR_INVALID;
  if (r) set_error("af_forward: unsupported clip dtype %d", dtype);
  if (!r) r = run_trunk(h, batch, feed, logits_dev, nullptr, features_dev, s);
  h->launches += g_launches - before;
  return (af_status)r;
}

af_status af_forward_frames(af_handle h, const void* clip_dev, int32_t dtype, const int64_t strides[5], int32_t batch,
                            float* logits_dev, float* frame_features_dev, void* stream) {
  if (!h || !frame_features_dev) { set_error("af_forward_frames: null pointer"); return AF_ERR_INVALID; }
  h->frame_feat_out = frame_features_dev;
  af_status rc = af_forward(h, clip_dev, dtype, strides, batch, logits_dev, nullptr, stream);
  h->frame_feat_out = nullptr;
  return rc;
}

af_status af_infer_u8(af_handle h, const uint8_t* clips_dev, int32_t batch, const float mean255[3],
                      const float std255[3], float* logits_dev, float* scores_dev, float* features_dev,
                      void* stream) {
  af_status rc = check_batch(h, batch, "af_infer_u8");
  if (rc) return rc;
  if (!clips_dev || !mean255 || !std255) { set_error("af_infer_u8: null pointer"); return AF_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  AFB_CUDA(cudaSetDevice(h->device));
  const long long before = g_launches;
  const size_t clip_bytes = (size_t)h->T * h->S * h->S * 3;
  Feeder feed = [&](int f0, int fB, cudaStream_t st) {
    return pack_u8_launch(clips_dev + (size_t)f0 * clip_bytes, fB, mean255, std255, clip_at(h, f0), st);
  };
  int r = run_trunk(h, batch, feed, logits_dev, scores_dev, features_dev, s);
  h->launches += g_launches - before;
  return (af_status)r;
}

af_status af_infer_u8_host(af_handle h, const uint8_t* clips_host, int32_t batch, const float mean255[3],
                           const float std255[3], float* logits_host, float* scores_host, void* stream) {
  af_status rc = check_batch(h, batch, "af_infer_u8_host");
  if (rc) return rc;
  if (!clips_host || !mean255 || !std255) { set_error("af_infer_u8_host: null pointer"); return AF_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  AFB_CUDA(cudaSetDevice(h->device));
  const size_t clip_bytes = (size_t)h->T * h->S * h->S * 3;
  if (!h->u8_stage) AFB_CUDA(cudaMalloc(&h->u8_stage, (size_t)h->max_batch * clip_bytes));
  if (!h->copy_stream) AFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  // H2D chunk by chunk on a side stream; the trunk waits per chunk, so copies hide behind compute.
  // Host-fed calls use front chunks of at most 8 clips so that there is something to overlap with.
  struct ChunkGuard {
    af_engine* e; int saved;
    ChunkGuard(af_engine* e_) : e(e_), saved(e_->cb_front) { if (e->cb_front > 8) e->cb_front = 8; }
    ~ChunkGuard() { e->cb_front = saved; }
  } chunk_guard(h);
  const auto chunks = front_chunks(h, batch);
  while (h->copy_events.size() < chunks.size() + 1) {
    cudaEvent_t ev;
    AFB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    h->copy_events.push_back(ev);
  }
  AFB_CUDA(cudaEventRecord(h->copy_events[chunks.size()], s));          // order after earlier work on s
  AFB_CUDA(cudaStreamWaitEvent(h->copy_stream, h->copy_events[chunks.size()], 0));
  for (size_t i = 0; i < chunks.size(); ++i) {
    AFB_CUDA(cudaMemcpyAsync(h->u8_stage + (size_t)chunks[i].first * clip_bytes, clips_host + (size_t)chunks[i].first * clip_bytes,
                             (size_t)chunks[i].second * clip_bytes, cudaMemcpyHostToDevice, h->copy_stream));
    AFB_CUDA(cudaEventRecord(h->copy_events[i], h->copy_stream));
  }
  {
    const long long before = g_launches;
    size_t next = 0;
    Feeder feed = [&](int f0, int fB, cudaStream_t st) {
      if (next >= chunks.size() || chunks[next].first != f0) { set_error("internal: chunk order"); return (int)AF_ERR_INVALID; }
      if (cudaStreamWaitEvent(st, h->copy_events[next], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); return (int)AF_ERR_CUDA; }
      ++next;
      return pack_u8_launch(h->u8_stage + (size_t)f0 * clip_bytes, fB, mean255, std255, clip_at(h, f0), st);
    };
    int r = run_trunk(h, batch, feed, h->out_stage, h->out_stage + h->max_batch, nullptr, s);
    h->launches += g_launches - before;
    if (r) return (af_status)r;
  }
  if (logits_host)
    AFB_CUDA(cudaMemcpyAsync(logits_host, h->out_stage, batch * sizeof(float), cudaMemcpyDeviceToHost, s));
  if (scores_host)
    AFB_CUDA(cudaMemcpyAsync(scores_host, h->out_stage + h->max_batch, batch * sizeof(float), cudaMemcpyDeviceToHost, s));
  AFB_CUDA(cudaStreamSynchronize(s));
  return AF_OK;
}

af_status af_submit_u8_host(af_handle h, const uint8_t* clips_host, int32_t batch, const float mean255[3],
                            const float std255[3], void* stream, int32_t* ticket) {
  af_status rc = check_batch(h, batch, "af_submit_u8_host");
  if (rc) return rc;
  if (!clips_host || !mean255 || !std255 || !ticket) { set_error("af_submit_u8_host: null pointer"); return AF_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  AFB_CUDA(cudaSetDevice(h->device));
  const int slot = h->next_slot;
  af_engine::HostSlot& hs = h->host_slot[slot];
  if (hs.batch > 0) {
    set_error("af_submit_u8_host: both slots are outstanding; call af_wait on ticket %d first", slot);
    return AF_ERR_INVALID;
  }
  const size_t clip_bytes = (size_t)h->T * h->S * h->S * 3;
  if (!hs.u8_dev) {
    AFB_CUDA(cudaMalloc(&hs.u8_dev, (size_t)h->max_batch * clip_bytes));
    AFB_CUDA(cudaMalloc(&hs.out_dev, (size_t)2 * h->max_batch * sizeof(float)));
    AFB_CUDA(cudaMallocHost(&hs.out_pinned, (size_t)2 * h->max_batch * sizeof(float)));
    AFB_CUDA(cudaEventCreateWithFlags(&hs.copied, cudaEventDisableTiming));
    AFB_CUDA(cudaEventCreateWithFlags(&hs.done, cudaEventDisableTiming));
  }
  if (!h->copy_stream) AFB_CUDA(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
  // The upload runs on the copy stream, so it overlaps whatever the compute stream is still doing for the previous
  // submission; the slot's previous user finished before its af_wait returned, so the staging buffer is free.
  AFB_CUDA(cudaMemcpyAsync(hs.u8_dev, clips_host, (size_t)batch * clip_bytes, cudaMemcpyHostToDevice, h->copy_stream));
  AFB_CUDA(cudaEventRecord(hs.copied, h->copy_stream));
  AFB_CUDA(cudaStreamWaitEvent(s, hs.copied, 0));
  {
    const long long before = g_launches;
    Feeder feed = [&](int f0, int fB, cudaStream_t st) {
      return pack_u8_launch(hs.u8_dev + (size_t)f0 * clip_bytes, fB, mean255, std255, clip_at(h, f0), st);
    };
    int r = run_trunk(h, batch, feed, hs.out_dev, hs.out_dev + h->max_batch, nullptr, s);
    h->launches += g_launches - before;
    if (r) return (af_status)r;
  }
  AFB_CUDA(cudaMemcpyAsync(hs.out_pinned, hs.out_dev, (size_t)2 * h->max_batch * sizeof(float), cudaMemcpyDeviceToHost, s));
  AFB_CUDA(cudaEventRecord(hs.done, s));
  hs.batch = batch;
  h->next_slot = slot ^ 1;
  *ticket = slot;
  return AF_OK;
}

af_status af_wait(af_handle h, int32_t ticket, float* logits_host, float* scores_host) {
  if (!h || ticket < 0 || ticket > 1) { set_error("af_wait: invalid handle or ticket"); return AF_ERR_INVALID; }
  af_engine::HostSlot& hs = h->host_slot[ticket];
  if (hs.batch <= 0) { set_error("af_wait: ticket %d is not outstanding", ticket); return AF_ERR_INVALID; }
  AFB_CUDA(cudaSetDevice(h->device));
  const int batch = hs.batch;
  hs.batch = 0;                                   // the slot is free again even if the wait reports an error
  AFB_CUDA(cudaEventSynchronize(hs.done));
  if (logits_host) memcpy(logits_host, hs.out_pinned, batch * sizeof(float));
  if (scores_host) memcpy(scores_host, hs.out_pinned + h->max_batch, batch * sizeof(float));
  return AF_OK;
}

af_status af_crop_u8(const af_frame_desc* frames_dev, const af_clip_geom* geom_dev, int32_t batch,
                     int32_t frames_per_clip, int32_t size, int32_t bgr, uint8_t* out_dev, void* stream) {
  if (!frames_dev || !geom_dev || !out_dev || batch <= 0 || frames_per_clip <= 0 || size <= 0) {
    set_error("af_crop_u8: invalid arguments");
    return AF_ERR_INVALID;
  }
  static_assert(sizeof(af_frame_desc) == sizeof(FrameDesc), "frame desc layout");
  static_assert(sizeof(af_clip_geom) == sizeof(ClipGeom), "clip geom layout");
  return (af_status)crop_launch((const FrameDesc*)frames_dev, (const ClipGeom*)geom_dev, batch, frames_per_clip, size,
                                bgr, out_dev, nullptr, nullptr, nullptr, (cudaStream_t)stream);
}

af_status af_ring_put_rows(uint8_t* ring_dev, int64_t slot_stride, int64_t pitch, int32_t n, const int32_t* slots,
                           const uint8_t* const* frames_host, const int32_t* row0, const int32_t* row1, void* stream) {
  if (!ring_dev || n < 0 || (n > 0 && (!slots || !frames_host || !row0 || !row1)) || pitch <= 0 || slot_stride < pitch) {
    set_error("af_ring_put_rows: invalid arguments");
    return AF_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < n; ++i) {
    if (row1[i] <= row0[i]) continue;
    if (row0[i] < 0 || (int64_t)row1[i] * pitch > slot_stride || slots[i] < 0 || !frames_host[i]) {
      set_error("af_ring_put_rows: item %d out of range (slot %d, rows %d..%d)", i, slots[i], row0[i], row1[i]);
      return AF_ERR_INVALID;
    }
    const int64_t off = (int64_t)row0[i] * pitch;
    AFB_CUDA(cudaMemcpyAsync(ring_dev + (int64_t)slots[i] * slot_stride + off, frames_host[i] + off,
                             (size_t)((int64_t)(row1[i] - row0[i]) * pitch), cudaMemcpyHostToDevice, s));
  }
  return AF_OK;
}

af_status af_ring_put_boxes(uint8_t* ring_dev, int64_t slot_stride, int64_t pitch, int32_t n, const int32_t* slots,
                            const uint8_t* const* frames_host, const int32_t* boxes_xyxy, void* stream) {
  if (!ring_dev || n < 0 || (n > 0 && (!slots || !frames_host || !boxes_xyxy)) || pitch <= 0 || slot_stride < pitch) {
    set_error("af_ring_put_boxes: invalid arguments");
    return AF_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  for (int i = 0; i < n; ++i) {
    const int32_t* b = boxes_xyxy + 4 * (size_t)i;
    if (b[2] <= b[0] || b[3] <= b[1]) continue;
    // byte columns of the box, widened to 16-byte multiples (K1's interior path fetches aligned 16-byte words)
    int64_t c0 = (int64_t)b[0] * 3 / 16 * 16, c1 = ((int64_t)b[2] * 3 + 15) / 16 * 16;
    if (c1 > pitch) c1 = pitch;
    if (b[0] < 0 || b[1] < 0 || c0 >= c1 || (int64_t)b[3] * pitch > slot_stride || slots[i] < 0 || !frames_host[i]) {
      set_error("af_ring_put_boxes: item %d out of range (slot %d, box %d,%d..%d,%d)", i, slots[i], b[0], b[1], b[2], b[3]);
      return AF_ERR_INVALID;
    }
    const int64_t off = (int64_t)b[1] * pitch + c0;
    AFB_CUDA(cudaMemcpy2DAsync(ring_dev + (int64_t)slots[i] * slot_stride + off, (size_t)pitch, frames_host[i] + off, (size_t)pitch,
                               (size_t)(c1 - c0), (size_t)(b[3] - b[1]), cudaMemcpyHostToDevice, s));
  }
  return AF_OK;
}

af_status af_crop_pack(const af_frame_desc* frames_dev, const af_clip_geom* geom_dev, int32_t batch,
                       int32_t frames_per_clip, int32_t size, int32_t bgr, const float mean255[3], const float std255[3],
                       void* clip_out_dev, int32_t out_dtype, const int64_t out_strides[5], void* stream) {
  if (!frames_dev || !geom_dev || !clip_out_dev || !mean255 || !std255 || !out_strides || batch <= 0 ||
      frames_per_clip <= 0 || size <= 0) {
    set_error("af_crop_pack: invalid arguments");
    return AF_ERR_INVALID;
  }
  if (out_dtype != AF_F32 && out_dtype != AF_BF16) { set_error("af_crop_pack: clip dtype must be AF_F32 or AF_BF16"); return AF_ERR_INVALID; }
  if (out_strides[1] == 0) { set_error("af_crop_pack: channel stride must be non-zero"); return AF_ERR_INVALID; }
  ClipLayout dst;
  dst.base = clip_out_dev;
  dst.sB = out_strides[0]; dst.sC = out_strides[1]; dst.sT = out_strides[2]; dst.sH = out_strides[3]; dst.sW = out_strides[4];
  dst.T = frames_per_clip; dst.S = size; dst.is_bf16 = out_dtype == AF_BF16;
  return (af_status)crop_launch((const FrameDesc*)frames_dev, (const ClipGeom*)geom_dev, batch, frames_per_clip, size, bgr,
                                nullptr, &dst, mean255, std255, (cudaStream_t)stream);
}

af_status af_crop_infer(af_handle h, const af_frame_desc* frames_dev, const af_clip_geom* geom_dev, int32_t batch,
                        int32_t bgr, const float mean255[3], const float std255[3], float* logits_dev,
                        float* scores_dev, float* features_dev, void* stream) {
  af_status rc = check_batch(h, batch, "af_crop_infer");
  if (rc) return rc;
  if (!frames_dev || !geom_dev || !mean255 || !std255) { set_error("af_crop_infer: null pointer"); return AF_ERR_INVALID; }
  cudaStream_t s = (cudaStream_t)stream;
  AFB_CUDA(cudaSetDevice(h->device));
  const long long before = g_launches;
  Feeder feed = [&](int f0, int fB, cudaStream_t st) {
    const ClipLayout dst = clip_at(h, f0);
    return crop_launch((const FrameDesc*)frames_dev + (size_t)f0 * h->T, (const ClipGeom*)geom_dev + f0, fB, h->T, h->S, bgr,
                       nullptr, &dst, mean255, std255, st);
  };
  int r = run_trunk(h, batch, feed, logits_dev, scores_dev, features_dev, s);
  h->launches += g_launches - before;
  return (af_status)r;
}

af_status af_conv_ndhwc(const void* x_dev, const af_conv_desc* conv_host, const void* residual_dev, void* y_dev,
                        int32_t batch, int32_t t, int32_t hgt, int32_t wid, int32_t relu, int32_t precision,
                        int32_t impl, void* stream) {
  if (!x_dev || !conv_host || !y_dev || batch <= 0) { set_error("af_conv_ndhwc: invalid arguments"); return AF_ERR_INVALID; }
  if (conv_host->cin % 4 != 0 || conv_host->cout % 64 != 0) {
    set_error("af_conv_ndhwc: needs cin %% 4 == 0 and cout %% 64 == 0");
    return AF_ERR_INVALID;
  }
  const bool is_bf16 = precision == AF_PREC_BF16;
  if (is_bf16) { int rc = conv_umma_init(); if (!rc) rc = conv_rows_init(); if (!rc) rc = conv_tsweep_init(); if (rc) return (af_status)rc; }
  if (precision == AF_PREC_TF32) { int rc = conv_tf32_init(); if (rc) return (af_status)rc; }
  ConvLayer L;
  int rc = upload_layer(*conv_host, is_bf16, L, precision == AF_PREC_TF32);
  if (!rc) {
    Dims in = {t, hgt, wid, L.cin_p};
    const long long sW = in.C, sH = (long long)in.W * in.C, sT = sH * in.H, sB = sT * in.T;
    const int pool = impl == 4 ? 1 : 0;
    const int tpool = impl == 5 ? 1 : 0;
    if (pool) {
      const Dims o = conv_out(L, in);
      rc = cudaMemsetAsync(y_dev, 0, (size_t)batch * o.T * (o.H / 2) * (o.W / 2) * o.C * 2, (cudaStream_t)stream) == cudaSuccess ? 0 : AF_ERR_CUDA;
    }
    if (!rc)
      rc = run_conv(nullptr, L, x_dev, in, sB, sT, sH, sW, batch, residual_dev, y_dev, relu != 0, (cudaStream_t)stream,
                    pool ? 3 : (tpool ? 2 : impl), pool, tpool);
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
      set_error("af_conv_ndhwc: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = AF_ERR_CUDA;
    }
  }
  free_layer(L);
  return (af_status)rc;
}

af_status af_conv_shortcut_ndhwc(const void* x_dev, const af_conv_desc* conv_host, const void* x2_dev,
                                 const af_conv_desc* shortcut_host, void* y_dev, int32_t batch, int32_t t, int32_t hgt,
                                 int32_t wid, int32_t hgt2, int32_t wid2, int32_t relu, void* stream) {
  if (!x_dev || !conv_host || !x2_dev || !shortcut_host || !y_dev || batch <= 0) {
    set_error("af_conv_shortcut_ndhwc: invalid arguments");
    return AF_ERR_INVALID;
  }
  if (conv_host->cin % 64 != 0 || shortcut_host->cin % 64 != 0 || conv_host->cout % 64 != 0 ||
      shortcut_host->cout != conv_host->cout) {
    set_error("af_conv_shortcut_ndhwc: needs cin %% 64 == 0 on both convs and equal cout %% 64 == 0");
    return AF_ERR_INVALID;
  }
  int rc = conv_umma_init();
  if (rc) return (af_status)rc;
  ConvLayer L, L2;
  float* bias = nullptr;
  rc = upload_layer(*conv_host, true, L);
  if (!rc) rc = upload_layer(*shortcut_host, true, L2);
  if (!rc) {
    std::vector<float> sum(conv_host->cout);
    for (int i = 0; i < conv_host->cout; ++i) sum[i] = conv_host->bias[i] + shortcut_host->bias[i];
    if (cudaMalloc(&bias, sum.size() * sizeof(float)) != cudaSuccess ||
        cudaMemcpy(bias, sum.data(), sum.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
      set_error("af_conv_shortcut_ndhwc: bias upload failed");
      rc = AF_ERR_CUDA;
    }
  }
  if (!rc) {
    Dims in = {t, hgt, wid, L.cin_p};
    const long long sW = in.C, sH = (long long)in.W * in.C, sT = sH * in.H, sB = sT * in.T;
    FusedShortcut sc = {&L2, x2_dev, Dims{t, hgt2, wid2, L2.cin_p}, bias};
    rc = run_conv(nullptr, L, x_dev, in, sB, sT, sH, sW, batch, nullptr, y_dev, relu != 0, (cudaStream_t)stream, 2, 0, 0, &sc);
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
      set_error("af_conv_shortcut_ndhwc: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = AF_ERR_CUDA;
    }
  }
  if (bias) cudaFree(bias);
  free_layer(L);
  free_layer(L2);
  return (af_status)rc;
}

af_status af_stem_pool_ndhwc4(const void* clip_dev, const af_conv_desc* stem_host, void* y_dev, int32_t batch, int32_t t,
                              int32_t s_, int32_t per_frame_kernel, void* stream) {
  if (!clip_dev || !stem_host || !y_dev || batch <= 0 || t <= 0 || s_ <= 0) { set_error("af_stem_pool_ndhwc4: invalid arguments"); return AF_ERR_INVALID; }
  const af_conv_desc& d = *stem_host;
  if (d.cin != 3 || d.cout != 64 || d.kt != 5 || d.kh != 7 || d.kw != 7 || d.st != 1 || d.sh != 2 || d.sw != 2 || d.pt != 2 ||
      d.ph != 3 || d.pw != 3 || (s_ % 4) != 0) {
    set_error("af_stem_pool_ndhwc4: takes the reference stem (3->64, k[5,7,7], s[1,2,2], p[2,3,3]) and size %% 4 == 0");
    return AF_ERR_INVALID;
  }
  int rc = conv_rows_init();
  if (rc) return (af_status)rc;
  cudaStream_t st = (cudaStream_t)stream;
  bf16* w35 = nullptr;
  float* bias = nullptr;
  void* padded = nullptr;
  const long long Tp = t + 4, Hp = s_ + 6, Wp = s_ + 8;
  const size_t pad_bytes = (size_t)batch * Tp * Hp * Wp * 4 * 2;
  rc = upload_stem_direct(d, &w35);
  if (!rc && (cudaMalloc(&bias, 64 * sizeof(float)) != cudaSuccess ||
              cudaMemcpy(bias, d.bias, 64 * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
              cudaMalloc(&padded, pad_bytes) != cudaSuccess || cudaMemsetAsync(padded, 0, pad_bytes, st) != cudaSuccess)) {
    set_error("af_stem_pool_ndhwc4: device allocation failed");
    rc = AF_ERR_CUDA;
  }
  if (!rc) {
    ClipLayout cl;
    cl.sW = 4; cl.sH = Wp * 4; cl.sT = Hp * Wp * 4; cl.sB = Tp * Hp * Wp * 4; cl.T = t; cl.S = s_; cl.is_bf16 = true;
    cl.base = (char*)padded + (2 * cl.sT + 3 * cl.sH + 3 * cl.sW) * 2;
    const long long q[5] = {(long long)t * s_ * s_ * 4, 1, (long long)s_ * s_ * 4, (long long)s_ * 4, 4};   // [B,3,T,S,S] view of NDHWC4
    rc = pack_clip_launch(clip_dev, AF_BF16, q, batch, cl, st);
  }
  if (!rc && cudaMemsetAsync(y_dev, 0, (size_t)batch * t * (s_ / 4) * (s_ / 4) * 64 * 2, st) != cudaSuccess) rc = AF_ERR_CUDA;
  if (!rc) rc = conv_stem_direct_launch(padded, batch, t, s_, w35, d.bias, y_dev, 1, st, per_frame_kernel);
  if (!rc && cudaStreamSynchronize(st) != cudaSuccess) {
    set_error("af_stem_pool_ndhwc4: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
    rc = AF_ERR_CUDA;
  }
  if (w35) cudaFree(w35);
  if (bias) cudaFree(bias);
  if (padded) cudaFree(padded);
  return (af_status)rc;
}

static af_status bc_fused_entry(const void* x_dev, const af_conv_desc* conv_b_host, const af_conv_desc* conv_c_host,
                                const void* residual_dev, const void* x2_dev, const af_conv_desc* shortcut_host, void* y_dev,
                                int32_t batch, int32_t t, int32_t hgt, int32_t wid, int pool_t, void* stream) {
  if (!x_dev || !conv_b_host || !conv_c_host || !y_dev || batch <= 0 || (!residual_dev == !x2_dev) || (!x2_dev != !shortcut_host)) {
    set_error("af_conv_bc_fused_ndhwc: invalid arguments (give either a residual or a shortcut input + conv)");
    return AF_ERR_INVALID;
  }
  int rc = conv_umma_init();
  if (!rc) rc = conv_bc_fused_init();
  if (rc) return (af_status)rc;
  ConvLayer Lb, Lc, Ls;
  float* bias = nullptr;
  std::vector<float> bias_sum_h;
  rc = upload_layer(*conv_b_host, true, Lb);
  if (!rc) rc = upload_layer(*conv_c_host, true, Lc);
  if (!rc && shortcut_host) {
    rc = shortcut_host->cout == conv_c_host->cout ? upload_layer(*shortcut_host, true, Ls) : (int)AF_ERR_INVALID;
    if (!rc) {
      std::vector<float> sum(conv_c_host->cout);
      for (int i = 0; i < conv_c_host->cout; ++i) sum[i] = conv_c_host->bias[i] + shortcut_host->bias[i];
      bias_sum_h = sum;
      if (cudaMalloc(&bias, sum.size() * sizeof(float)) != cudaSuccess ||
          cudaMemcpy(bias, sum.data(), sum.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("af_conv_bc_fused_ndhwc: bias upload failed");
        rc = AF_ERR_CUDA;
      }
    }
  }
  if (!rc) {
    const Dims in = {t, hgt, wid, Lb.cin_p};
    const Dims dmid = conv_out(Lb, in);
    const ConvProblem pb = dense_problem(Lb, x_dev, in, batch, nullptr, nullptr, true);
    FusedShortcut sc = {&Ls, x2_dev, Dims{t, hgt, wid, Ls.cin_p}, bias, bias_sum_h.data()};
    const long long sW = dmid.C, sH = (long long)dmid.W * dmid.C, sT = sH * dmid.H, sB = sT * dmid.T;
    const ConvProblem pc = make_problem(Lc, nullptr, dmid, sB, sT, sH, sW, batch, residual_dev, y_dev, true, 0, pool_t,
                                        shortcut_host ? &sc : nullptr);
    if (!conv_bc_fused_supported(pb, pc)) {
      set_error("af_conv_bc_fused_ndhwc: takes b = 1x3x3 s1 p[0,1,1] 64->64, c = 1x1x1 64->256, width %% 8 == 0 (pooled form: "
                "residual only, even T)");
      rc = AF_ERR_INVALID;
    } else {
      rc = conv_bc_fused_launch(pb, pc, (cudaStream_t)stream);
    }
    if (!rc && cudaStreamSynchronize((cudaStream_t)stream) != cudaSuccess) {
      set_error("af_conv_bc_fused_ndhwc: kernel failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = AF_ERR_CUDA;
    }
  }
  if (bias) cudaFree(bias);
  free_layer(Lb);
  free_layer(Lc);
  free_layer(Ls);
  return (af_status)rc;
}

af_status af_conv_bc_fused_ndhwc(const void* x_dev, const af_conv_desc* conv_b_host, const af_conv_desc* conv_c_host,
                                 const void* residual_dev, const void* x2_dev, const af_conv_desc* shortcut_host, void* y_dev,
                                 int32_t batch, int32_t t, int32_t hgt, int32_t wid, void* stream) {
  return bc_fused_entry(x_dev, conv_b_host, conv_c_host, residual_dev, x2_dev, shortcut_host, y_dev, batch, t, hgt, wid, 0, stream);
}

af_status af_conv_bc_fused_tpool_ndhwc(const void* x_dev, const af_conv_desc* conv_b_host, const af_conv_desc* conv_c_host,
                                       const void* residual_dev, void* y_dev, int32_t batch, int32_t t, int32_t hgt,
                                       int32_t wid, void* stream) {
  return bc_fused_entry(x_dev, conv_b_host, conv_c_host, residual_dev, nullptr, nullptr, y_dev, batch, t, hgt, wid, 1, stream);
}

af_status af_get_stage(af_handle h, int32_t which, float* out_dev, int64_t capacity_elems, int32_t dims_out[5],
                       void* stream) {
  if (!h || which < 1 || which > 5 || !dims_out) { set_error("af_get_stage: invalid arguments"); return AF_ERR_INVALID; }
  if (!h->keep_stages || !h->stage_buf[which - 1]) {
    set_error("af_get_stage: stage %d not kept (set option keep_stages=1 before the forward)", which);
    return AF_ERR_INVALID;
  }
  const Dims d = h->stage_dims[which - 1];
  dims_out[0] = h->stage_batch; dims_out[1] = d.C; dims_out[2] = d.T; dims_out[3] = d.H; dims_out[4] = d.W;
  const long long n = (long long)h->stage_batch * d.elems();
  if (out_dev) {
    if (capacity_elems < n) { set_error("af_get_stage: buffer too small (%lld < %lld)", (long long)capacity_elems, n); return AF_ERR_INVALID; }
    AFB_CUDA(cudaMemcpyAsync(out_dev, h->stage_buf[which - 1], n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  }
  return AF_OK;
}

}  // extern "C"
