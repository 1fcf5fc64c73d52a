// Exported C ABI (include/smilecode_b200.h): argument validation + dispatch to the launchers.
#include <cstdarg>
#include <cstdio>

#include "../../include/smilecode_b200.h"
#include <cstdlib>
#include "common.cuh"

#include "kernels.h"

namespace smile {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: kernel launch failed: %s", what, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  return SMILE_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace smile

using namespace smile;

#define REQUIRE(cond, ...)          \
  do {                              \
    if (!(cond)) {                  \
      set_error(__VA_ARGS__);       \
      return SMILE_ERR_INVALID_ARG; \
    }                               \
  } while (0)

#define REQUIRE_PTR(p) REQUIRE((p) != nullptr && aligned16(p), "%s: pointer `" #p "` is NULL or not 16-byte aligned", __func__)
#define REQUIRE_VOL(B, D, H, W)                                                                              \
  REQUIRE((B) > 0 && (D) > 0 && (H) > 0 && (W) > 0 && (long long)(D) * (H) * (W) < (1LL << 31),              \
          "%s: bad volume B=%d D=%d H=%d W=%d (each > 0, D*H*W < 2^31)", __func__, (B), (D), (H), (W))

extern "C" {

int smile_version(void) { return 100; }
const char* smile_last_error(void) { return g_err; }

int smile_modet_attn_fwd(const float* q, const float* k, const float* rpb, float* out, int B, int D, int H, int W,
                         int heads, int head_dim, float scale, smile_stream_t stream) {
  REQUIRE_PTR(q);
  REQUIRE_PTR(k);
  REQUIRE_PTR(out);
  REQUIRE(rpb == nullptr || aligned16(rpb), "%s: rpb not 16-byte aligned", __func__);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(heads > 0 && head_dim > 0 && heads <= 256, "%s: heads=%d head_dim=%d out of range", __func__, heads, head_dim);
  return launch_modet_attn(q, k, rpb, out, B, D, H, W, heads, head_dim, scale, (cudaStream_t)stream);
}

int smile_warp3d_fwd(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W,
                     smile_stream_t stream) {
  REQUIRE_PTR(src);
  REQUIRE_PTR(flow);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0, "%s: C=%d", __func__, C);
  REQUIRE(src != out, "%s: out must not alias src", __func__);
  return launch_warp3d(src, flow, out, B, C, D, H, W, (cudaStream_t)stream);
}

int smile_upsample2x_fwd(const float* x, float* out, int B, int C, int D, int H, int W, float premul,
                         smile_stream_t stream) {
  REQUIRE_PTR(x);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, 2 * D, 2 * H, 2 * W);
  REQUIRE(C > 0, "%s: C=%d", __func__, C);
  return launch_upsample2x(x, out, B, C, D, H, W, premul, (cudaStream_t)stream);
}

int smile_flow_compose_fwd(const float* flow, const float* w, float* out, int B, int D, int H, int W, float postmul,
                           smile_stream_t stream) {
  REQUIRE_PTR(flow);
  REQUIRE_PTR(w);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(out != flow, "%s: out must not alias flow", __func__);
  return launch_compose(flow, w, out, B, D, H, W, postmul, (cudaStream_t)stream);
}

int smile_modet_fused_fwd(const float* q, const float* k, const float* rpb, const float* ln_gamma, const float* ln_beta,
                          const float* flow_in, const float* moving, float* flow_out, float* moved, int B, int D, int H,
                          int W, int head_dim, float scale, float postmul, int Cmov, smile_stream_t stream) {
  REQUIRE_PTR(q);
  REQUIRE((ln_gamma == nullptr) == (ln_beta == nullptr), "%s: ln_gamma and ln_beta must be given together", __func__);
  REQUIRE_PTR(k);
  REQUIRE_PTR(flow_in);
  REQUIRE_PTR(flow_out);
  REQUIRE(rpb == nullptr || aligned16(rpb), "%s: rpb not 16-byte aligned", __func__);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(head_dim > 0, "%s: head_dim=%d", __func__, head_dim);
  REQUIRE(flow_out != flow_in, "%s: flow_out must not alias flow_in", __func__);
  if (moved != nullptr) {
    REQUIRE_PTR(moving);
    REQUIRE(aligned16(moved) && Cmov > 0, "%s: moved misaligned or Cmov=%d", __func__, Cmov);
  }
  return launch_modet_fused(q, k, rpb, ln_gamma, ln_beta, flow_in, moving, flow_out, moved, B, D, H, W, head_dim, scale,
                            postmul, Cmov, (cudaStream_t)stream);
}

int smile_proj_ln_fwd(const float* feat, const float* weight, const float* bias, const float* gamma, const float* beta,
                      float* out, int B, int Cin, int C, long long N, float eps, smile_stream_t stream) {
  REQUIRE_PTR(feat);
  REQUIRE_PTR(weight);
  REQUIRE_PTR(bias);
  REQUIRE_PTR(gamma);
  REQUIRE_PTR(beta);
  REQUIRE_PTR(out);
  REQUIRE(B > 0 && Cin > 0 && C > 0 && N > 0 && N < (1LL << 31), "%s: bad sizes B=%d Cin=%d C=%d N=%lld", __func__, B, Cin,
          C, N);
  return launch_proj_ln(feat, weight, bias, gamma, beta, out, B, Cin, C, N, eps, (cudaStream_t)stream);
}

int smile_warp_proj_ln_fwd(const float* src, const float* flow, const float* weight, const float* bias,
                           const float* gamma, const float* beta, float* out, int B, int Cin, int C, int D, int H,
                           int W, float eps, smile_stream_t stream) {
  REQUIRE_PTR(src);
  REQUIRE_PTR(flow);
  REQUIRE_PTR(weight);
  REQUIRE_PTR(bias);
  REQUIRE_PTR(gamma);
  REQUIRE_PTR(beta);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(Cin > 0 && C > 0, "%s: Cin=%d C=%d", __func__, Cin, C);
  bool handled = false;
  int rc = launch_warp_proj_ln(src, flow, weight, bias, gamma, beta, out, B, Cin, C, D, H, W, eps, (cudaStream_t)stream,
                               &handled);
  if (handled) return rc;
  set_error("%s: no fused kernel for Cin=%d C=%d D=%d H=%d W=%d (use smile_warp3d_fwd + smile_proj_ln_fwd)", __func__, Cin,
            C, D, H, W);
  return SMILE_ERR_UNSUPPORTED;
}

int smile_conv3d_fwd(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                     double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                     smile_stream_t stream) {
  REQUIRE_PTR(in);
  REQUIRE_PTR(weight);
  REQUIRE_PTR(bias);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(Cin > 0 && Cout > 0 && Cin <= 4096 && Cout <= 4096, "%s: Cin=%d Cout=%d out of range", __func__, Cin, Cout);
  REQUIRE(in != out, "%s: out must not alias in", __func__);
  REQUIRE((long long)B <= 65535, "%s: B=%d exceeds grid.z", __func__, B);
  return launch_conv3d(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps,
                       (cudaStream_t)stream);
}

int smile_conv3d_bf16_fwd(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                          double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                          smile_stream_t stream) {
  REQUIRE_PTR(in);
  REQUIRE_PTR(weight);
  REQUIRE_PTR(bias);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(Cin > 0 && Cout > 0 && Cin <= 4096 && Cout <= 4096, "%s: Cin=%d Cout=%d out of range", __func__, Cin, Cout);
  REQUIRE(in != out, "%s: out must not alias in", __func__);
  REQUIRE((long long)B <= 65535, "%s: B=%d exceeds grid.z", __func__, B);
  bool handled = false;
  // few channels (the wide levels): depth-marching kernel; otherwise the plane-tiled kernel where it wins
  const char* march = getenv("SMILE_CONV_MARCH");
  if (march == nullptr || march[0] != '0') {
    const int rc = launch_conv3d_march_bf16(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps,
                                            (cudaStream_t)stream, &handled);
    if (handled) return rc;
  }
  const int rc = launch_conv3d_bf16(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps,
                                    (cudaStream_t)stream, &handled);
  if (handled) return rc;
  // shapes the bf16 tensor-core kernel does not take (one input channel, very wide rows) run at full precision
  return launch_conv3d(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps,
                       (cudaStream_t)stream);
}

long long smile_conv3d_tc_prep_floats(int Cin, int Cout) {
  if (Cin <= 0 || Cout <= 0 || Cin > 4096 || Cout > 4096) return 0;
  return conv3d_tc_prep_floats(Cin, Cout);
}

int smile_conv3d_tc_prep(const float* weight, float* wprep, int Cin, int Cout, smile_stream_t stream) {
  REQUIRE_PTR(weight);
  REQUIRE_PTR(wprep);
  REQUIRE(Cin > 0 && Cout > 0 && Cin <= 4096 && Cout <= 4096, "%s: Cin=%d Cout=%d out of range", __func__, Cin, Cout);
  return launch_conv3d_tc_prep(weight, wprep, Cin, Cout, (cudaStream_t)stream);
}

int smile_conv3d_prepped_fwd(const float* in, const float* weight, const float* wprep, const float* bias, float* out,
                             const double* in_stats, double* out_stats, int B, int Cin, int Cout, int D, int H, int W,
                             int act_out, float eps, smile_stream_t stream) {
  REQUIRE_PTR(in);
  REQUIRE_PTR(weight);
  REQUIRE_PTR(bias);
  REQUIRE_PTR(out);
  REQUIRE(wprep == nullptr || aligned16(wprep), "%s: wprep not 16-byte aligned", __func__);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(Cin > 0 && Cout > 0 && Cin <= 4096 && Cout <= 4096, "%s: Cin=%d Cout=%d out of range", __func__, Cin, Cout);
  REQUIRE(in != out, "%s: out must not alias in", __func__);
  REQUIRE((long long)B <= 65535, "%s: B=%d exceeds grid.z", __func__, B);
  return launch_conv3d(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps,
                       (cudaStream_t)stream, wprep);
}

int smile_instnorm_lrelu_pool_fwd(const float* raw, const double* stats, float* out, float* pooled, int B, int C, int D,
                                  int H, int W, float eps, smile_stream_t stream) {
  REQUIRE_PTR(raw);
  REQUIRE_PTR(stats);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0 && (long long)B * C <= 65535, "%s: B*C=%lld out of range", __func__, (long long)B * C);
  REQUIRE(pooled == nullptr || (D >= 2 && H >= 2 && W >= 2), "%s: pooling needs every dim >= 2", __func__);
  return launch_in_finalize(raw, stats, out, pooled, B, C, D, H, W, eps, (cudaStream_t)stream);
}

int smile_cwm_fuse_fwd(const float* fields, const float* logits, float* out, int B, int F, long long N,
                       smile_stream_t stream) {
  REQUIRE_PTR(fields);
  REQUIRE_PTR(logits);
  REQUIRE_PTR(out);
  REQUIRE(B > 0 && F > 0 && N > 0 && N < (1LL << 31), "%s: bad sizes B=%d F=%d N=%lld", __func__, B, F, N);
  return launch_cwm_fuse(fields, logits, out, B, F, N, (cudaStream_t)stream);
}

int smile_modet_qkrpb_fwd(const float* q, const float* kpad, const float* rpb, float* attn, int B, int heads, int H,
                          int W, int T, int head_dim, smile_stream_t stream) {
  REQUIRE_PTR(q);
  REQUIRE_PTR(kpad);
  REQUIRE_PTR(attn);
  REQUIRE(rpb == nullptr || aligned16(rpb), "%s: rpb not 16-byte aligned", __func__);
  REQUIRE(B > 0 && heads > 0 && H > 0 && W > 0 && T > 0 && head_dim > 0, "%s: bad sizes B=%d heads=%d H=%d W=%d T=%d hd=%d",
          __func__, B, heads, H, W, T, head_dim);
  return launch_qkrpb_fwd(q, kpad, rpb, attn, B, heads, H, W, T, head_dim, (cudaStream_t)stream);
}

int smile_modet_qkrpb_bwd(const float* d_attn, const float* q, const float* kpad, float* d_q, float* d_kpad,
                          float* d_rpb, int B, int heads, int H, int W, int T, int head_dim, smile_stream_t stream) {
  REQUIRE_PTR(d_attn);
  REQUIRE_PTR(q);
  REQUIRE_PTR(kpad);
  REQUIRE_PTR(d_q);
  REQUIRE_PTR(d_kpad);
  REQUIRE(d_rpb == nullptr || aligned16(d_rpb), "%s: d_rpb not 16-byte aligned", __func__);
  REQUIRE(B > 0 && heads > 0 && H > 0 && W > 0 && T > 0 && head_dim > 0, "%s: bad sizes B=%d heads=%d H=%d W=%d T=%d hd=%d",
          __func__, B, heads, H, W, T, head_dim);
  return launch_qkrpb_bwd(d_attn, q, kpad, d_q, d_kpad, d_rpb, B, heads, H, W, T, head_dim, (cudaStream_t)stream);
}

long long smile_ncc_vxm_work_bytes(int B, int D, int H, int W) {
  return 10LL * B * D * H * W * (long long)sizeof(float) + 16;
}

int smile_ncc_vxm_fwd(const float* y_true, const float* y_pred, float* out, void* work, int B, int D, int H, int W,
                      int win, smile_stream_t stream) {
  REQUIRE_PTR(y_true);
  REQUIRE_PTR(y_pred);
  REQUIRE_PTR(out);
  REQUIRE_PTR(work);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(win >= 1 && (win & 1) && win <= 63, "%s: win=%d must be odd and <= 63", __func__, win);
  return launch_ncc_vxm(y_true, y_pred, out, reinterpret_cast<float*>(work), B, D, H, W, win, (cudaStream_t)stream);
}

int smile_grad3d_l2_fwd(const float* flow, float* out, void* work, int B, int C, int D, int H, int W,
                        smile_stream_t stream) {
  REQUIRE_PTR(flow);
  REQUIRE_PTR(out);
  REQUIRE_PTR(work);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0 && D > 1 && H > 1 && W > 1, "%s: needs C > 0 and every spatial dim > 1", __func__);
  return launch_grad3d_l2(flow, out, reinterpret_cast<double*>(work), B, C, D, H, W, (cudaStream_t)stream);
}

int smile_warp3d_bwd(const float* g, const float* src, const float* flow, float* d_src, float* d_flow, int B, int C,
                     int D, int H, int W, smile_stream_t stream) {
  REQUIRE_PTR(g);
  REQUIRE_PTR(src);
  REQUIRE_PTR(flow);
  REQUIRE(d_src != nullptr || d_flow != nullptr, "%s: both gradient outputs are NULL", __func__);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0, "%s: C=%d", __func__, C);
  return launch_warp3d_bwd(g, src, flow, d_src, d_flow, B, C, D, H, W, (cudaStream_t)stream);
}

int smile_upsample2x_bwd(const float* g, float* d_x, int B, int C, int D, int H, int W, float premul,
                         smile_stream_t stream) {
  REQUIRE_PTR(g);
  REQUIRE_PTR(d_x);
  REQUIRE_VOL(B, 2 * D, 2 * H, 2 * W);
  REQUIRE(C > 0, "%s: C=%d", __func__, C);
  return launch_upsample2x_bwd(g, d_x, B, C, D, H, W, premul, (cudaStream_t)stream);
}

int smile_modet_attn_bwd(const float* g, const float* q, const float* k, const float* rpb, float* d_q, float* d_k,
                         float* d_rpb, float* work, int B, int D, int H, int W, int heads, int head_dim, float scale,
                         smile_stream_t stream) {
  REQUIRE_PTR(g);
  REQUIRE_PTR(q);
  REQUIRE_PTR(k);
  REQUIRE_PTR(d_q);
  REQUIRE_PTR(d_k);
  REQUIRE_PTR(work);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(heads > 0 && head_dim > 0 && heads <= 65535, "%s: heads=%d head_dim=%d", __func__, heads, head_dim);
  return launch_modet_attn_bwd(g, q, k, rpb, d_q, d_k, d_rpb, work, B, D, H, W, heads, head_dim, scale,
                               (cudaStream_t)stream);
}

int smile_proj_ln_bwd(const float* g, const float* feat, const float* weight, const float* bias, const float* gamma,
                      float* d_feat, float* d_weight, float* d_bias, float* d_gamma, float* d_beta, int B, int Cin, int C,
                      long long N, float eps, smile_stream_t stream) {
  REQUIRE_PTR(g);
  REQUIRE_PTR(feat);
  REQUIRE_PTR(weight);
  REQUIRE_PTR(bias);
  REQUIRE_PTR(gamma);
  REQUIRE_PTR(d_weight);
  REQUIRE_PTR(d_bias);
  REQUIRE_PTR(d_gamma);
  REQUIRE_PTR(d_beta);
  REQUIRE(B > 0 && Cin > 0 && C > 0 && N > 0 && N < (1LL << 31), "%s: bad sizes", __func__);
  return launch_proj_ln_bwd(g, feat, weight, bias, gamma, d_feat, d_weight, d_bias, d_gamma, d_beta, B, Cin, C, N, eps,
                            (cudaStream_t)stream);
}

int smile_cwm_fuse_bwd(const float* g, const float* fields, const float* logits, float* d_fields, float* d_logits, int B,
                       int F, long long N, smile_stream_t stream) {
  REQUIRE_PTR(g);
  REQUIRE_PTR(fields);
  REQUIRE_PTR(logits);
  REQUIRE_PTR(d_fields);
  REQUIRE_PTR(d_logits);
  REQUIRE(B > 0 && F > 0 && N > 0, "%s: bad sizes", __func__);
  return launch_cwm_fuse_bwd(g, fields, logits, d_fields, d_logits, B, F, N, (cudaStream_t)stream);
}

int smile_conv3d_flip_weights(const float* w, float* wT, int Cout, int Cin, smile_stream_t stream) {
  REQUIRE_PTR(w);
  REQUIRE_PTR(wT);
  REQUIRE(Cout > 0 && Cin > 0 && w != wT, "%s: bad arguments", __func__);
  return launch_conv3d_flip_weights(w, wT, Cout, Cin, (cudaStream_t)stream);
}

int smile_conv3d_wgrad(const float* in, const float* d_out, float* d_w, float* d_b, int B, int Cin, int Cout, int D, int H,
                       int W, smile_stream_t stream) {
  REQUIRE_PTR(in);
  REQUIRE_PTR(d_out);
  REQUIRE_PTR(d_w);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(Cin > 0 && Cout > 0 && B <= 65535, "%s: bad sizes", __func__);
  return launch_conv3d_wgrad(in, d_out, d_w, d_b, B, Cin, Cout, D, H, W, (cudaStream_t)stream);
}

int smile_conv3d_wgrad_bf16(const float* in, const float* d_out, float* d_w, float* d_b, int B, int Cin, int Cout, int D,
                            int H, int W, smile_stream_t stream) {
  REQUIRE_PTR(in);
  REQUIRE_PTR(d_out);
  REQUIRE_PTR(d_w);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(Cin > 0 && Cout > 0 && Cin <= 4096 && Cout <= 4096, "%s: Cin=%d Cout=%d out of range", __func__, Cin, Cout);
  REQUIRE(B <= 65535, "%s: B=%d exceeds grid.z of the full-precision fallback", __func__, B);
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(d_w, 0, (size_t)Cout * Cin * 27 * sizeof(float), st) != cudaSuccess ||
      (d_b != nullptr && cudaMemsetAsync(d_b, 0, (size_t)Cout * sizeof(float), st) != cudaSuccess)) {
    set_error("%s: memset failed", __func__);
    return SMILE_ERR_CUDA;
  }
  // The im2col build of the tensor-core kernel reads the activations with scattered 4-byte loads (L1-bound): it wins from the
  // 80-wide level down (16->16 @80x96x80 0.71 -> 0.47 ms, 128->128 @10x12x10 0.51 -> < 0.1) and loses on the 160-wide level
  // (8->8 1.10 -> 1.60), which stays on the SIMT kernels.  SMILE_WGRAD_TC=2 forces it everywhere, =0 switches it off.
  static const int knob = [] { const char* e = getenv("SMILE_WGRAD_TC"); return e ? atoi(e) : 1; }();
  if (knob == 2 || (knob == 1 && (long long)D * H * W <= 80LL * 96 * 80)) {
    bool handled = false;
    const int rc = launch_conv3d_wgrad_tc(in, d_out, d_w, d_b, B, Cin, Cout, D, H, W, st, &handled);
    if (handled) return rc;
  }
  return launch_conv3d_wgrad(in, d_out, d_w, d_b, B, Cin, Cout, D, H, W, st);
}

int smile_in_lrelu_bwd(const float* d_act, const float* act, const double* fwd_stats, void* work, float* d_raw, int B, int C,
                       long long N, float eps, int mode, smile_stream_t stream) {
  REQUIRE_PTR(d_act);
  REQUIRE_PTR(act);
  REQUIRE_PTR(d_raw);
  REQUIRE(mode == 1 || (fwd_stats != nullptr && work != nullptr), "%s: mode 0 needs fwd_stats and work", __func__);
  REQUIRE(B > 0 && C > 0 && N > 0 && (long long)B * C <= 65535, "%s: bad sizes", __func__);
  return launch_in_lrelu_bwd(d_act, act, fwd_stats, reinterpret_cast<double*>(work), d_raw, B, C, N, eps, mode,
                             (cudaStream_t)stream);
}

int smile_avgpool2_bwd_add(const float* d_pooled, float* d_full, int B, int C, int D, int H, int W, smile_stream_t stream) {
  REQUIRE_PTR(d_pooled);
  REQUIRE_PTR(d_full);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0 && (long long)B * C <= 65535, "%s: bad sizes", __func__);
  return launch_pool_bwd_add(d_pooled, d_full, B, C, D, H, W, (cudaStream_t)stream);
}

int smile_ncc_vxm_bwd(const float* y_true, const float* y_pred, float* d_true, void* work, const float* gscale, int B, int D,
                      int H, int W, int win, smile_stream_t stream) {
  REQUIRE_PTR(y_true);
  REQUIRE_PTR(y_pred);
  REQUIRE_PTR(d_true);
  REQUIRE_PTR(work);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(win >= 1 && (win & 1) && win <= 63, "%s: win=%d must be odd and <= 63", __func__, win);
  return launch_ncc_vxm_bwd(y_true, y_pred, d_true, reinterpret_cast<float*>(work), gscale, B, D, H, W, win,
                            (cudaStream_t)stream);
}

int smile_grad3d_l2_bwd(const float* flow, float* d_flow, const float* gscale, int B, int C, int D, int H, int W,
                        smile_stream_t stream) {
  REQUIRE_PTR(flow);
  REQUIRE_PTR(d_flow);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0 && D > 1 && H > 1 && W > 1, "%s: needs C > 0 and every spatial dim > 1", __func__);
  return launch_grad3d_l2_bwd(flow, d_flow, gscale, B, C, D, H, W, (cudaStream_t)stream);
}

int smile_adam_amsgrad_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                            long long n, float lr, float beta1, float beta2, float eps, int step, smile_stream_t stream) {
  REQUIRE_PTR(param);
  REQUIRE_PTR(grad);
  REQUIRE_PTR(exp_avg);
  REQUIRE_PTR(exp_avg_sq);
  REQUIRE_PTR(max_exp_avg_sq);
  REQUIRE(n > 0 && step >= 1, "%s: n=%lld step=%d", __func__, n, step);
  return launch_adam_amsgrad(param, grad, exp_avg, exp_avg_sq, max_exp_avg_sq, n, lr, beta1, beta2, eps, step,
                             (cudaStream_t)stream);
}

int smile_warp3d_nearest_fwd(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W,
                             smile_stream_t stream) {
  REQUIRE_PTR(src);
  REQUIRE_PTR(flow);
  REQUIRE_PTR(out);
  REQUIRE_VOL(B, D, H, W);
  REQUIRE(C > 0, "%s: C=%d", __func__, C);
  REQUIRE(out != src, "%s: out must not alias src", __func__);
  return launch_warp3d_nearest(src, flow, out, B, C, D, H, W, (cudaStream_t)stream);
}

int smile_dice_counts_fwd(const float* pred, const float* truth, const int* labels, int nlabels,
                          unsigned long long* counts, long long n, smile_stream_t stream) {
  REQUIRE_PTR(pred);
  REQUIRE_PTR(truth);
  REQUIRE(labels != nullptr && counts != nullptr, "%s: labels / counts is NULL", __func__);
  REQUIRE(nlabels > 0 && n > 0, "%s: nlabels=%d n=%lld", __func__, nlabels, n);
  return launch_dice_counts(pred, truth, labels, nlabels, counts, n, (cudaStream_t)stream);
}

int smile_jacdet_fwd(const float* flow, double* det, unsigned long long* nonpos, int D, int H, int W,
                     smile_stream_t stream) {
  REQUIRE_PTR(flow);
  REQUIRE(nonpos != nullptr, "%s: nonpos is NULL", __func__);
  REQUIRE_VOL(1, D, H, W);
  REQUIRE(D >= 2 && H >= 2 && W >= 2, "%s: np.gradient needs at least 2 samples per axis (D=%d H=%d W=%d)", __func__, D, H, W);
  return launch_jacdet(flow, det, nonpos, D, H, W, (cudaStream_t)stream);
}

}  // extern "C"
