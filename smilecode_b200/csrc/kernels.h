// Internal launch entry points (C++ linkage); the exported C ABI lives in capi.cu / include/smilecode_b200.h.
#pragma once
#include <cuda_runtime.h>

namespace smile {

// warp.cu
int launch_warp3d(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W, cudaStream_t st);
int launch_warp_proj_ln(const float* src, const float* flow, const float* weight, const float* bias, const float* gamma,
                        const float* beta, float* out, int B, int Cin, int C, int D, int H, int W, float eps,
                        cudaStream_t st, bool* handled);
int launch_compose(const float* flow, const float* w, float* out, int B, int D, int H, int W, float post, cudaStream_t st);
int launch_upsample2x(const float* x, float* out, int B, int C, int D, int H, int W, float pre, cudaStream_t st);

// attn.cu
int launch_modet_attn(const float* q, const float* k, const float* rpb, float* out, int B, int D, int H, int W,
                      int heads, int hd, float scale, cudaStream_t st);
// fused heads==1 level: w = attn(q,k); flow_out = post*(T(flow_in, w) + w); optionally moved = T(moving, flow_out)
// ln_gamma / ln_beta: optional device pointers to the LayerNorm affine parameters that produced q and k (maximum-free
// softmax when the logit bound they imply is small; null = always the online-maximum softmax)
int launch_modet_fused(const float* q, const float* k, const float* rpb, const float* ln_gamma, const float* ln_beta,
                       const float* flow_in, const float* moving, float* flow_out, float* moved, int B, int D, int H, int W,
                       int hd, float scale, float post, int Cmov, cudaStream_t st);

// proj.cu
int launch_proj_ln(const float* feat, const float* weight, const float* bias, const float* gamma, const float* beta,
                   float* out, int B, int Cin, int C, long long N, float eps, cudaStream_t st);

// conv.cu
int launch_conv3d(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                  double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                  cudaStream_t st, const float* wprep = nullptr);
int launch_conv3d_tc(const float* in, const float* weight, const float* wprep, const float* bias, float* out,
                     const double* in_stats, double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out,
                     float eps, cudaStream_t st, bool* handled);
int launch_conv3d_bf16(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                       double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                       cudaStream_t st, bool* handled);
int launch_conv3d_march_bf16(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                             double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                             cudaStream_t st, bool* handled);
int launch_conv3d_march_split(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                             double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                             cudaStream_t st, bool* handled);
long long conv3d_tc_prep_floats(int Cin, int Cout);
int launch_conv3d_tc_prep(const float* weight, float* wprep, int Cin, int Cout, cudaStream_t st);
int launch_conv3d_tma_flat(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                           double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                           cudaStream_t st, bool* handled);
int launch_in_finalize(const float* raw, const double* stats, float* out, float* pooled, int B, int C, int D, int H,
                       int W, float eps, cudaStream_t st);
int launch_cwm_fuse(const float* fields, const float* logits, float* out, int B, int F, long long N, cudaStream_t st);


// qkrpb.cu (a3: twins of the reference's modet_fw / modet_bw)
int launch_qkrpb_fwd(const float* q, const float* kpad, const float* rpb, float* attn, int B, int heads, int H, int W,
                     int T, int hd, cudaStream_t st);
int launch_qkrpb_bwd(const float* d_attn, const float* q, const float* kpad, float* dq, float* dk, float* drpb, int B,
                     int heads, int H, int W, int T, int hd, cudaStream_t st);

// losses.cu (a10)
int launch_ncc_vxm(const float* y_true, const float* y_pred, float* out, float* work, int B, int D, int H, int W, int win,
                   cudaStream_t st);
int launch_grad3d_l2(const float* flow, float* out, double* work, int B, int C, int D, int H, int W, cudaStream_t st);


// metrics.cu (evaluation path of infer.py)
int launch_warp3d_nearest(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W,
                          cudaStream_t st);
int launch_dice_counts(const float* pred, const float* truth, const int* labels, int nlabels, unsigned long long* counts,
                       long long n, cudaStream_t st);
int launch_jacdet(const float* flow, double* det, unsigned long long* nonpos, int D, int H, int W, cudaStream_t st);

// backward.cu (training path)
int launch_warp3d_bwd(const float* g, const float* src, const float* flow, float* d_src, float* d_flow, int B, int C, int D,
                      int H, int W, cudaStream_t st);
int launch_upsample2x_bwd(const float* g, float* dx, int B, int C, int D, int H, int W, float pre, cudaStream_t st);
int launch_modet_attn_bwd(const float* g, const float* q, const float* k, const float* rpb, float* dq, float* dk,
                          float* drpb, float* dl_work, int B, int D, int H, int W, int heads, int hd, float scale,
                          cudaStream_t st);
int launch_proj_ln_bwd(const float* gout, const float* feat, const float* weight, const float* bias, const float* gamma,
                       float* dfeat, float* dweight, float* dbias, float* dgamma, float* dbeta, int B, int Cin, int C,
                       long long N, float eps, cudaStream_t st);
int launch_cwm_fuse_bwd(const float* g, const float* fields, const float* logits, float* dfields, float* dlogits, int B,
                        int F, long long N, cudaStream_t st);


// backward_conv.cu
int launch_conv3d_flip_weights(const float* w, float* wT, int Cout, int Cin, cudaStream_t st);
int launch_conv3d_wgrad(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int D, int H,
                        int W, cudaStream_t st);
int launch_conv3d_wgrad_tc(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int D, int H,
                           int W, cudaStream_t st, bool* handled);
int launch_in_lrelu_bwd(const float* da, const float* act, const double* fwd_stats, double* sums_work, float* dy, int B,
                        int C, long long N, float eps, int mode, cudaStream_t st);
int launch_pool_bwd_add(const float* dpooled, float* dfull, int B, int C, int D, int H, int W, cudaStream_t st);
int launch_grad3d_l2_bwd(const float* flow, float* dflow, const float* gscale, int B, int C, int D, int H, int W,
                         cudaStream_t st);
int launch_ncc_vxm_bwd(const float* y_true, const float* y_pred, float* d_true, float* work, const float* gscale, int B,
                       int D, int H, int W, int win, cudaStream_t st);
int launch_adam_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float b1,
                        float b2, float eps, int step, cudaStream_t st);

}  // namespace smile
