/* smilecode_b200 -- C ABI of the B200-native (sm_100a) ModeT registration hot path.
 *
 * Drop-in boundary.  The reference's only native interface is the pybind pair
 *     modet_fw(query, key, rpb) -> attn          (ModeT-cu/modet/modet.cpp:4-18, include/utils.h:23-27)
 *     modet_bw(d_attn, query, key, biasEnabled)  (ModeT-cu/modet/modet.cpp:20-31, include/utils.h:39-44)
 * called from ModeT-cu/functional.py:5-28; every other step of the path is a torch library call
 * inside ModeT/models.py.  This header is what a binding for the whole path binds instead: one
 * entry point per reference function on the path, plain pointers and
 * sizes, no torch types.
 *
 * Conventions (all entry points):
 *   - every pointer is a DEVICE pointer to contiguous fp32 (fp64 where stated), 16-byte aligned;
 *   - outputs are caller-allocated (the caller's allocator owns memory; the reference's callee-
 *     allocated torch::zeros at modet_kernel.cu:115 is replaced by this);
 *   - work is enqueued on `stream` (a cudaStream_t) and never synchronises the host
 *     (reference: c10::cuda::getCurrentCUDAStream(), modet_kernel.cu:119,362);
 *   - returns SMILE_OK (0) or a negative error code; smile_last_error() gives the message of the
 *     last failure on the calling thread (reference: TORCH_CHECK -> RuntimeError, utils.h:7-14);
 *   - no global mutable state besides the per-process TMA descriptor cache; re-entrant.
 *   - volumes are [B, C, D, H, W] channels-first unless stated; "channels-last" is [B, D, H, W, C].
 */
#ifndef SMILECODE_B200_H
#define SMILECODE_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define SMILE_OK 0
#define SMILE_ERR_INVALID_ARG (-1)
#define SMILE_ERR_CUDA (-2)
#define SMILE_ERR_UNSUPPORTED (-3)

typedef void* smile_stream_t; /* cudaStream_t */

/* Library version (major*10000 + minor*100 + patch) and last error text (thread-local). */
int smile_version(void);
const char* smile_last_error(void);

/* a2  ModeTransformer.forward (ModeT/models.py:308-334; supersedes modet_fw + softmax + attn@v,
 * ModeT-cu/models.py:304-314).  q, k: channels-last [B,D,H,W,heads*head_dim]; rpb: [heads,3,3,3]
 * or NULL; out: [B,3*heads,D,H,W].  The key volume is zero padded and the padded taps stay in
 * the softmax. */
int smile_modet_attn_fwd(const float* q, const float* k, const float* rpb, float* out, int B, int D, int H, int W,
                         int heads, int head_dim, float scale, smile_stream_t stream);

/* a5  SpatialTransformer.forward (ModeT/models.py:49-67): out[b,c] = trilinear sample of src[b,c]
 * at voxel + flow[b,:,voxel]; zeros padding, align_corners=True, integer corner indices bit-exact
 * with torch.  src/out: [B,C,D,H,W]; flow: [B,3,D,H,W]. */
int smile_warp3d_fwd(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W,
                     smile_stream_t stream);

/* a6  nn.Upsample(scale_factor=2, 'trilinear', align_corners=True) (ModeT/models.py:354, 257-261)
 * with the power-of-two pre-scale folded in: out = premul * up2(x).  x: [B,C,D,H,W] -> [B,C,2D,2H,2W]. */
int smile_upsample2x_fwd(const float* x, float* out, int B, int C, int D, int H, int W, float premul,
                         smile_stream_t stream);

/* a6  flow composition (ModeT/models.py:392, 398, 403, 408):
 * out = postmul * (SpatialTransformer(flow, w) + w); all [B,3,D,H,W]. */
int smile_flow_compose_fwd(const float* flow, const float* w, float* out, int B, int D, int H, int W, float postmul,
                           smile_stream_t stream);

/* a2+a6+a5 fused, for the heads==1 pyramid levels (ModeT/models.py:401-403 and 406-410):
 *   w        = ModeTransformer(q, k)                      (heads = 1)
 *   flow_out = postmul * (SpatialTransformer(flow_in, w) + w)
 *   moved    = SpatialTransformer(moving, flow_out)       (skipped when moved == NULL)
 * q, k: channels-last [B,D,H,W,head_dim]; flow_in/flow_out: [B,3,D,H,W]; moving/moved: [B,Cmov,D,H,W].
 * ln_gamma, ln_beta: optional (both or neither) device pointers to the head_dim LayerNorm affine parameters of the
 * ProjectionLayer that produced BOTH q and k (models.py:233, 240).  They are a promise about the inputs, not an operand:
 * |LayerNorm(x)|_2 <= sqrt(head_dim), so |logit| <= |scale| * (max|gamma| * sqrt(head_dim) + |beta|_2)^2 + max|rpb|; the
 * kernel evaluates that bound and, when it is far inside the fp32 exponent range, sums the exponentials without a running
 * maximum.  NULL (or a large bound, or non-finite parameters) selects the online-maximum softmax, valid for any q, k.
 * Numerics: ex2.approx exponentials accumulated tap plane by tap plane, the two trilinear samples are separable lerps --
 * equal to the three separate calls within ~1e-5 absolute (tolerance-checked, 1e-4), not bit for bit; smile_warp3d_fwd
 * and smile_flow_compose_fwd are the bit-exact forms. */
int smile_modet_fused_fwd(const float* q, const float* k, const float* rpb, const float* ln_gamma, const float* ln_beta,
                          const float* flow_in, const float* moving, float* flow_out, float* moved, int B, int D, int H,
                          int W, int head_dim, float scale, float postmul, int Cmov, smile_stream_t stream);

/* a7  ProjectionLayer.forward (ModeT/models.py:238-241): channels-first feat [B,Cin,N] ->
 * LayerNorm(Linear(feat)) channels-last [B,N,C].  weight: [C,Cin]; bias, gamma, beta: [C]. */
int smile_proj_ln_fwd(const float* feat, const float* weight, const float* bias, const float* gamma, const float* beta,
                      float* out, int B, int Cin, int C, long long N, float eps, smile_stream_t stream);

/* a5+a7 fused (the pattern `M = transformer(M, flow); k = projblock(M)` of ModeT/models.py:388-389, 394-395,
 * 400-401, 405-406): out = LayerNorm(Linear(SpatialTransformer(src, flow))) channels-last [B,D,H,W,C]
 * without materialising the warped feature volume.  src: [B,Cin,D,H,W]; flow: [B,3,D,H,W]. */
int smile_warp_proj_ln_fwd(const float* src, const float* flow, const float* weight, const float* bias,
                           const float* gamma, const float* beta, float* out, int B, int Cin, int C, int D, int H,
                           int W, float eps, smile_stream_t stream);

/* a8/a4 Conv3d(kernel 3, stride 1, padding 1) (ModeT/models.py:127, 143, 253).
 *   in_stats  (fp64 [B*Cin][2] = sum, sum of squares of `in`) non-NULL: `in` is a raw conv output and
 *             InstanceNorm3d(eps)+LeakyReLU(0.1) of it is applied on load (models.py:148-150);
 *   out_stats (fp64 [B*Cout][2], zeroed by the caller) non-NULL: sum / sum of squares of the raw
 *             output are accumulated into it for the following InstanceNorm;
 *   act_out != 0: LeakyReLU(0.1) on the stored output (ConvBlock, models.py:131-132).
 * weight: [Cout,Cin,3,3,3]; bias: [Cout]. */
int smile_conv3d_fwd(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                     double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                     smile_stream_t stream);

/* Same contract with bf16 MMA operands: tcgen05 kind::f16, fp32 accumulation in TMEM (csrc/conv_bf16.cu) -- the
 * reduced-precision encoder of BASELINE.json configs[2..3].  in / out / statistics stay fp32; only the operands of the
 * products are rounded (activations after the normalise-on-load, weights), so a layer output differs from
 * smile_conv3d_fwd by ~1e-3 relative.  Layers with fewer than 4 input channels (and rows too wide for the staged halo)
 * run on the fp32 kernels inside this call. */
int smile_conv3d_bf16_fwd(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                          double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                          smile_stream_t stream);

/* Optional prepared weights for the tensor-core convolution (layers with >= 16 input and >= 12 output channels on
 * volumes up to 48 voxels wide run on tcgen05 with a 3xTF32 split, csrc/conv_tc.cu).  smile_conv3d_fwd prepares the
 * split weights on every call into stream-ordered scratch; a caller with constant weights can do it once:
 *   n = smile_conv3d_tc_prep_floats(Cin, Cout);  allocate n floats;  smile_conv3d_tc_prep(weight, wprep, Cin, Cout, s);
 *   smile_conv3d_prepped_fwd(..., wprep, ...)   -- same contract as smile_conv3d_fwd; wprep may be NULL. */
long long smile_conv3d_tc_prep_floats(int Cin, int Cout);
int smile_conv3d_tc_prep(const float* weight, float* wprep, int Cin, int Cout, smile_stream_t stream);
int smile_conv3d_prepped_fwd(const float* in, const float* weight, const float* wprep, const float* bias, float* out,
                             const double* in_stats, double* out_stats, int B, int Cin, int Cout, int D, int H, int W,
                             int act_out, float eps, smile_stream_t stream);

/* a8  InstanceNorm3d + LeakyReLU(0.1) from fp64 sums (models.py:149-150), and AvgPool3d(2)
 * (models.py:198) of the result into `pooled` [B,C,D/2,H/2,W/2] when pooled != NULL.  out may alias raw. */
int smile_instnorm_lrelu_pool_fwd(const float* raw, const double* stats, float* out, float* pooled, int B, int C, int D,
                                  int H, int W, float eps, smile_stream_t stream);

/* a4  CWM tail (ModeT/models.py:254, 268-275): out = 2 * sum_f fields[:,3f:3f+3] * softmax_f(logits).
 * fields: [B,3F,N]; logits: [B,F,N]; out: [B,3,N]. */
int smile_cwm_fuse_fwd(const float* fields, const float* logits, float* out, int B, int F, long long N,
                       smile_stream_t stream);

/* a3  Twins of the reference's pybind pair (ModeT-cu/modet/modet.cpp:4-37; kernels modet_kernel.cu:17-381;
 * Python caller ModeT-cu/functional.py:5-28).  Same tensors and layouts as the reference:
 *   q      [B,heads,H,W,T,head_dim]        (already multiplied by scale, ModeT-cu/models.py:304)
 *   kpad   [B,heads,H+2,W+2,T+2,head_dim]  (zero padded by the caller, models.py:305-306)
 *   rpb    [heads,3,3,3] or NULL           (modet.cpp:13 substitutes zeros for None)
 *   attn, d_attn [B,heads,H,W,T,27]        tap t = (ti*3 + tj)*3 + tk
 * modet_fw:  attn[...,t] = <q, kpad[.. + (ti,tj,tk)]> + rpb[head,t]        (pre-softmax logits)
 * modet_bw:  d_q, d_kpad (PADDED shape; the caller's pad-backward crops it), d_rpb (may be NULL when the
 *            forward had no bias).  d_rpb is zeroed inside the call (the reference allocates torch::zeros). */
int smile_modet_qkrpb_fwd(const float* q, const float* kpad, const float* rpb, float* attn, int B, int heads, int H,
                          int W, int T, int head_dim, smile_stream_t stream);
int smile_modet_qkrpb_bwd(const float* d_attn, const float* q, const float* kpad, float* d_q, float* d_kpad,
                          float* d_rpb, int B, int heads, int H, int W, int T, int head_dim, smile_stream_t stream);

/* a10 NCC_vxm.forward (ModeT/losses.py:43-95): out[0] = -mean(cc), cc from separable win^3 box sums of
 * I, J, I^2, J^2, IJ with zero padding.  y_true, y_pred: [B,1,D,H,W].  work: device scratch of at least
 * smile_ncc_vxm_work_bytes(B,D,H,W) bytes, 16-byte aligned. */
long long smile_ncc_vxm_work_bytes(int B, int D, int H, int W);
int smile_ncc_vxm_fwd(const float* y_true, const float* y_pred, float* out, void* work, int B, int D, int H, int W,
                      int win, smile_stream_t stream);

/* a10 Grad3d(penalty='l2').forward (ModeT/losses.py:16-31): out[0] = (mean(dD^2) + mean(dH^2) + mean(dW^2)) / 3 over
 * forward differences of flow [B,C,D,H,W].  work: 3 doubles of device scratch (24 bytes, 16-byte aligned). */
int smile_grad3d_l2_fwd(const float* flow, float* out, void* work, int B, int C, int D, int H, int W,
                        smile_stream_t stream);

/* ---------------------------------------------------------------------------------------------------
 * Backward entry points (training path).  The reference gets these from torch autograd over the library
 * ops it calls; here each is a hand-written kernel.  Gradient outputs are caller-allocated
 * and are (re)initialised inside the call where the kernel accumulates with atomics.
 * ------------------------------------------------------------------------------------------------- */

/* SpatialTransformer backward (ModeT/models.py:49-67 through grid_sample): g = d(loss)/d(out) [B,C,D,H,W];
 * d_src [B,C,D,H,W] and/or d_flow [B,3,D,H,W]; either may be NULL. */
int smile_warp3d_bwd(const float* g, const float* src, const float* flow, float* d_src, float* d_flow, int B, int C,
                     int D, int H, int W, smile_stream_t stream);

/* Adjoint of smile_upsample2x_fwd: g [B,C,2D,2H,2W] -> d_x [B,C,D,H,W] (premul folded in). */
int smile_upsample2x_bwd(const float* g, float* d_x, int B, int C, int D, int H, int W, float premul,
                         smile_stream_t stream);

/* ModeTransformer backward (ModeT/models.py:308-334): g [B,3*heads,D,H,W]; q, k, d_q, d_k channels-last
 * [B,D,H,W,heads*head_dim]; d_rpb [heads,3,3,3] or NULL; work: B*D*H*W*heads*27 floats of scratch (the
 * d_logits, which the reference materialises as the gradient of its attn tensor). */
int smile_modet_attn_bwd(const float* g, const float* q, const float* k, const float* rpb, float* d_q, float* d_k,
                         float* d_rpb, float* work, int B, int D, int H, int W, int heads, int head_dim, float scale,
                         smile_stream_t stream);

/* ProjectionLayer backward (ModeT/models.py:238-241): g channels-last [B,N,C]; d_feat [B,Cin,N] (may be NULL);
 * d_weight [C,Cin], d_bias, d_gamma, d_beta [C]. */
int smile_proj_ln_bwd(const float* g, const float* feat, const float* weight, const float* bias, const float* gamma,
                      float* d_feat, float* d_weight, float* d_bias, float* d_gamma, float* d_beta, int B, int Cin, int C,
                      long long N, float eps, smile_stream_t stream);

/* CWM tail backward (ModeT/models.py:268-275): g [B,3,N] -> d_fields [B,3F,N], d_logits [B,F,N]. */
int smile_cwm_fuse_bwd(const float* g, const float* fields, const float* logits, float* d_fields, float* d_logits, int B,
                       int F, long long N, smile_stream_t stream);

/* Conv3d backward pieces (nn.Conv3d k=3 s=1 p=1; ModeT/models.py:127, 143, 253).
 *   data gradient:   d_in = smile_conv3d_fwd(d_out, wT, zero bias) with wT = smile_conv3d_flip_weights(w):
 *                    wT[ci][co][26-t] = w[co][ci][t]                       (wT: [Cin,Cout,3,3,3])
 *   weight gradient: d_w[co][ci][t] = sum_{b,v} d_out[b,co,v] * in[b,ci,v+off(t)], d_b[co] = sum d_out (d_b may be NULL) */
int smile_conv3d_flip_weights(const float* w, float* wT, int Cout, int Cin, smile_stream_t stream);
int smile_conv3d_wgrad(const float* in, const float* d_out, float* d_w, float* d_b, int B, int Cin, int Cout, int D, int H,
                       int W, smile_stream_t stream);
/* The same weight / bias gradient with the products on bf16 tensor cores (tcgen05, fp32 accumulation; operands rounded to
 * bf16: ~3e-3 relative on a gradient element) -- the weight-gradient half of the bf16 training mode, next to
 * smile_conv3d_bf16_fwd for the forward and data-gradient products. */
int smile_conv3d_wgrad_bf16(const float* in, const float* d_out, float* d_w, float* d_b, int B, int Cin, int Cout, int D,
                            int H, int W, smile_stream_t stream);

/* InstanceNorm3d + LeakyReLU(0.1) backward (models.py:148-150): act = lrelu(IN(raw)); given d_act, act and the forward
 * fp64 (sum, sumsq) of raw, writes d_raw.  mode 0: IN + LeakyReLU; mode 1: LeakyReLU only (ConvBlock, models.py:131-132;
 * stats / work may be NULL).  work: B*C*2 doubles of scratch.  d_raw may alias d_act. */
int smile_in_lrelu_bwd(const float* d_act, const float* act, const double* fwd_stats, void* work, float* d_raw, int B, int C,
                       long long N, float eps, int mode, smile_stream_t stream);

/* AvgPool3d(2) backward, accumulated: d_full[b,c,d,h,w] += d_pooled[b,c,d/2,h/2,w/2] / 8 (models.py:198). */
int smile_avgpool2_bwd_add(const float* d_pooled, float* d_full, int B, int C, int D, int H, int W, smile_stream_t stream);

/* Loss backward (ModeT/losses.py).  gscale: device pointer to the upstream scalar gradient (NULL = 1).
 * NCC_vxm: gradient w.r.t. the FIRST argument (train.py:126 passes the warped image there); work as in the forward. */
int smile_ncc_vxm_bwd(const float* y_true, const float* y_pred, float* d_true, void* work, const float* gscale, int B, int D,
                      int H, int W, int win, smile_stream_t stream);
int smile_grad3d_l2_bwd(const float* flow, float* d_flow, const float* gscale, int B, int C, int D, int H, int W,
                        smile_stream_t stream);

/* torch.optim.Adam(amsgrad=True, weight_decay=0) update of one flat fp32 buffer (ModeT/train.py:101); step counts from 1. */
int smile_adam_amsgrad_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq,
                            long long n, float lr, float beta1, float beta2, float eps, int step, smile_stream_t stream);

/* ---- evaluation path of ModeT/infer.py:86-92 -------------------------------------------------- */

/* utils.register_model(img_size, 'nearest') (ModeT/utils.py:74-83 -> SpatialTransformer(mode='nearest'), 30-72):
 * out[b,c,p] = src[b,c,nearbyint(p + flow[b,:,p])] with align_corners=True coordinates, 0 outside the volume.
 * src/out [B,C,D,H,W], flow [B,3,D,H,W]. */
int smile_warp3d_nearest_fwd(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W,
                             smile_stream_t stream);

/* utils.dice_val_VOI (ModeT/utils.py:86-106) without the device->host copy of the volumes: for every label l of
 * labels[0..nlabels) counts[3*i+0] = |pred==l & truth==l|, [3*i+1] = |pred==l|, [3*i+2] = |truth==l| (unsigned 64-bit,
 * zeroed by the call).  pred/truth hold label values as fp32, compared after truncation (`.long()`, infer.py:91);
 * labels must lie in [0, 1024), nlabels <= 256.  The caller forms mean_i 2*c0/(c1+c2+1e-5) in double. */
int smile_dice_counts_fwd(const float* pred, const float* truth, const int* labels, int nlabels,
                          unsigned long long* counts, long long n, smile_stream_t stream);

/* utils.jacobian_determinant_vxm (ModeT/utils.py:108-150) for one 3-D displacement field flow[3,D,H,W]:
 * np.gradient(disp + grid) in float64 (central differences, one-sided at the faces) and the 3x3 determinant,
 * evaluated operation by operation like numpy.  det (optional, may be NULL): [D,H,W] float64; nonpos: number of
 * voxels with det <= 0 (unsigned 64-bit, zeroed by the call) -- infer.py:90 divides it by D*H*W. */
int smile_jacdet_fwd(const float* flow, double* det, unsigned long long* nonpos, int D, int H, int W,
                     smile_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SMILECODE_B200_H */
