// a2 ModeTransformer.forward (reference ModeT/models.py:308-334; ModeT-cu/modet/modet_kernel.cu:17-87
// computes only the logits of this) -- generic kernels for any shape / head count.
//
// One thread owns one (voxel, head): the hd-float query row stays in registers, the 27 key rows
// are read through L1 (adjacent lanes read adjacent rows of the channels-last volume, so every
// tap is a fully used contiguous segment), the 27 logits + softmax + the three signed sums of
// "attn @ V" never leave the register file.  Zero-padded taps keep logit = rpb (models.py:319).
// The TMA-tiled fast path for the large heads==1 levels lives in attn_tma.cu.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace smile {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

constexpr float kLog2e = 1.4426950408889634f;

// Softmax over the 27 logits (already scaled by log2 e) and expectation of the tap offsets.
// EXACT: exp2f() (full-accuracy) instead of ex2.approx -- the bisect knob SMILE_ATTN_EXACT=1 (tools/parity_bisect.py).
template <bool EXACT = false>
__device__ __forceinline__ void softmax_expect(const float (&lg)[27], float& od, float& oh, float& ow) {
  float m = lg[0];
#pragma unroll
  for (int t = 1; t < 27; ++t) m = fmaxf(m, lg[t]);
  float sum = 0.f, sd = 0.f, sh = 0.f, sw = 0.f;
#pragma unroll
  for (int t = 0; t < 27; ++t) {
    float p = EXACT ? exp2f(lg[t] - m) : fast_exp2(lg[t] - m);
    sum += p;
    const int ti = t / 9 - 1, tj = (t / 3) % 3 - 1, tk = t % 3 - 1;
    if (ti != 0) sd += (ti > 0 ? p : -p);
    if (tj != 0) sh += (tj > 0 ? p : -p);
    if (tk != 0) sw += (tk > 0 ? p : -p);
  }
  float inv = 1.0f / sum;
  od = sd * inv;
  oh = sh * inv;
  ow = sw * inv;
}

// logits of one (voxel, head) straight from global memory.  HD > 0: compile-time head_dim
// (even, rows 8 B aligned); HD == 0: run-time head_dim, scalar loads.
template <int HD>
__device__ __forceinline__ void logits_from_global(const float* __restrict__ q, const float* __restrict__ k,
                                                   const float* __restrict__ rpb27, int hd, int C, int d, int h, int w,
                                                   int D, int H, int W, float qscale, float (&lg)[27]) {
  if (HD > 0) {
    float2 qv[HD > 0 ? HD / 2 : 1];
#pragma unroll
    for (int i = 0; i < HD / 2; ++i) {
      qv[i] = __ldg(reinterpret_cast<const float2*>(q) + i);
      qv[i].x *= qscale;
      qv[i].y *= qscale;
    }
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int dd = d + t / 9 - 1, hh = h + (t / 3) % 3 - 1, ww = w + t % 3 - 1;
      float acc = 0.f;
      if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const float2* kr = reinterpret_cast<const float2*>(k + ((long long)(t / 9 - 1) * H * W + ((t / 3) % 3 - 1) * W + (t % 3 - 1)) * C);
#pragma unroll
        for (int i = 0; i < HD / 2; ++i) {
          float2 kv = __ldg(kr + i);
          acc = fmaf(qv[i].x, kv.x, acc);
          acc = fmaf(qv[i].y, kv.y, acc);
        }
      }
      lg[t] = acc + rpb27[t];
    }
  } else {
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int dd = d + t / 9 - 1, hh = h + (t / 3) % 3 - 1, ww = w + t % 3 - 1;
      float acc = 0.f;
      if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const float* kr = k + ((long long)(t / 9 - 1) * H * W + ((t / 3) % 3 - 1) * W + (t % 3 - 1)) * C;
        for (int i = 0; i < hd; ++i) acc = fmaf(__ldg(q + i) * qscale, __ldg(kr + i), acc);
      }
      lg[t] = acc + rpb27[t];
    }
  }
}

// q,k [B,D,H,W,heads*hd] channels-last -> out [B,3*heads,D,H,W]
template <int HD, bool EXACT = false>
__global__ void __launch_bounds__(128) attn_generic_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                           const float* __restrict__ rpb, float* __restrict__ out, int D,
                                                           int H, int W, int heads, int hd, float qscale) {
  extern __shared__ float s_rpb[];  // heads*27, pre-multiplied by log2(e)
  for (int i = threadIdx.x; i < heads * 27; i += blockDim.x) s_rpb[i] = rpb ? rpb[i] * kLog2e : 0.f;
  __syncthreads();
  const int HW = H * W, C = heads * hd;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y;
  const float* qb = q + (long long)b * N * C;
  const float* kb = k + (long long)b * N * C;
  float* ob = out + (long long)b * 3 * heads * N;
  const long long total = N * heads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / heads;
    const int head = (int)(i - p * heads);
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    float lg[27];
    logits_from_global<HD>(qb + p * C + head * hd, kb + p * C + head * hd, s_rpb + head * 27, hd, C, d, h, w, D, H, W,
                           qscale, lg);
    float od, oh, ow;
    softmax_expect<EXACT>(lg, od, oh, ow);
    float* o = ob + (long long)head * 3 * N + p;
    o[0] = od;
    o[N] = oh;
    o[2 * N] = ow;
  }
}

// Generic (no TMA) fused heads==1 level: attention -> flow compose -> optional warp of `moving`.
//   w = attn(q,k);  f' = post * (T(flow_in, w) + w);  moved[c] = T(moving[c], f')
template <int HD, bool EXACT = false>
__global__ void __launch_bounds__(128) fused_generic_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ rpb, const float* __restrict__ flow_in,
                                                            const float* __restrict__ moving, float* __restrict__ flow_out,
                                                            float* __restrict__ moved, int D, int H, int W, int hd,
                                                            float qscale, float post, int Cmov) {
  __shared__ float s_rpb[27];
  if (threadIdx.x < 27) s_rpb[threadIdx.x] = rpb ? rpb[threadIdx.x] * kLog2e : 0.f;
  __syncthreads();
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y;
  const float* qb = q + (long long)b * N * hd;
  const float* kb = k + (long long)b * N * hd;
  const float* fb = flow_in + (long long)b * 3 * N;
  float* ob = flow_out + (long long)b * 3 * N;
  const float dm1 = (float)(D - 1), hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    float lg[27];
    logits_from_global<HD>(qb + p * hd, kb + p * hd, s_rpb, hd, hd, d, h, w, D, H, W, qscale, lg);
    float w0, w1, w2;
    softmax_expect<EXACT>(lg, w0, w1, w2);
    TriSample s;
    tri_setup(s, st_coord(d, w0, dm1), st_coord(h, w1, hm1), st_coord(w, w2, wm1), D, H, W);
    const float f0 = __fmul_rn(post, __fadd_rn(tri_gather(s, fb), w0));
    const float f1 = __fmul_rn(post, __fadd_rn(tri_gather(s, fb + N), w1));
    const float f2 = __fmul_rn(post, __fadd_rn(tri_gather(s, fb + 2 * N), w2));
    ob[p] = f0;
    ob[N + p] = f1;
    ob[2 * N + p] = f2;
    if (moved != nullptr) {
      tri_setup(s, st_coord(d, f0, dm1), st_coord(h, f1, hm1), st_coord(w, f2, wm1), D, H, W);
      for (int c = 0; c < Cmov; ++c)
        moved[((long long)b * Cmov + c) * N + p] = tri_gather(s, moving + ((long long)b * Cmov + c) * N);
    }
  }
}

int launch_modet_attn_tma(const float* q, const float* k, const float* rpb, const float* flow_in, const float* moving,
                          float* w_out, float* flow_out, float* moved, int B, int D, int H, int W, float scale, float post,
                          int Cmov, cudaStream_t st, bool* handled);
int launch_modet_attn_tma2(const float* q, const float* k, const float* rpb, const float* ln_gamma, const float* ln_beta,
                           const float* flow_in, const float* moving, float* w_out, float* flow_out, float* moved, int B,
                           int D, int H, int W, float scale, float post, int Cmov, cudaStream_t st, bool* handled);

// A/B knob: SMILE_FUSED_V1=1 runs the first-generation kernel (one voxel per thread, attn_tma.cu)
static inline bool use_v1() {
  static const bool v1 = [] { const char* e = getenv("SMILE_FUSED_V1"); return e != nullptr && e[0] == '1'; }();
  return v1;
}

static inline int grid1d(long long n, int block) {
  long long g = ceil_div_ll(n, block);
  long long cap = (long long)kNumSMs * 64;
  return (int)(g < cap ? g : cap);
}

// bisect knob, read per call: 1 = generic kernels with full-accuracy exp2f (no TMA path, no ex2.approx)
static inline bool attn_exact() {
  const char* e = getenv("SMILE_ATTN_EXACT");
  return e != nullptr && e[0] == '1';
}

int launch_modet_attn(const float* q, const float* k, const float* rpb, float* out, int B, int D, int H, int W, int heads,
                      int hd, float scale, cudaStream_t st) {
  const bool exact = attn_exact();
  if (heads == 1 && hd == 6 && !exact) {
    bool handled = false;
    int rc = use_v1() ? launch_modet_attn_tma(q, k, rpb, nullptr, nullptr, out, nullptr, nullptr, B, D, H, W, scale, 1.0f, 0,
                                              st, &handled)
                      : launch_modet_attn_tma2(q, k, rpb, nullptr, nullptr, nullptr, nullptr, out, nullptr, nullptr, B, D, H,
                                               W, scale, 1.0f, 0, st, &handled);
    if (handled) return rc;
  }
  const long long total = (long long)D * H * W * heads;
  dim3 grid(grid1d(total, 128), B);
  size_t smem = (size_t)heads * 27 * sizeof(float);
  if (exact && hd == 6)
    attn_generic_kernel<6, true><<<grid, 128, smem, st>>>(q, k, rpb, out, D, H, W, heads, hd, scale * kLog2e);
  else if (exact)
    attn_generic_kernel<0, true><<<grid, 128, smem, st>>>(q, k, rpb, out, D, H, W, heads, hd, scale * kLog2e);
  else if (hd == 6)
    attn_generic_kernel<6><<<grid, 128, smem, st>>>(q, k, rpb, out, D, H, W, heads, hd, scale * kLog2e);
  else if (hd == 4)
    attn_generic_kernel<4><<<grid, 128, smem, st>>>(q, k, rpb, out, D, H, W, heads, hd, scale * kLog2e);
  else if (hd == 8)
    attn_generic_kernel<8><<<grid, 128, smem, st>>>(q, k, rpb, out, D, H, W, heads, hd, scale * kLog2e);
  else
    attn_generic_kernel<0><<<grid, 128, smem, st>>>(q, k, rpb, out, D, H, W, heads, hd, scale * kLog2e);
  return check_launch("modet_attn");
}

int launch_modet_fused(const float* q, const float* k, const float* rpb, const float* ln_gamma, const float* ln_beta,
                       const float* flow_in, const float* moving, float* flow_out, float* moved, int B, int D, int H, int W,
                       int hd, float scale, float post, int Cmov, cudaStream_t st) {
  if (hd == 6 && !attn_exact()) {
    bool handled = false;
    int rc = use_v1() ? launch_modet_attn_tma(q, k, rpb, flow_in, moving, nullptr, flow_out, moved, B, D, H, W, scale, post,
                                              Cmov, st, &handled)
                      : launch_modet_attn_tma2(q, k, rpb, ln_gamma, ln_beta, flow_in, moving, nullptr, flow_out, moved, B, D,
                                               H, W, scale, post, Cmov, st, &handled);
    if (handled) return rc;
  }
  const long long N = (long long)D * H * W;
  dim3 grid(grid1d(N, 128), B);
  if (attn_exact())
    fused_generic_kernel<0, true><<<grid, 128, 0, st>>>(q, k, rpb, flow_in, moving, flow_out, moved, D, H, W, hd,
                                                        scale * kLog2e, post, Cmov);
  else if (hd == 6)
    fused_generic_kernel<6><<<grid, 128, 0, st>>>(q, k, rpb, flow_in, moving, flow_out, moved, D, H, W, hd, scale * kLog2e,
                                                  post, Cmov);
  else
    fused_generic_kernel<0><<<grid, 128, 0, st>>>(q, k, rpb, flow_in, moving, flow_out, moved, D, H, W, hd, scale * kLog2e,
                                                  post, Cmov);
  return check_launch("modet_fused");
}

}  // namespace smile
