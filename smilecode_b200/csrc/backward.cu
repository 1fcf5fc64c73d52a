// Backward kernels of the decoder ops (training path; SURVEY.md section 8 rows a2-a7 "+ bwd", A6):
//   warp3d_bwd        SpatialTransformer backward: d_src (scatter) and d_flow        (models.py:49-67)
//   upsample2x_bwd    adjoint of trilinear x2, align_corners=True                     (models.py:354)
//   modet_attn_bwd    ModeTransformer backward: recompute softmax, d_logits, dq, dk, d_rpb (models.py:308-334)
//   proj_ln_bwd       ProjectionLayer backward: LayerNorm + Linear                    (models.py:230-241)
//   cwm_fuse_bwd      softmax-weighted field fusion backward                          (models.py:268-275)
// All are HBM/L2-bound streaming or gather/scatter kernels: one thread per voxel, lanes along W,
// parameter gradients reduced warp -> CTA -> one atomicAdd per CTA and element.
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

inline int grid_for(long long n, int block, int per_sm = 16) {
  long long g = ceil_div_ll(n, block);
  const long long cap = (long long)kNumSMs * per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ------------------------------------------------------------------------------------------------
// warp backward.  out[c,p] = sum_corners src[c,corner] * wx*wy*wz (in-volume corners only).
//   d_src[c,corner] += w * g[c,p]
//   d_coord_x = sum_c g[c,p] * sum_{y,z corners} wy*wz * (src[x1] - src[x0])   (masked), same for y, z
//   d_flow[a,p] = d_coord_a   (the normalise / un-normalise pair of models.py:56 and GridSampler.h:31 has
//                              derivative (S-1)/2 * 2/(S-1) = 1)
// ------------------------------------------------------------------------------------------------
struct Axis {
  int i0, i1;       // clamped corner indices
  float w0, w1;     // weights, 0 when the corner is outside the volume
  float m0, m1;     // 1 / 0 in-volume masks
};
__device__ __forceinline__ Axis axis_of(float c, int S) {
  Axis a;
  const float f = floorf(c);
  const int j0 = __float2int_rd(c), j1 = j0 + 1;
  a.m0 = ((unsigned)j0 < (unsigned)S) ? 1.f : 0.f;
  a.m1 = ((unsigned)j1 < (unsigned)S) ? 1.f : 0.f;
  a.w1 = (c - f) * a.m1;
  a.w0 = ((f + 1.0f) - c) * a.m0;
  a.i0 = min(max(j0, 0), S - 1);
  a.i1 = min(max(j1, 0), S - 1);
  return a;
}

__global__ void __launch_bounds__(256) warp3d_bwd_kernel(const float* __restrict__ g, const float* __restrict__ src,
                                                         const float* __restrict__ flow, float* __restrict__ d_src,
                                                         float* __restrict__ d_flow, int C, int D, int H, int W) {
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y;
  const float* fl = flow + (long long)b * 3 * N;
  const float* sb = src + (long long)b * C * N;
  const float* gb = g + (long long)b * C * N;
  const float dm1 = (float)(D - 1), hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    const Axis az = axis_of(st_coord(d, __ldg(fl + p), dm1), D);
    const Axis ay = axis_of(st_coord(h, __ldg(fl + N + p), hm1), H);
    const Axis ax = axis_of(st_coord(w, __ldg(fl + 2 * N + p), wm1), W);
    const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1}, xi[2] = {ax.i0, ax.i1};
    const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1}, wx[2] = {ax.w0, ax.w1};
    const float mz[2] = {az.m0, az.m1}, my[2] = {ay.m0, ay.m1}, mx[2] = {ax.m0, ax.m1};
    float gz = 0.f, gy = 0.f, gx = 0.f;
    for (int c = 0; c < C; ++c) {
      const float gv = __ldg(gb + (long long)c * N + p);
      const float* sc = sb + (long long)c * N;
      float* dc = d_src ? d_src + ((long long)b * C + c) * N : nullptr;
#pragma unroll
      for (int cz = 0; cz < 2; ++cz)
#pragma unroll
        for (int cy = 0; cy < 2; ++cy)
#pragma unroll
          for (int cx = 0; cx < 2; ++cx) {
            const long long o = ((long long)zi[cz] * H + yi[cy]) * W + xi[cx];
            const float wgt = wx[cx] * wy[cy] * wz[cz];
            if (dc != nullptr && wgt != 0.f) atomicAdd(dc + o, wgt * gv);
            if (d_flow != nullptr) {
              const float v = __ldg(sc + o) * gv;
              const float sx = cx ? mx[1] : -mx[0], sy = cy ? my[1] : -my[0], sz = cz ? mz[1] : -mz[0];
              gx = fmaf(v * sx, wy[cy] * wz[cz], gx);
              gy = fmaf(v * sy, wx[cx] * wz[cz], gy);
              gz = fmaf(v * sz, wx[cx] * wy[cy], gz);
            }
          }
    }
    if (d_flow != nullptr) {
      float* df = d_flow + (long long)b * 3 * N;
      df[p] = gz;
      df[N + p] = gy;
      df[2 * N + p] = gx;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// upsample2x backward: d_x[i] = pre * sum over outputs o that read i of weight(o,i) * g[o]  (scatter)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ g, float* __restrict__ dx, int C,
                                                             int D, int H, int W, float pre) {
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  const long long ON = (long long)OD * OH * OW;
  const long long IN = (long long)D * H * W;
  const int b = blockIdx.y;
  const float rd = up2_ratio(D), rh = up2_ratio(H), rw = up2_ratio(W);
  const float* gb = g + (long long)b * C * ON;
  float* db = dx + (long long)b * C * IN;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < ON; p += (long long)gridDim.x * blockDim.x) {
    const int od = (int)(p / ((long long)OH * OW));
    const int r = (int)(p - (long long)od * OH * OW);
    const int oh = r / OW, ow = r - oh * OW;
    int d0, d1, h0, h1, w0, w1;
    float ld, lh, lw;
    up2_index(od, D, rd, d0, d1, ld);
    up2_index(oh, H, rh, h0, h1, lh);
    up2_index(ow, W, rw, w0, w1, lw);
    const float wd[2] = {1.0f - ld, ld}, wh[2] = {1.0f - lh, lh}, ww[2] = {1.0f - lw, lw};
    const int di[2] = {d0, d1}, hi[2] = {h0, h1}, wi[2] = {w0, w1};
    for (int c = 0; c < C; ++c) {
      const float gv = pre * __ldg(gb + (long long)c * ON + p);
      float* dc = db + (long long)c * IN;
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int e = 0; e < 2; ++e)
#pragma unroll
          for (int f = 0; f < 2; ++f) {
            const float wt = wd[a] * wh[e] * ww[f];
            if (wt != 0.f) atomicAdd(dc + ((long long)di[a] * H + hi[e]) * W + wi[f], wt * gv);
          }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// attention backward, pass 1 (one thread per (voxel, head)): recompute logits + softmax, then
//   gV[t] = sum_a g[h*3+a] * V[t,a];  dl[t] = p[t] * (gV[t] - sum_t' p[t'] gV[t'])
//   dq[d] = scale * sum_t dl[t] * k[n+off(t), d]  (in-volume taps);  drpb[h,t] += dl[t];  dl stored for pass 2.
// pass 2 (one thread per (voxel, head) of the KEY volume): dk[m,d] = scale * sum_t dl[m-off(t)][t] * q[m-off(t), d].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const float* __restrict__ g, const float* __restrict__ q,
                                                          const float* __restrict__ k, const float* __restrict__ rpb,
                                                          float* __restrict__ dq, float* __restrict__ dl_out,
                                                          float* __restrict__ drpb, int D, int H, int W, int heads,
                                                          int hd, float scale) {
  __shared__ float s_part[4][27];
  const int HW = H * W, Cc = heads * hd;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y, head = blockIdx.z;
  const float* qb = q + (long long)b * N * Cc;
  const float* kb = k + (long long)b * N * Cc;
  const float* gb = g + (long long)b * 3 * heads * N;
  float* dqb = dq + (long long)b * N * Cc;
  float* dlb = dl_out + (long long)b * N * heads * 27;
  const bool even = (hd % 2 == 0);  // rows are 8-byte aligned: float2 loads
  float racc[27];  // this thread's share of d_rpb[head, :]
#pragma unroll
  for (int t = 0; t < 27; ++t) racc[t] = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    const float* qr = qb + p * Cc + head * hd;
    float lg[27];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int dd = d + t / 9 - 1, hh = h + (t / 3) % 3 - 1, ww = w + t % 3 - 1;
      float acc = 0.f;
      if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const float* kr = kb + (((long long)dd * H + hh) * W + ww) * Cc + head * hd;
        if (even) {
          for (int c = 0; c < hd; c += 2) {
            const float2 qv = __ldg(reinterpret_cast<const float2*>(qr + c)), kv = __ldg(reinterpret_cast<const float2*>(kr + c));
            acc = fmaf(qv.x, kv.x, fmaf(qv.y, kv.y, acc));
          }
        } else {
          for (int c = 0; c < hd; ++c) acc = fmaf(__ldg(qr + c), __ldg(kr + c), acc);
        }
      }
      lg[t] = acc * scale + (rpb ? __ldg(rpb + head * 27 + t) : 0.f);
      m = fmaxf(m, lg[t]);
    }
    float sum = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      lg[t] = __expf(lg[t] - m);
      sum += lg[t];
    }
    const float inv = 1.0f / sum;
    const float g0 = __ldg(gb + ((long long)head * 3 + 0) * N + p), g1 = __ldg(gb + ((long long)head * 3 + 1) * N + p),
                g2 = __ldg(gb + ((long long)head * 3 + 2) * N + p);
    float dot = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      lg[t] *= inv;  // p[t]
      const float gv = g0 * (float)(t / 9 - 1) + g1 * (float)((t / 3) % 3 - 1) + g2 * (float)(t % 3 - 1);
      dot = fmaf(lg[t], gv, dot);
    }
    float* dlr = dlb + (long long)head * 27 * N + p;   // [heads][27][N]: coalesced here and in the dk gather
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float gv = g0 * (float)(t / 9 - 1) + g1 * (float)((t / 3) % 3 - 1) + g2 * (float)(t % 3 - 1);
      lg[t] = lg[t] * (gv - dot);  // d_logit
      dlr[(long long)t * N] = lg[t];
      racc[t] += lg[t];
    }
    if (even) {
      for (int c = 0; c < hd; c += 2) {
        float a0 = 0.f, a1 = 0.f;
#pragma unroll
        for (int t = 0; t < 27; ++t) {
          const int dd = d + t / 9 - 1, hh = h + (t / 3) % 3 - 1, ww = w + t % 3 - 1;
          if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) {
            const float2 kv = __ldg(reinterpret_cast<const float2*>(kb + (((long long)dd * H + hh) * W + ww) * Cc + head * hd + c));
            a0 = fmaf(lg[t], kv.x, a0);
            a1 = fmaf(lg[t], kv.y, a1);
          }
        }
        *reinterpret_cast<float2*>(dqb + p * Cc + head * hd + c) = make_float2(a0 * scale, a1 * scale);
      }
    } else {
      for (int c = 0; c < hd; ++c) {
        float acc = 0.f;
#pragma unroll
        for (int t = 0; t < 27; ++t) {
          const int dd = d + t / 9 - 1, hh = h + (t / 3) % 3 - 1, ww = w + t % 3 - 1;
          if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W)
            acc = fmaf(lg[t], __ldg(kb + (((long long)dd * H + hh) * W + ww) * Cc + head * hd + c), acc);
        }
        dqb[p * Cc + head * hd + c] = acc * scale;
      }
    }
  }
  if (drpb != nullptr) {
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float v = warp_sum(racc[t]);
      if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5][t] = v;
    }
    __syncthreads();
    if (threadIdx.x < 27)
      atomicAdd(drpb + head * 27 + threadIdx.x,
                s_part[0][threadIdx.x] + s_part[1][threadIdx.x] + s_part[2][threadIdx.x] + s_part[3][threadIdx.x]);
  }
}

__global__ void __launch_bounds__(128) attn_bwd_dk_kernel(const float* __restrict__ dl, const float* __restrict__ q,
                                                          float* __restrict__ dk, int D, int H, int W, int heads, int hd,
                                                          float scale) {
  const int HW = H * W, Cc = heads * hd;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y;
  const float* qb = q + (long long)b * N * Cc;
  const float* dlb = dl + (long long)b * N * heads * 27;
  float* dkb = dk + (long long)b * N * Cc;
  const long long total = N * heads;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i / heads;
    const int head = (int)(i - p * heads);
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    float acc[16];
    for (int c = 0; c < hd; ++c) acc[c] = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      // query voxel n whose tap t is this key voxel: n = m - off(t)
      const int dd = d - (t / 9 - 1), hh = h - ((t / 3) % 3 - 1), ww = w - (t % 3 - 1);
      if (dd >= 0 && dd < D && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const long long n = ((long long)dd * H + hh) * W + ww;
        const float gv = __ldg(dlb + ((long long)head * 27 + t) * N + n);
        const float* qr = qb + n * Cc + head * hd;
        if (hd % 2 == 0) {
          for (int c = 0; c < hd; c += 2) {
            const float2 qv = __ldg(reinterpret_cast<const float2*>(qr + c));
            acc[c] = fmaf(gv, qv.x, acc[c]);
            acc[c + 1] = fmaf(gv, qv.y, acc[c + 1]);
          }
        } else {
          for (int c = 0; c < hd; ++c) acc[c] = fmaf(gv, __ldg(qr + c), acc[c]);
        }
      }
    }
    for (int c = 0; c < hd; ++c) dkb[p * Cc + head * hd + c] = acc[c] * scale;
  }
}

// heads == 1, head_dim == 6 (the two large levels): accumulators in registers, 32-bit indices, taps fully unrolled.
// (The generic kernel above indexes acc[] with a run-time head_dim: local memory; 0.46 ms per call at 160x192x160.)
__global__ void __launch_bounds__(128) attn_bwd_dk6_kernel(const float* __restrict__ dl, const float* __restrict__ q,
                                                           float* __restrict__ dk, int D, int H, int W, float scale) {
  const int HW = H * W, N = D * HW;
  const int b = blockIdx.y;
  const float* qb = q + (long long)b * N * 6;
  const float* dlb = dl + (long long)b * N * 27;
  float* dkb = dk + (long long)b * N * 6;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < N; p += gridDim.x * blockDim.x) {
    const int d = p / HW;
    const int r = p - d * HW;
    const int h = r / W, w = r - h * W;
    float2 a0 = make_float2(0.f, 0.f), a1 = a0, a2 = a0;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      // query voxel n whose tap t is this key voxel: n = m - off(t)
      const int od = t / 9 - 1, oh = (t / 3) % 3 - 1, ow = t % 3 - 1;
      const int dd = d - od, hh = h - oh, ww = w - ow;
      if ((unsigned)dd < (unsigned)D && (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W) {
        const int n = p - (od * HW + oh * W + ow);
        const float gv = __ldg(dlb + t * N + n);
        const float2* qr = reinterpret_cast<const float2*>(qb + (long long)n * 6);
        const float2 q0 = __ldg(qr), q1 = __ldg(qr + 1), q2 = __ldg(qr + 2);
        a0.x = fmaf(gv, q0.x, a0.x); a0.y = fmaf(gv, q0.y, a0.y);
        a1.x = fmaf(gv, q1.x, a1.x); a1.y = fmaf(gv, q1.y, a1.y);
        a2.x = fmaf(gv, q2.x, a2.x); a2.y = fmaf(gv, q2.y, a2.y);
      }
    }
    float2* o = reinterpret_cast<float2*>(dkb + (long long)p * 6);
    o[0] = make_float2(a0.x * scale, a0.y * scale);
    o[1] = make_float2(a1.x * scale, a1.y * scale);
    o[2] = make_float2(a2.x * scale, a2.y * scale);
  }
}

// ------------------------------------------------------------------------------------------------
// ProjectionLayer backward.  z = W x + b; y = gamma * (z - mean) * rstd + beta.
// ------------------------------------------------------------------------------------------------
// Phase A (thread = voxel of a 128-voxel tile): recompute z and the LayerNorm statistics, form dz and the
// per-voxel contributions to d_gamma / d_beta, park them (and the x tile) in shared memory, write d_feat.
// Phase B (thread = output element): every parameter-gradient element sums its 128 products from shared
// memory into a register that lives across the CTA's grid-stride loop; one atomicAdd per element and CTA at
// the end.  No warp shuffles, no per-voxel atomics.
template <int C>
__global__ void __launch_bounds__(128) proj_ln_bwd_kernel(const float* __restrict__ gout, const float* __restrict__ feat,
                                                          const float* __restrict__ weight, const float* __restrict__ bias,
                                                          const float* __restrict__ gamma, float* __restrict__ dfeat,
                                                          float* __restrict__ dweight, float* __restrict__ dbias,
                                                          float* __restrict__ dgamma, float* __restrict__ dbeta, int Cin,
                                                          long long N, float eps) {
  constexpr int TV = 128, XP = TV + 1;
  extern __shared__ float smem[];
  float* s_w = smem;                   // [Cin][C]
  float* s_b = s_w + Cin * C;          // bias, gamma
  float* s_dz = s_b + 2 * C;           // [TV][C]
  float* s_gx = s_dz + TV * C;         // [TV][C]  gy * xhat
  float* s_gy = s_gx + TV * C;         // [TV][C]  gy
  float* s_x = s_gy + TV * C;          // [Cin][XP]
  for (int i = threadIdx.x; i < Cin * C; i += blockDim.x) {
    const int ci = i / C, c = i - ci * C;
    s_w[i] = weight[c * Cin + ci];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_b[i] = bias[i];
    s_b[C + i] = gamma[i];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* fb = feat + (long long)b * Cin * N;
  const float* gb = gout + (long long)b * N * C;
  float* dfb = dfeat ? dfeat + (long long)b * Cin * N : nullptr;
  const int nout = Cin * C;
  constexpr int MAXO = 48;             // outputs per thread: Cin*C <= 128*48 (128 x 48 at the coarsest level)
  float wacc[MAXO];
#pragma unroll
  for (int i = 0; i < MAXO; ++i) wacc[i] = 0.f;
  float vacc[3] = {0.f, 0.f, 0.f};     // d_bias / d_gamma / d_beta element of thread c < C
  for (long long p0 = (long long)blockIdx.x * TV; p0 < N; p0 += (long long)gridDim.x * TV) {
    const long long p = p0 + threadIdx.x;
    const bool ok = p < N;
    float z[C], dz[C];
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = s_b[c];
    for (int ci = 0; ci < Cin; ++ci) {
      const float x = ok ? __ldg(fb + (long long)ci * N + p) : 0.f;
      s_x[ci * XP + threadIdx.x] = x;
#pragma unroll
      for (int c = 0; c < C; ++c) z[c] = fmaf(x, s_w[ci * C + c], z[c]);
    }
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mean += z[c];
    mean *= (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) var = fmaf(z[c] - mean, z[c] - mean, var);
    const float rstd = rsqrtf(var * (1.0f / C) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float gy = ok ? __ldg(gb + p * C + c) : 0.f;
      const float xh = (z[c] - mean) * rstd;
      s_gx[threadIdx.x * C + c] = gy * xh;
      s_gy[threadIdx.x * C + c] = gy;
      const float dxh = gy * s_b[C + c];
      dz[c] = dxh;
      z[c] = xh;
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
    s1 *= (1.0f / C);
    s2 *= (1.0f / C);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      dz[c] = ok ? rstd * (dz[c] - s1 - z[c] * s2) : 0.f;
      s_dz[threadIdx.x * C + c] = dz[c];
    }
    if (ok && dfb != nullptr)
      for (int ci = 0; ci < Cin; ++ci) {
        float dxv = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) dxv = fmaf(dz[c], s_w[ci * C + c], dxv);
        dfb[(long long)ci * N + p] = dxv;
      }
    __syncthreads();
    // phase B (when Cin*C < 128 the 128 voxels are split over 128 / (Cin*C) thread groups)
    if (nout >= 128) {
#pragma unroll
      for (int i = 0; i < MAXO; ++i) {
        const int o = threadIdx.x + i * 128;
        if (o < nout) {
          const int ci = o / C, c = o - ci * C;
          float a = wacc[i];
          for (int v = 0; v < TV; ++v) a = fmaf(s_dz[v * C + c], s_x[ci * XP + v], a);
          wacc[i] = a;
        }
      }
    } else {
      const int ngroups = 128 / nout;
      const int o = threadIdx.x % nout, grp = threadIdx.x / nout;
      if (grp < ngroups) {
        const int ci = o / C, c = o - ci * C;
        float a = wacc[0];
        for (int v = grp; v < TV; v += ngroups) a = fmaf(s_dz[v * C + c], s_x[ci * XP + v], a);
        wacc[0] = a;
      }
    }
    if (threadIdx.x < C) {
      const int c = threadIdx.x;
      for (int v = 0; v < TV; ++v) {
        vacc[0] += s_dz[v * C + c];
        vacc[1] += s_gx[v * C + c];
        vacc[2] += s_gy[v * C + c];
      }
    }
    __syncthreads();
  }
  if (nout >= 128) {
#pragma unroll
    for (int i = 0; i < MAXO; ++i) {
      const int o = threadIdx.x + i * 128;
      if (o < nout) {
        const int ci = o / C, c = o - ci * C;
        atomicAdd(dweight + c * Cin + ci, wacc[i]);
      }
    }
  } else if ((int)threadIdx.x / nout < 128 / nout) {
    const int o = threadIdx.x % nout;
    const int ci = o / C, c = o - ci * C;
    atomicAdd(dweight + c * Cin + ci, wacc[0]);
  }
  if (threadIdx.x < C) {
    atomicAdd(dbias + threadIdx.x, vacc[0]);
    atomicAdd(dgamma + threadIdx.x, vacc[1]);
    atomicAdd(dbeta + threadIdx.x, vacc[2]);
  }
}

// ------------------------------------------------------------------------------------------------
// CWM tail backward: out[a] = 2 * sum_f u[3f+a] * p[f],  p = softmax_f(logits)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cwm_fuse_bwd_kernel(const float* __restrict__ g, const float* __restrict__ fields,
                                                           const float* __restrict__ logits, float* __restrict__ dfields,
                                                           float* __restrict__ dlogits, int F, long long N) {
  const int b = blockIdx.y;
  const float* fb = fields + (long long)b * 3 * F * N;
  const float* lb = logits + (long long)b * F * N;
  const float* gb = g + (long long)b * 3 * N;
  float* dfb = dfields + (long long)b * 3 * F * N;
  float* dlb = dlogits + (long long)b * F * N;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    float m = -INFINITY;
    for (int f = 0; f < F; ++f) m = fmaxf(m, __ldg(lb + (long long)f * N + p));
    float sum = 0.f;
    for (int f = 0; f < F; ++f) sum += expf(__ldg(lb + (long long)f * N + p) - m);
    const float inv = 1.0f / sum;
    const float g0 = 2.f * __ldg(gb + p), g1 = 2.f * __ldg(gb + N + p), g2 = 2.f * __ldg(gb + 2 * N + p);
    float dot = 0.f;
    for (int f = 0; f < F; ++f) {
      const float pf = expf(__ldg(lb + (long long)f * N + p) - m) * inv;
      const float t = g0 * __ldg(fb + (long long)(3 * f) * N + p) + g1 * __ldg(fb + (long long)(3 * f + 1) * N + p) +
                      g2 * __ldg(fb + (long long)(3 * f + 2) * N + p);
      dot = fmaf(pf, t, dot);
    }
    for (int f = 0; f < F; ++f) {
      const float pf = expf(__ldg(lb + (long long)f * N + p) - m) * inv;
      const float t = g0 * __ldg(fb + (long long)(3 * f) * N + p) + g1 * __ldg(fb + (long long)(3 * f + 1) * N + p) +
                      g2 * __ldg(fb + (long long)(3 * f + 2) * N + p);
      dlb[(long long)f * N + p] = pf * (t - dot);
      dfb[(long long)(3 * f) * N + p] = g0 * pf;
      dfb[(long long)(3 * f + 1) * N + p] = g1 * pf;
      dfb[(long long)(3 * f + 2) * N + p] = g2 * pf;
    }
  }
}

// The same with warp-aggregated scatter (volumes below 2^31 voxels).  Lanes run along W, and with a smooth flow the x1
// corner of lane L is the x0 corner of lane L+1 in the same (z, y) row: lane L adds its neighbour's x0 contribution to
// its own x1 contribution and issues ONE atomic for the pair, the neighbour skips its x0 atomic.  Equality of the linear
// indices is all that is tested, so row ends, clamped corners and rough flows just fall back to separate atomics.  Half
// the atomics of the kernel above where it matters (64 per voxel at 8 channels: 1.3 ms at 160x192x160).
__global__ void __launch_bounds__(256) warp3d_bwd_agg_kernel(const float* __restrict__ g, const float* __restrict__ src,
                                                             const float* __restrict__ flow, float* __restrict__ d_src,
                                                             float* __restrict__ d_flow, int C, int D, int H, int W) {
  const int HW = H * W;
  const int N = D * HW;
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31;
  const float* fl = flow + (long long)b * 3 * N;
  const float* sb = src + (long long)b * C * N;
  const float* gb = g + (long long)b * C * N;
  const float dm1 = (float)(D - 1), hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + (threadIdx.x & ~31); base < N; base += stride) {
    const bool valid = base + lane < N;
    const int p = valid ? (int)(base + lane) : N - 1;
    const int d = p / HW;
    const int r = p - d * HW;
    const int h = r / W, w = r - h * W;
    const Axis az = axis_of(st_coord(d, __ldg(fl + p), dm1), D);
    const Axis ay = axis_of(st_coord(h, __ldg(fl + N + p), hm1), H);
    const Axis ax = axis_of(st_coord(w, __ldg(fl + 2 * N + p), wm1), W);
    const int zi[2] = {az.i0, az.i1}, yi[2] = {ay.i0, ay.i1};
    const float wz[2] = {az.w0, az.w1}, wy[2] = {ay.w0, ay.w1}, wx[2] = {ax.w0, ax.w1};
    const float mz[2] = {az.m0, az.m1}, my[2] = {ay.m0, ay.m1}, mx[2] = {ax.m0, ax.m1};
    // which of this lane's four (z, y) rows continue into the next lane's row, and which are continued by the previous
    bool merge[4], absorbed[4];
    int o0[4], o1[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = (zi[q >> 1] * H + yi[q & 1]) * W;
      o0[q] = row + ax.i0;
      o1[q] = row + ax.i1;
      const int o0n = __shfl_down_sync(0xffffffffu, o0[q], 1), o1p = __shfl_up_sync(0xffffffffu, o1[q], 1);
      merge[q] = lane < 31 && o0n == o1[q];
      absorbed[q] = lane > 0 && o1p == o0[q];
    }
    float gz = 0.f, gy = 0.f, gx = 0.f;
    for (int c = 0; c < C; ++c) {
      const float gv = valid ? __ldg(gb + (long long)c * N + p) : 0.f;
      const float* sc = sb + (long long)c * N;
      float* dc = d_src ? d_src + ((long long)b * C + c) * N : nullptr;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int cz = q >> 1, cy = q & 1;
        if (dc != nullptr) {     // uniform
          const float a = (wx[0] * wy[cy] * wz[cz]) * gv;
          float bb = (wx[1] * wy[cy] * wz[cz]) * gv;
          const float an = __shfl_down_sync(0xffffffffu, a, 1);
          if (merge[q]) bb += an;
          if (bb != 0.f) atomicAdd(dc + o1[q], bb);
          if (!absorbed[q] && a != 0.f) atomicAdd(dc + o0[q], a);
        }
        if (d_flow != nullptr) {
          const float v0 = __ldg(sc + o0[q]) * gv, v1 = __ldg(sc + o1[q]) * gv;
          const float sy = cy ? my[1] : -my[0], sz = cz ? mz[1] : -mz[0];
          gx = fmaf(v0 * -mx[0], wy[cy] * wz[cz], gx);
          gy = fmaf(v0 * sy, wx[0] * wz[cz], gy);
          gz = fmaf(v0 * sz, wx[0] * wy[cy], gz);
          gx = fmaf(v1 * mx[1], wy[cy] * wz[cz], gx);
          gy = fmaf(v1 * sy, wx[1] * wz[cz], gy);
          gz = fmaf(v1 * sz, wx[1] * wy[cy], gz);
        }
      }
    }
    if (d_flow != nullptr && valid) {
      float* df = d_flow + (long long)b * 3 * N;
      df[p] = gz;
      df[N + p] = gy;
      df[2 * N + p] = gx;
    }
  }
}

}  // namespace

int launch_warp3d_bwd(const float* g, const float* src, const float* flow, float* d_src, float* d_flow, int B, int C, int D,
                      int H, int W, cudaStream_t st) {
  const long long N = (long long)D * H * W;
  if (d_src != nullptr) {
    cudaError_t e = cudaMemsetAsync(d_src, 0, (size_t)B * C * N * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("warp3d_bwd: memset failed: %s", cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
  }
  static const bool no_agg = getenv("SMILE_WARP_BWD_PLAIN") != nullptr;   // A/B knob
  if (N < (1LL << 31) && !no_agg)
    warp3d_bwd_agg_kernel<<<dim3(grid_for(N, 256, 32), B), 256, 0, st>>>(g, src, flow, d_src, d_flow, C, D, H, W);
  else
    warp3d_bwd_kernel<<<dim3(grid_for(N, 256, 32), B), 256, 0, st>>>(g, src, flow, d_src, d_flow, C, D, H, W);
  return check_launch("warp3d_bwd");
}

int launch_upsample2x_bwd(const float* g, float* dx, int B, int C, int D, int H, int W, float pre, cudaStream_t st) {
  const long long IN = (long long)D * H * W;
  cudaError_t e = cudaMemsetAsync(dx, 0, (size_t)B * C * IN * sizeof(float), st);
  if (e != cudaSuccess) {
    set_error("upsample2x_bwd: memset failed: %s", cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  upsample2x_bwd_kernel<<<dim3(grid_for(8 * IN, 256, 32), B), 256, 0, st>>>(g, dx, C, D, H, W, pre);
  return check_launch("upsample2x_bwd");
}

int launch_attn_bwd_dq_tma(const float* g, const float* q, const float* k, const float* rpb, float* dq, float* dl,
                           float* drpb, int B, int D, int H, int W, float scale, cudaStream_t st, bool* handled);

int launch_modet_attn_bwd(const float* g, const float* q, const float* k, const float* rpb, float* dq, float* dk,
                          float* drpb, float* dl_work, int B, int D, int H, int W, int heads, int hd, float scale,
                          cudaStream_t st) {
  if (hd > 16) {
    set_error("modet_attn_bwd: head_dim %d > 16 is not supported", hd);
    return SMILE_ERR_UNSUPPORTED;
  }
  const long long total = (long long)D * H * W * heads;
  if (drpb != nullptr) {
    cudaError_t e = cudaMemsetAsync(drpb, 0, (size_t)heads * 27 * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("modet_attn_bwd: memset failed: %s", cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
  }
  dim3 grid(grid_for(total, 128, 32), B);
  int rc = SMILE_OK;
  bool handled = false;
  if (heads == 1 && hd == 6) {
    rc = launch_attn_bwd_dq_tma(g, q, k, rpb, dq, dl_work, drpb, B, D, H, W, scale, st, &handled);
    if (rc) return rc;
  }
  if (!handled) {
    dim3 grid_q(grid_for((long long)D * H * W, 128, 16), B, heads);
    attn_bwd_dq_kernel<<<grid_q, 128, 0, st>>>(g, q, k, rpb, dq, dl_work, drpb, D, H, W, heads, hd, scale);
    rc = check_launch("modet_attn_bwd(dq)");
    if (rc) return rc;
  }
  if (heads == 1 && hd == 6 && (long long)D * H * W * 27 < (1LL << 31))
    attn_bwd_dk6_kernel<<<grid, 128, 0, st>>>(dl_work, q, dk, D, H, W, scale);
  else
    attn_bwd_dk_kernel<<<grid, 128, 0, st>>>(dl_work, q, dk, D, H, W, heads, hd, scale);
  return check_launch("modet_attn_bwd(dk)");
}

// Narrow projections (the two fine levels: C = 6 from Cin = 8 / 16 -- 87 % of the voxels).  The generic kernel above
// round-trips dz / gy through shared memory and reduces the parameter gradients with a serial loop per 128-voxel tile
// (six threads walk 128 voxels while the CTA waits: 1.06 ms per call at 160x192x160).  Here every thread keeps its partial
// d_weight [CIN x C], d_bias, d_gamma, d_beta in registers across all its voxels and the CTA reduces ONCE at the end
// (shuffles -> shared memory -> one atomic per element); the loop has no barrier and no shared-memory traffic.
template <int C, int CIN>
__global__ void __launch_bounds__(128) proj_ln_bwd_small_kernel(const float* __restrict__ gout, const float* __restrict__ feat,
                                                                const float* __restrict__ weight, const float* __restrict__ bias,
                                                                const float* __restrict__ gamma, float* __restrict__ dfeat,
                                                                float* __restrict__ dweight, float* __restrict__ dbias,
                                                                float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                                long long N, float eps) {
  __shared__ float s_w[CIN * C], s_b[2 * C];
  __shared__ float s_red[4][CIN * C + 3 * C];
  for (int i = threadIdx.x; i < CIN * C; i += blockDim.x) {
    const int ci = i / C, c = i - ci * C;
    s_w[i] = weight[c * CIN + ci];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_b[i] = bias[i];
    s_b[C + i] = gamma[i];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* fb = feat + (long long)b * CIN * N;
  const float* gb = gout + (long long)b * N * C;
  float* dfb = dfeat ? dfeat + (long long)b * CIN * N : nullptr;
  float wacc[CIN * C], vb[C], vg[C], vbeta[C];
#pragma unroll
  for (int i = 0; i < CIN * C; ++i) wacc[i] = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) vb[c] = vg[c] = vbeta[c] = 0.f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    float x[CIN], z[C], dz[C];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) x[ci] = __ldg(fb + (long long)ci * N + p);
#pragma unroll
    for (int c = 0; c < C; ++c) z[c] = s_b[c];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int c = 0; c < C; ++c) z[c] = fmaf(x[ci], s_w[ci * C + c], z[c]);
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mean += z[c];
    mean *= (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) var = fmaf(z[c] - mean, z[c] - mean, var);
    const float rstd = rsqrtf(var * (1.0f / C) + eps);
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float gy = __ldg(gb + p * C + c);
      const float xh = (z[c] - mean) * rstd;
      vg[c] = fmaf(gy, xh, vg[c]);
      vbeta[c] += gy;
      const float dxh = gy * s_b[C + c];
      dz[c] = dxh;
      z[c] = xh;
      s1 += dxh;
      s2 = fmaf(dxh, xh, s2);
    }
    s1 *= (1.0f / C);
    s2 *= (1.0f / C);
#pragma unroll
    for (int c = 0; c < C; ++c) {
      dz[c] = rstd * (dz[c] - s1 - z[c] * s2);
      vb[c] += dz[c];
    }
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci) {
      float dxv = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        dxv = fmaf(dz[c], s_w[ci * C + c], dxv);
        wacc[ci * C + c] = fmaf(dz[c], x[ci], wacc[ci * C + c]);
      }
      if (dfb != nullptr) dfb[(long long)ci * N + p] = dxv;
    }
  }
  // one reduction per CTA
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  auto wsum = [](float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
  };
#pragma unroll
  for (int i = 0; i < CIN * C; ++i) {
    const float v = wsum(wacc[i]);
    if (lane == 0) s_red[warp][i] = v;
  }
#pragma unroll
  for (int c = 0; c < C; ++c) {
    const float a = wsum(vb[c]), g2 = wsum(vg[c]), be = wsum(vbeta[c]);
    if (lane == 0) {
      s_red[warp][CIN * C + c] = a;
      s_red[warp][CIN * C + C + c] = g2;
      s_red[warp][CIN * C + 2 * C + c] = be;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < CIN * C + 3 * C; i += blockDim.x) {
    const float v = s_red[0][i] + s_red[1][i] + s_red[2][i] + s_red[3][i];
    if (i < CIN * C) {
      const int ci = i / C, c = i - ci * C;
      atomicAdd(dweight + c * CIN + ci, v);
    } else if (i < CIN * C + C) {
      atomicAdd(dbias + (i - CIN * C), v);
    } else if (i < CIN * C + 2 * C) {
      atomicAdd(dgamma + (i - CIN * C - C), v);
    } else {
      atomicAdd(dbeta + (i - CIN * C - 2 * C), v);
    }
  }
}

template <int C>
static int launch_pl_bwd(const float* gout, const float* feat, const float* weight, const float* bias, const float* gamma,
                         float* dfeat, float* dweight, float* dbias, float* dgamma, float* dbeta, int B, int Cin, long long N,
                         float eps, cudaStream_t st) {
  if ((long long)Cin * C > 128LL * 48) {
    set_error("proj_ln_bwd: Cin*C = %d exceeds the compiled limit %d", Cin * C, 128 * 48);
    return SMILE_ERR_UNSUPPORTED;
  }
  const size_t smem = (size_t)(Cin * C + 2 * C + 3 * 128 * C + Cin * 129) * sizeof(float);
  auto kern = proj_ln_bwd_kernel<C>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("proj_ln_bwd: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
  }
  cudaMemsetAsync(dweight, 0, (size_t)Cin * C * sizeof(float), st);
  cudaMemsetAsync(dbias, 0, C * sizeof(float), st);
  cudaMemsetAsync(dgamma, 0, C * sizeof(float), st);
  cudaMemsetAsync(dbeta, 0, C * sizeof(float), st);
  if constexpr (C == 6) {
   if ((Cin == 8 || Cin == 16) && N >= 65536) {
    dim3 sgrid(grid_for(N, 128, 8), B);
    if (Cin == 8)
      proj_ln_bwd_small_kernel<C, 8><<<sgrid, 128, 0, st>>>(gout, feat, weight, bias, gamma, dfeat, dweight, dbias, dgamma,
                                                            dbeta, N, eps);
    else
      proj_ln_bwd_small_kernel<C, 16><<<sgrid, 128, 0, st>>>(gout, feat, weight, bias, gamma, dfeat, dweight, dbias, dgamma,
                                                             dbeta, N, eps);
    return check_launch("proj_ln_bwd(small)");
   }
  }
  dim3 grid(grid_for(N, 128, 8), B);
  kern<<<grid, 128, smem, st>>>(gout, feat, weight, bias, gamma, dfeat, dweight, dbias, dgamma, dbeta, Cin, N, eps);
  return check_launch("proj_ln_bwd");
}

int launch_proj_ln_bwd(const float* gout, const float* feat, const float* weight, const float* bias, const float* gamma,
                       float* dfeat, float* dweight, float* dbias, float* dgamma, float* dbeta, int B, int Cin, int C,
                       long long N, float eps, cudaStream_t st) {
#define SMILE_PLB(CC) \
  case CC:            \
    return launch_pl_bwd<CC>(gout, feat, weight, bias, gamma, dfeat, dweight, dbias, dgamma, dbeta, B, Cin, N, eps, st);
  switch (C) {
    SMILE_PLB(4)
    SMILE_PLB(6)
    SMILE_PLB(8)
    SMILE_PLB(12)
    SMILE_PLB(16)
    SMILE_PLB(18)
    SMILE_PLB(24)
    SMILE_PLB(30)
    SMILE_PLB(32)
    SMILE_PLB(36)
    SMILE_PLB(42)
    SMILE_PLB(48)
    default:
      set_error("proj_ln_bwd: projection width C=%d is not compiled in", C);
      return SMILE_ERR_UNSUPPORTED;
  }
#undef SMILE_PLB
}

int launch_cwm_fuse_bwd(const float* g, const float* fields, const float* logits, float* dfields, float* dlogits, int B,
                        int F, long long N, cudaStream_t st) {
  cwm_fuse_bwd_kernel<<<dim3(grid_for(N, 256, 16), B), 256, 0, st>>>(g, fields, logits, dfields, dlogits, F, N);
  return check_launch("cwm_fuse_bwd");
}

}  // namespace smile
