// Backward of the conv / InstanceNorm / LeakyReLU / AvgPool blocks (Encoder and CWM, reference
// ModeT/models.py:119-151, 186-228, 250-254), the losses (losses.py:6-95) and the Adam(amsgrad) update
// (train.py:101).  Training path, first version: correct and deterministic where cheap, tuned later.
//
//   conv3d dgrad   = smile_conv3d_fwd on the gradient with flipped, transposed weights (conv3d_flip_weights)
//   conv3d wgrad   = conv3d_wgrad_kernel: lanes along W, each thread marches a depth chunk accumulating
//                    4 output channels x 27 taps for one input channel in registers, then warp -> atomics
//   IN+LReLU bwd   = two passes: per-(b,c) sums S1 = sum(dn), S2 = sum(dn * n) then
//                    dy = rstd * (dn - S1/N - n * S2/N), with n recovered from the stored activation
//                    (LeakyReLU is invertible) and dn = da * lrelu'(n)
#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

inline int grid_for(long long n, int block, int per_sm = 16) {
  long long g = ceil_div_ll(n, block);
  const long long cap = (long long)kNumSMs * per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// wT[ci][co][26 - t] = w[co][ci][t]
__global__ void flip_weights_kernel(const float* __restrict__ w, float* __restrict__ wT, int Cout, int Cin) {
  const int total = Cout * Cin * 27;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int t = i % 27;
    const int r = i / 27;
    const int ci = r % Cin, co = r / Cin;
    wT[((long long)ci * Cout + co) * 27 + (26 - t)] = w[i];
  }
}

// d_weight[co][ci][t] += sum_{b,v} dy[b,co,v] * x[b,ci,v + off(t)];  d_bias[co] += sum dy (ci == 0 CTAs)
//
// CTA = 4 warps = 4 rows x 32 columns of the (H, W) plane, marching a depth chunk; blockIdx.y = (group of 4
// output channels, input channel).  The input neighbourhood comes from a 4-plane shared-memory ring of
// (4+2) x (32+2) halo tiles, zero filled outside the volume, so the 27 taps are immediate-offset LDS with no
// bounds logic; the 4 x 27 accumulators of a thread are packed channel pairs (fma.rn.f32x2).
constexpr int WG_CO = 4;
constexpr int WG_PW = 36;            // ring row pitch (34 used)
constexpr int WG_PLANE = 6 * WG_PW;  // floats per ring plane
__global__ void __launch_bounds__(128) conv3d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dw, float* __restrict__ db, int B, int Cin,
                                                           int Cout, int D, int H, int W, int dchunk, int tiles_h,
                                                           int tiles_w) {
  __shared__ float s_x[4][WG_PLANE];
  __shared__ float s_part[4][WG_CO * 27 + WG_CO];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int t = blockIdx.x;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  t /= tiles_h;
  const int dc = t;
  const int b = blockIdx.z;
  const int cog = blockIdx.y / Cin, ci = blockIdx.y % Cin;
  const int co0 = cog * WG_CO;
  const int h0 = th * 4, w0 = tw * 32;
  const int h = h0 + warp, w = w0 + lane;
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const bool inside = (h < H) && (w < W);
  const float* xb = x + ((long long)b * Cin + ci) * N;
  const float* dyb = dy + ((long long)b * Cout + co0) * N;
  float2 acc[WG_CO / 2][27];
  float bsum[WG_CO];
#pragma unroll
  for (int c = 0; c < WG_CO; ++c) bsum[c] = 0.f;
#pragma unroll
  for (int c = 0; c < WG_CO / 2; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[c][k] = make_float2(0.f, 0.f);
  const int d_begin = dc * dchunk, d_end = min(D, d_begin + dchunk);

  // cooperative load of one halo plane (6 x 34 = 204 values, <= 2 per thread) into ring slot (dd & 3):
  // global loads are issued early into registers (fetch) and parked in shared memory after the math (commit)
  const int e0 = tid, e1 = tid + 128;
  const int r0 = e0 / 34, c0 = e0 - r0 * 34, r1 = e1 / 34, c1 = e1 - r1 * 34;
  const bool ok0 = (h0 - 1 + r0 >= 0) && (h0 - 1 + r0 < H) && (w0 - 1 + c0 >= 0) && (w0 - 1 + c0 < W);
  const bool ok1 = (e1 < 204) && (h0 - 1 + r1 >= 0) && (h0 - 1 + r1 < H) && (w0 - 1 + c1 >= 0) && (w0 - 1 + c1 < W);
  const long long o0 = (long long)(h0 - 1 + r0) * W + (w0 - 1 + c0), o1 = (long long)(h0 - 1 + r1) * W + (w0 - 1 + c1);
  const int s0 = r0 * WG_PW + c0, s1 = r1 * WG_PW + c1;
  float pre0 = 0.f, pre1 = 0.f;
  auto fetch = [&](int dd) {
    const bool okd = dd >= 0 && dd < D;
    pre0 = (okd && ok0) ? __ldg(xb + (long long)dd * HW + o0) : 0.f;
    pre1 = (okd && ok1) ? __ldg(xb + (long long)dd * HW + o1) : 0.f;
  };
  auto commit = [&](int dd) {
    float* dst = s_x[dd & 3];
    dst[s0] = pre0;
    if (e1 < 204) dst[s1] = pre1;
  };
  fetch(d_begin - 1);
  commit(d_begin - 1);
  fetch(d_begin);
  commit(d_begin);
  fetch(d_begin + 1);
  commit(d_begin + 1);
  __syncthreads();
  const int toff = warp * WG_PW + lane;  // tile position of tap (kh=0, kw=0)
  // the output gradients are streamed from DRAM exactly once: fetch them one depth ahead as well
  float g[WG_CO], gn[WG_CO];
  auto fetch_g = [&](int dd, float (&dst)[WG_CO]) {
#pragma unroll
    for (int c = 0; c < WG_CO; ++c)
      dst[c] = (inside && dd < d_end && co0 + c < Cout) ? __ldg(dyb + (long long)c * N + (long long)dd * HW + h * W + w) : 0.f;
  };
  fetch_g(d_begin, g);
  for (int d = d_begin; d < d_end; ++d) {
    fetch(d + 2);  // lands in slot (d+2)&3 == (d-2)&3, which nobody reads in this iteration
    fetch_g(d + 1, gn);
#pragma unroll
    for (int c = 0; c < WG_CO; ++c) bsum[c] += g[c];
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float* pl = s_x[(d - 1 + kd) & 3] + toff;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float xv = pl[(j / 3) * WG_PW + (j % 3)];
        const float2 xb2 = make_float2(xv, xv);
#pragma unroll
        for (int c = 0; c < WG_CO / 2; ++c)
          acc[c][kd * 9 + j] = fma2(xb2, make_float2(g[2 * c], g[2 * c + 1]), acc[c][kd * 9 + j]);
      }
    }
    commit(d + 2);
#pragma unroll
    for (int c = 0; c < WG_CO; ++c) g[c] = gn[c];
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < WG_CO; ++c) {
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float v = warp_sum((c & 1) ? acc[c / 2][k].y : acc[c / 2][k].x);
      if (lane == 0) s_part[warp][c * 27 + k] = v;
    }
    const float v = warp_sum(bsum[c]);
    if (lane == 0) s_part[warp][WG_CO * 27 + c] = v;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < WG_CO * 27 + WG_CO; i += blockDim.x) {
    const float v = s_part[0][i] + s_part[1][i] + s_part[2][i] + s_part[3][i];
    if (i < WG_CO * 27) {
      const int c = i / 27, k = i % 27;
      if (co0 + c < Cout) atomicAdd(dw + ((long long)(co0 + c) * Cin + ci) * 27 + k, v);
    } else if (ci == 0 && db != nullptr) {
      const int c = i - WG_CO * 27;
      if (co0 + c < Cout) atomicAdd(db + co0 + c, v);
    }
  }
}

// ---- InstanceNorm + LeakyReLU backward ------------------------------------------------------------
// act = lrelu(n), n = (y - mean) * rstd.  mode 0: IN + LReLU;  mode 1: LReLU only (ConvBlock)
__device__ __forceinline__ void act_to_n(float a, float& n, float& slope) {
  const bool pos = a >= 0.f;
  n = pos ? a : a * 10.0f;
  slope = pos ? 1.0f : 0.1f;
}

__global__ void __launch_bounds__(256) in_bwd_reduce_kernel(const float* __restrict__ da, const float* __restrict__ act,
                                                            double* __restrict__ sums, long long N) {
  const int bc = blockIdx.y;
  const float* g = da + (long long)bc * N;
  const float* a = act + (long long)bc * N;
  float s1 = 0.f, s2 = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float n, sl;
    act_to_n(__ldg(a + i), n, sl);
    const float dn = __ldg(g + i) * sl;
    s1 += dn;
    s2 = fmaf(dn, n, s2);
  }
  __shared__ float s_w[8][2];
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if ((threadIdx.x & 31) == 0) {
    s_w[threadIdx.x >> 5][0] = s1;
    s_w[threadIdx.x >> 5][1] = s2;
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += (double)s_w[i][threadIdx.x];
    atomicAdd(sums + 2 * bc + threadIdx.x, tot);
  }
}

__global__ void __launch_bounds__(256) in_bwd_apply_kernel(const float* __restrict__ da, const float* __restrict__ act,
                                                           const double* __restrict__ fwd_stats,
                                                           const double* __restrict__ sums, float* __restrict__ dy,
                                                           long long N, float eps, int mode) {
  const int bc = blockIdx.y;
  const float* g = da + (long long)bc * N;
  const float* a = act + (long long)bc * N;
  float* o = dy + (long long)bc * N;
  float rstd = 1.f, m1 = 0.f, m2 = 0.f;
  if (mode == 0) {
    const double s = fwd_stats[2 * bc], ss = fwd_stats[2 * bc + 1];
    const double mean = s / (double)N;
    rstd = (float)(1.0 / sqrt(fmax(ss / (double)N - mean * mean, 0.0) + (double)eps));
    m1 = (float)(sums[2 * bc] / (double)N);
    m2 = (float)(sums[2 * bc + 1] / (double)N);
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    float n, sl;
    act_to_n(__ldg(a + i), n, sl);
    const float dn = __ldg(g + i) * sl;
    o[i] = (mode == 0) ? rstd * (dn - m1 - n * m2) : dn;
  }
}

// d_full[v] += d_pooled[v / 2] / 8   (AvgPool3d(2) backward added onto an existing gradient)
__global__ void __launch_bounds__(256) pool_bwd_add_kernel(const float* __restrict__ dp, float* __restrict__ dfull, int D,
                                                           int H, int W) {
  const int bc = blockIdx.y;
  const int PD = D / 2, PH = H / 2, PW = W / 2;
  const long long N = (long long)D * H * W;
  const float* p = dp + (long long)bc * PD * PH * PW;
  float* f = dfull + (long long)bc * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    const long long r = i / W;
    const int h = (int)(r % H), d = (int)(r / H);
    if (d / 2 < PD && h / 2 < PH && w / 2 < PW) f[i] += 0.125f * __ldg(p + ((long long)(d / 2) * PH + h / 2) * PW + w / 2);
  }
}

// ---- losses backward --------------------------------------------------------------------------------
// Grad3d 'l2': L = (mean(dD^2) + mean(dH^2) + mean(dW^2)) / 3
__global__ void __launch_bounds__(256) grad3d_bwd_kernel(const float* __restrict__ f, float* __restrict__ df, int D, int H,
                                                         int W, long long planes, float cd, float ch, float cw,
                                                         const float* __restrict__ gscale) {
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const long long total = planes * N;
  const float gs = gscale ? __ldg(gscale) : 1.0f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % N;
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    const float v = __ldg(f + i);
    float acc = 0.f;
    if (d >= 1) acc += cd * (v - __ldg(f + i - HW));
    if (d + 1 < D) acc -= cd * (__ldg(f + i + HW) - v);
    if (h >= 1) acc += ch * (v - __ldg(f + i - W));
    if (h + 1 < H) acc -= ch * (__ldg(f + i + W) - v);
    if (w >= 1) acc += cw * (v - __ldg(f + i - 1));
    if (w + 1 < W) acc -= cw * (__ldg(f + i + 1) - v);
    df[i] = gs * acc;
  }
}

// NCC backward w.r.t. the first argument I.  With S = box sums (n = win^3):
//   cross = IJ - I_s J_s / n,  Iv = I2 - I_s^2 / n,  Jv = J2 - J_s^2 / n,  cc = cross^2 / (Iv Jv + eps)
//   A = d cc / d cross = 2 cross / den,   Cc = d cc / d Iv = -cross^2 Jv / den^2
//   dL/dI = -(1/M) [ J box(A) - box(A J_s / n) + 2 I box(Cc) - 2 box(Cc I_s / n) ]
// pass kinds: 0 = first box pass over W forming the five products; 1 = middle pass (H); 2 = last pass (D) that
// turns the five sums into the four fields; then 3 box passes over the four fields; the last one combines.
template <int AXIS, int NF_IN, int MODE>
__global__ void __launch_bounds__(256)
ncc_bwd_box_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ in,
                   float* __restrict__ out, int D, int H, int W, int R, float win_size, float coef,
                   const float* __restrict__ gscale) {
  // MODE 0: inputs a,b -> 5 sums;  1: 5 -> 5;  2: 5 sums -> 4 fields (A, A*uJ, Cc, Cc*uI);  3: 4 -> 4;
  // MODE 4: 4 -> d_I using a (= I) and b (= J)
  const long long N = (long long)D * H * W;
  const long long BN = N * gridDim.y;
  const int bz = blockIdx.y;
  const int HW = H * W;
  const float gs = (MODE == 4 && gscale) ? __ldg(gscale) : 1.0f;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    const int pos = AXIS == 0 ? w : (AXIS == 1 ? h : d);
    const int len = AXIS == 0 ? W : (AXIS == 1 ? H : D);
    const long long stride = AXIS == 0 ? 1 : (AXIS == 1 ? W : HW);
    const long long base = (long long)bz * N + p;
    float s[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    const int lo = max(-R, -pos), hi = min(R, len - 1 - pos);
    for (int o = lo; o <= hi; ++o) {
      const long long idx = base + o * stride;
      if (MODE == 0) {
        const float x = __ldg(a + idx), y = __ldg(b + idx);
        s[0] += x;
        s[1] += y;
        s[2] = fmaf(x, x, s[2]);
        s[3] = fmaf(y, y, s[3]);
        s[4] = fmaf(x, y, s[4]);
      } else {
#pragma unroll
        for (int f = 0; f < NF_IN; ++f) s[f] += __ldg(in + f * BN + idx);
      }
    }
    if (MODE == 0 || MODE == 1) {
#pragma unroll
      for (int f = 0; f < 5; ++f) out[f * BN + base] = s[f];
    } else if (MODE == 2) {
      const float I_s = s[0], J_s = s[1];
      const float uI = I_s / win_size, uJ = J_s / win_size;
      const float cross = s[4] - uJ * I_s - uI * J_s + uI * uJ * win_size;
      const float Iv = s[2] - 2.f * uI * I_s + uI * uI * win_size;
      const float Jv = s[3] - 2.f * uJ * J_s + uJ * uJ * win_size;
      const float den = Iv * Jv + 1e-5f;
      const float A = 2.f * cross / den;
      const float Cc = -cross * cross * Jv / (den * den);
      out[0 * BN + base] = A;
      out[1 * BN + base] = A * uJ;
      out[2 * BN + base] = Cc;
      out[3 * BN + base] = Cc * uI;
    } else if (MODE == 3) {
#pragma unroll
      for (int f = 0; f < 4; ++f) out[f * BN + base] = s[f];
    } else {
      const float I = __ldg(a + base), J = __ldg(b + base);
      out[base] = gs * coef * (J * s[0] - s[1] + 2.f * I * s[2] - 2.f * s[3]);
    }
  }
}

// Adam with amsgrad (torch.optim.Adam semantics, weight_decay = 0; train.py:101)
__global__ void __launch_bounds__(256) adam_amsgrad_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                           float* __restrict__ m, float* __restrict__ v,
                                                           float* __restrict__ vmax, long long n, float lr, float b1,
                                                           float b2, float eps, float bc1, float bc2_sqrt) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float gi = g[i];
    const float mi = b1 * m[i] + (1.f - b1) * gi;
    const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
    const float vm = fmaxf(vmax[i], vi);
    m[i] = mi;
    v[i] = vi;
    vmax[i] = vm;
    const float denom = sqrtf(vm) / bc2_sqrt + eps;
    p[i] -= (lr / bc1) * (mi / denom);
  }
}

}  // namespace

int launch_conv3d_flip_weights(const float* w, float* wT, int Cout, int Cin, cudaStream_t st) {
  flip_weights_kernel<<<grid_for((long long)Cout * Cin * 27, 256, 4), 256, 0, st>>>(w, wT, Cout, Cin);
  return check_launch("conv3d_flip_weights");
}

int launch_conv3d_wgrad_tma(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int D, int H,
                            int W, cudaStream_t st, bool* handled);

int launch_conv3d_wgrad(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int D, int H,
                        int W, cudaStream_t st) {
  cudaMemsetAsync(dw, 0, (size_t)Cout * Cin * 27 * sizeof(float), st);
  if (db != nullptr) cudaMemsetAsync(db, 0, (size_t)Cout * sizeof(float), st);
  {
    bool handled = false;
    int rc = launch_conv3d_wgrad_tma(x, dy, dw, db, B, Cin, Cout, D, H, W, st, &handled);
    if (handled) return rc;
  }
  const int tiles_h = ceil_div(H, 4), tiles_w = ceil_div(W, 32);
  const long long groups = (long long)ceil_div(Cout, WG_CO) * Cin;
  if (groups > 65535) {
    set_error("conv3d_wgrad: Cout/4 * Cin = %lld exceeds grid.y", groups);
    return SMILE_ERR_UNSUPPORTED;
  }
  // depth chunks: enough CTAs to fill the GPU, but at least 8 planes per chunk to amortise the reduction
  long long per_plane = (long long)tiles_h * tiles_w * groups * B;
  int chunks = (int)ceil_div_ll(4LL * kNumSMs * 4, per_plane);
  if (chunks < 1) chunks = 1;
  int dchunk = ceil_div(D, chunks);
  if (dchunk < 8) dchunk = D < 8 ? D : 8;
  chunks = ceil_div(D, dchunk);
  dim3 grid(tiles_h * tiles_w * chunks, (unsigned)groups, B);
  conv3d_wgrad_kernel<<<grid, 128, 0, st>>>(x, dy, dw, db, B, Cin, Cout, D, H, W, dchunk, tiles_h, tiles_w);
  return check_launch("conv3d_wgrad");
}

int launch_in_lrelu_bwd(const float* da, const float* act, const double* fwd_stats, double* sums_work, float* dy, int B,
                        int C, long long N, float eps, int mode, cudaStream_t st) {
  dim3 grid(grid_for(N, 256, 8), B * C);
  if (mode == 0) {
    cudaMemsetAsync(sums_work, 0, (size_t)B * C * 2 * sizeof(double), st);
    in_bwd_reduce_kernel<<<grid, 256, 0, st>>>(da, act, sums_work, N);
    int rc = check_launch("in_lrelu_bwd(reduce)");
    if (rc) return rc;
  }
  in_bwd_apply_kernel<<<grid, 256, 0, st>>>(da, act, fwd_stats, sums_work, dy, N, eps, mode);
  return check_launch("in_lrelu_bwd(apply)");
}

int launch_pool_bwd_add(const float* dpooled, float* dfull, int B, int C, int D, int H, int W, cudaStream_t st) {
  dim3 grid(grid_for((long long)D * H * W, 256, 8), B * C);
  pool_bwd_add_kernel<<<grid, 256, 0, st>>>(dpooled, dfull, D, H, W);
  return check_launch("pool_bwd_add");
}

int launch_grad3d_l2_bwd(const float* flow, float* dflow, const float* gscale, int B, int C, int D, int H, int W,
                         cudaStream_t st) {
  const long long planes = (long long)B * C;
  const double nd = (double)planes * (D - 1) * H * W, nh = (double)planes * D * (H - 1) * W,
               nw = (double)planes * D * H * (W - 1);
  grad3d_bwd_kernel<<<grid_for(planes * D * H * W, 256, 16), 256, 0, st>>>(flow, dflow, D, H, W, planes,
                                                                          (float)(2.0 / (3.0 * nd)), (float)(2.0 / (3.0 * nh)),
                                                                          (float)(2.0 / (3.0 * nw)), gscale);
  return check_launch("grad3d_l2_bwd");
}

// work: 10 * B * N floats (two ping-pong sets of five planes)
int launch_ncc_vxm_bwd(const float* y_true, const float* y_pred, float* d_true, float* work, const float* gscale, int B,
                       int D, int H, int W, int win, cudaStream_t st) {
  const long long N = (long long)D * H * W, BN = N * B;
  float* s1 = work;
  float* s2 = work + 5 * BN;
  const int R = win / 2;
  const float ws = (float)win * win * win;
  const float coef = (float)(-1.0 / (double)BN);
  dim3 grid(grid_for(N, 256, 16), B);
  ncc_bwd_box_kernel<0, 5, 0><<<grid, 256, 0, st>>>(y_true, y_pred, nullptr, s1, D, H, W, R, ws, coef, nullptr);
  ncc_bwd_box_kernel<1, 5, 1><<<grid, 256, 0, st>>>(nullptr, nullptr, s1, s2, D, H, W, R, ws, coef, nullptr);
  ncc_bwd_box_kernel<2, 5, 2><<<grid, 256, 0, st>>>(nullptr, nullptr, s2, s1, D, H, W, R, ws, coef, nullptr);
  ncc_bwd_box_kernel<0, 4, 3><<<grid, 256, 0, st>>>(nullptr, nullptr, s1, s2, D, H, W, R, ws, coef, nullptr);
  ncc_bwd_box_kernel<1, 4, 3><<<grid, 256, 0, st>>>(nullptr, nullptr, s2, s1, D, H, W, R, ws, coef, nullptr);
  ncc_bwd_box_kernel<2, 4, 4><<<grid, 256, 0, st>>>(y_true, y_pred, s1, d_true, D, H, W, R, ws, coef, gscale);
  return check_launch("ncc_vxm_bwd");
}

int launch_adam_amsgrad(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float b1,
                        float b2, float eps, int step, cudaStream_t st) {
  const float bc1 = 1.0f - powf(b1, (float)step);
  const float bc2 = 1.0f - powf(b2, (float)step);
  adam_amsgrad_kernel<<<grid_for(n, 256, 8), 256, 0, st>>>(p, g, m, v, vmax, n, lr, b1, b2, eps, bc1, sqrtf(bc2));
  return check_launch("adam_amsgrad");
}

}  // namespace smile
