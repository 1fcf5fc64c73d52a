// a5 SpatialTransformer.forward (reference ModeT/models.py:49-67) and a6 flow upsample + compose
// (models.py:354, 392, 398, 403, 408) as stand-alone kernels, plus the a5+a7 fusion
// k = ProjectionLayer(SpatialTransformer(M, flow)) used at every decoder level (models.py:388-389).
//
// HBM/L2-bound gathers: one thread per voxel, lanes along W so every plane read/write is a
// coalesced 128 B row segment.  The sample is set up once per voxel and reused for every channel:
// exact coordinate replay (Markstein division == IEEE division, common.cuh), corner indices clamped
// into the volume with the weight of an out-of-volume corner forced to 0 (adds +-0 where torch
// skips the corner), four row offsets + two column indices, eight weights in torch's order.
#include "common.cuh"
#include "kernels.h"

namespace smile {

namespace {

struct Sample8 {
  unsigned r00, r01, r10, r11;  // (z0,y0) (z0,y1) (z1,y0) (z1,y1) row offsets
  unsigned x0, x1;
  float w[8];                   // tnw tne tsw tse bnw bne bsw bse
};

__device__ __forceinline__ void axis_corners(float c, int S, int& i0, int& i1, float& w0, float& w1) {
  const float f = floorf(c);
  const int j0 = __float2int_rd(c), j1 = j0 + 1;
  w1 = __fsub_rn(c, f);
  w0 = __fsub_rn(__fadd_rn(f, 1.0f), c);
  w0 = ((unsigned)j0 < (unsigned)S) ? w0 : 0.f;
  w1 = ((unsigned)j1 < (unsigned)S) ? w1 : 0.f;
  i0 = min(max(j0, 0), S - 1);
  i1 = min(max(j1, 0), S - 1);
}

struct VolDims {
  int D, H, W;
  float dm1, hm1, wm1, rd, rh, rw;  // S-1 and RN(1/(S-1))
};

__device__ __forceinline__ float coord(float idx, float f, float sm1, float rc) {
  // p = idx + f; n = 2*(p/(S-1) - 0.5); x = ((n+1)/2)*(S-1)   (models.py:51,56; GridSampler.h:31)
  const float p = __fadd_rn(idx, f);
  const float q0 = __fmul_rn(p, rc);
  const float r = __fmaf_rn(-q0, sm1, p);
  const float q = __fmaf_rn(r, rc, q0);
  return __fmul_rn(__fadd_rn(__fsub_rn(q, 0.5f), 0.5f), sm1);
}

__device__ __forceinline__ void sample_setup(Sample8& s, const VolDims& v, int d, int h, int w, float fd, float fh,
                                             float fw) {
  int z0, z1, y0, y1, x0, x1;
  float uz0, uz1, uy0, uy1, ux0, ux1;
  axis_corners(coord((float)d, fd, v.dm1, v.rd), v.D, z0, z1, uz0, uz1);
  axis_corners(coord((float)h, fh, v.hm1, v.rh), v.H, y0, y1, uy0, uy1);
  axis_corners(coord((float)w, fw, v.wm1, v.rw), v.W, x0, x1, ux0, ux1);
  const float u00 = __fmul_rn(ux0, uy0), u10 = __fmul_rn(ux1, uy0), u01 = __fmul_rn(ux0, uy1), u11 = __fmul_rn(ux1, uy1);
  s.w[0] = __fmul_rn(u00, uz0);
  s.w[1] = __fmul_rn(u10, uz0);
  s.w[2] = __fmul_rn(u01, uz0);
  s.w[3] = __fmul_rn(u11, uz0);
  s.w[4] = __fmul_rn(u00, uz1);
  s.w[5] = __fmul_rn(u10, uz1);
  s.w[6] = __fmul_rn(u01, uz1);
  s.w[7] = __fmul_rn(u11, uz1);
  s.r00 = (unsigned)((z0 * v.H + y0) * v.W);
  s.r01 = (unsigned)((z0 * v.H + y1) * v.W);
  s.r10 = (unsigned)((z1 * v.H + y0) * v.W);
  s.r11 = (unsigned)((z1 * v.H + y1) * v.W);
  s.x0 = (unsigned)x0;
  s.x1 = (unsigned)x1;
}

__device__ __forceinline__ float sample_gather(const Sample8& s, const float* __restrict__ plane) {
  const float v0 = __ldg(plane + (s.r00 + s.x0)), v1 = __ldg(plane + (s.r00 + s.x1));
  const float v2 = __ldg(plane + (s.r01 + s.x0)), v3 = __ldg(plane + (s.r01 + s.x1));
  const float v4 = __ldg(plane + (s.r10 + s.x0)), v5 = __ldg(plane + (s.r10 + s.x1));
  const float v6 = __ldg(plane + (s.r11 + s.x0)), v7 = __ldg(plane + (s.r11 + s.x1));
  float t = __fmul_rn(v0, s.w[0]);
  t = __fadd_rn(t, __fmul_rn(v1, s.w[1]));
  t = __fadd_rn(t, __fmul_rn(v2, s.w[2]));
  t = __fadd_rn(t, __fmul_rn(v3, s.w[3]));
  t = __fadd_rn(t, __fmul_rn(v4, s.w[4]));
  t = __fadd_rn(t, __fmul_rn(v5, s.w[5]));
  t = __fadd_rn(t, __fmul_rn(v6, s.w[6]));
  t = __fadd_rn(t, __fmul_rn(v7, s.w[7]));
  return t;
}

VolDims make_dims(int D, int H, int W) {
  VolDims v;
  v.D = D; v.H = H; v.W = W;
  v.dm1 = (float)(D - 1); v.hm1 = (float)(H - 1); v.wm1 = (float)(W - 1);
  v.rd = 1.0f / v.dm1; v.rh = 1.0f / v.hm1; v.rw = 1.0f / v.wm1;
  return v;
}

// Work decomposition of the two fast kernels: a CTA owns an (8*32/TWL rows) x (TWL columns) patch of the (H, W)
// plane and marches `dchunk` depths.  The z1 rows of one step are the z0 rows of the next and the
// y-neighbours are the rows of the neighbouring warps, so the 8-corner gathers hit L1 instead of
// re-fetching every row from L2 for every voxel row (a flat 1-D mapping re-reads each source row ~4x).
struct Tiling {
  int tiles_h, tiles_w, dchunk;
};

template <int TWL>
__device__ __forceinline__ bool tile_coords(const Tiling& tl, const VolDims& v, int& h, int& w, int& d_begin, int& d_end) {
  constexpr int LH = 32 / TWL;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int t = blockIdx.x;
  const int tw = t % tl.tiles_w;
  t /= tl.tiles_w;
  const int th = t % tl.tiles_h;
  const int tc = t / tl.tiles_h;
  h = th * (8 * LH) + warp * LH + lane / TWL;
  w = tw * TWL + lane % TWL;
  d_begin = tc * tl.dchunk;
  d_end = min(v.D, d_begin + tl.dchunk);
  return h < v.H && w < v.W;
}

// out[b,c,p] = trilinear(src[b,c], p + flow[b,:,p])            (zeros padding)
template <int TWL>
__global__ void __launch_bounds__(256) warp3d_fast_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                          float* __restrict__ out, int C, const VolDims v,
                                                          const Tiling tl) {
  const int HW = v.H * v.W;
  const int N = v.D * HW;
  const int b = blockIdx.y;
  const float* fl = flow + (long long)b * 3 * N;
  const float* sb = src + (long long)b * C * N;
  float* ob = out + (long long)b * C * N;
  int h, w, d_begin, d_end;
  if (!tile_coords<TWL>(tl, v, h, w, d_begin, d_end)) return;
  for (int d = d_begin; d < d_end; ++d) {
    const int p = d * HW + h * v.W + w;
    Sample8 s;
    sample_setup(s, v, d, h, w, __ldg(fl + p), __ldg(fl + N + p), __ldg(fl + 2 * N + p));
    int c = 0;
    for (; c + 2 <= C; c += 2) {  // two channels in flight: 16 independent gathers
      const float a0 = sample_gather(s, sb + (long long)c * N);
      const float a1 = sample_gather(s, sb + (long long)(c + 1) * N);
      ob[(long long)c * N + p] = a0;
      ob[(long long)(c + 1) * N + p] = a1;
    }
    if (c < C) ob[(long long)c * N + p] = sample_gather(s, sb + (long long)c * N);
  }
}

// k[b,p,:] = LayerNorm(Linear(trilinear(src[b,:], p + flow[b,:,p])))  -> channels-last [B,N,C]
template <int CIN, int C, int TWL>
__global__ void __launch_bounds__(256, 3) warp_proj_ln_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                           const float* __restrict__ weight, const float* __restrict__ bias,
                                                           const float* __restrict__ gamma, const float* __restrict__ beta,
                                                           float* __restrict__ out, const VolDims v, const Tiling tl,
                                                           float eps) {
  __shared__ float s_w[CIN * C + 3 * C];  // [ci][c] weights, then bias, gamma, beta
  for (int i = threadIdx.x; i < CIN * C; i += blockDim.x) {
    const int ci = i / C, c = i - ci * C;
    s_w[i] = weight[c * CIN + ci];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_w[CIN * C + i] = bias[i];
    s_w[CIN * C + C + i] = gamma[i];
    s_w[CIN * C + 2 * C + i] = beta[i];
  }
  __syncthreads();
  const float* s_b = s_w + CIN * C;
  const int HW = v.H * v.W;
  const int N = v.D * HW;
  const int b = blockIdx.y;
  const float* fl = flow + (long long)b * 3 * N;
  const float* sb = src + (long long)b * CIN * N;
  float* ob = out + (long long)b * N * C;
  int h, w, d_begin, d_end;
  if (!tile_coords<TWL>(tl, v, h, w, d_begin, d_end)) return;
  for (int d = d_begin; d < d_end; ++d) {
    const int p = d * HW + h * v.W + w;
    Sample8 s;
    sample_setup(s, v, d, h, w, __ldg(fl + p), __ldg(fl + N + p), __ldg(fl + 2 * N + p));
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = s_b[c];
    // gathers of eight channels in flight at once (64 loads): the kernel is latency bound, not issue bound
    // (8->6 @160x192x160: 0.22 -> 0.14 ms).  Batches of GB channels; 4 where 8 made ptxas spill (32 -> 12).
    constexpr int GB = (CIN == 32) ? 4 : 8;
#pragma unroll 1
    for (int ci0 = 0; ci0 < CIN; ci0 += GB) {
      float x[GB];
#pragma unroll
      for (int u = 0; u < GB; ++u) x[u] = sample_gather(s, sb + (long long)(ci0 + u) * N);
#pragma unroll
      for (int u = 0; u < GB; ++u) {
        const float* wr = s_w + (ci0 + u) * C;
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = fmaf(x[u], wr[c], acc[c]);
      }
    }
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mean += acc[c];
    mean *= (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      const float dlt = acc[c] - mean;
      var = fmaf(dlt, dlt, var);
    }
    const float rstd = rsqrtf(var * (1.0f / C) + eps);
    float* o = ob + (long long)p * C;
#pragma unroll
    for (int c = 0; c < C; c += 2) {
      float2 t;
      t.x = (acc[c] - mean) * rstd * s_b[C + c] + s_b[2 * C + c];
      t.y = (acc[c + 1] - mean) * rstd * s_b[C + c + 1] + s_b[2 * C + c + 1];
      *reinterpret_cast<float2*>(o + c) = t;
    }
  }
}

// tiles of (8 * 32/TWL) x TWL, depth chunks sized so the grid is a few waves of 148 SMs x 8 CTAs
template <int TWL>
Tiling make_tiling(int B, int D, int H, int W, int& grid_x) {
  Tiling tl;
  tl.tiles_h = ceil_div(H, 8 * (32 / TWL));
  tl.tiles_w = ceil_div(W, TWL);
  const int plane_tiles = tl.tiles_h * tl.tiles_w * B;
  int chunks = ceil_div(2 * kNumSMs * 8, plane_tiles);  // aim at ~2 full waves
  if (chunks < 1) chunks = 1;
  if (chunks > D) chunks = D;
  tl.dchunk = ceil_div(D, chunks);
  grid_x = tl.tiles_h * tl.tiles_w * ceil_div(D, tl.dchunk);
  return tl;
}

}  // namespace

// Generic kernels (any size, incl. dimensions of extent 1 where the reference divides by zero).
__global__ void __launch_bounds__(256) warp3d_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                     float* __restrict__ out, int C, int D, int H, int W) {
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y;
  const float* fl = flow + (long long)b * 3 * N;
  const float* sb = src + (long long)b * C * N;
  float* ob = out + (long long)b * C * N;
  const float dm1 = (float)(D - 1), hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    int d = (int)(p / HW);
    int r = (int)(p - (long long)d * HW);
    int h = r / W, w = r - h * W;
    TriSample s;
    tri_setup(s, st_coord(d, __ldg(fl + p), dm1), st_coord(h, __ldg(fl + N + p), hm1),
              st_coord(w, __ldg(fl + 2 * N + p), wm1), D, H, W);
    for (int c = 0; c < C; ++c) ob[(long long)c * N + p] = tri_gather(s, sb + (long long)c * N);
  }
}

// out[b,a,p] = post * ( trilinear(flow[b,a], p + w[b,:,p]) + w[b,a,p] )      a in 0..2
// (the "flow = T(flow, w) + w" compose step; post=2 folds the 2* of models.py:403)
__global__ void __launch_bounds__(256) compose_kernel(const float* __restrict__ flow, const float* __restrict__ wf,
                                                      float* __restrict__ out, int D, int H, int W, float post) {
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int b = blockIdx.y;
  const float* fb = flow + (long long)b * 3 * N;
  const float* wb = wf + (long long)b * 3 * N;
  float* ob = out + (long long)b * 3 * N;
  const float dm1 = (float)(D - 1), hm1 = (float)(H - 1), wm1 = (float)(W - 1);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    int d = (int)(p / HW);
    int r = (int)(p - (long long)d * HW);
    int h = r / W, w = r - h * W;
    float w0 = __ldg(wb + p), w1 = __ldg(wb + N + p), w2 = __ldg(wb + 2 * N + p);
    TriSample s;
    tri_setup(s, st_coord(d, w0, dm1), st_coord(h, w1, hm1), st_coord(w, w2, wm1), D, H, W);
    ob[p] = __fmul_rn(post, __fadd_rn(tri_gather(s, fb), w0));
    ob[N + p] = __fmul_rn(post, __fadd_rn(tri_gather(s, fb + N), w1));
    ob[2 * N + p] = __fmul_rn(post, __fadd_rn(tri_gather(s, fb + 2 * N), w2));
  }
}

// out[b,c] = pre * up2(x[b,c]),  [D,H,W] -> [2D,2H,2W], trilinear, align_corners=True.
// Scaling by a power of two commutes with the interpolation bit for bit, so pre=2 reproduces
// upsample_trilin(2*flow) (models.py:392) and CWM's trailing 2* alike.
// One CTA per output row (od, oh): the depth / height source indices and weights are uniform per CTA, threads run
// along ow (no per-voxel 64-bit divisions; the 93 us of the first version at 80x96x80 -> 160x192x160 were index math).
__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, float* __restrict__ out, int C,
                                                         int D, int H, int W, float pre) {
  const int OH = 2 * H, OW = 2 * W;
  const long long ON = 8LL * D * H * W;
  const long long IN = (long long)D * H * W;
  const int b = blockIdx.y;
  const int od = blockIdx.x / OH, oh = blockIdx.x - od * OH;
  const float rd = up2_ratio(D), rh = up2_ratio(H), rw = up2_ratio(W);
  int d0, d1, h0, h1;
  float ld, lh;
  up2_index(od, D, rd, d0, d1, ld);
  up2_index(oh, H, rh, h0, h1, lh);
  const float md = 1.0f - ld, mh = 1.0f - lh;
  const int o00 = (d0 * H + h0) * W, o01 = (d0 * H + h1) * W, o10 = (d1 * H + h0) * W, o11 = (d1 * H + h1) * W;
  const float* xb = x + (long long)b * C * IN;
  float* ob = out + (long long)b * C * ON + ((long long)od * OH + oh) * OW;
  for (int ow = threadIdx.x; ow < OW; ow += blockDim.x) {
    int w0, w1;
    float lw;
    up2_index(ow, W, rw, w0, w1, lw);
    const float mw = 1.0f - lw;
    for (int c = 0; c < C; ++c) {
      const float* xc = xb + (long long)c * IN;
      float a00 = mw * __ldg(xc + o00 + w0) + lw * __ldg(xc + o00 + w1);
      float a01 = mw * __ldg(xc + o01 + w0) + lw * __ldg(xc + o01 + w1);
      float a10 = mw * __ldg(xc + o10 + w0) + lw * __ldg(xc + o10 + w1);
      float a11 = mw * __ldg(xc + o11 + w0) + lw * __ldg(xc + o11 + w1);
      float v = md * (mh * a00 + lh * a01) + ld * (mh * a10 + lh * a11);
      ob[(long long)c * ON + ow] = pre * v;
    }
  }
}

static inline int grid_for(long long n, int block) {
  long long g = ceil_div_ll(n, block);
  long long cap = (long long)kNumSMs * 32;
  return (int)(g < cap ? g : cap);
}

int launch_warp3d(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W, cudaStream_t st) {
  long long N = (long long)D * H * W;
  if (D >= 2 && H >= 2 && W >= 2) {
    const VolDims v = make_dims(D, H, W);
    int gx;
    if (W > 16) {
      const Tiling tl = make_tiling<32>(B, D, H, W, gx);
      warp3d_fast_kernel<32><<<dim3(gx, B), 256, 0, st>>>(src, flow, out, C, v, tl);
    } else if (W > 8) {
      const Tiling tl = make_tiling<16>(B, D, H, W, gx);
      warp3d_fast_kernel<16><<<dim3(gx, B), 256, 0, st>>>(src, flow, out, C, v, tl);
    } else {
      const Tiling tl = make_tiling<8>(B, D, H, W, gx);
      warp3d_fast_kernel<8><<<dim3(gx, B), 256, 0, st>>>(src, flow, out, C, v, tl);
    }
    return check_launch("warp3d");
  }
  warp3d_kernel<<<dim3(grid_for(N, 256), B), 256, 0, st>>>(src, flow, out, C, D, H, W);
  return check_launch("warp3d");
}

template <int CI, int CC>
static int launch_wp(const float* src, const float* flow, const float* weight, const float* bias, const float* gamma,
                     const float* beta, float* out, int B, int D, int H, int W, float eps, cudaStream_t st) {
  const VolDims v = make_dims(D, H, W);
  int gx;
  if (W > 16) {
    const Tiling tl = make_tiling<32>(B, D, H, W, gx);
    warp_proj_ln_kernel<CI, CC, 32><<<dim3(gx, B), 256, 0, st>>>(src, flow, weight, bias, gamma, beta, out, v, tl, eps);
  } else if (W > 8) {
    const Tiling tl = make_tiling<16>(B, D, H, W, gx);
    warp_proj_ln_kernel<CI, CC, 16><<<dim3(gx, B), 256, 0, st>>>(src, flow, weight, bias, gamma, beta, out, v, tl, eps);
  } else {
    const Tiling tl = make_tiling<8>(B, D, H, W, gx);
    warp_proj_ln_kernel<CI, CC, 8><<<dim3(gx, B), 256, 0, st>>>(src, flow, weight, bias, gamma, beta, out, v, tl, eps);
  }
  return check_launch("warp_proj_ln");
}

int launch_warp_proj_ln(const float* src, const float* flow, const float* weight, const float* bias, const float* gamma,
                        const float* beta, float* out, int B, int Cin, int C, int D, int H, int W, float eps,
                        cudaStream_t st, bool* handled) {
  *handled = false;
  if (D < 2 || H < 2 || W < 2) return SMILE_OK;
#define SMILE_WP(CI, CC)                                                                           \
  if (Cin == CI && C == CC) {                                                                      \
    *handled = true;                                                                               \
    return launch_wp<CI, CC>(src, flow, weight, bias, gamma, beta, out, B, D, H, W, eps, st);      \
  }
  SMILE_WP(8, 6)
  SMILE_WP(16, 6)
  SMILE_WP(32, 12)
  SMILE_WP(64, 24)
#undef SMILE_WP
  return SMILE_OK;
}

int launch_compose(const float* flow, const float* w, float* out, int B, int D, int H, int W, float post, cudaStream_t st) {
  long long N = (long long)D * H * W;
  compose_kernel<<<dim3(grid_for(N, 256), B), 256, 0, st>>>(flow, w, out, D, H, W, post);
  return check_launch("flow_compose");
}
int launch_upsample2x(const float* x, float* out, int B, int C, int D, int H, int W, float pre, cudaStream_t st) {
  const int threads = 2 * W >= 256 ? 256 : (2 * W > 64 ? 128 : 64);
  upsample2x_kernel<<<dim3((unsigned)(4 * D * H), B), threads, 0, st>>>(x, out, C, D, H, W, pre);
  return check_launch("upsample2x");
}

}  // namespace smile
