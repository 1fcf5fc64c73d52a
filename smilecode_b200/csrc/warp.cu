// a5 SpatialTransformer.forward (reference ModeT/models.py:49-67) and a6 flow upsample + compose
// (models.py:354, 392, 398, 403, 408) as stand-alone kernels.  HBM-bound gathers: one thread per
// voxel, lanes along W so every plane read/write is a coalesced 128 B row segment; the eight
// corner offsets / weights are computed once per voxel and reused for every channel.
#include "common.cuh"
#include "kernels.h"

namespace smile {

// out[b,c,p] = trilinear(src[b,c], p + flow[b,:,p])            (zeros padding)
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
__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, float* __restrict__ out, int C,
                                                         int D, int H, int W, float pre) {
  const int OD = 2 * D, OH = 2 * H, OW = 2 * W;
  const long long ON = (long long)OD * OH * OW;
  const long long IN = (long long)D * H * W;
  const int b = blockIdx.y;
  const float rd = up2_ratio(D), rh = up2_ratio(H), rw = up2_ratio(W);
  const float* xb = x + (long long)b * C * IN;
  float* ob = out + (long long)b * C * ON;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < ON; p += (long long)gridDim.x * blockDim.x) {
    int od = (int)(p / ((long long)OH * OW));
    int r = (int)(p - (long long)od * OH * OW);
    int oh = r / OW, ow = r - oh * OW;
    int d0, d1, h0, h1, w0, w1;
    float ld, lh, lw;
    up2_index(od, D, rd, d0, d1, ld);
    up2_index(oh, H, rh, h0, h1, lh);
    up2_index(ow, W, rw, w0, w1, lw);
    const float md = 1.0f - ld, mh = 1.0f - lh, mw = 1.0f - lw;
    const int o00 = (d0 * H + h0) * W, o01 = (d0 * H + h1) * W, o10 = (d1 * H + h0) * W, o11 = (d1 * H + h1) * W;
    for (int c = 0; c < C; ++c) {
      const float* xc = xb + (long long)c * IN;
      float a00 = mw * __ldg(xc + o00 + w0) + lw * __ldg(xc + o00 + w1);
      float a01 = mw * __ldg(xc + o01 + w0) + lw * __ldg(xc + o01 + w1);
      float a10 = mw * __ldg(xc + o10 + w0) + lw * __ldg(xc + o10 + w1);
      float a11 = mw * __ldg(xc + o11 + w0) + lw * __ldg(xc + o11 + w1);
      float v = md * (mh * a00 + lh * a01) + ld * (mh * a10 + lh * a11);
      ob[(long long)c * ON + p] = pre * v;
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
  warp3d_kernel<<<dim3(grid_for(N, 256), B), 256, 0, st>>>(src, flow, out, C, D, H, W);
  return check_launch("warp3d");
}
int launch_compose(const float* flow, const float* w, float* out, int B, int D, int H, int W, float post, cudaStream_t st) {
  long long N = (long long)D * H * W;
  compose_kernel<<<dim3(grid_for(N, 256), B), 256, 0, st>>>(flow, w, out, D, H, W, post);
  return check_launch("flow_compose");
}
int launch_upsample2x(const float* x, float* out, int B, int C, int D, int H, int W, float pre, cudaStream_t st) {
  long long ON = 8LL * D * H * W;
  upsample2x_kernel<<<dim3(grid_for(ON, 256), B), 256, 0, st>>>(x, out, C, D, H, W, pre);
  return check_launch("upsample2x");
}

}  // namespace smile
