// Evaluation path of ModeT/infer.py on the GPU: the three things infer.py:86-92 does with every
// registered pair, which in the reference cost a 59 MB device->host copy of the flow plus numpy work:
//   * utils.register_model(img_size, 'nearest')  (utils.py:74-83 -> SpatialTransformer(mode='nearest'), 30-72):
//     warp of the moving segmentation with nearest-neighbour sampling;
//   * utils.dice_val_VOI (utils.py:86-106): per-label intersection / cardinalities of two label volumes;
//   * utils.jacobian_determinant_vxm (utils.py:108-150): np.gradient of (disp + grid) in float64, 3x3 determinant,
//     and the count of non-positive determinants infer.py:90 reports.
#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

// grid_sample(mode='nearest', align_corners=True, padding_mode='zeros'): nearbyint of the un-normalised
// coordinate (round half to even), zero outside the volume (torch ATen/native/cuda/GridSampler.cu, nearest branch)
__global__ void __launch_bounds__(256) warp3d_nearest_kernel(const float* __restrict__ src, const float* __restrict__ flow,
                                                            float* __restrict__ out, int C, int D, int H, int W) {
  const long long N = (long long)D * H * W;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int b = blockIdx.y;
  if (n >= N) return;
  const int w = (int)(n % W);
  const long long t = n / W;
  const int h = (int)(t % H), d = (int)(t / H);
  const float* fb = flow + (long long)b * 3 * N;
  const float z = st_coord(d, fb[n], (float)(D - 1));
  const float y = st_coord(h, fb[N + n], (float)(H - 1));
  const float x = st_coord(w, fb[2 * N + n], (float)(W - 1));
  const int iz = __float2int_rn(z), iy = __float2int_rn(y), ix = __float2int_rn(x);
  const bool in = fabsf(z) < 1e9f && fabsf(y) < 1e9f && fabsf(x) < 1e9f && iz >= 0 && iz < D && iy >= 0 && iy < H &&
                  ix >= 0 && ix < W;
  const long long off = in ? ((long long)iz * H + iy) * W + ix : 0;
  for (int c = 0; c < C; ++c) {
    const float* sp = src + ((long long)b * C + c) * N;
    out[((long long)b * C + c) * N + n] = in ? __ldg(sp + off) : 0.f;
  }
}

constexpr int kMaxLabel = 1024;
constexpr int kMaxLabels = 256;

// counts[l] = {|pred == l and true == l|, |pred == l|, |true == l|}; volumes hold label values as floats, compared
// after truncation to integer like the reference's `.long()` (infer.py:91)
__global__ void __launch_bounds__(256) dice_counts_kernel(const float* __restrict__ pred, const float* __restrict__ truth,
                                                         const int* __restrict__ labels, int nlabels,
                                                         unsigned long long* __restrict__ counts, long long n) {
  __shared__ short lut[kMaxLabel];
  __shared__ unsigned int cnt[kMaxLabels * 3];
  for (int i = threadIdx.x; i < kMaxLabel; i += blockDim.x) lut[i] = -1;
  for (int i = threadIdx.x; i < nlabels * 3; i += blockDim.x) cnt[i] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < nlabels; i += blockDim.x) {
    const int l = labels[i];
    if (l >= 0 && l < kMaxLabel) lut[l] = (short)i;  // a label listed twice counts under its last index
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long p = (long long)pred[i], t = (long long)truth[i];
    const int ip = (p >= 0 && p < kMaxLabel) ? lut[p] : -1;
    const int it = (t >= 0 && t < kMaxLabel) ? lut[t] : -1;
    if (ip >= 0) {
      atomicAdd(&cnt[ip * 3 + 1], 1u);
      if (p == t) atomicAdd(&cnt[ip * 3], 1u);
    }
    if (it >= 0) atomicAdd(&cnt[it * 3 + 2], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < nlabels * 3; i += blockDim.x)
    if (cnt[i]) atomicAdd(&counts[i], (unsigned long long)cnt[i]);
}

// np.gradient along one axis of f = disp_c + grid_c in float64 (edge_order 1): central difference / 2 inside,
// one-sided difference at the two ends.  `self` adds the identity (the grid's own gradient) along its axis.
__device__ __forceinline__ double grad1(const float* __restrict__ f, long long idx, int i, int S, long long stride,
                                        bool self) {
  double g;
  if (i == 0) g = __dsub_rn((double)f[idx + stride], (double)f[idx]);
  else if (i == S - 1) g = __dsub_rn((double)f[idx], (double)f[idx - stride]);
  else g = __dmul_rn(__dsub_rn((double)f[idx + stride], (double)f[idx - stride]), 0.5);
  // (disp + grid) differences: the grid contributes exactly 1 per step (2 over a central difference, halved)
  return self ? __dadd_rn(g, 1.0) : g;
}

__global__ void __launch_bounds__(256) jacdet_kernel(const float* __restrict__ flow, double* __restrict__ det,
                                                    unsigned long long* __restrict__ nonpos, int D, int H, int W) {
  const long long N = (long long)D * H * W;
  const long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int bad = 0;
  if (n < N) {
    const int w = (int)(n % W);
    const long long t = n / W;
    const int h = (int)(t % H), d = (int)(t / H);
    const long long sD = (long long)H * W, sH = W, sW = 1;
    double dx[3], dy[3], dz[3];  // gradients along axis 0 (D), 1 (H), 2 (W) of the three components
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float* f = flow + (long long)c * N;
      dx[c] = grad1(f, n, d, D, sD, c == 0);
      dy[c] = grad1(f, n, h, H, sH, c == 1);
      dz[c] = grad1(f, n, w, W, sW, c == 2);
    }
    // utils.py:137-141, evaluated operation by operation in double without contraction
    const double j0 = __dmul_rn(dx[0], __dsub_rn(__dmul_rn(dy[1], dz[2]), __dmul_rn(dy[2], dz[1])));
    const double j1 = __dmul_rn(dx[1], __dsub_rn(__dmul_rn(dy[0], dz[2]), __dmul_rn(dy[2], dz[0])));
    const double j2 = __dmul_rn(dx[2], __dsub_rn(__dmul_rn(dy[0], dz[1]), __dmul_rn(dy[1], dz[0])));
    const double v = __dadd_rn(__dsub_rn(j0, j1), j2);
    if (det != nullptr) det[n] = v;
    bad = v <= 0.0 ? 1u : 0u;
  }
  const unsigned int m = __ballot_sync(0xffffffffu, bad);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nonpos, (unsigned long long)__popc(m));
}

}  // namespace

int launch_warp3d_nearest(const float* src, const float* flow, float* out, int B, int C, int D, int H, int W,
                          cudaStream_t st) {
  const long long N = (long long)D * H * W;
  dim3 grid((unsigned)ceil_div_ll(N, 256), B);
  warp3d_nearest_kernel<<<grid, 256, 0, st>>>(src, flow, out, C, D, H, W);
  return check_launch("warp3d_nearest");
}

int launch_dice_counts(const float* pred, const float* truth, const int* labels, int nlabels, unsigned long long* counts,
                       long long n, cudaStream_t st) {
  if (nlabels > kMaxLabels) {
    set_error("dice_counts: at most %d labels (got %d)", kMaxLabels, nlabels);
    return SMILE_ERR_UNSUPPORTED;
  }
  cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)nlabels * 3 * sizeof(unsigned long long), st);
  if (e != cudaSuccess) {
    set_error("dice_counts: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  const int blocks = (int)(ceil_div_ll(n, 256) < 4 * kNumSMs ? ceil_div_ll(n, 256) : 4 * kNumSMs);
  dice_counts_kernel<<<blocks, 256, 0, st>>>(pred, truth, labels, nlabels, counts, n);
  return check_launch("dice_counts");
}

int launch_jacdet(const float* flow, double* det, unsigned long long* nonpos, int D, int H, int W, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(nonpos, 0, sizeof(unsigned long long), st);
  if (e != cudaSuccess) {
    set_error("jacdet: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  const long long N = (long long)D * H * W;
  jacdet_kernel<<<(unsigned)ceil_div_ll(N, 256), 256, 0, st>>>(flow, det, nonpos, D, H, W);
  return check_launch("jacdet");
}

}  // namespace smile
