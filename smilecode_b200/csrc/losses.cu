// a10 training losses, forward: NCC_vxm (reference ModeT/losses.py:34-95) and Grad3d 'l2'
// (losses.py:6-31).  HBM-bound streaming kernels.
//
// NCC_vxm: the reference runs five dense 9x9x9 all-ones conv3d (729 MACs per voxel each).  The box
// filter is separable, so it is three 9-tap passes (W, H, D) over the five fields I, J, I^2, J^2,
// IJ; the first pass forms the products on the fly, the last pass evaluates
// cc = cross^2 / (I_var * J_var + 1e-5) and reduces it (warp shuffle -> CTA -> one fp64 atomicAdd).
// Each output sums its 9 taps directly (no running sum), so rounding does not accumulate along a
// row.  Out-of-volume taps contribute 0 (the conv's zero padding).
#include "common.cuh"
#include "kernels.h"

namespace smile {

namespace {

template <int AXIS, bool FIRST, bool LAST>
__global__ void __launch_bounds__(256)
ncc_box_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ in5,
               float* __restrict__ out5, double* __restrict__ acc, int D, int H, int W, int R, float win_size) {
  // in5 / out5: [5][B*N] planes; a, b: [B*N] (FIRST only).  One thread per voxel, lanes along W.
  const long long N = (long long)D * H * W;
  const long long BN = N * gridDim.y;
  const int bz = blockIdx.y;
  const int HW = H * W;
  float local = 0.f;
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
      if (FIRST) {
        const float x = __ldg(a + idx), y = __ldg(b + idx);
        s[0] += x;
        s[1] += y;
        s[2] = fmaf(x, x, s[2]);
        s[3] = fmaf(y, y, s[3]);
        s[4] = fmaf(x, y, s[4]);
      } else {
#pragma unroll
        for (int f = 0; f < 5; ++f) s[f] += __ldg(in5 + f * BN + idx);
      }
    }
    if (!LAST) {
#pragma unroll
      for (int f = 0; f < 5; ++f) out5[f * BN + base] = s[f];
    } else {
      // losses.py:85-93
      const float I_sum = s[0], J_sum = s[1], I2_sum = s[2], J2_sum = s[3], IJ_sum = s[4];
      const float u_I = I_sum / win_size, u_J = J_sum / win_size;
      const float cross = IJ_sum - u_J * I_sum - u_I * J_sum + u_I * u_J * win_size;
      const float I_var = I2_sum - 2.f * u_I * I_sum + u_I * u_I * win_size;
      const float J_var = J2_sum - 2.f * u_J * J_sum + u_J * u_J * win_size;
      local += cross * cross / (I_var * J_var + 1e-5f);
    }
  }
  if (LAST) {
    __shared__ float s_w[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
      double tot = 0.0;
      for (int i = 0; i < 8; ++i) tot += (double)s_w[i];
      atomicAdd(acc, tot);
    }
  }
}

__global__ void ncc_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, double count) {
  out[0] = (float)(-acc[0] / count);
}

// Grad3d 'l2': mean of squared forward differences along D ("dy"), H ("dx"), W ("dz"), / 3.
__global__ void __launch_bounds__(256) grad3d_kernel(const float* __restrict__ f, double* __restrict__ acc, int D, int H,
                                                     int W, long long planes) {
  const int HW = H * W;
  const long long N = (long long)D * HW;
  float sd = 0.f, sh = 0.f, sw = 0.f;
  const long long total = planes * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % N;
    const int d = (int)(p / HW);
    const int r = (int)(p - (long long)d * HW);
    const int h = r / W, w = r - h * W;
    const float v = __ldg(f + i);
    if (d + 1 < D) {
      const float t = __ldg(f + i + HW) - v;
      sd = fmaf(t, t, sd);
    }
    if (h + 1 < H) {
      const float t = __ldg(f + i + W) - v;
      sh = fmaf(t, t, sh);
    }
    if (w + 1 < W) {
      const float t = __ldg(f + i + 1) - v;
      sw = fmaf(t, t, sw);
    }
  }
  __shared__ float s_w[8][3];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sd += __shfl_xor_sync(0xffffffffu, sd, o);
    sh += __shfl_xor_sync(0xffffffffu, sh, o);
    sw += __shfl_xor_sync(0xffffffffu, sw, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_w[threadIdx.x >> 5][0] = sd;
    s_w[threadIdx.x >> 5][1] = sh;
    s_w[threadIdx.x >> 5][2] = sw;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double tot = 0.0;
    for (int i = 0; i < 8; ++i) tot += (double)s_w[i][threadIdx.x];
    atomicAdd(acc + threadIdx.x, tot);
  }
}

__global__ void grad3d_finalize_kernel(const double* __restrict__ acc, float* __restrict__ out, double nd, double nh,
                                       double nw) {
  out[0] = (float)((acc[0] / nd + acc[1] / nh + acc[2] / nw) / 3.0);
}

inline int grid_for(long long n, int block) {
  long long g = ceil_div_ll(n, block);
  const long long cap = (long long)kNumSMs * 16;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

// work: 10 * B * N floats followed (8-byte aligned) by one double; see include/smilecode_b200.h
int launch_ncc_vxm(const float* y_true, const float* y_pred, float* out, float* work, int B, int D, int H, int W, int win,
                   cudaStream_t st) {
  const long long N = (long long)D * H * W;
  const long long BN = N * B;
  float* s1 = work;
  float* s2 = work + 5 * BN;
  double* acc = reinterpret_cast<double*>(work + 10 * BN);
  cudaError_t e = cudaMemsetAsync(acc, 0, sizeof(double), st);
  if (e != cudaSuccess) {
    set_error("ncc_vxm: memset failed: %s", cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  const int R = win / 2;
  const float ws = (float)win * win * win;
  dim3 grid(grid_for(N, 256), B);
  ncc_box_kernel<0, true, false><<<grid, 256, 0, st>>>(y_true, y_pred, nullptr, s1, nullptr, D, H, W, R, ws);
  ncc_box_kernel<1, false, false><<<grid, 256, 0, st>>>(nullptr, nullptr, s1, s2, nullptr, D, H, W, R, ws);
  ncc_box_kernel<2, false, true><<<grid, 256, 0, st>>>(nullptr, nullptr, s2, nullptr, acc, D, H, W, R, ws);
  ncc_finalize_kernel<<<1, 1, 0, st>>>(acc, out, (double)BN);
  return check_launch("ncc_vxm");
}

// work: three doubles
int launch_grad3d_l2(const float* flow, float* out, double* work, int B, int C, int D, int H, int W, cudaStream_t st) {
  cudaError_t e = cudaMemsetAsync(work, 0, 3 * sizeof(double), st);
  if (e != cudaSuccess) {
    set_error("grad3d: memset failed: %s", cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  const long long planes = (long long)B * C;
  const long long total = planes * D * H * W;
  grad3d_kernel<<<grid_for(total, 256), 256, 0, st>>>(flow, work, D, H, W, planes);
  const double nd = (double)planes * (D - 1) * H * W, nh = (double)planes * D * (H - 1) * W,
               nw = (double)planes * D * H * (W - 1);
  grad3d_finalize_kernel<<<1, 1, 0, st>>>(work, out, nd, nh, nw);
  return check_launch("grad3d_l2");
}

}  // namespace smile
