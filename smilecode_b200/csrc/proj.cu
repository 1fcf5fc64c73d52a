// a7 ProjectionLayer.forward (reference ModeT/models.py:238-241): channels-first feature volume
// -> Linear(Cin -> C) -> LayerNorm(C) -> channels-last [B, D, H, W, C], one fused pass.
// HBM-bound (4*(Cin + C) bytes per voxel): one thread per voxel, plane reads coalesced across
// the warp, the C x Cin weight matrix broadcast from shared memory, the C outputs and the
// LayerNorm statistics in registers.
#include "common.cuh"
#include "kernels.h"

namespace smile {

template <int C>
__global__ void __launch_bounds__(128) proj_ln_kernel(const float* __restrict__ feat, const float* __restrict__ weight,
                                                      const float* __restrict__ bias, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta, float* __restrict__ out, int Cin,
                                                      long long N, float eps) {
  extern __shared__ float smem[];
  float* s_w = smem;             // [Cin][C]  (transposed so one voxel's C weights are contiguous)
  float* s_b = s_w + Cin * C;    // bias, gamma, beta
  for (int i = threadIdx.x; i < Cin * C; i += blockDim.x) {
    int ci = i / C, c = i - ci * C;
    s_w[i] = weight[c * Cin + ci];
  }
  for (int i = threadIdx.x; i < C; i += blockDim.x) {
    s_b[i] = bias[i];
    s_b[C + i] = gamma[i];
    s_b[2 * C + i] = beta[i];
  }
  __syncthreads();
  const int b = blockIdx.y;
  const float* fb = feat + (long long)b * Cin * N;
  float* ob = out + (long long)b * N * C;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = s_b[c];
    for (int ci = 0; ci < Cin; ++ci) {
      const float x = __ldg(fb + (long long)ci * N + p);
      const float* wr = s_w + ci * C;
#pragma unroll
      for (int c = 0; c < C; ++c) acc[c] = fmaf(x, wr[c], acc[c]);
    }
    float mean = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) mean += acc[c];
    mean *= (1.0f / C);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < C; ++c) {
      float dlt = acc[c] - mean;
      var = fmaf(dlt, dlt, var);
    }
    const float rstd = rsqrtf(var * (1.0f / C) + eps);
    float* o = ob + p * C;
    if (C % 2 == 0) {
#pragma unroll
      for (int c = 0; c < C; c += 2) {
        float2 v;
        v.x = (acc[c] - mean) * rstd * s_b[C + c] + s_b[2 * C + c];
        v.y = (acc[c + 1] - mean) * rstd * s_b[C + c + 1] + s_b[2 * C + c + 1];
        *reinterpret_cast<float2*>(o + c) = v;
      }
    } else {
#pragma unroll
      for (int c = 0; c < C; ++c) o[c] = (acc[c] - mean) * rstd * s_b[C + c] + s_b[2 * C + c];
    }
  }
}

template <int C>
static int launch_c(const float* feat, const float* weight, const float* bias, const float* gamma, const float* beta,
                    float* out, int B, int Cin, long long N, float eps, cudaStream_t st) {
  size_t smem = (size_t)(Cin * C + 3 * C) * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(proj_ln_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("proj_ln: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
  }
  long long g = ceil_div_ll(N, 128);
  long long cap = (long long)kNumSMs * 16;
  dim3 grid((unsigned)(g < cap ? g : cap), B);
  proj_ln_kernel<C><<<grid, 128, smem, st>>>(feat, weight, bias, gamma, beta, out, Cin, N, eps);
  return check_launch("proj_ln");
}

int launch_proj_ln(const float* feat, const float* weight, const float* bias, const float* gamma, const float* beta,
                   float* out, int B, int Cin, int C, long long N, float eps, cudaStream_t st) {
#define SMILE_PROJ_CASE(CC) \
  case CC:                  \
    return launch_c<CC>(feat, weight, bias, gamma, beta, out, B, Cin, N, eps, st);
  switch (C) {
    SMILE_PROJ_CASE(4)
    SMILE_PROJ_CASE(6)
    SMILE_PROJ_CASE(8)
    SMILE_PROJ_CASE(12)
    SMILE_PROJ_CASE(16)
    SMILE_PROJ_CASE(18)
    SMILE_PROJ_CASE(24)
    SMILE_PROJ_CASE(30)
    SMILE_PROJ_CASE(32)
    SMILE_PROJ_CASE(36)
    SMILE_PROJ_CASE(42)
    SMILE_PROJ_CASE(48)
    SMILE_PROJ_CASE(64)
    default:
      set_error("proj_ln: projection width C=%d is not compiled in (supported: 4,6,8,12,16,18,24,30,32,36,42,48,64)", C);
      return SMILE_ERR_UNSUPPORTED;
  }
#undef SMILE_PROJ_CASE
}

}  // namespace smile
