// Shared device helpers for the ModeT hot-path kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define SMILE_OK 0
#define SMILE_ERR_INVALID_ARG (-1)
#define SMILE_ERR_CUDA (-2)
#define SMILE_ERR_UNSUPPORTED (-3)

namespace smile {

void set_error(const char* fmt, ...);
int check_launch(const char* what);

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// Sampling coordinate of SpatialTransformer + grid_sample(align_corners=True), replayed op by op
// in round-to-nearest fp32 with no FMA contraction (reference ModeT/models.py:51,56 and torch
// ATen/native/GridSampler.h:31).  p = idx + f; n = 2*(p/(S-1) - 0.5); x = ((n+1)/2)*(S-1).
// The round trip changes floor(x) for ~12 % of exact-integer coordinates, so it is NOT
// simplified to idx + f.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float st_coord(int idx, float f, float size_minus_1) {
  float p = __fadd_rn((float)idx, f);
  float n = __fmul_rn(2.0f, __fsub_rn(__fdiv_rn(p, size_minus_1), 0.5f));
  return __fmul_rn(__fmul_rn(__fadd_rn(n, 1.0f), 0.5f), size_minus_1);
}

// Corner bookkeeping for one trilinear sample (zeros padding).  Corner order and weight
// products follow torch's grid_sampler_3d: tnw,tne,tsw,tse,bnw,bne,bsw,bse with
// w = (wx * wy) * wz, accumulated sequentially from 0 with separate mul and add.
struct TriSample {
  int off[8];      // linear offsets into one D*H*W plane (valid only where the mask bit is set)
  float w[8];
  unsigned mask;   // bit c set <=> corner c lies inside the volume
};

__device__ __forceinline__ void tri_setup(TriSample& s, float z, float y, float x, int D, int H, int W) {
  float z0f = floorf(z), y0f = floorf(y), x0f = floorf(x);
  // saturating conversions keep wild / non-finite coordinates out of bounds instead of UB
  int z0 = __float2int_rd(z), y0 = __float2int_rd(y), x0 = __float2int_rd(x);
  float wz1 = __fsub_rn(z, z0f), wy1 = __fsub_rn(y, y0f), wx1 = __fsub_rn(x, x0f);
  float wz0 = __fsub_rn(__fadd_rn(z0f, 1.0f), z), wy0 = __fsub_rn(__fadd_rn(y0f, 1.0f), y),
        wx0 = __fsub_rn(__fadd_rn(x0f, 1.0f), x);
  bool finite = (fabsf(z) < 1e9f) && (fabsf(y) < 1e9f) && (fabsf(x) < 1e9f);
  bool zin0 = finite && z0 >= 0 && z0 < D, zin1 = finite && z0 + 1 >= 0 && z0 + 1 < D;
  bool yin0 = y0 >= 0 && y0 < H, yin1 = y0 + 1 >= 0 && y0 + 1 < H;
  bool xin0 = x0 >= 0 && x0 < W, xin1 = x0 + 1 >= 0 && x0 + 1 < W;
  int base = (z0 * H + y0) * W + x0;
  int HW = H * W;
  float wxy00 = __fmul_rn(wx0, wy0), wxy10 = __fmul_rn(wx1, wy0), wxy01 = __fmul_rn(wx0, wy1),
        wxy11 = __fmul_rn(wx1, wy1);
  s.off[0] = base;              s.w[0] = __fmul_rn(wxy00, wz0);
  s.off[1] = base + 1;          s.w[1] = __fmul_rn(wxy10, wz0);
  s.off[2] = base + W;          s.w[2] = __fmul_rn(wxy01, wz0);
  s.off[3] = base + W + 1;      s.w[3] = __fmul_rn(wxy11, wz0);
  s.off[4] = base + HW;         s.w[4] = __fmul_rn(wxy00, wz1);
  s.off[5] = base + HW + 1;     s.w[5] = __fmul_rn(wxy10, wz1);
  s.off[6] = base + HW + W;     s.w[6] = __fmul_rn(wxy01, wz1);
  s.off[7] = base + HW + W + 1; s.w[7] = __fmul_rn(wxy11, wz1);
  s.mask = (unsigned)(zin0 && yin0 && xin0) | ((unsigned)(zin0 && yin0 && xin1) << 1) |
           ((unsigned)(zin0 && yin1 && xin0) << 2) | ((unsigned)(zin0 && yin1 && xin1) << 3) |
           ((unsigned)(zin1 && yin0 && xin0) << 4) | ((unsigned)(zin1 && yin0 && xin1) << 5) |
           ((unsigned)(zin1 && yin1 && xin0) << 6) | ((unsigned)(zin1 && yin1 && xin1) << 7);
}

__device__ __forceinline__ float tri_gather(const TriSample& s, const float* __restrict__ plane) {
  float acc = 0.0f;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (s.mask & (1u << c)) acc = __fadd_rn(acc, __fmul_rn(__ldg(plane + s.off[c]), s.w[c]));
  }
  return acc;
}

// ---------------------------------------------------------------------------------------------
// nn.Upsample(scale_factor=2, trilinear, align_corners=True) source index / lambda for one axis
// (torch ATen/native/UpSample.h:277-296 and 451-475): ratio=(S-1)/(2S-1) in fp32,
// real = ratio*dst, i0 = (int)real, lambda1 = clamp(real - i0, 0, 1), i1 = i0 + (i0 < S-1).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void up2_index(int dst, int S, float ratio, int& i0, int& i1, float& l1) {
  float real = __fmul_rn(ratio, (float)dst);
  i0 = min((int)real, S - 1);
  l1 = fminf(fmaxf(__fsub_rn(real, (float)i0), 0.0f), 1.0f);
  i1 = i0 + (i0 < S - 1 ? 1 : 0);
}
__host__ __device__ inline float up2_ratio(int S) { return (2 * S > 1) ? (float)(S - 1) / (float)(2 * S - 1) : 0.0f; }

__device__ __forceinline__ float lrelu01(float v) { return v >= 0.0f ? v : 0.1f * v; }

}  // namespace smile
