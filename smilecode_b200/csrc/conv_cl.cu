// Conv3d 3x3x3 / pad 1 for SMALL volumes with many channels (W <= 16: the coarsest encoder level
// 10x12x10 with 64/128 channels, reference ModeT/models.py:216-218, and the tiny levels of small inputs).
//
// With ~1e3 voxels and 1e2 channels the work is weight-heavy, so the lanes of a warp run over OUTPUT
// CHANNELS instead of W: lane L owns channels (co0 + L, co0 + L + 32) as one packed fp32x2 accumulator
// per voxel of its row, a warp owns one output row (all W voxels), a CTA four rows of one plane and 64
// output channels.  Per (ci, kd, kh) a thread reads the input row once as broadcast LDS.128 and three
// weight taps per channel (conflict-free LDS: per-channel stride 109 words), then issues 3*W FFMA2.
// Inputs and weights of CIC = 4 channels are staged per step; the producer's InstanceNorm + LeakyReLU
// is applied while staging, out-of-volume elements are written as 0 (== zero padding).
#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

constexpr int CIC = 4;            // input channels per step
constexpr int R = 4;              // output rows per CTA (one per warp)
constexpr int COB = 64;           // output channels per CTA
constexpr int WS = CIC * 27 + 1;  // per-output-channel weight stride in shared memory (odd: conflict-free)

template <int WMAX, bool NORM>
__global__ void __launch_bounds__(128) conv3d_cl_kernel(const float* __restrict__ in, const float* __restrict__ weight,
                                                        const float* __restrict__ bias, float* __restrict__ out,
                                                        const double* __restrict__ in_stats, double* __restrict__ out_stats,
                                                        int Cin, int Cout, int D, int H, int W, int tiles_h, int act_out,
                                                        float eps) {
  constexpr int XP = WMAX + 4;                 // row pitch; voxel w lives at index w + 1
  constexpr int IN_ELEMS = CIC * 3 * (R + 2) * XP;
  extern __shared__ __align__(16) float smem[];
  float* s_in = smem;                          // [2][CIC][3][R+2][XP]
  float* s_w = s_in + 2 * IN_ELEMS;            // [2][COB][WS]
  float* s_mr = s_w + 2 * COB * WS;            // [Cin][2] rstd, -mean*rstd
  __shared__ double s_red[R][COB][2];

  const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
  const int d = blockIdx.x / tiles_h, h0 = (blockIdx.x % tiles_h) * R;
  const int co0 = blockIdx.y * COB;
  const int b = blockIdx.z;
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int nchunks = (Cin + CIC - 1) / CIC;

  if (NORM) {
    for (int c = tid; c < Cin; c += 128) {
      const double s = in_stats[((long long)b * Cin + c) * 2], ss = in_stats[((long long)b * Cin + c) * 2 + 1];
      const double mean = s / (double)N;
      const double var = fmax(ss / (double)N - mean * mean, 0.0);
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      s_mr[2 * c] = rstd;
      s_mr[2 * c + 1] = -(float)mean * rstd;
    }
    __syncthreads();
  }

  const float* inb = in + (long long)b * Cin * N;
  constexpr int ITERS = (IN_ELEMS + 127) / 128;
  float pre[ITERS];  // input elements of the next step, in flight while the current step is multiplied

  // phase 1 of staging: global loads of the next step's inputs into registers, weights by cp.async
  auto stage_issue = [&](int chunk, int buf) {
    const int ci0 = chunk * CIC;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int e = tid + it * 128;
      const int x = e % XP;
      int t = e / XP;
      const int y = t % (R + 2);
      t /= (R + 2);
      const int z = t % 3;
      const int ci = ci0 + t / 3;
      const int gd = d - 1 + z, gh = h0 - 1 + y, gw = x - 1;
      float v = 0.f;
      if (e < IN_ELEMS && ci < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
        v = __ldg(inb + (long long)ci * N + (long long)gd * HW + gh * W + gw);
        if (NORM) {
          v = fmaf(v, s_mr[2 * ci], s_mr[2 * ci + 1]);
          v = fmaxf(v, 0.1f * v);
        }
      }
      pre[it] = v;
    }
    // weights: global [co][ci][27] -> shared [co_local][ci_local*27 + tap]: straight 4-byte cp.async copies
    // (zero-filled past the valid run), no registers, all in flight at once
    // consecutive lanes copy consecutive floats of one channel's contiguous (ci, tap) run: coalesced
    constexpr int RUN = CIC * 27;              // 108 floats per output channel and step
    constexpr int PER = COB * RUN / 128;       // 54 elements per thread
    const int nvalid = min(CIC, Cin - ci0) * 27;
    const uint32_t dst0 = (uint32_t)__cvta_generic_to_shared(s_w + buf * COB * WS);
#pragma unroll 6
    for (int i = 0; i < PER; ++i) {
      const int e = tid + i * 128;
      const int col = e / RUN, rem = e - col * RUN;
      const int co = co0 + col;
      const int ok = (co < Cout) && (rem < nvalid);
      const float* src = weight + ((long long)(ok ? co : 0) * Cin + ci0) * 27 + (ok ? rem : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst0 + 4 * (col * WS + rem)), "l"(src),
                   "r"(ok ? 4 : 0)
                   : "memory");
    }
  };
  // phase 2: registers -> shared, wait for the cp.async group
  auto stage_commit = [&](int buf) {
    float* si = s_in + buf * IN_ELEMS;
#pragma unroll
    for (int it = 0; it < ITERS; ++it) {
      const int e = tid + it * 128;
      if (e < IN_ELEMS) si[e] = pre[it];
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
  };

  float2 acc[WMAX];
  {
    const float b0 = (co0 + lane < Cout) ? __ldg(bias + co0 + lane) : 0.f;
    const float b1 = (co0 + lane + 32 < Cout) ? __ldg(bias + co0 + lane + 32) : 0.f;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) acc[w] = make_float2(b0, b1);
  }

  stage_issue(0, 0);
  stage_commit(0);
  __syncthreads();
  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int buf = chunk & 1;
    const bool more = chunk + 1 < nchunks;
    if (more) stage_issue(chunk + 1, buf ^ 1);  // the other buffer was released by the previous barrier
    const float* si = s_in + buf * IN_ELEMS + r * XP;
    const float* sw0 = s_w + buf * COB * WS + lane * WS;
    const float* sw1 = sw0 + 32 * WS;
#pragma unroll 1
    for (int c = 0; c < CIC; ++c) {
#pragma unroll
      for (int kd = 0; kd < 3; ++kd) {
#pragma unroll
        for (int kh = 0; kh < 3; ++kh) {
          float x[XP];
          const float4* xr = reinterpret_cast<const float4*>(si + ((c * 3 + kd) * (R + 2) + kh) * XP);
#pragma unroll
          for (int i = 0; i < XP / 4; ++i) {
            const float4 v = xr[i];
            x[4 * i] = v.x;
            x[4 * i + 1] = v.y;
            x[4 * i + 2] = v.z;
            x[4 * i + 3] = v.w;
          }
          const int wo = c * 27 + kd * 9 + kh * 3;
#pragma unroll
          for (int kw = 0; kw < 3; ++kw) {
            const float2 wv = make_float2(sw0[wo + kw], sw1[wo + kw]);
#pragma unroll
            for (int w = 0; w < WMAX; ++w) acc[w] = fma2(make_float2(x[w + kw], x[w + kw]), wv, acc[w]);
          }
        }
      }
    }
    if (more) stage_commit(buf ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const int gh = h0 + r;
  const int c_a = co0 + lane, c_b = co0 + lane + 32;
  float ps0 = 0.f, pq0 = 0.f, ps1 = 0.f, pq1 = 0.f;
  if (gh < H) {
    float* oa = out + ((long long)b * Cout + c_a) * N + (long long)d * HW + gh * W;
    float* obp = out + ((long long)b * Cout + c_b) * N + (long long)d * HW + gh * W;
#pragma unroll
    for (int w = 0; w < WMAX; ++w) {
      if (w < W) {
        const float v0 = acc[w].x, v1 = acc[w].y;
        if (c_a < Cout) {
          ps0 += v0;
          pq0 = fmaf(v0, v0, pq0);
          oa[w] = act_out ? lrelu01(v0) : v0;
        }
        if (c_b < Cout) {
          ps1 += v1;
          pq1 = fmaf(v1, v1, pq1);
          obp[w] = act_out ? lrelu01(v1) : v1;
        }
      }
    }
  }
  if (out_stats != nullptr) {
    s_red[r][lane][0] = (double)ps0;
    s_red[r][lane][1] = (double)pq0;
    s_red[r][lane + 32][0] = (double)ps1;
    s_red[r][lane + 32][1] = (double)pq1;
    __syncthreads();
    const int col = tid >> 1, which = tid & 1;  // 64 channels x {sum, sumsq}
    if (co0 + col < Cout) {
      double tot = 0.0;
#pragma unroll
      for (int rr = 0; rr < R; ++rr) tot += s_red[rr][col][which];
      atomicAdd(out_stats + ((long long)b * Cout + co0 + col) * 2 + which, tot);
    }
  }
}

template <int WMAX>
int launch_w(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
             double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps, cudaStream_t st) {
  constexpr int XP = WMAX + 4;
  const int tiles_h = ceil_div(H, R);
  const size_t smem = (size_t)(2 * CIC * 3 * (R + 2) * XP + 2 * COB * WS + 2 * Cin) * sizeof(float);
  dim3 grid(D * tiles_h, ceil_div(Cout, COB), B);
  auto run = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv3d(small): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
    kern<<<grid, 128, smem, st>>>(in, weight, bias, out, in_stats, out_stats, Cin, Cout, D, H, W, tiles_h, act_out, eps);
    return check_launch("conv3d(small)");
  };
  if (in_stats != nullptr) return run(conv3d_cl_kernel<WMAX, true>);
  return run(conv3d_cl_kernel<WMAX, false>);
}

}  // namespace

int launch_conv3d_small(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                        double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                        cudaStream_t st, bool* handled) {
  *handled = false;
  if (W > 16 || Cout < 16) return SMILE_OK;  // narrow outputs waste the channel lanes: generic kernel
  *handled = true;
  if (W <= 4) return launch_w<4>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
  if (W <= 8) return launch_w<8>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
  if (W <= 12) return launch_w<12>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
  return launch_w<16>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
}

}  // namespace smile
