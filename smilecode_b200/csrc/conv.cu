// Conv3d 3x3x3 / pad 1 building block of a8 Encoder (reference ModeT/models.py:119-151, 186-228)
// and a4 CWM (models.py:250-254), plus the InstanceNorm / LeakyReLU / AvgPool glue around it and
// the CWM softmax-fuse (models.py:268-275).  fp32 SIMT direct convolution (the 1e-4 parity path):
//
//   * a CTA owns a V x TH x TWL output tile for CO output channels; the (V+2)(TH+2)(TWL+2) input
//     halo tile of CIC input channels is staged in shared memory per step (zero fill == padding);
//   * lanes run along W (conflict-free LDS, coalesced 128 B stores), each thread slides a
//     (V+2)-deep register window along D and keeps V x CO accumulators, so one input LDS feeds
//     3*CO FMAs and one broadcast weight LDS.128 feeds 4*V FMAs;
//   * the previous layer's InstanceNorm + LeakyReLU is applied while the tile is staged
//     ("normalise on load"), and this layer's per-(b,c) sum / sum-of-squares are reduced
//     warp -> CTA -> one fp64 atomicAdd per channel, so a ConvInsBlock costs one read of its
//     input and one write of its raw output.
#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace smile {

template <int CO, int V, int TWL, int NW, int CIC>
struct ConvCfg {
  static constexpr int LH = 32 / TWL;       // rows of H covered by one warp
  static constexpr int TH = NW * LH;        // tile extent along H
  static constexpr int TD = V;              // tile extent along D
  static constexpr int TWP = TWL + 2;       // smem row pitch
  static constexpr int IN_ELEMS = CIC * (TD + 2) * (TH + 2) * TWP;
  static constexpr int W_ELEMS = CIC * 27 * CO;
  static constexpr int THREADS = NW * 32;
};

template <int CO, int V, int TWL, int NW, int CIC>
__global__ void __launch_bounds__(NW * 32)
conv3d_kernel(const float* __restrict__ in, const float* __restrict__ weight, const float* __restrict__ bias,
              float* __restrict__ out, const double* __restrict__ in_stats, double* __restrict__ out_stats, int Cin,
              int Cout, int D, int H, int W, int tiles_h, int tiles_w, int act_out, float eps) {
  using Cfg = ConvCfg<CO, V, TWL, NW, CIC>;
  constexpr int TH = Cfg::TH, TD = Cfg::TD, TWP = Cfg::TWP, LH = Cfg::LH;
  extern __shared__ float smem[];
  float* s_in = smem;                        // [CIC][TD+2][TH+2][TWP]
  float* s_w = s_in + Cfg::IN_ELEMS;         // [CIC][27][CO]
  float* s_mr = s_w + Cfg::W_ELEMS;          // [Cin][2] mean, rstd of the producer (if in_stats)
  __shared__ double s_red[NW][CO][2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = lane % TWL, ty = warp * LH + lane / TWL;
  int t = blockIdx.x;
  const int tw_i = t % tiles_w;
  t /= tiles_w;
  const int th_i = t % tiles_h;
  const int td_i = t / tiles_h;
  const int d0 = td_i * TD, h0 = th_i * TH, w0 = tw_i * TWL;
  const int co0 = blockIdx.y * CO;
  const int b = blockIdx.z;
  const long long N = (long long)D * H * W;
  const int HW = H * W;

  if (in_stats != nullptr) {
    for (int c = tid; c < Cin; c += Cfg::THREADS) {
      const double s = in_stats[((long long)b * Cin + c) * 2], ss = in_stats[((long long)b * Cin + c) * 2 + 1];
      const double mean = s / (double)N;
      const double var = fmax(ss / (double)N - mean * mean, 0.0);
      s_mr[2 * c] = (float)mean;
      s_mr[2 * c + 1] = (float)(1.0 / sqrt(var + (double)eps));
    }
  }

  float acc[V][CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) {
    const float bv = (co0 + co < Cout) ? __ldg(bias + co0 + co) : 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v][co] = bv;
  }

  const float* inb = in + (long long)b * Cin * N;
  for (int ci0 = 0; ci0 < Cin; ci0 += CIC) {
    __syncthreads();  // previous step's readers are done (and s_mr is visible on the first pass)
    for (int e = tid; e < Cfg::IN_ELEMS; e += Cfg::THREADS) {
      const int x = e % TWP;
      int r = e / TWP;
      const int y = r % (TH + 2);
      r /= (TH + 2);
      const int z = r % (TD + 2);
      const int c = r / (TD + 2);
      const int gd = d0 - 1 + z, gh = h0 - 1 + y, gw = w0 - 1 + x, ci = ci0 + c;
      float val = 0.f;
      if (ci < Cin && gd >= 0 && gd < D && gh >= 0 && gh < H && gw >= 0 && gw < W) {
        val = __ldg(inb + (long long)ci * N + (long long)gd * HW + gh * W + gw);
        if (in_stats != nullptr) val = lrelu01((val - s_mr[2 * ci]) * s_mr[2 * ci + 1]);
      }
      s_in[e] = val;
    }
    for (int e = tid; e < Cfg::W_ELEMS; e += Cfg::THREADS) {
      const int co = e % CO;
      const int r = e / CO;
      const int tap = r % 27, c = r / 27;
      const int ci = ci0 + c;
      s_w[e] = (ci < Cin && co0 + co < Cout) ? __ldg(weight + ((long long)(co0 + co) * Cin + ci) * 27 + tap) : 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int c = 0; c < CIC; ++c) {
      const float* sc = s_in + c * (TD + 2) * (TH + 2) * TWP + ty * TWP + tx;
      const float* wc = s_w + c * 27 * CO;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float xin[V + 2];
#pragma unroll
          for (int z = 0; z < V + 2; ++z) xin[z] = sc[(z * (TH + 2) + kh) * TWP + kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            float wv[CO];
            const float4* wp = reinterpret_cast<const float4*>(wc + (kd * 9 + kh * 3 + kw) * CO);
#pragma unroll
            for (int i = 0; i < CO / 4; ++i) {
              float4 w4 = wp[i];
              wv[4 * i] = w4.x;
              wv[4 * i + 1] = w4.y;
              wv[4 * i + 2] = w4.z;
              wv[4 * i + 3] = w4.w;
            }
#pragma unroll
            for (int v = 0; v < V; ++v)
#pragma unroll
              for (int co = 0; co < CO; ++co) acc[v][co] = fmaf(xin[v + kd], wv[co], acc[v][co]);
          }
        }
      }
    }
  }

  // ---- epilogue: store (+ optional LeakyReLU) and InstanceNorm statistics of the raw output ----
  const int gh = h0 + ty, gw = w0 + tx;
  const bool hw_ok = gh < H && gw < W;
  float psum[CO], psq[CO];
#pragma unroll
  for (int co = 0; co < CO; ++co) psum[co] = psq[co] = 0.f;
  float* ob = out + ((long long)b * Cout + co0) * N + (long long)gh * W + gw;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int gd = d0 + v;
    if (hw_ok && gd < D) {
#pragma unroll
      for (int co = 0; co < CO; ++co) {
        if (co0 + co < Cout) {
          const float val = acc[v][co];
          psum[co] += val;
          psq[co] = fmaf(val, val, psq[co]);
          ob[(long long)co * N + (long long)gd * HW] = act_out ? lrelu01(val) : val;
        }
      }
    }
  }
  if (out_stats != nullptr) {
#pragma unroll
    for (int co = 0; co < CO; ++co) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        psum[co] += __shfl_xor_sync(0xffffffffu, psum[co], o);
        psq[co] += __shfl_xor_sync(0xffffffffu, psq[co], o);
      }
      if (lane == 0) {
        s_red[warp][co][0] = (double)psum[co];
        s_red[warp][co][1] = (double)psq[co];
      }
    }
    __syncthreads();
    if (tid < CO * 2) {
      const int co = tid >> 1, which = tid & 1;
      if (co0 + co < Cout) {
        double tot = 0.0;
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) tot += s_red[wi][co][which];
        atomicAdd(out_stats + ((long long)b * Cout + co0 + co) * 2 + which, tot);
      }
    }
  }
}

template <int CO, int V, int TWL, int NW, int CIC>
static int launch_cfg(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                      double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                      cudaStream_t st) {
  using Cfg = ConvCfg<CO, V, TWL, NW, CIC>;
  const int tiles_d = ceil_div(D, Cfg::TD), tiles_h = ceil_div(H, Cfg::TH), tiles_w = ceil_div(W, TWL);
  const size_t smem = (size_t)(Cfg::IN_ELEMS + Cfg::W_ELEMS + 2 * Cin) * sizeof(float);
  auto kern = conv3d_kernel<CO, V, TWL, NW, CIC>;
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) {
      set_error("conv3d: cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
  }
  dim3 grid(tiles_d * tiles_h * tiles_w, ceil_div(Cout, CO), B);
  kern<<<grid, Cfg::THREADS, smem, st>>>(in, weight, bias, out, in_stats, out_stats, Cin, Cout, D, H, W, tiles_h, tiles_w,
                                         act_out, eps);
  return check_launch("conv3d");
}

int launch_conv3d_tma(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                      double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                      cudaStream_t st, bool* handled);

int launch_conv3d_small(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                        double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                        cudaStream_t st, bool* handled);

int launch_conv3d(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                  double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                  cudaStream_t st, const float* wprep) {
  {
    // the wide 8-channel layers: fp16-split tensor-core kernel with fp32-class accuracy (conv_march.cu)
    bool split = false;
    int rc = launch_conv3d_march_split(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st, &split);
    if (split) return rc;
  }
  {
    // layers wide enough for an MMA (>= 12 output channels) run on the tensor cores (conv_tc.cu); SMILE_CONV_TC=0
    // keeps everything on the SIMT kernels
    bool tc = false;
    int rc = launch_conv3d_tc(in, weight, wprep, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st, &tc);
    if (tc) return rc;
  }
  {
    static const bool no_tma = getenv("SMILE_CONV_NO_TMA") != nullptr;  // profiling knob: force the generic kernel
    bool handled = false;
    if (!no_tma) {
      int rc = launch_conv3d_tma(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st,
                                 &handled);
      if (handled) return rc;
      rc = launch_conv3d_small(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st,
                               &handled);
      if (handled) return rc;
    }
  }
#define SMILE_CONV(CO, V, TWL, NW, CIC) \
  return launch_cfg<CO, V, TWL, NW, CIC>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st)
  const bool narrow = Cout <= 4;
  if (W >= 24) {
    if (Cin == 1) {
      if (narrow) SMILE_CONV(4, 4, 32, 8, 1);
      SMILE_CONV(8, 4, 32, 8, 1);
    }
    if (narrow) SMILE_CONV(4, 4, 32, 8, 4);
    SMILE_CONV(8, 4, 32, 8, 4);
  } else if (W >= 12) {
    if (narrow) SMILE_CONV(4, 4, 16, 4, 4);
    SMILE_CONV(8, 4, 16, 4, 4);
  } else {
    if (narrow) SMILE_CONV(4, 2, 8, 2, 4);
    SMILE_CONV(8, 2, 8, 2, 4);
  }
#undef SMILE_CONV
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm3d(affine=False, eps) + LeakyReLU(0.1) of a raw conv output from its fp64 sums, and
// (optionally) AvgPool3d(2) of the result for the next pyramid level in the same pass
// (models.py:144-150 and 198, 204, 210, 216).  One thread per 2x2x2 cell.
// ---------------------------------------------------------------------------------------------
template <bool EVEN>
__global__ void __launch_bounds__(256) in_finalize_kernel(const float* __restrict__ raw, const double* __restrict__ stats,
                                                          float* __restrict__ out, float* __restrict__ pooled, int D, int H,
                                                          int W, float eps) {
  const int bc = blockIdx.y;
  const long long N = (long long)D * H * W;
  const double s = stats[2 * bc], ss = stats[2 * bc + 1];
  const double meand = s / (double)N;
  const float mean = (float)meand;
  const float rstd = (float)(1.0 / sqrt(fmax(ss / (double)N - meand * meand, 0.0) + (double)eps));
  const int CD = (D + 1) / 2, CH = (H + 1) / 2, CW = (W + 1) / 2;
  const int PD = D / 2, PH = H / 2, PW = W / 2;
  const float* rb = raw + (long long)bc * N;
  float* ob = out + (long long)bc * N;
  const long long cells = (long long)CD * CH * CW;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += (long long)gridDim.x * blockDim.x) {
    const int cw = (int)(i % CW);
    long long r = i / CW;
    const int ch = (int)(r % CH);
    const int cd = (int)(r / CH);
    float sum = 0.f;
    if (EVEN) {  // all extents even: the cell is four aligned float2 rows -> coalesced 8-byte loads and stores
      const long long o00 = ((long long)(2 * cd) * H + 2 * ch) * W + 2 * cw;
      const long long offs[4] = {o00, o00 + W, o00 + (long long)H * W, o00 + (long long)H * W + W};
      float2 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float2*>(rb + offs[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k].x = lrelu01((v[k].x - mean) * rstd);
        v[k].y = lrelu01((v[k].y - mean) * rstd);
        *reinterpret_cast<float2*>(ob + offs[k]) = v[k];
      }
      // same summation order as the scalar path: (dz, dy, dx) lexicographic
      sum = ((((((v[0].x + v[0].y) + v[1].x) + v[1].y) + v[2].x) + v[2].y) + v[3].x) + v[3].y;
      if (pooled != nullptr) pooled[(long long)bc * PD * PH * PW + ((long long)cd * PH + ch) * PW + cw] = sum * 0.125f;
      continue;
    }
#pragma unroll
    for (int dz = 0; dz < 2; ++dz)
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int d = 2 * cd + dz, h = 2 * ch + dy, w = 2 * cw + dx;
          if (d < D && h < H && w < W) {
            const long long o = ((long long)d * H + h) * W + w;
            const float v = lrelu01((rb[o] - mean) * rstd);
            ob[o] = v;
            sum += v;
          }
        }
    if (pooled != nullptr && cd < PD && ch < PH && cw < PW)
      pooled[(long long)bc * PD * PH * PW + ((long long)cd * PH + ch) * PW + cw] = sum * 0.125f;
  }
}

int launch_in_finalize(const float* raw, const double* stats, float* out, float* pooled, int B, int C, int D, int H, int W,
                       float eps, cudaStream_t st) {
  const long long cells = (long long)((D + 1) / 2) * ((H + 1) / 2) * ((W + 1) / 2);
  long long g = ceil_div_ll(cells, 256);
  const long long cap = 4096;
  dim3 grid((unsigned)(g < cap ? g : cap), B * C);
  const bool even = D % 2 == 0 && H % 2 == 0 && W % 2 == 0 && (reinterpret_cast<uintptr_t>(raw) & 7u) == 0 &&
                    (reinterpret_cast<uintptr_t>(out) & 7u) == 0;
  if (even)
    in_finalize_kernel<true><<<grid, 256, 0, st>>>(raw, stats, out, pooled, D, H, W, eps);
  else
    in_finalize_kernel<false><<<grid, 256, 0, st>>>(raw, stats, out, pooled, D, H, W, eps);
  return check_launch("in_finalize");
}

// ---------------------------------------------------------------------------------------------
// CWM tail (models.py:254, 268-275): softmax over the F weight logits, out = 2 * sum_f field_f * p_f
// fields [B,3F,N], logits [B,F,N] -> out [B,3,N]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cwm_fuse_kernel(const float* __restrict__ fields, const float* __restrict__ logits,
                                                       float* __restrict__ out, int F, long long N) {
  const int b = blockIdx.y;
  const float* fb = fields + (long long)b * 3 * F * N;
  const float* lb = logits + (long long)b * F * N;
  float* ob = out + (long long)b * 3 * N;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < N; p += (long long)gridDim.x * blockDim.x) {
    float m = -INFINITY;
    for (int f = 0; f < F; ++f) m = fmaxf(m, __ldg(lb + (long long)f * N + p));
    float sum = 0.f, a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int f = 0; f < F; ++f) {
      const float e = expf(__ldg(lb + (long long)f * N + p) - m);
      sum += e;
      a0 = fmaf(e, __ldg(fb + (long long)(3 * f) * N + p), a0);
      a1 = fmaf(e, __ldg(fb + (long long)(3 * f + 1) * N + p), a1);
      a2 = fmaf(e, __ldg(fb + (long long)(3 * f + 2) * N + p), a2);
    }
    const float inv = 2.0f / sum;
    ob[p] = a0 * inv;
    ob[N + p] = a1 * inv;
    ob[2 * N + p] = a2 * inv;
  }
}

int launch_cwm_fuse(const float* fields, const float* logits, float* out, int B, int F, long long N, cudaStream_t st) {
  long long g = ceil_div_ll(N, 256);
  const long long cap = (long long)kNumSMs * 16;
  dim3 grid((unsigned)(g < cap ? g : cap), B);
  cwm_fuse_kernel<<<grid, 256, 0, st>>>(fields, logits, out, F, N);
  return check_launch("cwm_fuse");
}

}  // namespace smile
