// Conv3d 3x3x3 / pad 1, fp32 SIMT with TMA-staged input tiles -- the fast path of a8 Encoder and
// a4 CWM convolutions (reference ModeT/models.py:119-151, 186-228, 250-254) for W % 4 == 0.
//
//   * The input halo tile of CIC channels is one 5-D TMA box [W+8, TH+2, TD+2, CIC, 1] per K-chunk
//     (out-of-volume elements zero-filled by TMA == the conv's zero padding), double buffered on
//     two mbarriers so the next chunk lands while the current one is multiplied.  The box starts
//     4 floats left of the tile because TMA needs a 16-byte aligned start coordinate.
//   * A thread owns V consecutive depths x CO output channels of one (h, w); accumulators are
//     packed channel pairs and every multiply-add is fma.rn.f32x2 with the input value broadcast:
//     per (ci, kh, kw) it issues V+2 scalar LDS + 3*CO/4 broadcast LDS.128 for 3*V*CO/2 FFMA2.
//   * When the input is a raw conv output (in_stats != NULL) its InstanceNorm + LeakyReLU(0.1)
//     is applied in place on the landed tile (in-volume elements only: the padding must stay 0
//     in the activation domain); per-(b,c) fp64 sum / sum-of-squares of this layer's raw output
//     are reduced warp -> CTA -> one atomicAdd per channel, as in conv.cu.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <mutex>

#pragma once
#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace conv_tma_detail {


__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}

// FW == 0: a warp covers TWL (16 or 32) voxels of W and 32 / TWL rows -- W is cut into TWL-wide tiles.
// FW  > 0: "flat" tile for volumes whose rows are not a multiple of 16/32 voxels (W = 20, 40, 80 of the LPBA
//          pyramid): the tile is the FULL row of FW voxels times THREADS / FW rows and thread t owns voxel
//          (t / FW, t % FW), so 94 % of the lanes do useful work where W-tiling reaches 62-83 %.
template <int CO, int V, int TWL, int NW, int CIC, int FW = 0>
struct TCfg {
  static constexpr int LH = 32 / TWL;
  static constexpr int TH = FW > 0 ? (NW * 32) / FW : NW * LH;
  static constexpr int TD = V;
  static constexpr int TWP = (FW > 0 ? FW : TWL) + 8;     // box row: floats w0-4 .. w0+TWL+3
  static constexpr int PLANE = (TH + 2) * TWP;            // floats per (c, z)
  static constexpr int CH = (TD + 2) * PLANE;             // floats per channel
  static constexpr int IN_ELEMS = CIC * CH;
  static constexpr int IN_BYTES = IN_ELEMS * 4;
  static constexpr int IN_STRIDE = (IN_BYTES + 127) / 128 * 128;
  static constexpr int W_ELEMS = CIC * 27 * CO;
  static constexpr int W_STRIDE = (W_ELEMS * 4 + 127) / 128 * 128;
  static constexpr int THREADS = NW * 32;
  static constexpr int OFF_IN = 0;
  static constexpr int OFF_W = 2 * IN_STRIDE;
  static constexpr int OFF_BAR = OFF_W + 2 * W_STRIDE;
  static constexpr int OFF_RED = OFF_BAR + 64;            // double [NW][CO][2]
  static constexpr int OFF_MR = OFF_RED + NW * CO * 2 * 8;  // float [Cin][2] rstd, -mean * rstd
  static int smem_bytes(int Cin) { return OFF_MR + 2 * Cin * 4 + 16; }
};

template <int CO, int V, int TWL, int NW, int CIC, bool NORM, int FW = 0>
__global__ void __launch_bounds__(NW * 32, (NW * 32 <= 128) ? 4 : ((CO == 4 && CIC == 1) ? 3 : 2))
conv3d_tma_kernel(const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ weight,
                  const float* __restrict__ bias, float* __restrict__ out, const double* __restrict__ in_stats,
                  double* __restrict__ out_stats, int Cin, int Cout, int D, int H, int W, int tiles_h, int tiles_w,
                  int act_out, float eps) {
  using C = TCfg<CO, V, TWL, NW, CIC, FW>;
  constexpr int TH = C::TH, TD = C::TD, TWP = C::TWP, LH = C::LH;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + C::OFF_BAR;
  double* s_red = reinterpret_cast<double*>(smem + C::OFF_RED);
  float* s_mr = reinterpret_cast<float*>(smem + C::OFF_MR);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // 16-wide tiles put two rows in a warp; rows 2 apart (pitch 24 floats -> 48 = 16 mod 32 banks) keep the
  // two half-warps on disjoint shared-memory banks, adjacent rows would 2-way conflict
  // flat tiles: threads past FW * TH shadow voxel (0, 0) and store nothing
  const bool active = FW == 0 || tid < FW * TH;
  const int tx = FW > 0 ? (active ? tid % (FW > 0 ? FW : 1) : 0) : lane % TWL;
  const int ty = FW > 0 ? (active ? tid / (FW > 0 ? FW : 1) : 0)
                        : ((LH == 2) ? ((warp >> 1) * 4 + (warp & 1) + 2 * (lane / TWL)) : (warp * LH + lane / TWL));
  int t = blockIdx.x;
  const int tw_i = t % tiles_w;
  t /= tiles_w;
  const int th_i = t % tiles_h;
  const int td_i = t / tiles_h;
  const int d0 = td_i * TD, h0 = th_i * TH, w0 = FW > 0 ? 0 : tw_i * TWL;
  const int co0 = blockIdx.y * CO;
  const int b = blockIdx.z;
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int nchunks = (Cin + CIC - 1) / CIC;

  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_expect_tx(bar0, C::IN_BYTES);
    tma_load_5d(sbase + C::OFF_IN, &tm_in, bar0, w0 - 4, h0 - 1, d0 - 1, 0, b);
  }
  if (NORM) {
    for (int c = tid; c < Cin; c += C::THREADS) {
      const double s = in_stats[((long long)b * Cin + c) * 2], ss = in_stats[((long long)b * Cin + c) * 2 + 1];
      const double mean = s / (double)N;
      const double var = fmax(ss / (double)N - mean * mean, 0.0);
      const float rstd = (float)(1.0 / sqrt(var + (double)eps));
      s_mr[2 * c] = rstd;
      s_mr[2 * c + 1] = -(float)mean * rstd;
    }
  }

  // weights of one K-chunk: global [co][ci][27] -> shared [ci][27][co], as fire-and-forget 4-byte cp.async
  // copies (zero-filled outside the tensor) so the global-load latency overlaps the multiply phase
  auto stage_weights = [&](int chunk, int buf) {
    const uint32_t sw = smem_u32(smem + C::OFF_W + buf * C::W_STRIDE);
    const int ci0 = chunk * CIC;
    for (int e = tid; e < C::W_ELEMS; e += C::THREADS) {
      const int co = e / (CIC * 27);
      const int rem = e - co * (CIC * 27);  // ci_local * 27 + tap: contiguous in global memory
      const int ci = ci0 + rem / 27;
      const bool ok = ci < Cin && co0 + co < Cout;
      const float* src = weight + (ok ? ((long long)(co0 + co) * Cin + ci0) * 27 + rem : 0);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(sw + 4 * (rem * CO + co)), "l"(src),
                   "r"(ok ? 4 : 0)
                   : "memory");
    }
  };
  stage_weights(0, 0);
  asm volatile("cp.async.wait_all;" ::: "memory");

  float2 acc[V][CO / 2];
#pragma unroll
  for (int cp = 0; cp < CO / 2; ++cp) {
    const float b0 = (co0 + 2 * cp < Cout) ? __ldg(bias + co0 + 2 * cp) : 0.f;
    const float b1 = (co0 + 2 * cp + 1 < Cout) ? __ldg(bias + co0 + 2 * cp + 1) : 0.f;
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v][cp] = make_float2(b0, b1);
  }
  __syncthreads();  // barriers initialised, s_mr + weights of chunk 0 visible

  // in-volume window of this tile in tile coordinates (for the normalise-on-load pass)
  const int zlo = max(0, 1 - d0), zhi = min(TD + 2, D - d0 + 1);
  const int ylo = max(0, 1 - h0), yhi = min(TH + 2, H - h0 + 1);
  const int xlo = max(0, 4 - w0), xhi = min(TWP, W - w0 + 4);

  for (int chunk = 0; chunk < nchunks; ++chunk) {
    const int buf = chunk & 1;
    if (chunk + 1 < nchunks) {
      // buffer buf^1 was last read in iteration chunk-1, which ended with __syncthreads
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        mbar_expect_tx(bar0 + 8 * (buf ^ 1), C::IN_BYTES);
        tma_load_5d(sbase + C::OFF_IN + (buf ^ 1) * C::IN_STRIDE, &tm_in, bar0 + 8 * (buf ^ 1), w0 - 4, h0 - 1, d0 - 1,
                    (chunk + 1) * CIC, b);
      }
      stage_weights(chunk + 1, buf ^ 1);
    }
    mbar_wait(bar0 + 8 * buf, (chunk >> 1) & 1);
    float* s_in = reinterpret_cast<float*>(smem + C::OFF_IN + buf * C::IN_STRIDE);
    if (NORM) {
      // x <- LeakyReLU((x - mean_c) * rstd_c) on the in-volume part of the tile.  A thread owns one
      // float4 column of the tile (fixed x mask) and walks rows RPP at a time.
      constexpr int R4 = TWP / 4;
      constexpr int RPP = C::THREADS / R4;
      constexpr int ROWS = CIC * (TD + 2) * (TH + 2);
      const int tx4 = tid % R4, trow = tid / R4;
      unsigned xm = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) xm |= (unsigned)((4 * tx4 + i >= xlo) && (4 * tx4 + i < xhi)) << i;
      if (trow < RPP) {
        float4* s4 = reinterpret_cast<float4*>(s_in) + tx4;
        for (int row = trow; row < ROWS; row += RPP) {
          const int y = row % (TH + 2);
          const int rz = row / (TH + 2);
          const int z = rz % (TD + 2);
          const int ci = chunk * CIC + rz / (TD + 2);
          if (ci < Cin && z >= zlo && z < zhi && y >= ylo && y < yhi) {
            const float2 mr = *reinterpret_cast<const float2*>(s_mr + 2 * ci);  // (rstd, -mean * rstd)
            float4 v = s4[row * R4];
            v.x = fmaf(v.x, mr.x, mr.y);
            v.y = fmaf(v.y, mr.x, mr.y);
            v.z = fmaf(v.z, mr.x, mr.y);
            v.w = fmaf(v.w, mr.x, mr.y);
            v.x = fmaxf(v.x, 0.1f * v.x);
            v.y = fmaxf(v.y, 0.1f * v.y);
            v.z = fmaxf(v.z, 0.1f * v.z);
            v.w = fmaxf(v.w, 0.1f * v.w);
            if (xm != 0xFu) {
              if (!(xm & 1u)) v.x = 0.f;
              if (!(xm & 2u)) v.y = 0.f;
              if (!(xm & 4u)) v.z = 0.f;
              if (!(xm & 8u)) v.w = 0.f;
            }
            s4[row * R4] = v;
          }
        }
      }
    }
    __syncthreads();  // transformed tile + this chunk's weights visible to every warp

    const float* swb = reinterpret_cast<const float*>(smem + C::OFF_W + buf * C::W_STRIDE);
#pragma unroll 1
    for (int c = 0; c < CIC; ++c) {
      const float* sc = s_in + c * C::CH + ty * TWP + tx + 3;  // column of voxel (w - 1)
      const float* wc = swb + c * 27 * CO;
#pragma unroll
      for (int kh = 0; kh < 3; ++kh) {
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) {
          float xin[V + 2];
#pragma unroll
          for (int z = 0; z < V + 2; ++z) xin[z] = sc[z * C::PLANE + kh * TWP + kw];
#pragma unroll
          for (int kd = 0; kd < 3; ++kd) {
            float2 wv[CO / 2];
            const float4* wp = reinterpret_cast<const float4*>(wc + (kd * 9 + kh * 3 + kw) * CO);
#pragma unroll
            for (int i = 0; i < CO / 4; ++i) {
              const float4 w4 = wp[i];
              wv[2 * i] = make_float2(w4.x, w4.y);
              wv[2 * i + 1] = make_float2(w4.z, w4.w);
            }
#pragma unroll
            for (int v = 0; v < V; ++v) {
              const float2 xb = make_float2(xin[v + kd], xin[v + kd]);
#pragma unroll
              for (int cp = 0; cp < CO / 2; ++cp) acc[v][cp] = fma2(xb, wv[cp], acc[v][cp]);
            }
          }
        }
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");  // next chunk's weights have landed (issued before the math)
    __syncthreads();  // everyone is done with buffer `buf` before it is refilled two chunks later
  }

  // ---- epilogue: store (+ optional LeakyReLU) and InstanceNorm statistics of the raw output ----
  const int gh = h0 + ty, gw = w0 + tx;
  const bool hw_ok = active && gh < H && gw < W;
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
          const float val = (co & 1) ? acc[v][co / 2].y : acc[v][co / 2].x;
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
        s_red[(warp * CO + co) * 2] = (double)psum[co];
        s_red[(warp * CO + co) * 2 + 1] = (double)psq[co];
      }
    }
    __syncthreads();
    if (tid < CO * 2) {
      const int co = tid >> 1, which = tid & 1;
      if (co0 + co < Cout) {
        double tot = 0.0;
#pragma unroll
        for (int wi = 0; wi < NW; ++wi) tot += s_red[(wi * CO + co) * 2 + which];
        atomicAdd(out_stats + ((long long)b * Cout + co0 + co) * 2 + which, tot);
      }
    }
  }
}

inline PFN_cuTensorMapEncodeTiled get_encode() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  });
  return fn;
}

template <int CO, int V, int TWL, int NW, int CIC, int FW = 0>
int launch_tcfg(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps, cudaStream_t st) {
  using C = TCfg<CO, V, TWL, NW, CIC, FW>;
  CUtensorMap tm;
  const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)Cin, (cuuint64_t)B};
  const cuuint64_t str[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4,
                             (cuuint64_t)Cin * D * H * W * 4};
  const cuuint32_t box[5] = {(cuuint32_t)C::TWP, (cuuint32_t)(C::TH + 2), (cuuint32_t)(C::TD + 2), (cuuint32_t)CIC, 1};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult rc = get_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(in), dims, str, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("conv3d(TMA): cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return SMILE_ERR_CUDA;
  }
  const int tiles_d = ceil_div(D, C::TD), tiles_h = ceil_div(H, C::TH), tiles_w = FW > 0 ? 1 : ceil_div(W, TWL);
  const int smem = C::smem_bytes(Cin);
  dim3 grid(tiles_d * tiles_h * tiles_w, ceil_div(Cout, CO), B);
  auto run = [&](auto kern) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) {
      set_error("conv3d(TMA): cannot reserve %d B of shared memory: %s", smem, cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
    kern<<<grid, C::THREADS, smem, st>>>(tm, weight, bias, out, in_stats, out_stats, Cin, Cout, D, H, W, tiles_h, tiles_w,
                                         act_out, eps);
    return check_launch("conv3d(TMA)");
  };
  if (in_stats != nullptr) return run(conv3d_tma_kernel<CO, V, TWL, NW, CIC, true, FW>);
  return run(conv3d_tma_kernel<CO, V, TWL, NW, CIC, false, FW>);
}

}  // namespace conv_tma_detail
}  // namespace smile
