// Conv3d weight gradient, TMA-fed (training path; reference: autograd of nn.Conv3d, ModeT/models.py:127,143,253).
//
//   d_w[co][ci][t] = sum_{b,v} d_out[b,co,v] * in[b,ci,v + off(t)],   d_b[co] = sum d_out
//
// CTA = 4 warps = 4 rows x 32 columns of the (H, W) plane marching a depth chunk for one input channel and
// four output channels.  Per depth step ONE thread issues two TMA boxes -- the (4+2) x 40 input halo plane
// (zero-filled outside the volume = the conv's padding) and the 4 x 4 x 32 output-gradient plane -- into an
// 8-deep mbarrier ring, six planes ahead, so the multiply loop touches no global memory: 27 immediate-offset LDS
// + 4 LDS feed 54 packed FFMA2 (channel pairs) per voxel.  Partial sums live in registers for the whole chunk
// and are reduced warp -> CTA -> one atomicAdd per element at the end.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
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
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int CO = 4;                      // output channels per CTA
constexpr int NSL = 8;                     // ring depth
constexpr int AHEAD = 6;                   // planes in flight
constexpr int XPW = 40;                    // input plane row pitch: floats w0-4 .. w0+35
constexpr int X_BYTES = 6 * XPW * 4;       // 960
constexpr int X_STRIDE = 1024;
constexpr int G_BYTES = CO * 4 * 32 * 4;   // 2048
constexpr int SLOT = X_STRIDE + G_BYTES;   // 3072
constexpr int OFF_BAR = NSL * SLOT;
constexpr int SMEM = OFF_BAR + NSL * 8;
// two-row variant (a thread owns the voxel pair (h, w), (h + 1, w): 8 rows x 32 columns per CTA)
constexpr int X2_BYTES = 10 * XPW * 4;      // 1600
constexpr int X2_STRIDE = 1664;             // 128-byte multiple
constexpr int G2_BYTES = CO * 8 * 32 * 4;   // 4096
constexpr int SLOT2 = X2_STRIDE + G2_BYTES; // 5760
constexpr int OFF_BAR2 = NSL * SLOT2;
constexpr int SMEM2 = OFF_BAR2 + NSL * 8;

__global__ void __launch_bounds__(128) conv3d_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                               const __grid_constant__ CUtensorMap tm_g,
                                                               float* __restrict__ dw, float* __restrict__ db, int Cin,
                                                               int Cout, int D, int dchunk, int tiles_h, int tiles_w) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ float s_part[4][CO * 27 + CO];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int t = blockIdx.x;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  const int dc = t / tiles_h;
  const int b = blockIdx.z;
  const int cog = blockIdx.y / Cin, ci = blockIdx.y % Cin;
  const int co0 = cog * CO;
  const int h0 = th * 4, w0 = tw * 32;
  const int d_begin = dc * dchunk, d_end = min(D, d_begin + dchunk);
  const int p_first = d_begin - 1, p_last = d_end;  // planes staged: input needs d-1 .. d+1

  if (tid == 0) {
    for (int i = 0; i < NSL; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int p) {  // called by thread 0 only
    const int q = p - p_first;
    const uint32_t slot = sbase + (q % NSL) * SLOT, bar = bar0 + 8 * (q % NSL);
    mbar_expect_tx(bar, X_BYTES + G_BYTES);
    tma_load_5d(slot, &tm_x, bar, w0 - 4, h0 - 1, p, ci, b);
    tma_load_5d(slot + X_STRIDE, &tm_g, bar, w0, h0, p, co0, b);
  };
  if (tid == 0)
    for (int p = p_first; p <= p_last && p < p_first + AHEAD; ++p) issue(p);
  int next_p = p_first + AHEAD;

  float2 acc[CO / 2][27];
  float bsum[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) bsum[c] = 0.f;
#pragma unroll
  for (int c = 0; c < CO / 2; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[c][k] = make_float2(0.f, 0.f);

  // planes d_begin-1 and d_begin must have landed before the first step
  mbar_wait(bar0, 0);
  if (p_first + 1 <= p_last) mbar_wait(bar0 + 8, 0);
  for (int d = d_begin; d < d_end; ++d) {
    const int q1 = d + 1 - p_first;  // ring index of plane d+1
    mbar_wait(bar0 + 8 * (q1 % NSL), (q1 / NSL) & 1);
    const int qm = q1 - 2, q0 = q1 - 1;
    const float* gs = reinterpret_cast<const float*>(smem + (q0 % NSL) * SLOT + X_STRIDE) + warp * 32 + lane;
    float g[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) g[c] = gs[c * 128];
#pragma unroll
    for (int c = 0; c < CO; ++c) bsum[c] += g[c];
    const float2 g01 = make_float2(g[0], g[1]), g23 = make_float2(g[2], g[3]);
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float* pl = reinterpret_cast<const float*>(smem + ((qm + kd) % NSL) * SLOT) + warp * XPW + lane + 3;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float xv = pl[(j / 3) * XPW + (j % 3)];
        const float2 xb2 = make_float2(xv, xv);
        acc[0][kd * 9 + j] = fma2(xb2, g01, acc[0][kd * 9 + j]);
        acc[1][kd * 9 + j] = fma2(xb2, g23, acc[1][kd * 9 + j]);
      }
    }
    __syncthreads();  // every warp is done with plane d-1
    // one new plane per step keeps AHEAD planes in flight; its slot held plane (next_p - NSL) <= d - 1: free
    if (tid == 0 && next_p <= p_last) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(next_p);
    }
    ++next_p;
  }
#pragma unroll
  for (int c = 0; c < CO; ++c) {
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float v = warp_sum((c & 1) ? acc[c / 2][k].y : acc[c / 2][k].x);
      if (lane == 0) s_part[warp][c * 27 + k] = v;
    }
    const float v = warp_sum(bsum[c]);
    if (lane == 0) s_part[warp][CO * 27 + c] = v;
  }
  __syncthreads();
  for (int i = tid; i < CO * 27 + CO; i += 128) {
    const float v = s_part[0][i] + s_part[1][i] + s_part[2][i] + s_part[3][i];
    if (i < CO * 27) {
      const int c = i / 27, k = i % 27;
      if (co0 + c < Cout) atomicAdd(dw + ((long long)(co0 + c) * Cin + ci) * 27 + k, v);
    } else if (ci == 0 && db != nullptr) {
      const int c = i - CO * 27;
      if (co0 + c < Cout) atomicAdd(db + co0 + c, v);
    }
  }
}

// Two voxels (h, w), (h + 1, w) per thread; otherwise the kernel above.
__global__ void __launch_bounds__(128) conv3d_wgrad_tma2_kernel(const __grid_constant__ CUtensorMap tm_x,
                                                               const __grid_constant__ CUtensorMap tm_g,
                                                               float* __restrict__ dw, float* __restrict__ db, int Cin,
                                                               int Cout, int D, int dchunk, int tiles_h, int tiles_w) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ float s_part[4][CO * 27 + CO];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR2;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int t = blockIdx.x;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  const int dc = t / tiles_h;
  const int b = blockIdx.z;
  const int cog = blockIdx.y / Cin, ci = blockIdx.y % Cin;
  const int co0 = cog * CO;
  const int h0 = th * 8, w0 = tw * 32;
  const int d_begin = dc * dchunk, d_end = min(D, d_begin + dchunk);
  const int p_first = d_begin - 1, p_last = d_end;  // planes staged: input needs d-1 .. d+1

  if (tid == 0) {
    for (int i = 0; i < NSL; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](int p) {  // called by thread 0 only
    const int q = p - p_first;
    const uint32_t slot = sbase + (q % NSL) * SLOT2, bar = bar0 + 8 * (q % NSL);
    mbar_expect_tx(bar, X2_BYTES + G2_BYTES);
    tma_load_5d(slot, &tm_x, bar, w0 - 4, h0 - 1, p, ci, b);
    tma_load_5d(slot + X2_STRIDE, &tm_g, bar, w0, h0, p, co0, b);
  };
  if (tid == 0)
    for (int p = p_first; p <= p_last && p < p_first + AHEAD; ++p) issue(p);
  int next_p = p_first + AHEAD;

  float2 acc[CO / 2][27];
  float bsum[CO];
#pragma unroll
  for (int c = 0; c < CO; ++c) bsum[c] = 0.f;
#pragma unroll
  for (int c = 0; c < CO / 2; ++c)
#pragma unroll
    for (int k = 0; k < 27; ++k) acc[c][k] = make_float2(0.f, 0.f);

  // planes d_begin-1 and d_begin must have landed before the first step
  mbar_wait(bar0, 0);
  if (p_first + 1 <= p_last) mbar_wait(bar0 + 8, 0);
  for (int d = d_begin; d < d_end; ++d) {
    const int q1 = d + 1 - p_first;  // ring index of plane d+1
    mbar_wait(bar0 + 8 * (q1 % NSL), (q1 / NSL) & 1);
    const int qm = q1 - 2, q0 = q1 - 1;
    // the pair's output gradients: rows 2 * warp (A) and 2 * warp + 1 (B) of the 8-row tile, [co][row][col]
    const float* gs = reinterpret_cast<const float*>(smem + (q0 % NSL) * SLOT2 + X2_STRIDE) + (2 * warp) * 32 + lane;
    float gA[CO], gB[CO];
#pragma unroll
    for (int c = 0; c < CO; ++c) {
      gA[c] = gs[c * 256];
      gB[c] = gs[c * 256 + 32];
    }
#pragma unroll
    for (int c = 0; c < CO; ++c) bsum[c] += gA[c] + gB[c];
    const float2 gA01 = make_float2(gA[0], gA[1]), gA23 = make_float2(gA[2], gA[3]);
    const float2 gB01 = make_float2(gB[0], gB[1]), gB23 = make_float2(gB[2], gB[3]);
    // 36 input values (4 rows x 3 columns per plane) feed 108 packed FMAs: the one-voxel kernel reads 27 + 4 values per 54
    // and is bound by the issue of its shared-memory loads
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float* pl = reinterpret_cast<const float*>(smem + ((qm + kd) % NSL) * SLOT2) + (2 * warp) * XPW + lane + 3;
      float xr[4][3];
#pragma unroll
      for (int rr = 0; rr < 4; ++rr)
#pragma unroll
        for (int kw = 0; kw < 3; ++kw) xr[rr][kw] = pl[rr * XPW + kw];
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float xa = xr[j / 3][j % 3], xb = xr[j / 3 + 1][j % 3];
        acc[0][kd * 9 + j] = fma2(make_float2(xa, xa), gA01, acc[0][kd * 9 + j]);
        acc[1][kd * 9 + j] = fma2(make_float2(xa, xa), gA23, acc[1][kd * 9 + j]);
        acc[0][kd * 9 + j] = fma2(make_float2(xb, xb), gB01, acc[0][kd * 9 + j]);
        acc[1][kd * 9 + j] = fma2(make_float2(xb, xb), gB23, acc[1][kd * 9 + j]);
      }
    }
    __syncthreads();  // every warp is done with plane d-1
    // one new plane per step keeps AHEAD planes in flight; its slot held plane (next_p - NSL) <= d - 1: free
    if (tid == 0 && next_p <= p_last) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(next_p);
    }
    ++next_p;
  }
#pragma unroll
  for (int c = 0; c < CO; ++c) {
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float v = warp_sum((c & 1) ? acc[c / 2][k].y : acc[c / 2][k].x);
      if (lane == 0) s_part[warp][c * 27 + k] = v;
    }
    const float v = warp_sum(bsum[c]);
    if (lane == 0) s_part[warp][CO * 27 + c] = v;
  }
  __syncthreads();
  for (int i = tid; i < CO * 27 + CO; i += 128) {
    const float v = s_part[0][i] + s_part[1][i] + s_part[2][i] + s_part[3][i];
    if (i < CO * 27) {
      const int c = i / 27, k = i % 27;
      if (co0 + c < Cout) atomicAdd(dw + ((long long)(co0 + c) * Cin + ci) * 27 + k, v);
    } else if (ci == 0 && db != nullptr) {
      const int c = i - CO * 27;
      if (co0 + c < Cout) atomicAdd(db + co0 + c, v);
    }
  }
}

PFN_cuTensorMapEncodeTiled get_encode() {
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

bool encode5(CUtensorMap* tm, const float* base, int W, int H, int D, int C, int B, const cuuint32_t (&box)[5]) {
  const cuuint64_t dims[5] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t str[4] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4,
                             (cuuint64_t)C * D * H * W * 4};
  const cuuint32_t es[5] = {1, 1, 1, 1, 1};
  CUresult rc = get_encode()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, str, box, es,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("conv3d_wgrad(TMA): cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return false;
  }
  return true;
}

}  // namespace

// dw / db must already be zeroed by the caller (launch_conv3d_wgrad does it).
int launch_conv3d_wgrad_tma(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int D, int H,
                            int W, cudaStream_t st, bool* handled) {
  *handled = false;
  if (W % 4 != 0 || W < 16 || get_encode() == nullptr) return SMILE_OK;
  const long long groups = (long long)ceil_div(Cout, CO) * Cin;
  if (groups > 65535) return SMILE_OK;
  *handled = true;
  static const bool one_row = getenv("SMILE_WGRAD_ONE_ROW") != nullptr;   // A/B knob
  const bool two = H >= 8 && !one_row;
  const int TROWS = two ? 8 : 4;
  CUtensorMap tx, tg;
  const cuuint32_t bx[5] = {XPW, (cuuint32_t)(TROWS + 2), 1, 1, 1}, bg[5] = {32, (cuuint32_t)TROWS, 1, CO, 1};
  if (!encode5(&tx, x, W, H, D, Cin, B, bx) || !encode5(&tg, dy, W, H, D, Cout, B, bg)) return SMILE_ERR_CUDA;
  const int tiles_h = ceil_div(H, TROWS), tiles_w = ceil_div(W, 32);
  const long long per_plane = (long long)tiles_h * tiles_w * groups * B;
  int chunks = (int)ceil_div_ll(4LL * kNumSMs * 4, per_plane);
  if (chunks < 1) chunks = 1;
  int dchunk = ceil_div(D, chunks);
  if (dchunk < 8) dchunk = D < 8 ? D : 8;
  chunks = ceil_div(D, dchunk);
  const int smem = two ? SMEM2 : SMEM;
  auto kern = two ? conv3d_wgrad_tma2_kernel : conv3d_wgrad_tma_kernel;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) {
    set_error("conv3d_wgrad(TMA): cannot reserve %d B of shared memory: %s", smem, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  dim3 grid(tiles_h * tiles_w * chunks, (unsigned)groups, B);
  kern<<<grid, 128, smem, st>>>(tx, tg, dw, db, Cin, Cout, D, dchunk, tiles_h, tiles_w);
  return check_launch("conv3d_wgrad(TMA)");
}

}  // namespace smile
