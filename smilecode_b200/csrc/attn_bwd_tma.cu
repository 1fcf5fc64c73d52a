// ModeTransformer backward for the heads==1, head_dim==6 levels (the two large ones), TMA-fed.
// Reference: autograd of ModeT/models.py:308-334 (in ModeT-cu: modet_bw + softmax / matmul backward).
//
//   p = softmax_t(scale <q, k[n+off(t)]> + rpb[t]);  gV[t] = <g, off(t)>;  dl[t] = p[t] (gV[t] - sum_t' p[t'] gV[t'])
//   dq = scale * sum_t dl[t] k[n+off(t)]      d_rpb += dl      dl is stored tap-major for the dk gather
//
// A CTA marches 8 rows x 32 columns along D.  Per plane one thread issues three TMA boxes (key plane with halo,
// query plane, upstream-gradient plane) into a 6-deep mbarrier ring, four planes ahead; the 27 key rows of a
// voxel are read twice from shared memory (logits, then dq) as conflict-free 8-byte LDS, all math is packed
// fma.rn.f32x2 where it pairs naturally.  Zero-filled halo rows give logit = rpb and contribute nothing to dq,
// exactly like the zero padding of models.py:319.
#include <cuda.h>
#include <cudaTypedefs.h>

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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
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
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

constexpr int TH = 8, TW = 32, HD = 6;
constexpr int KW = TW + 4;                       // key row: voxels w0-2 .. w0+33 (16-byte aligned TMA start)
constexpr int K_BYTES = (TH + 2) * KW * HD * 4;  // 8640
constexpr int K_STRIDE = 8704;
constexpr int Q_BYTES = TH * TW * HD * 4;        // 6144
constexpr int G_BYTES = 3 * TH * TW * 4;         // 3072
constexpr int SLOT = K_STRIDE + Q_BYTES + G_BYTES;
constexpr int NSL = 6, AHEAD = 4;
constexpr int OFF_BAR = NSL * SLOT;
constexpr int OFF_RPB = OFF_BAR + 64;
constexpr int SMEM = OFF_RPB + 128;
constexpr float kLog2e = 1.4426950408889634f;

__global__ void __launch_bounds__(TH * 32, 2)
attn_bwd_march_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_q,
                      const __grid_constant__ CUtensorMap tm_g, const float* __restrict__ rpb, float* __restrict__ dq,
                      float* __restrict__ dl_out, float* __restrict__ drpb, int D, int H, int W, int dchunk, int tiles_h,
                      int tiles_w, float scale) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + OFF_BAR;
  float* s_rpb = reinterpret_cast<float*>(smem + OFF_RPB);
  const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
  int t = blockIdx.x;
  const int tw = t % tiles_w;
  t /= tiles_w;
  const int th = t % tiles_h;
  const int dc = t / tiles_h;
  const int b = blockIdx.y;
  const int h0 = th * TH, w0 = tw * TW;
  const int h = h0 + r, w = w0 + lane;
  const bool valid = h < H && w < W;
  const int HW = H * W;
  const long long N = (long long)D * HW;
  const int d_begin = dc * dchunk, d_end = min(D, d_begin + dchunk);
  const int p_first = d_begin - 1, p_last = d_end;

  if (tid == 0) {
    for (int i = 0; i < NSL; ++i) mbar_init(bar0 + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < 28) s_rpb[tid] = (tid < 27 && rpb != nullptr) ? rpb[tid] * kLog2e : 0.f;
  __syncthreads();
  auto issue = [&](int p) {
    const int q = p - p_first;
    const uint32_t slot = sbase + (q % NSL) * SLOT, bar = bar0 + 8 * (q % NSL);
    mbar_expect_tx(bar, K_BYTES + Q_BYTES + G_BYTES);
    tma_load_4d(slot, &tm_k, bar, (w0 - 2) * HD, h0 - 1, p, b);
    tma_load_4d(slot + K_STRIDE, &tm_q, bar, w0 * HD, h0, p, b);
    tma_load_4d(slot + K_STRIDE + Q_BYTES, &tm_g, bar, w0, h0, p, b * 3);
  };
  if (tid == 0)
    for (int p = p_first; p <= p_last && p < p_first + AHEAD; ++p) issue(p);
  int next_p = p_first + AHEAD;

  const float qs = scale * kLog2e;
  float* dqb = dq + (long long)b * N * HD;
  float* dlb = dl_out + (long long)b * 27 * N;

  mbar_wait(bar0, 0);
  if (p_first + 1 <= p_last) mbar_wait(bar0 + 8, 0);
  for (int d = d_begin; d < d_end; ++d) {
    const int q1 = d + 1 - p_first;
    mbar_wait(bar0 + 8 * (q1 % NSL), (q1 / NSL) & 1);
    const int qm = q1 - 2, q0 = q1 - 1;
    const uint8_t* s0 = smem + (q0 % NSL) * SLOT;
    const float2* qp = reinterpret_cast<const float2*>(s0 + K_STRIDE) + (r * TW + lane) * 3;
    const float2 qa = qp[0], qb = qp[1], qc = qp[2];
    const float* gp = reinterpret_cast<const float*>(s0 + K_STRIDE + Q_BYTES) + r * TW + lane;
    const float g0 = gp[0], g1 = gp[TH * TW], g2 = gp[2 * TH * TW];
    // logits in the log2 domain
    float lg[27];
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float2* kp = reinterpret_cast<const float2*>(smem + ((qm + kd) % NSL) * SLOT) + (r * KW + lane + 1) * 3;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float2* kr = kp + ((j / 3) * KW + (j % 3)) * 3;
        float2 a = mul2(qa, kr[0]);
        a = fma2(qb, kr[1], a);
        a = fma2(qc, kr[2], a);
        lg[kd * 9 + j] = fmaf(a.x + a.y, qs, s_rpb[kd * 9 + j]);
      }
    }
    float m = lg[0];
#pragma unroll
    for (int k = 1; k < 27; ++k) m = fmaxf(m, lg[k]);
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      lg[k] = ex2(lg[k] - m);
      sum += lg[k];
    }
    const float inv = 1.0f / sum;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      lg[k] *= inv;
      const float gv = g0 * (float)(k / 9 - 1) + g1 * (float)((k / 3) % 3 - 1) + g2 * (float)(k % 3 - 1);
      dot = fmaf(lg[k], gv, dot);
    }
    const long long vo = (long long)d * HW + (long long)h * W + w;
#pragma unroll
    for (int k = 0; k < 27; ++k) {
      const float gv = g0 * (float)(k / 9 - 1) + g1 * (float)((k / 3) % 3 - 1) + g2 * (float)(k % 3 - 1);
      lg[k] = lg[k] * (gv - dot);  // d_logit
      if (valid) dlb[(long long)k * N + vo] = lg[k];
    }
    // dq = scale * sum_t dl[t] * k[n + off(t)]
    float2 da = make_float2(0.f, 0.f), dbv = da, dcv = da;
#pragma unroll
    for (int kd = 0; kd < 3; ++kd) {
      const float2* kp = reinterpret_cast<const float2*>(smem + ((qm + kd) % NSL) * SLOT) + (r * KW + lane + 1) * 3;
#pragma unroll
      for (int j = 0; j < 9; ++j) {
        const float2* kr = kp + ((j / 3) * KW + (j % 3)) * 3;
        const float2 l2 = make_float2(lg[kd * 9 + j], lg[kd * 9 + j]);
        da = fma2(l2, kr[0], da);
        dbv = fma2(l2, kr[1], dbv);
        dcv = fma2(l2, kr[2], dcv);
      }
    }
    if (valid) {
      float2* o = reinterpret_cast<float2*>(dqb + vo * HD);
      o[0] = make_float2(da.x * scale, da.y * scale);
      o[1] = make_float2(dbv.x * scale, dbv.y * scale);
      o[2] = make_float2(dcv.x * scale, dcv.y * scale);
    }
    __syncthreads();
    if (tid == 0 && next_p <= p_last) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      issue(next_p);
    }
    ++next_p;
  }
}

// d_rpb[t] = sum over batch and voxels of the stored d_logits plane t (27 streaming reductions)
__global__ void __launch_bounds__(256) drpb_reduce_kernel(const float* __restrict__ dl, float* __restrict__ drpb, int B,
                                                          long long N) {
  const int tap = blockIdx.y;
  float acc = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* p = dl + ((long long)b * 27 + tap) * N;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < N; i += (long long)gridDim.x * blockDim.x)
      acc += __ldg(p + i);
  }
  __shared__ float s_w[8];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float tot = 0.f;
    for (int i = 0; i < 8; ++i) tot += s_w[i];
    atomicAdd(drpb + tap, tot);
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

bool encode4(CUtensorMap* map, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides)[3],
             const cuuint32_t (&box)[4]) {
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = get_encode()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                             CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("modet_attn_bwd(TMA): cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return false;
  }
  return true;
}

}  // namespace

// heads == 1, head_dim == 6, W % 4 == 0: computes dq, d_logits (tap-major [B][27][N]) and d_rpb (pre-zeroed).
int launch_attn_bwd_dq_tma(const float* g, const float* q, const float* k, const float* rpb, float* dq, float* dl,
                           float* drpb, int B, int D, int H, int W, float scale, cudaStream_t st, bool* handled) {
  *handled = false;
  if (W % 4 != 0 || W < 8 || get_encode() == nullptr) return SMILE_OK;
  if ((long long)D * H * W * HD >= (1LL << 31)) return SMILE_OK;
  *handled = true;
  CUtensorMap mk, mq, mg;
  const cuuint64_t qk_dims[4] = {(cuuint64_t)W * HD, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t qk_str[3] = {(cuuint64_t)W * HD * 4, (cuuint64_t)H * W * HD * 4, (cuuint64_t)D * H * W * HD * 4};
  const cuuint32_t k_box[4] = {KW * HD, TH + 2, 1, 1}, q_box[4] = {TW * HD, TH, 1, 1};
  const cuuint64_t g_dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * 3};
  const cuuint64_t g_str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
  const cuuint32_t g_box[4] = {TW, TH, 1, 3};
  if (!encode4(&mk, k, qk_dims, qk_str, k_box) || !encode4(&mq, q, qk_dims, qk_str, q_box) ||
      !encode4(&mg, g, g_dims, g_str, g_box))
    return SMILE_ERR_CUDA;
  const int tiles_h = ceil_div(H, TH), tiles_w = ceil_div(W, TW);
  const long long per_plane = (long long)tiles_h * tiles_w * B;
  int chunks = (int)ceil_div_ll(2LL * kNumSMs * 2, per_plane);
  if (chunks < 1) chunks = 1;
  int dchunk = ceil_div(D, chunks);
  if (dchunk < 8) dchunk = D < 8 ? D : 8;
  chunks = ceil_div(D, dchunk);
  cudaError_t e = cudaFuncSetAttribute(attn_bwd_march_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  if (e != cudaSuccess) {
    set_error("modet_attn_bwd(TMA): cannot reserve %d B of shared memory: %s", SMEM, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  dim3 grid(tiles_h * tiles_w * chunks, B);
  attn_bwd_march_kernel<<<grid, TH * 32, SMEM, st>>>(mk, mq, mg, rpb, dq, dl, drpb, D, H, W, dchunk, tiles_h, tiles_w,
                                                     scale);
  int rc = check_launch("modet_attn_bwd(TMA dq)");
  if (rc != SMILE_OK || drpb == nullptr) return rc;
  const long long N = (long long)D * H * W;
  long long gx = ceil_div_ll(N, 256 * 8);
  if (gx > 64) gx = 64;
  drpb_reduce_kernel<<<dim3((unsigned)gx, 27), 256, 0, st>>>(dl, drpb, B, N);
  return check_launch("modet_attn_bwd(d_rpb)");
}

}  // namespace smile
