// Conv3d weight gradient on the tensor cores (bf16 operands, fp32 accumulation in TMEM) -- the weight-gradient half of
// the bf16 training mode (BASELINE.json configs[2..3]); reference: autograd of nn.Conv3d, ModeT/models.py:127,143,253.
//
//   d_w[co][ci][t] = sum_{b,v} d_out[b,co,v] * in[b,ci,v + off(t)],   d_b[co] = sum d_out
//
// As a GEMM the reduction runs over the POSITIONS (K), the rows are (input channel, tap) pairs (M = 4 channels x 27 taps =
// 108 of 128 rows) and the columns the output channels (N <= 64 per CTA):
//     D[(ci, t)][co] += A[(ci, t)][k] * B[co][k],   A = im2col(in) built in shared memory, B = d_out.
// Both operands are K-major with the positions contiguous -- the NCDHW order of the tensors -- so B is a straight
// fp32 -> bf16 conversion of 8-position groups and a row of A is the same data shifted by the tap offset; the shifts are not
// 16-byte multiples, which is why A is materialised (8 positions = one 16-byte core-matrix row per store) instead of being
// aliased by descriptors as in the forward kernels.
//
// A CTA owns one (4-channel, <= 64-output-channel) tile and a contiguous range of 64-position chunks.  Four builder warps
// fill a ring of four operand stages -- lane = (8-position group, channel): the 3 x 3 rows x 10 values its group touches are
// loaded once and the 27 shifted windows are cut from registers -- and a fifth warp's elected thread issues the four K = 16
// MMAs of each stage; tcgen05.commit hands the stage back.  The accumulator stays in TMEM for the whole range and is
// drained once into d_w with atomics (the ranges of one tile are spread over several CTAs).
//
// Why it exists: the SIMT weight-gradient kernels march a 4-row x 32-column tile per (input channel, four output channels)
// -- on the coarse levels (10..40 voxels wide) 70 % of the lanes idle and the activations are re-read Cout / 4 times:
// 128->128 @10x12x10 ran at 4 TFLOP/s, and weight gradients were 5.9 of the 22 ms training step.
#include <cuda_bf16.h>

#include <cstdint>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

constexpr int KC = 64;                 // positions per chunk (four K = 16 MMAs)
constexpr int KG = KC / 8;             // 8-position groups per chunk
constexpr int CIT = 4;                 // input channels per CTA tile
constexpr int MROWS = 128;             // MMA M; rows (ci, tap) = ci * 27 + tap, 108 used
constexpr int NMAX = 64;               // output channels per CTA tile
constexpr int STAGES = 4;              // one per builder warp
constexpr int A_BYTES = KG * MROWS * 16;          // 16 KB
constexpr int B_BYTES = KG * NMAX * 16;           // 8 KB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int OFF_BAR = STAGES * STAGE_BYTES;     // full[4], free[4], done
constexpr int OFF_TMEM = OFF_BAR + 9 * 8 + 8;
constexpr int OFF_DB = OFF_TMEM + 16;             // float[NMAX]
constexpr int SMEM = OFF_DB + NMAX * 4 + 16;
constexpr int THREADS = (STAGES + 1) * 32;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
  }
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint64_t desc64(uint32_t lo) {   // SBO = 128 B, descriptor version 1 in the high word
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(lo), "r"(0x4008u));
  return d;
}

struct WDims {
  int B, Cin, Cout, D, H, W;
  int HW, N;            // voxels per channel
  int P;                // B * N positions
  int chunks;           // ceil(P / 64)
  int chunks_per_cta;
  int ntile;            // output channels of this launch's tiles (multiple of 16, <= 64)
};

__global__ void __launch_bounds__(THREADS, 2) conv3d_wgrad_tc_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                                  float* __restrict__ dw, float* __restrict__ db,
                                                                  const WDims dm) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  float* s_db = reinterpret_cast<float*>(smem + OFF_DB);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ci0 = blockIdx.y * CIT, co0 = blockIdx.z * dm.ntile;
  const int NT = dm.ntile;
  const int c_begin = blockIdx.x * dm.chunks_per_cta;
  const int c_end = min(dm.chunks, c_begin + dm.chunks_per_cta);
  const int nchunks = c_end - c_begin;
  if (nchunks <= 0) return;   // uniform per CTA

  if (tid == 0) {
    for (int i = 0; i < 9; ++i) mbar_init(sbase + OFF_BAR + 8 * i, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == STAGES) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(NMAX));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  // rows 108..127 of every A stage stay zero; partial tiles write zeros themselves
  for (int i = tid; i < STAGES * STAGE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  if (tid < NMAX) s_db[tid] = 0.f;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;
  const uint32_t bar_full = sbase + OFF_BAR, bar_free = sbase + OFF_BAR + 32, bar_done = sbase + OFF_BAR + 64;

  if (warp < STAGES) {
    // ------------------------------------------------------------------ builders
    const int D = dm.D, H = dm.H, W = dm.W, HW = dm.HW, N = dm.N;
    uint8_t* stA = smem + warp * STAGE_BYTES;
    uint8_t* stB = stA + A_BYTES;
    const int g = lane >> 2, cl = lane & 3;          // A: 8-position group, local input channel
    const int ci = ci0 + cl;
    const bool ci_ok = ci < dm.Cin;
    const bool do_db = db != nullptr && blockIdx.y == 0;
    float dbacc[NMAX * KG / 32];     // per-lane partial sums of d_out (bias gradient), one per item slot
#pragma unroll
    for (int t = 0; t < NMAX * KG / 32; ++t) dbacc[t] = 0.f;
    int use = 0;
    for (int q = warp; q < nchunks; q += STAGES, ++use) {
      if (use > 0) mbar_wait(bar_free + 8 * warp, (uint32_t)((use - 1) & 1));
      const int p0 = (c_begin + q) * KC;
      // ---- A: rows (cl, tap) of group g
      {
        const int pg = p0 + g * 8;
        int w = pg % W, t = pg / W;
        int h = t % H;
        t /= H;
        int d = t % D, b = t / D;
        uint4* dst = reinterpret_cast<uint4*>(stA + g * (MROWS * 16) + (cl * 27) * 16);
        if (!ci_ok) {
          // channel beyond Cin (partial tile): its rows keep the zeros written at kernel start
        } else if (pg + 7 < dm.P && w + 8 <= W) {
          // the eight positions share a row: 3 x 3 rows of 10 values -- all 90 loads are issued before the first is
          // converted (the builders are latency-bound) -- and the 27 shifted windows are cut from registers
          const float* xc = x + ((long long)b * dm.Cin + ci) * N;
          float v[9][10];
          const bool vec = (W & 7) == 0;     // rows and groups start on 32-byte boundaries: two 16-byte loads + the halo
#pragma unroll
          for (int r9 = 0; r9 < 9; ++r9) {
            const int dz = d + r9 / 3 - 1, hy = h + r9 % 3 - 1;
            const bool rok = (unsigned)dz < (unsigned)D && (unsigned)hy < (unsigned)H;
            const float* row = xc + (rok ? (dz * HW + hy * W) : 0);
            if (vec) {
              float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
              if (rok) {
                a = __ldg(reinterpret_cast<const float4*>(row + w));
                c = __ldg(reinterpret_cast<const float4*>(row + w + 4));
              }
              v[r9][0] = (rok && w > 0) ? __ldg(row + w - 1) : 0.f;
              v[r9][1] = a.x; v[r9][2] = a.y; v[r9][3] = a.z; v[r9][4] = a.w;
              v[r9][5] = c.x; v[r9][6] = c.y; v[r9][7] = c.z; v[r9][8] = c.w;
              v[r9][9] = (rok && w + 8 < W) ? __ldg(row + w + 8) : 0.f;
            } else {
#pragma unroll
              for (int j = 0; j < 10; ++j) {
                const int wx = w + j - 1;
                v[r9][j] = (rok && (unsigned)wx < (unsigned)W) ? __ldg(row + wx) : 0.f;
              }
            }
          }
#pragma unroll
          for (int r9 = 0; r9 < 9; ++r9) {
#pragma unroll
            for (int kw = 0; kw < 3; ++kw)
              dst[r9 * 3 + kw] = make_uint4(pack_bf16(v[r9][kw], v[r9][kw + 1]), pack_bf16(v[r9][kw + 2], v[r9][kw + 3]),
                                            pack_bf16(v[r9][kw + 4], v[r9][kw + 5]), pack_bf16(v[r9][kw + 6], v[r9][kw + 7]));
          }
        } else {
          // generic: the group crosses a row / plane / batch end, or the end of the data
          int bj[8], dj[8], hj[8], wj[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            bj[j] = b; dj[j] = d; hj[j] = h; wj[j] = w;
            if (++w == W) {
              w = 0;
              if (++h == H) {
                h = 0;
                if (++d == D) {
                  d = 0;
                  ++b;
                }
              }
            }
          }
#pragma unroll 1
          for (int tap = 0; tap < 27; ++tap) {
            const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int dz = dj[j] + kd - 1, hy = hj[j] + kh - 1, wx = wj[j] + kw - 1;
              const bool ok = ci_ok && pg + j < dm.P && (unsigned)dz < (unsigned)D && (unsigned)hy < (unsigned)H &&
                              (unsigned)wx < (unsigned)W;
              v[j] = ok ? __ldg(x + ((long long)bj[j] * dm.Cin + ci) * N + dz * HW + hy * W + wx) : 0.f;
            }
            dst[tap] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
          }
        }
      }
      // ---- B: d_out rows (output channel n) of the eight groups; items (n, g) over the lanes, four items per round with
      // all eight 16-byte loads of a round in flight (d_out is streamed from DRAM: these are the coldest loads of the step)
#pragma unroll 1
      for (int t0 = 0; t0 < NT * KG / 32; t0 += 4) {
        float v[4][8];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int it = lane + 32 * (t0 + u);
          const int n = it / KG, gg = it % KG;
          const int co = co0 + n;
          const int pg = p0 + gg * 8;
          if (t0 + u >= NT * KG / 32) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[u][j] = 0.f;
          } else if (co < dm.Cout && pg + 7 < dm.P && (N & 7) == 0) {      // one batch item, 32-byte aligned
            const int bb = pg / N, r = pg - bb * N;
            const float4* src = reinterpret_cast<const float4*>(dy + ((long long)bb * dm.Cout + co) * N + r);
            const float4 a = __ldg(src), c = __ldg(src + 1);
            v[u][0] = a.x; v[u][1] = a.y; v[u][2] = a.z; v[u][3] = a.w;
            v[u][4] = c.x; v[u][5] = c.y; v[u][6] = c.z; v[u][7] = c.w;
          } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int p = pg + j;
              const int bb = p / N, r = p - bb * N;
              v[u][j] = (co < dm.Cout && p < dm.P) ? __ldg(dy + ((long long)bb * dm.Cout + co) * N + r) : 0.f;
            }
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (t0 + u >= NT * KG / 32) break;
          const int it = lane + 32 * (t0 + u);
          const int n = it / KG, gg = it % KG;
          *reinterpret_cast<uint4*>(stB + gg * (NT * 16) + n * 16) = make_uint4(
              pack_bf16(v[u][0], v[u][1]), pack_bf16(v[u][2], v[u][3]), pack_bf16(v[u][4], v[u][5]), pack_bf16(v[u][6], v[u][7]));
          // bias gradient: item t of a lane always belongs to output channel (lane + 32 t) / 8
          dbacc[t0 + u] += ((v[u][0] + v[u][1]) + (v[u][2] + v[u][3])) + ((v[u][4] + v[u][5]) + (v[u][6] + v[u][7]));
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * warp);
    }
    if (do_db) {       // the eight lanes of a (n) row -> one shared-memory add per output channel and warp
#pragma unroll
      for (int t = 0; t < NMAX * KG / 32; ++t) {
        float sacc = dbacc[t];
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 1);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 2);
        sacc += __shfl_xor_sync(0xffffffffu, sacc, 4);
        const int n = (lane + 32 * t) / KG;
        if ((lane & 7) == 0 && n < NT) atomicAdd(s_db + n, sacc);
      }
    }
    // ---- drain: TMEM lane = row (ci, tap); this warp owns lanes 32 * warp ..
    mbar_wait(bar_done, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int m = warp * 32 + lane;
    const int mc = m / 27, mt = m - mc * 27;
    const bool row_ok = m < CIT * 27 && ci0 + mc < dm.Cin;
    for (int n0 = 0; n0 < NT; n0 += 16) {
      uint32_t r[16];
      const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)n0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row_ok) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const int co = co0 + n0 + i;
          if (co < dm.Cout) atomicAdd(dw + ((long long)co * dm.Cin + ci0 + mc) * 27 + mt, __uint_as_float(r[i]));
        }
      }
    }
  } else {
    // ------------------------------------------------------------------ issuer warp (one elected thread)
    uint32_t leader = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      // instruction descriptor: D = f32, A = B = bf16, M = 128, N = NT
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(MROWS >> 4) << 24);
      for (int q = 0; q < nchunks; ++q) {
        const int st = q % STAGES, use = q / STAGES;
        mbar_wait(bar_full + 8 * st, (uint32_t)(use & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a16 = (sbase + st * STAGE_BYTES) >> 4, b16 = (sbase + st * STAGE_BYTES + A_BYTES) >> 4;
#pragma unroll
        for (int kk = 0; kk < KC / 16; ++kk) {
          // K chunk j of MMA kk = 8-position group 2 kk + j: LBO = one group = MROWS (NT) rows x 16 B
          const uint64_t da = desc64((a16 + (uint32_t)(2 * kk * MROWS)) | ((uint32_t)MROWS << 16));
          const uint64_t dbv = desc64((b16 + (uint32_t)(2 * kk * NT)) | ((uint32_t)NT << 16));
          if (q == 0 && kk == 0)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(dbv), "r"(idesc) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem), "l"(da), "l"(dbv), "r"(idesc) : "memory");
        }
        // the stage is free again when these MMAs have read it
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_free + 8 * st)
                     : "memory");
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_done) : "memory");
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == STAGES) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(NMAX));
  if (db != nullptr && blockIdx.y == 0 && tid < NT && co0 + tid < dm.Cout) atomicAdd(db + co0 + tid, s_db[tid]);
}

}  // namespace

// dw / db must already be zeroed by the caller.  *handled = false for shapes the kernel does not take.
int launch_conv3d_wgrad_tc(const float* x, const float* dy, float* dw, float* db, int B, int Cin, int Cout, int D, int H, int W,
                           cudaStream_t st, bool* handled) {
  *handled = false;
  const long long N = (long long)D * H * W;
  if (N * B >= (1LL << 31) || N * Cin * B >= (1LL << 31) || N * Cout * B >= (1LL << 31)) return SMILE_OK;
  *handled = true;
  WDims dm;
  dm.B = B; dm.Cin = Cin; dm.Cout = Cout; dm.D = D; dm.H = H; dm.W = W;
  dm.HW = H * W; dm.N = (int)N; dm.P = (int)(N * B);
  dm.chunks = ceil_div(dm.P, KC);
  dm.ntile = Cout >= NMAX ? NMAX : ceil_div(Cout, 16) * 16;
  const int ci_tiles = ceil_div(Cin, CIT), co_tiles = ceil_div(Cout, dm.ntile);
  // K splits: about four CTAs per SM over the whole launch, at least 8 chunks each (the drain is 108 x N atomics per CTA)
  int splits = ceil_div(4 * kNumSMs, ci_tiles * co_tiles);
  if (splits > dm.chunks / 8) splits = dm.chunks / 8;
  if (splits < 1) splits = 1;
  dm.chunks_per_cta = ceil_div(dm.chunks, splits);
  splits = ceil_div(dm.chunks, dm.chunks_per_cta);
  if (ci_tiles > 65535 || co_tiles > 65535) {
    *handled = false;
    return SMILE_OK;
  }
  cudaError_t e = cudaFuncSetAttribute(conv3d_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
  if (e != cudaSuccess) {
    set_error("conv3d_wgrad(tcgen05): cannot reserve %d B of shared memory: %s", SMEM, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  conv3d_wgrad_tc_kernel<<<dim3(splits, ci_tiles, co_tiles), THREADS, SMEM, st>>>(x, dy, dw, db, dm);
  return check_launch("conv3d_wgrad(tcgen05)");
}

}  // namespace smile
