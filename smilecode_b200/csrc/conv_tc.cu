// Conv3d 3x3x3 / pad 1 on the 5th-generation tensor cores (tcgen05, accumulators in TMEM) with fp32-class
// accuracy -- the a8 Encoder / a4 CWM convolutions with >= 12 output channels (reference ModeT/models.py:119-151,
// 186-228, 250-254).  Same contract as the SIMT kernels (conv.cu): NCDHW fp32 in/out, optional producer
// InstanceNorm + LeakyReLU applied on load, fp64 sum / sum-of-squares of the raw output for the next layer.
//
// Implicit GEMM without im2col (formulation verified in tools/probe/umma_probe.cu):
//   * an output tile is M = 128 CONSECUTIVE positions of the zero-padded (H+2) x (W+2) plane, so a tap (kh, kw) is a
//     constant shift of (kh-1)*(W+2) + (kw-1) positions and (kd) selects one of three staged planes;
//   * activations are staged position-major, 4 fp32 channels = 16 bytes per position: a K-major no-swizzle UMMA
//     descriptor with a 16-byte row pitch whose START ADDRESS carries the tap shift; the second K chunk of an MMA
//     (K = 8 for kind::tf32) is the next block of 4 channels, LBO = one staged plane;
//   * N = NT (16 or 32) output channels per CTA, weights of one 8-channel K stage arrive as one bulk copy of a
//     pre-arranged block (conv3d_tc_prep_kernel).
// Accuracy (tools/probe/tf32x3_probe.cu): a tf32 MMA keeps 11 significant bits per operand and the TMEM
// accumulation error grows with the number of accumulated MMAs.  So (i) every operand is split x = hi + lo
// (hi = x with the low 13 mantissa bits cleared) and each tap computes hi*hi + lo*hi + hi*lo, and (ii) the 27 taps of a
// stage accumulate into FOUR groups of TMEM accumulators (<= 14 MMAs each) that are drained after every stage and
// summed in registers with round-to-nearest fp32 adds.  hi*hi and hi*lo share one MMA (B operand [B_hi | B_lo], N = 2*NT),
// so a tap costs two MMAs and the A_hi tile is read from shared memory once instead of twice.
#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

constexpr int M = 128;        // positions per tile (UMMA M)
constexpr int KC = 8;         // input channels per stage (UMMA K for tf32)
constexpr int G = 4;          // TMEM accumulator groups per tile (each 2*NT columns: [hi*hi + lo*hi | hi*lo])
__device__ __constant__ int kGroupStart[G + 1] = {0, 7, 14, 21, 27};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell); base offset 0, no swizzle
  return d;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// weight [Cout][Cin][27]  ->  wprep[ntile][stage][tap][cb(2)][hi/lo][NT][4]; channel ci = stage*8 + cb*4 + j,
// output channel co = ntile*NT + n; zero outside the tensor.  Per (tap, cb) the NT hi rows are followed by the NT lo
// rows, so one B descriptor with N = 2*NT covers [B_hi | B_lo] and one with N = NT covers B_hi alone.
__global__ void conv3d_tc_prep_kernel(const float* __restrict__ w, float* __restrict__ wprep, int Cout, int Cin, int NT,
                                      int nstage, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long t = e;
  const int j = (int)(t % 4); t /= 4;
  const int n = (int)(t % NT); t /= NT;
  const int hl = (int)(t % 2); t /= 2;
  const int cb = (int)(t % 2); t /= 2;
  const int tap = (int)(t % 27); t /= 27;
  const int stage = (int)(t % nstage); t /= nstage;
  const int ntile = (int)t;
  const int ci = stage * KC + cb * 4 + j, co = ntile * NT + n;
  float v = 0.f;
  if (ci < Cin && co < Cout) v = w[((long long)co * Cin + ci) * 27 + tap];
  const float h = tf32_hi(v);
  wprep[e] = hl == 0 ? h : v - h;
}

constexpr int THREADS = 256;  // 8 warps: warps w and w + 4 share TMEM lane quadrant w and split the NT columns

template <int NT, bool NORM, int SU>
__global__ void __launch_bounds__(THREADS)
conv3d_tc_kernel(const float* __restrict__ in, const float* __restrict__ wprep, const float* __restrict__ bias,
                 float* __restrict__ out, const double* __restrict__ in_stats, double* __restrict__ out_stats, int Cin,
                 int Cout, int D, int H, int W, int tiles_plane, int nstage, int SEG, int act_out, float eps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // [hi/lo][3 planes][2 channel blocks][SEG positions][4 channels]
  float4* sA = reinterpret_cast<float4*>(smem);
  const int a_plane = 3 * 2 * SEG;                       // float4 elements of one hi or lo copy
  float* sB = reinterpret_cast<float*>(smem + (size_t)2 * a_plane * 16);   // [27][2][hi/lo][NT][4]
  constexpr int B_BYTES = 2 * 27 * 2 * NT * 16;
  uint8_t* tail = reinterpret_cast<uint8_t*>(sB) + B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);    // [0] weights landed, [1] MMAs done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 16);
  double* s_part = reinterpret_cast<double*>(tail + 32); // [THREADS / NT parts][NT][2]
  float* s_mr = reinterpret_cast<float*>(tail + 32 + THREADS * 2 * 8);  // [nstage * 8][2] rstd, -mean*rstd

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wp = W + 2, HW = H * W;
  const long long N = (long long)D * HW;
  const int d = blockIdx.x / tiles_plane, tile = blockIdx.x - d * tiles_plane;
  const int ntile = blockIdx.y, co0 = ntile * NT;
  const int b = blockIdx.z;
  const int q_lo = Wp + 1, q_hi = H * Wp + W;             // padded index of voxel (0,0) and (H-1,W-1)
  const int q0 = q_lo + tile * M;
  const float inv_wp = 1.0f / (float)Wp;

  if (tid == 0) {
    mbar_init(smem_u32(bars), 1);
    mbar_init(smem_u32(bars + 1), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(G * 2 * NT));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (NORM) {
    for (int c = tid; c < nstage * KC; c += THREADS) {
      float rstd = 0.f, shift = 0.f;
      if (c < Cin) {
        const double s = in_stats[((long long)b * Cin + c) * 2], ss = in_stats[((long long)b * Cin + c) * 2 + 1];
        const double mean = s / (double)N;
        const double var = fmax(ss / (double)N - mean * mean, 0.0);
        rstd = (float)(1.0 / sqrt(var + (double)eps));
        shift = -(float)mean * rstd;
      }
      s_mr[2 * c] = rstd;
      s_mr[2 * c + 1] = shift;
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  constexpr int NH = NT / 2;             // columns owned by this thread
  const int row = (warp & 3) * 32 + lane;  // tile row (TMEM lane) of this thread
  const int chalf = warp >> 2;             // which half of the NT columns
  float acc[NH];
#pragma unroll
  for (int n = 0; n < NH; ++n) acc[n] = 0.f;

  const float* inb = in + (long long)b * Cin * N;
  const uint32_t idesc_base = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(M >> 4) << 24);
  const uint32_t idesc_n = idesc_base | ((uint32_t)(NT >> 3) << 17), idesc_2n = idesc_base | ((uint32_t)(2 * NT >> 3) << 17);
  const int items = 3 * 2 * SEG;  // float4 items per stage
  const int s_base = q0 - (Wp + 1);  // padded index of staged position 0

  for (int stage = 0; stage < nstage; ++stage) {
    const uint32_t ph = (uint32_t)(stage & 1);
    // the previous stage's MMAs are complete (waited for below), so both operand buffers are free
    if (tid == 0) {
      mbar_expect_tx(smem_u32(bars), B_BYTES);
      bulk_g2s(smem_u32(sB), wprep + ((long long)ntile * nstage + stage) * (B_BYTES / 4), B_BYTES, smem_u32(bars));
    }
    // ---- stage the activations: global NCDHW -> (normalise) -> hi/lo -> position-major float4
    const int ci0 = stage * KC;
    // SU items per thread in flight (template: 4 or 6, whichever covers the 6 * SEG items of a stage in one round)
    for (int i0 = tid; i0 < items; i0 += THREADS * SU) {
      float v[SU][4];
      bool ok[SU];
      int cbs[SU];
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int i = i0 + u * THREADS;
        const int s = i % SEG;
        const int t = i / SEG;
        const int cb = t & 1, kd = t >> 1;
        cbs[u] = cb;
        const int q = s_base + s;
        int hp = (int)(((float)q + 0.5f) * inv_wp);
        if (hp * Wp > q) --hp;
        else if ((hp + 1) * Wp <= q) ++hp;
        const int wp = q - hp * Wp;
        const int dd = d + kd - 1, h = hp - 1, w = wp - 1;
        ok[u] = i < items && dd >= 0 && dd < D && h >= 0 && h < H && w >= 0 && w < W;
        const float* p = inb + (long long)(ci0 + cb * 4) * N + (long long)dd * HW + h * W + w;
#pragma unroll
        for (int j = 0; j < 4; ++j) v[u][j] = (ok[u] && ci0 + cb * 4 + j < Cin) ? __ldg(p + (long long)j * N) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < SU; ++u) {
        const int i = i0 + u * THREADS;
        if (i < items) {
          float hi[4], lo[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float x = v[u][j];
            if (NORM) {
              const int c = ci0 + cbs[u] * 4 + j;
              x = fmaf(x, s_mr[2 * c], s_mr[2 * c + 1]);
              x = fmaxf(x, 0.1f * x);
              if (!ok[u]) x = 0.f;  // the padding is zero in the activation domain
            }
            hi[j] = tf32_hi(x);
            lo[j] = x - hi[j];
          }
          sA[i] = make_float4(hi[0], hi[1], hi[2], hi[3]);
          sA[a_plane + i] = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();

    // ---- one thread issues the 54 MMAs of the stage (two per tap).  It is chosen with elect.sync -- the compiler then feeds
    // UTCHMMA from uniform registers instead of expanding every tcgen05.mma into a per-lane ELECT / R2UR / branch loop -- and a
    // descriptor is a 32-bit add on its low word (start address; LBO in bits 16-29), the high word (SBO = 128 B, version) is
    // the constant 0x4008 (see conv_march.cu).
    if (warp == 0) {
      uint32_t leader = 0;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      if (leader) {
        mbar_wait(smem_u32(bars), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi16 = (smem_u32(sA) >> 4) | ((uint32_t)SEG << 16);           // LBO = one staged plane of SEG positions
        const uint32_t a_lo16 = a_hi16 + (uint32_t)a_plane;
        const uint32_t b16 = (smem_u32(sB) >> 4) | ((uint32_t)(2 * NT) << 16);         // LBO = 2 * NT rows x 16 B
        auto desc64 = [](uint32_t lo) {
          uint64_t dsc;
          asm("mov.b64 %0, {%1, %2};" : "=l"(dsc) : "r"(lo), "r"(0x4008u));
          return dsc;
        };
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
          const int g = tap < 7 ? 0 : tap < 14 ? 1 : tap < 21 ? 2 : 3;                 // kGroupStart = {0, 7, 14, 21, 27}
          const int gstart = g * 7;
          const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
          const uint32_t a_off = (uint32_t)(kd * 2 * SEG + (Wp + 1) + (kh - 1) * Wp + (kw - 1));
          const uint32_t b_off = (uint32_t)(tap * 2 * 2 * NT);   // [tap][cb][hi/lo][NT][4]: cb stride = 2*NT*16 B
          const uint32_t dcol = tmem + (uint32_t)(g * 2 * NT);
          // A_hi x [B_hi | B_lo]  (N = 2*NT: hi*hi into columns [0,NT), hi*lo into [NT,2NT)), then A_lo x B_hi into [0,NT)
          const uint64_t db = desc64(b16 + b_off);
          if (tap == gstart)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(dcol), "l"(desc64(a_hi16 + a_off)), "l"(db), "r"(idesc_2n) : "memory");
          else
            asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(dcol), "l"(desc64(a_hi16 + a_off)), "l"(db), "r"(idesc_2n) : "memory");
          asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                       ::"r"(dcol), "l"(desc64(a_lo16 + a_off)), "l"(db), "r"(idesc_n) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + 1))
                     : "memory");
      }
      __syncwarp();
    }
    // ---- drain the accumulator groups into registers (round-to-nearest adds)
    mbar_wait(smem_u32(bars + 1), ph);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
    for (int gb = 0; gb < 2 * G; ++gb) {   // per group: the [hi*hi + lo*hi] block, then the [hi*lo] block
      const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(gb * NT + chalf * NH);
      if (NH == 16) {
        uint32_t r[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[j % NH] += __uint_as_float(r[j]);
      } else {
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j % NH] += __uint_as_float(r[j]);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    // the next stage's __syncthreads orders these TMEM reads before the MMAs that overwrite the accumulators
  }

  // ---- epilogue: bias, store (+ optional LeakyReLU), InstanceNorm statistics of the raw output
  __syncthreads();  // every warp is done with TMEM and the operand buffers
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(G * 2 * NT));
  const int q = q0 + row;
  int hp = (int)(((float)q + 0.5f) * inv_wp);
  if (hp * Wp > q) --hp;
  else if ((hp + 1) * Wp <= q) ++hp;
  const int wp = q - hp * Wp;
  const bool valid = q <= q_hi && wp >= 1 && wp <= W;  // rows are in range whenever q is
  const int cbase = co0 + chalf * NH;
  float* ob = out + ((long long)b * Cout + cbase) * N + (long long)d * HW + (hp - 1) * W + (wp - 1);
  float* s_t = reinterpret_cast<float*>(smem);           // [NT][128] transposition buffer (operand buffer reused)
#pragma unroll
  for (int n = 0; n < NH; ++n) {
    float val = 0.f;
    if (valid && cbase + n < Cout) {
      val = acc[n] + __ldg(bias + cbase + n);
      ob[(long long)n * N] = act_out ? lrelu01(val) : val;
    }
    s_t[(chalf * NH + n) * 128 + row] = val;
  }
  if (out_stats != nullptr) {
    __syncthreads();
    constexpr int PARTS = THREADS / NT;   // each (channel, part) thread sums 128 / PARTS rows
    constexpr int RP = 128 / PARTS;
    const int n = tid % NT, part = tid / NT;
    float ps = 0.f, pq = 0.f;
#pragma unroll 8
    for (int m = 0; m < RP; ++m) {
      const float x = s_t[n * 128 + part * RP + ((m + tid) & (RP - 1))];  // rotated start: conflict-free
      ps += x;
      pq = fmaf(x, x, pq);
    }
    s_part[(part * NT + n) * 2] = (double)ps;
    s_part[(part * NT + n) * 2 + 1] = (double)pq;
    __syncthreads();
    if (tid < 2 * NT) {
      const int c = tid >> 1, which = tid & 1;
      if (co0 + c < Cout) {
        double tot = 0.0;
#pragma unroll
        for (int p = 0; p < PARTS; ++p) tot += s_part[(p * NT + c) * 2 + which];
        atomicAdd(out_stats + ((long long)b * Cout + co0 + c) * 2 + which, tot);
      }
    }
  }
}

template <int NT>
int launch_nt(const float* in, const float* weight, const float* wprep_in, const float* bias, float* out,
              const double* in_stats, double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
              cudaStream_t st) {
  const int Wp = W + 2;
  const int SEG = M + 2 * (Wp + 1);
  const int nstage = ceil_div(Cin, KC), ntiles_n = ceil_div(Cout, NT);
  const int tiles_plane = ceil_div((H - 1) * Wp + W, M);
  constexpr int B_BYTES = 2 * 27 * 2 * NT * 16;
  const size_t smem = (size_t)2 * 3 * 2 * SEG * 16 + B_BYTES + 32 + (size_t)THREADS * 2 * 8 +
                      (size_t)nstage * KC * 2 * 4 + 16;
  // weights split and re-arranged: prepared once by the caller (smile_conv3d_tc_prep), or per launch into
  // stream-ordered scratch
  const long long welems = (long long)ntiles_n * nstage * (B_BYTES / 4);
  float* scratch = nullptr;
  const float* wprep = wprep_in;
  if (wprep == nullptr) {
    // The default memory pool hands unused memory back to the driver at every synchronisation (release threshold 0),
    // which turns the next cudaMallocAsync into a multi-millisecond real allocation (measured: 18 ms per layer when
    // the caller alternates streams).  Keep up to 64 MB of scratch cached in the pool instead.
    static std::once_flag once[64];
    int dev = 0;
    cudaGetDevice(&dev);
    std::call_once(once[dev & 63], [dev] {
      cudaMemPool_t pool;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t cur = 0, want = 64ull << 20;
        cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &cur);
        if (cur < want) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &want);
      }
    });
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)welems * 4, st);
    if (e != cudaSuccess) {
      set_error("conv3d(tcgen05): cudaMallocAsync(%lld B) failed: %s", welems * 4, cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
    conv3d_tc_prep_kernel<<<(unsigned)ceil_div_ll(welems, 256), 256, 0, st>>>(weight, scratch, Cout, Cin, NT, nstage, welems);
    wprep = scratch;
  }
  dim3 grid(tiles_plane * D, ntiles_n, B);
  auto run = [&](auto kern) {
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) {
      set_error("conv3d(tcgen05): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e2));
      return SMILE_ERR_CUDA;
    }
    kern<<<grid, THREADS, smem, st>>>(in, wprep, bias, out, in_stats, out_stats, Cin, Cout, D, H, W, tiles_plane, nstage, SEG,
                                  act_out, eps);
    return check_launch("conv3d(tcgen05)");
  };
  const bool su6 = 3 * 2 * SEG > 4 * THREADS;  // more than four items per thread: keep six in flight (measured)
  const int rc = (in_stats != nullptr) ? (su6 ? run(conv3d_tc_kernel<NT, true, 6>) : run(conv3d_tc_kernel<NT, true, 4>))
                                       : (su6 ? run(conv3d_tc_kernel<NT, false, 6>) : run(conv3d_tc_kernel<NT, false, 4>));
  if (scratch != nullptr) cudaFreeAsync(scratch, st);
  return rc;
}

}  // namespace

// Prepared-weight interface: NT is a function of Cout only, so a caller can prepare once per weight tensor.
static int tc_nt(int Cout) { return Cout <= 16 ? 16 : 32; }

long long conv3d_tc_prep_floats(int Cin, int Cout) {
  const int NT = tc_nt(Cout);
  return (long long)ceil_div(Cout, NT) * ceil_div(Cin, KC) * (2 * 27 * 2 * NT * 4);
}

int launch_conv3d_tc_prep(const float* weight, float* wprep, int Cin, int Cout, cudaStream_t st) {
  const int NT = tc_nt(Cout);
  const long long welems = conv3d_tc_prep_floats(Cin, Cout);
  conv3d_tc_prep_kernel<<<(unsigned)ceil_div_ll(welems, 256), 256, 0, st>>>(weight, wprep, Cout, Cin, NT, ceil_div(Cin, KC),
                                                                            welems);
  return check_launch("conv3d_tc_prep");
}

// Tensor-core path for layers with enough output channels to fill an MMA (N >= 16 after padding).  *handled = false
// leaves the layer to the SIMT kernels.  wprep: weights prepared by launch_conv3d_tc_prep, or NULL.
int launch_conv3d_tc(const float* in, const float* weight, const float* wprep, const float* bias, float* out,
                     const double* in_stats, double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out,
                     float eps, cudaStream_t st, bool* handled) {
  *handled = false;
  static const int mode = [] { const char* e = getenv("SMILE_CONV_TC"); return e ? atoi(e) : 1; }();
  if (mode == 0 || Cout < 12 || H < 2 || W < 2) return SMILE_OK;
  // measured on the B200 (this version stages every key plane three times and is bound by staging on wide volumes):
  // ahead of the SIMT kernels from 16 input channels up on the coarse levels, behind them on the 80-wide level
  if (mode == 1 && (Cin < 16 || W > 48)) return SMILE_OK;
  const int Wp = W + 2;
  const int SEG = M + 2 * (Wp + 1);
  const int NT = tc_nt(Cout);
  const size_t smem = (size_t)2 * 3 * 2 * SEG * 16 + 2 * 27 * 2 * NT * 16 + 4096;
  if (smem > 200 * 1024 || (long long)(H + 2) * Wp >= (1 << 22)) return SMILE_OK;  // very wide rows: SIMT path
  *handled = true;
  if (NT == 16)
    return launch_nt<16>(in, weight, wprep, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
  return launch_nt<32>(in, weight, wprep, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
}

}  // namespace smile
