// Conv3d 3x3x3 / pad 1 with bf16 operands on the 5th-generation tensor cores (tcgen05 kind::f16, fp32 accumulation in
// TMEM) -- the reduced-precision encoder / CWM convolutions of BASELINE.json configs[2..3] ("bf16 ... Conv3d encoder on
// tensor cores"; reference layers ModeT/models.py:119-151, 186-228, 250-254).  Same contract as the fp32 kernels: NCDHW
// fp32 in / out, optional producer InstanceNorm + LeakyReLU applied on load, fp64 sum / sum-of-squares of the raw output
// for the next layer; only the MMA operands are rounded to bf16 (activations after the normalise-on-load, weights once
// per launch).  Expected error: ~2^-9 relative per product, i.e. ~1e-3 relative on a layer output (tests state it).
//
// Same implicit GEMM as conv_tc.cu (an output tile is M = 128 consecutive positions of the zero-padded plane, a tap is a
// shifted start address of a K-major no-swizzle A descriptor), with what bf16 makes possible:
//   * 8 bf16 channels = 16 bytes per staged position, K = 16 per MMA = two channel blocks; ONE MMA per tap (no hi/lo
//     split), 27 per 16-channel stage instead of 108;
//   * a single TMEM accumulator that lives across all taps and all stages and is drained once (fp32 accumulation of
//     bf16 products is exact per product; no grouped drains);
//   * the operand buffers are double buffered: the activations of stage s+1 are staged while the MMAs of stage s run;
//   * N = 16 / 32 / 64 output channels per CTA (64 from 64 output channels up: each CTA re-stages the activations, so
//     fewer, wider N tiles mean fewer redundant loads).
#include <cuda_bf16.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

constexpr int M = 128;        // positions per tile (UMMA M)
constexpr int KC = 16;        // input channels per stage (UMMA K for kind::f16)
constexpr int THREADS = 256;  // 8 warps: warps w and w + 4 share TMEM lane quadrant w and split the NT columns

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
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// weight [Cout][Cin][27] fp32 -> wprep bf16 [ntile][stage][tap][cb(2)][NT][8]; channel ci = stage*16 + cb*8 + j,
// output channel co = ntile*NT + n; zero outside the tensor.
__global__ void conv3d_bf16_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wprep, int Cout, int Cin,
                                        int NT, int nstage, long long total) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  long long t = e;
  const int j = (int)(t % 8); t /= 8;
  const int n = (int)(t % NT); t /= NT;
  const int cb = (int)(t % 2); t /= 2;
  const int tap = (int)(t % 27); t /= 27;
  const int stage = (int)(t % nstage); t /= nstage;
  const int ntile = (int)t;
  const int ci = stage * KC + cb * 8 + j, co = ntile * NT + n;
  float v = 0.f;
  if (ci < Cin && co < Cout) v = w[((long long)co * Cin + ci) * 27 + tap];
  wprep[e] = __float2bfloat16_rn(v);
}

template <int NT, bool NORM>
__global__ void __launch_bounds__(THREADS)
conv3d_bf16_kernel(const float* __restrict__ in, const __nv_bfloat16* __restrict__ wprep, const float* __restrict__ bias,
                   float* __restrict__ out, const double* __restrict__ in_stats, double* __restrict__ out_stats, int Cin,
                   int Cout, int D, int H, int W, int tiles_plane, int nstage, int SEG, int act_out, float eps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  // A: [2 buffers][3 planes][2 channel blocks][SEG positions] x 16 B (8 bf16 channels)
  const int a_items = 3 * 2 * SEG;
  uint4* sA = reinterpret_cast<uint4*>(smem);
  constexpr int B_BYTES = 27 * 2 * NT * 16;                 // [27][2][NT][8 bf16]
  uint8_t* sB = smem + (size_t)2 * a_items * 16;            // [2 buffers][B_BYTES]
  uint8_t* tail = sB + 2 * B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(tail);       // [0..1] weights landed, [2..3] MMAs of the buffer done
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tail + 32);
  double* s_part = reinterpret_cast<double*>(tail + 48);    // [THREADS / NT parts][NT][2]
  float* s_mr = reinterpret_cast<float*>(tail + 48 + THREADS * 2 * 8);  // [nstage * 16][2] rstd, -mean*rstd

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Wp = W + 2, HW = H * W;
  const long long N = (long long)D * HW;
  const int d = blockIdx.x / tiles_plane, tile = blockIdx.x - d * tiles_plane;
  const int ntile = blockIdx.y, co0 = ntile * NT;
  const int b = blockIdx.z;
  const int q_lo = Wp + 1, q_hi = H * Wp + W;             // padded index of voxel (0,0) and (H-1,W-1)
  const int q0 = q_lo + tile * M;
  const float inv_wp = 1.0f / (float)Wp;
  constexpr int TCOLS = NT < 32 ? 32 : NT;                // TMEM allocations are powers of two >= 32 columns

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(smem_u32(bars + i), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS));
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

  const float* inb = in + (long long)b * Cin * N;
  // instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), K-major both, N >> 3, M >> 4
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(NT >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  const int s_base = q0 - (Wp + 1);  // padded index of staged position 0

  for (int stage = 0; stage < nstage; ++stage) {
    const int buf = stage & 1;
    const uint32_t use = (uint32_t)(stage >> 1);          // this is the use-th time the buffer pair is filled
    uint4* sAb = sA + (size_t)buf * a_items;
    // the MMAs that read this buffer two stages ago are complete?
    if (stage >= 2) mbar_wait(smem_u32(bars + 2 + buf), (use - 1) & 1u);
    if (tid == 0) {
      mbar_expect_tx(smem_u32(bars + buf), B_BYTES);
      bulk_g2s(smem_u32(sB + buf * B_BYTES), reinterpret_cast<const uint8_t*>(wprep) + ((size_t)ntile * nstage + stage) * B_BYTES,
               B_BYTES, smem_u32(bars + buf));
    }
    // ---- stage the activations: global NCDHW fp32 -> (normalise) -> bf16 -> position-major, 8 channels per 16 bytes
    const int ci0 = stage * KC;
    for (int i = tid; i < a_items; i += THREADS) {
      const int s = i % SEG;
      const int t = i / SEG;
      const int cb = t & 1, kd = t >> 1;
      const int c0 = ci0 + cb * 8;
      uint4 packed = make_uint4(0u, 0u, 0u, 0u);
      if (c0 < Cin) {
        const int q = s_base + s;
        int hp = (int)(((float)q + 0.5f) * inv_wp);
        if (hp * Wp > q) --hp;
        else if ((hp + 1) * Wp <= q) ++hp;
        const int wp = q - hp * Wp;
        const int dd = d + kd - 1, h = hp - 1, w = wp - 1;
        if (dd >= 0 && dd < D && h >= 0 && h < H && w >= 0 && w < W) {
          const float* p = inb + (long long)c0 * N + (long long)dd * HW + h * W + w;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] = (c0 + j < Cin) ? __ldg(p + (long long)j * N) : 0.f;
          if (NORM) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (c0 + j < Cin) {
                const float x = fmaf(v[j], s_mr[2 * (c0 + j)], s_mr[2 * (c0 + j) + 1]);
                v[j] = fmaxf(x, 0.1f * x);
              }
            }
          }
          packed = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
        }
      }
      sAb[i] = packed;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();

    // ---- one thread issues the 27 MMAs of the stage; they run while the other buffer is being staged.  The thread is
    // chosen with elect.sync and a descriptor is a 32-bit add on its low word (see conv_march.cu: with `tid == 0` ptxas
    // expanded every tcgen05.mma into a per-lane ELECT / R2UR / branch loop).
    if (warp == 0) {
      uint32_t leader = 0;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
      if (leader) {
        mbar_wait(smem_u32(bars + buf), use & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a16 = (smem_u32(sAb) >> 4) | ((uint32_t)SEG << 16);                   // LBO = one plane of SEG positions
        const uint32_t b16 = (smem_u32(sB + buf * B_BYTES) >> 4) | ((uint32_t)NT << 16);     // LBO = NT rows x 16 B
        auto desc64 = [](uint32_t lo) {
          uint64_t dsc;
          asm("mov.b64 %0, {%1, %2};" : "=l"(dsc) : "r"(lo), "r"(0x4008u));
          return dsc;
        };
#pragma unroll
        for (int tap = 0; tap < 27; ++tap) {
          const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
          const uint32_t a_off = (uint32_t)(kd * 2 * SEG + (Wp + 1) + (kh - 1) * Wp + (kw - 1));
          const uint32_t b_off = (uint32_t)(tap * 2 * NT);
          const uint32_t accum = (stage > 0 || tap > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(desc64(a16 + a_off)), "l"(desc64(b16 + b_off)), "r"(idesc), "r"(accum)
              : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + 2 + buf))
                     : "memory");
      }
      __syncwarp();
    }
  }
  // ---- all MMAs done (the last commit tracks every MMA issued before it)
  {
    const int last = nstage - 1;
    mbar_wait(smem_u32(bars + 2 + (last & 1)), (uint32_t)(last >> 1) & 1u);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  constexpr int NH = NT / 2;               // columns owned by this thread
  const int row = (warp & 3) * 32 + lane;  // tile row (TMEM lane) of this thread
  const int chalf = warp >> 2;             // which half of the NT columns
  float acc[NH];
  {
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(chalf * NH);
#pragma unroll
    for (int c8 = 0; c8 < NH; c8 += 8) {
      uint32_t r[8];
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                   : "r"(taddr + (uint32_t)c8));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[c8 + j] = __uint_as_float(r[j]);
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");

  // ---- epilogue: bias, store (+ optional LeakyReLU), InstanceNorm statistics of the raw output
  __syncthreads();  // every warp is done with TMEM and the operand buffers
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS));
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

size_t smem_bytes(int NT, int SEG, int nstage) {
  return (size_t)2 * 3 * 2 * SEG * 16 + (size_t)2 * 27 * 2 * NT * 16 + 48 + (size_t)THREADS * 2 * 8 +
         (size_t)nstage * KC * 2 * 4 + 16;
}

template <int NT>
int launch_nt(const float* in, const float* weight, const float* bias, float* out, const double* in_stats, double* out_stats,
              int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps, cudaStream_t st) {
  const int Wp = W + 2;
  const int SEG = M + 2 * (Wp + 1);
  const int nstage = ceil_div(Cin, KC), ntiles_n = ceil_div(Cout, NT);
  const int tiles_plane = ceil_div((H - 1) * Wp + W, M);
  const size_t smem = smem_bytes(NT, SEG, nstage);
  // weights rounded to bf16 and re-arranged per launch into stream-ordered scratch (cached in the pool, see conv_tc.cu)
  const long long welems = (long long)ntiles_n * nstage * 27 * 2 * NT * 8;
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
  __nv_bfloat16* wprep = nullptr;
  cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&wprep), (size_t)welems * 2, st);
  if (e != cudaSuccess) {
    set_error("conv3d(bf16): cudaMallocAsync(%lld B) failed: %s", welems * 2, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  conv3d_bf16_prep_kernel<<<(unsigned)ceil_div_ll(welems, 256), 256, 0, st>>>(weight, wprep, Cout, Cin, NT, nstage, welems);
  dim3 grid(tiles_plane * D, ntiles_n, B);
  auto run = [&](auto kern) {
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e2 != cudaSuccess) {
      set_error("conv3d(bf16): cannot reserve %zu B of shared memory: %s", smem, cudaGetErrorString(e2));
      return SMILE_ERR_CUDA;
    }
    kern<<<grid, THREADS, smem, st>>>(in, wprep, bias, out, in_stats, out_stats, Cin, Cout, D, H, W, tiles_plane, nstage, SEG,
                                      act_out, eps);
    return check_launch("conv3d(bf16)");
  };
  const int rc = (in_stats != nullptr) ? run(conv3d_bf16_kernel<NT, true>) : run(conv3d_bf16_kernel<NT, false>);
  cudaFreeAsync(wprep, st);
  return rc;
}

}  // namespace

// bf16 tensor-core path.  *handled = false: the shape is left to the fp32 kernels (first layer with one input channel,
// rows too wide for the staged halo to fit in shared memory).
int launch_conv3d_bf16(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                       double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                       cudaStream_t st, bool* handled) {
  *handled = false;
  if (Cin < 4 || H < 2 || W < 2) return SMILE_OK;
  // Measured on the B200 (tools/conv_compare.py, profiles/r03e_conv_compare.txt): this kernel stages three padded planes
  // per 128 outputs, which only pays while rows are short -- 1.2-2.0x ahead of the fp32 kernels up to 48-wide volumes and
  // on 16+ input channels at 80-96 wide, 0.1-0.8x elsewhere.  SMILE_CONV_BF16=2 forces it wherever it is legal.
  const char* env = getenv("SMILE_CONV_BF16");   // read per call (tests flip it)
  const int mode = env ? atoi(env) : 1;
  if (mode != 2 && !(W <= 48 || (W <= 96 && Cin >= 16))) return SMILE_OK;
  const int Wp = W + 2;
  const int SEG = M + 2 * (Wp + 1);
  const int NT = Cout <= 16 ? 16 : (Cout < 64 ? 32 : 64);
  if (smem_bytes(NT, SEG, ceil_div(Cin, KC)) > 220 * 1024 || (long long)(H + 2) * Wp >= (1 << 22)) return SMILE_OK;
  *handled = true;
  if (NT == 16) return launch_nt<16>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
  if (NT == 32) return launch_nt<32>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
  return launch_nt<64>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st);
}

}  // namespace smile
