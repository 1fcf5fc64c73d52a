// Conv3d 3x3x3 / pad 1 for the WIDE, few-channel levels (Cin <= 16, Cout <= 16: the 160- and 80-wide encoder / CWM layers,
// 70 % of the convolution time) on tcgen05 -- the depth-marching counterpart of conv_bf16.cu / conv_tc.cu.  Two precisions
// (template parameter SP, struct Lay): bf16 operands for the bf16 mode (BASELINE.json configs[2..3]) and fp16-split operands
// (x = hi + 2^-11 lo, three products) with fp32-class accuracy for the default fp32 path.
//
// conv_bf16.cu stages three padded planes per 128 outputs; on a 160-wide volume that is 10 staged positions per output
// and the kernel is 5-10x slower than the fp32 SIMT path (profiles/r03e_conv_compare.txt).  Here a CTA owns a 16-row x
// 30-column tile of the (H, W) plane and marches it along D:
//   * every input plane of the tile (18 x 32 positions with halo, position-major, 8 channels = 16 bytes per position and
//     channel block) is staged ONCE into a ring of four planes -- 1.2 staged positions per output;
//   * a tap (kd, kh, kw) is ring slot kd and the shifted start address (kh * 32 + kw) * 16 B of a K-major no-swizzle A
//     descriptor, as in conv_tc.cu / conv_bf16.cu; four M tiles (4 rows x 32 padded columns each) per plane.  With more than 8
//     input channels the K = 16 MMA is one tap x 16 channels (27 MMAs per M tile), with 5..8 it is TWO taps x 8 channels (the
//     descriptor's leading-dimension offset is the distance between the taps: 15 MMAs), with <= 4 a staged position also
//     carries its right-hand neighbour's channels and one MMA covers a kernel row (9 MMAs);
//   * two TMEM accumulator buffers and a dedicated MMA-issuing warp (one elect.sync thread, 32-bit descriptor arithmetic on
//     the uniform datapath): the MMAs of plane d run while the eight worker warps stage plane d+2 and drain, bias-add and
//     store plane d-1; the hand-overs are mbarriers (plane staged -> issuer, MMAs done -> workers, buffer drained -> issuer),
//     there is no CTA-wide barrier in the loop;
//   * InstanceNorm statistics of the raw output: fp32 partials flushed every four planes into per-warp fp64 sums.
// Contract identical to smile_conv3d_fwd / smile_conv3d_bf16_fwd (NCDHW fp32 in / out, normalise-on-load, fp64 statistics).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

constexpr int P = 32;                 // padded pitch of a staged row (30 output columns + 2)
constexpr int TC = P - 2;             // output columns per tile
constexpr int MT = 4;                 // M tiles (128 positions = 4 rows) per plane
constexpr int TR = 4 * MT;            // output rows per tile
constexpr int SROWS = TR + 2;         // staged rows
constexpr int PLANE_POS = 592;        // staged positions per plane (18 * 32 = 576, + the overhang of the last taps)
constexpr int PLANE_BYTES = 2 * PLANE_POS * 16;   // two channel blocks of 8 bf16
constexpr int RING = 4;
constexpr int NT = 16;                // output channels per launch at most
constexpr int WORKERS = 256;          // warps 0-7 stage the planes and drain the accumulators
constexpr int THREADS = WORKERS + 32; // warp 8 only issues the MMAs (one lane): 108 per plane would otherwise delay warp 0
constexpr int OFF_B = RING * PLANE_BYTES;
// Precision / accumulator layout of a launch (template parameter SP):
//   0  bf16 operands, one accumulator of 16 columns per M tile (configs[2..3] of BASELINE.json)
//   1  fp16 split (fp32-class), Cout <= 8: MMA N = 16, two chains for the leading term, 32 columns per M tile
//   2  fp16 split (fp32-class), Cout <= 16: MMA N = 32, one chain, 32 columns per M tile [ hh(16) | corr(16) ]
template <int SP>
struct Lay {
  static constexpr int MMA_N = SP == 2 ? 32 : 16;
  static constexpr int BLK = 2 * MMA_N * 16;                      // bytes of one B block (two K chunks of MMA_N rows)
  static constexpr int NBLK = SP == 0 ? 27 : SP == 1 ? 32 : 30;   // SP 1: 2 x 15 + a zero block of N = 32 (two blocks)
  static constexpr int B_BYTES = NBLK * BLK;
  static constexpr int MCOLS = SP == 0 ? 16 : 32;                 // TMEM columns per M tile
  static constexpr int TCOLS = 2 * MT * MCOLS;                    // two accumulator buffers
  // [0] weights landed, [1..2] MMAs of TMEM buffer 0 / 1 done, [3..4] plane staged (by step parity), [5..6] buffer drained
  static constexpr int OFF_BAR = OFF_B + B_BYTES;
  static constexpr int OFF_TMEM = OFF_BAR + 64;
  static constexpr int OFF_MR = OFF_TMEM + 16;              // [16][2] rstd, -mean * rstd
  static constexpr int OFF_RED = OFF_MR + 128 + 64;         // (+ 16 bias values); [8 warps][16][2] doubles
  static constexpr int SMEM = OFF_RED + 8 * 16 * 2 * 8 + 16;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
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

// weight [Cout][Cin][27] fp32 -> bf16 B operands.
//   Cin > 8 : [tap 27][cb 2][NT][8], input channel ci = cb*8 + j (one K = 16 MMA per tap).
//   Cin <= 8: [mma 15][cb 2][NT][8]: two TAPS share a K = 16 MMA -- channel block 0 holds tap t, block 1 tap t+1 of the same
//             plane (A's second K chunk is the same staged plane shifted by the tap distance, see the issuer), the ninth
//             tap of a plane goes alone with a zero second block: 15 MMAs per M tile instead of 27.
__device__ __forceinline__ int pair_first_tap(int i) { return (i / 5) * 9 + (i % 5) * 2; }   // mma i of 15 -> its first tap
//   Cin <= 4: [mma 9][cb 2][NT][8]: MMA i = kernel row (kd, kh); block 0 holds taps kw = 0 (j < 4) and kw = 1 (j >= 4), block 1
//             tap kw = 2 (j < 4) and zeros.
__device__ __forceinline__ int row_tap(int i, int cb, int j) { return (j >= 4 && cb == 1) ? -1 : i * 3 + 2 * cb + (j >> 2); }
__global__ void conv_march_prep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wprep, int Cout, int Cin) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 27 * 2 * NT * 8) return;
  int t = e;
  const int j = t % 8; t /= 8;
  const int n = t % NT; t /= NT;
  const int cb = t % 2; t /= 2;
  float v = 0.f;
  if (Cin <= 4) {
    const int tap = t < 9 ? row_tap(t, cb, j) : -1;
    if (tap >= 0 && (j & 3) < Cin && n < Cout) v = w[((long long)n * Cin + (j & 3)) * 27 + tap];
  } else if (Cin > 8) {
    const int tap = t, ci = cb * 8 + j;
    if (ci < Cin && n < Cout) v = w[((long long)n * Cin + ci) * 27 + tap];
  } else if (t < 15) {
    const int t0 = pair_first_tap(t);
    const bool single = (t % 5) == 4;
    const int tap = t0 + cb;
    if (!(single && cb == 1) && j < Cin && n < Cout) v = w[((long long)n * Cin + j) * 27 + tap];
  }
  wprep[e] = __float2bfloat16_rn(v);
}

// ---------------------------------------------------------------------------------------------------------------------
// SPLIT mode: fp32-class accuracy on the tensor cores for the layers with at most 8 input and 8 output channels.
// Every operand is written x = hi + 2^-11 * lo with hi = fp16(x), lo = fp16((x - hi) * 2^11): 22 significant bits, the
// same class as the three-product TF32 split of conv_tc.cu, but with K = 16 per MMA and two taps per MMA.  Per tap pair
// two MMAs into a 24-column accumulator [ hh_a | corr | hh_b ]:
//     even pairs   A_hi x [ W_hi | W_lo ] at column 0      hh_a += hi*hi      corr += hi*lo
//     odd pairs    A_hi x [ W_lo | W_hi ] at column 8      corr += hi*lo      hh_b += hi*hi
//     every pair   A_lo x [  0   | W_hi ] at column 0                         corr += lo*hi
// and the epilogue returns (hh_a[n] + hh_b[n]) + 2^-11 * corr[n].  Products of two fp16 values are exact in fp32; what
// costs accuracy is the accumulation in TMEM (truncating adds, error ~ linear in the number of accumulated MMAs,
// tools/probe/tf32x3_probe.cu), so the leading term is spread over two chains of 8 and 7 MMAs that are added with a
// round-to-nearest fp32 add; the corrections are 2^-11 smaller and can share one column group.  One MMA against a zero
// B block (N = 32, accumulate off) clears the 32 columns of an M tile first: the chains overlap in `corr`, so no single
// MMA could be the "first" for all of its columns.
// |x| must stay below the fp16 range (65504) after the normalise-on-load: operands are clamped there.
// ---------------------------------------------------------------------------------------------------------------------
// CLAMP: off for values that went through InstanceNorm (|z| <= sqrt(voxels) < 2^15 for any volume the kernel takes)
template <bool CLAMP>
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  if (CLAMP) {
    a = fminf(fmaxf(a, -65504.f), 65504.f);
    b = fminf(fmaxf(b, -65504.f), 65504.f);
  }
  const __half2 h = __floats2half2_rn(a, b);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((a - hf.x) * 2048.f, (b - hf.y) * 2048.f);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// weight [Cout][CinT][27] fp32, input channels ci0 .. ci0 + Cin - 1 (Cin <= 8) -> fp16 B operands.
// SP 1 (Cout <= 8):  [set 2][mma 15][cb 2][16 rows][8] + two all-zero blocks
//   set 0, even pair: rows n < 8 = W_hi[n], rows n >= 8 = W_lo[n - 8];  odd pair: the two halves swapped
//   set 1: rows n < 8 = 0, rows n >= 8 = W_hi[n - 8]
// SP 2 (Cout <= 16): [set 2][mma 15][cb 2][32 rows][8]
//   set 0: rows n < 16 = W_hi[n], rows n >= 16 = W_lo[n - 16];   set 1: rows n < 16 = 0, rows n >= 16 = W_hi[n - 16]
template <int SP>
__global__ void conv_march_prep_split_kernel(const float* __restrict__ w, __half* __restrict__ wprep, int Cout, int Cin,
                                             int CinT, int ci0) {
  constexpr int NR = Lay<SP>::MMA_N, NG = NR / 2;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= Lay<SP>::B_BYTES / 2) return;
  if (e >= 30 * 2 * NR * 8) {
    wprep[e] = __float2half_rn(0.f);
    return;
  }
  int t = e;
  const int j = t % 8; t /= 8;
  const int n = t % NR; t /= NR;
  const int cb = t % 2; t /= 2;
  const int set = t / 15, i = t % 15;
  const int co = n % NG;
  float v = 0.f;
  if (Cin <= 4) {      // kernel rows: see conv_march_prep_kernel
    const int tap = i < 9 ? row_tap(i, cb, j) : -1;
    if (tap >= 0 && (j & 3) < Cin && co < Cout) v = w[((long long)co * CinT + ci0 + (j & 3)) * 27 + tap];
  } else {
    const int t0 = pair_first_tap(i);
    const bool single = (i % 5) == 4;
    const int tap = t0 + cb;
    if (!(single && cb == 1) && j < Cin && co < Cout) v = w[((long long)co * CinT + ci0 + j) * 27 + tap];
  }
  v = fminf(fmaxf(v, -65504.f), 65504.f);
  const __half h = __float2half_rn(v);
  const __half l = __float2half_rn((v - __half2float(h)) * 2048.f);
  const bool swap = SP == 1 && (i & 1) != 0;
  __half r = __float2half_rn(0.f);
  if (set == 0) r = ((n < NG) != swap) ? h : l;
  else if (n >= NG) r = h;
  wprep[e] = r;
}

// CIN8: at most 8 input channels (the second channel block stays zero); NORM: InstanceNorm + LeakyReLU on load
// SP > 0 (with CIN8): fp16 hi / lo operands, see above; the second channel-block plane of a ring slot holds A_lo.
// The input channels of this launch are ci0 .. ci0 + Cin - 1 of a tensor with CinT channels.  pass: 0 = the whole
// reduction; 1 = first of two launches over the input channels (bias + partial sums stored raw, no statistics);
// 2 = second (adds what pass 1 stored, then activation / statistics).
// CIN4 (with CIN8): at most 4 input channels.  A staged position holds its own 4 channels and those of its right-hand
// neighbour, so a 16-byte K chunk is TWO taps (kw, kw+1) and a K = 16 MMA covers a whole kernel row (kw = 0, 1 | 2, -):
// 9 MMAs per M tile and operand set instead of 15.
template <bool CIN8, bool NORM, int SP, bool CIN4>
__global__ void __launch_bounds__(THREADS, 2)   // 18 warps per SM = 5 on one scheduler: 16384 / (5 x 32) -> 96 registers
conv_march_kernel(const float* __restrict__ in, const void* __restrict__ wprep, const float* __restrict__ bias,
                  float* __restrict__ out, const double* __restrict__ in_stats, double* __restrict__ out_stats, int Cin,
                  int Cout, int D, int H, int W, int ntr, int ntc, int DS, int act_out, float eps, int CinT, int ci0, int pass) {
  extern __shared__ __align__(1024) uint8_t smem[];
  using L = Lay<SP>;
  constexpr bool SPLIT = SP != 0;
  constexpr int MCOLS = L::MCOLS, TCOLS = L::TCOLS, MMA_N = L::MMA_N;
  constexpr int OFF_BAR = L::OFF_BAR, OFF_TMEM = L::OFF_TMEM, OFF_MR = L::OFF_MR, OFF_RED = L::OFF_RED, B_BYTES = L::B_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_TMEM);
  float* s_mr = reinterpret_cast<float*>(smem + OFF_MR);
  double* s_red = reinterpret_cast<double*>(smem + OFF_RED);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int HW = H * W;
  const long long N = (long long)D * HW;
  int t = blockIdx.x;
  const int ds = t % DS; t /= DS;
  const int tc = t % ntc; t /= ntc;
  const int tr = t % ntr; t /= ntr;
  const int b = t;
  const int h0 = tr * TR, w0 = tc * TC;
  const int dlen = ceil_div(D, DS);
  const int d0 = ds * dlen, d1 = min(D, d0 + dlen);
  if (d0 >= d1) return;   // uniform per CTA

  if (tid == 0) {
    for (int i = 0; i < 3; ++i) mbar_init(smem_u32(bars + i), 1);
    for (int i = 3; i < 7; ++i) mbar_init(smem_u32(bars + i), WORKERS);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(smem_u32(bars), B_BYTES);
    bulk_g2s(smem_u32(smem + OFF_B), wprep, B_BYTES, smem_u32(bars));
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(TCOLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid < 16) {
    float rstd = 1.f, shift = 0.f;
    if (NORM && tid < Cin) {
      const double s = in_stats[((long long)b * CinT + ci0 + tid) * 2], ss = in_stats[((long long)b * CinT + ci0 + tid) * 2 + 1];
      const double mean = s / (double)N;
      const double var = fmax(ss / (double)N - mean * mean, 0.0);
      rstd = (float)(1.0 / sqrt(var + (double)eps));
      shift = -(float)mean * rstd;
    }
    s_mr[2 * tid] = rstd;
    s_mr[2 * tid + 1] = shift;
  }
  for (int i = tid; i < 8 * NT * 2; i += THREADS) s_red[i] = 0.0;
  // zero the ring once: the unused channel block, the overhang and the out-of-volume positions stay zero
  for (int i = tid; i < RING * PLANE_BYTES / 16; i += THREADS) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0u, 0u, 0u, 0u);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = *tmem_slot;

  const float* inb = in + ((long long)b * CinT + ci0) * N;
  constexpr int NCH = CIN8 ? 8 : 16;
  constexpr int PB = ((CIN8 && !SPLIT) ? 1 : 2) * PLANE_POS * 16;     // bytes of one staged plane (one or two blocks)

  // ---- stage one input plane (global depth dd) into its ring slot.  A thread stages the same (up to) three tile positions
  // of every plane: their in-plane offsets and border predicates are computed once; element offsets are 32-bit (the
  // launcher checks CinT * D * H * W < 2^31).  With at most 8 input channels the loads of all three positions are issued
  // before the first is converted: one exposed memory latency per plane instead of three (long_scoreboard was the top
  // stall of the workers, profiles/r04b_conv_split_full.txt).
  int s_off[3];
  bool s_ok[3];
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int i = tid + it * WORKERS;
    const int r = i >> 5, c = i & 31;
    const int h = h0 - 1 + r, w = w0 - 1 + c;
    s_ok[it] = i < SROWS * P && h >= 0 && h < H && w >= 0 && w < W;
    s_off[it] = h * W + w;
  }
  const unsigned N32 = (unsigned)N;
  auto convert_store = [&](uint4* slot, int i, float (&v)[NCH], bool ok) {
    if (NORM) {
#pragma unroll
      for (int j = 0; j < NCH; ++j) {
        const float x = fmaf(v[j], s_mr[2 * j], s_mr[2 * j + 1]);   // broadcast shared-memory reads
        v[j] = (ok && j < Cin) ? fmaxf(x, 0.1f * x) : 0.f;
      }
    }
    if (CIN4) {   // channels 4-7 of a position = channels 0-3 of its right-hand neighbour (lanes run along the row)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float nb = __shfl_down_sync(0xffffffffu, v[j], 1);
        v[4 + j] = (lane < 31) ? nb : 0.f;
      }
    }
    if (SPLIT) {
      uint4 hi4, lo4;
      split_pair<!NORM>(v[0], v[1], hi4.x, lo4.x);
      split_pair<!NORM>(v[2], v[3], hi4.y, lo4.y);
      split_pair<!NORM>(v[4], v[5], hi4.z, lo4.z);
      split_pair<!NORM>(v[6], v[7], hi4.w, lo4.w);
      slot[i] = hi4;
      slot[PLANE_POS + i] = lo4;
      return;
    }
    slot[i] = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
    if (!CIN8)
      slot[PLANE_POS + i] =
          make_uint4(pack_bf16(v[8 % NCH], v[9 % NCH]), pack_bf16(v[10 % NCH], v[11 % NCH]), pack_bf16(v[12 % NCH], v[13 % NCH]),
                     pack_bf16(v[14 % NCH], v[15 % NCH]));
  };
  auto stage_plane = [&](int dd) {
    uint4* slot = reinterpret_cast<uint4*>(smem + (size_t)((dd + RING) % RING) * PB);
    const bool plane_ok = dd >= 0 && dd < D;
    const unsigned e0 = plane_ok ? (unsigned)(dd * HW) : 0u;
    if (CIN8) {
      float v[3][NCH];
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const bool ok = plane_ok && s_ok[it];
        const float* p = inb + (e0 + (unsigned)s_off[it]);
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[it][j] = (ok && j < Cin && !(CIN4 && j >= 4)) ? __ldg(p + (unsigned)j * N32) : 0.f;
      }
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int i = tid + it * WORKERS;
        if (i < SROWS * P) convert_store(slot, i, v[it], plane_ok && s_ok[it]);
      }
    } else {
#pragma unroll
      for (int it = 0; it < 3; ++it) {
        const int i = tid + it * WORKERS;
        if (i >= SROWS * P) break;
        const bool ok = plane_ok && s_ok[it];
        float v[NCH];
        const float* p = inb + (e0 + (unsigned)s_off[it]);
#pragma unroll
        for (int j = 0; j < NCH; ++j) v[j] = (ok && j < Cin) ? __ldg(p + (unsigned)j * N32) : 0.f;
        convert_store(slot, i, v, ok);
      }
    }
  };

  // epilogue ownership: TMEM lane quadrant q = warp % 4 (row 4m + q of M tile m), M tiles 2*(warp / 4) and + 1
  const int q = warp & 3, mbase = 2 * (warp >> 2);
  const bool col_ok = lane < TC && (w0 + lane) < W;
  float* s_bias = s_mr + 32;
  if (tid < NT) s_bias[tid] = (tid < Cout) ? __ldg(bias + tid) : 0.f;
  __syncthreads();
  // InstanceNorm statistics of the raw output: fp32 partial sums per thread, flushed every four planes through a warp
  // reduction into per-warp fp64 accumulators in shared memory.  (Keeping the fp32 partials for the whole march -- up to
  // 2 x D/DS values per thread, 2560 per warp -- cost 1e-6 relative on the variance of a whole channel: a coherent error
  // that the 5-level cascade amplifies; with the fp16-split operands it alone moved the full-size flow error from 1.06x
  // to 1.75x the reference's own fp32 error.)
  float st_s[NT], st_q[NT];
#pragma unroll
  for (int n = 0; n < NT; ++n) st_s[n] = st_q[n] = 0.f;
  auto flush_stats = [&]() {
#pragma unroll
    for (int n = 0; n < NT; ++n) {
      if (n >= Cout) break;
      float sv = st_s[n], sq = st_q[n];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      if (lane == 0) {
        s_red[(warp * NT + n) * 2] += (double)sv;
        s_red[(warp * NT + n) * 2 + 1] += (double)sq;
      }
      st_s[n] = st_q[n] = 0.f;
    }
  };

  // one accumulator column group of 8 -> registers
  auto tld8 = [](uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
  };
  // pass 2 adds the partial sums the first launch stored: pull those lines into L1 BEFORE waiting for the MMAs, so that
  // the four dependent load rounds of the drain hit L1 instead of exposing an L2 round trip each
  auto prefetch_partial = [&](int dd) {
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
      const int h = h0 + 4 * (mbase + mm) + q;
      if (col_ok && h < H) {
        const float* ob = out + (long long)b * Cout * N + (long long)dd * HW + h * W + (w0 + lane);
#pragma unroll
        for (int n = 0; n < NT; ++n)
          if (n < Cout) asm volatile("prefetch.global.L1 [%0];" ::"l"(ob + (long long)n * N));
      }
    }
  };
  auto drain_plane = [&](int dd, int buf) {
#pragma unroll
    for (int mm = 0; mm < 2; ++mm) {
      const int m = mbase + mm;
      const int h = h0 + 4 * m + q;
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * (MT * MCOLS) + m * MCOLS);
      const bool live = col_ok && h < H;
      float* ob = out + (long long)b * Cout * N + (long long)dd * HW + h * W + (w0 + lane);
      // channel groups of 8: SP 0 columns [0-7 | 8-15]; SP 1 one group, columns hh_a 0-7, corr 8-15, hh_b 16-23;
      // SP 2 two groups, columns hh 0-15, corr 16-31
#pragma unroll
      for (int g = 0; g < (SP == 1 ? 1 : 2); ++g) {
        if (g * 8 >= Cout) break;
        uint32_t r[8], rc[8], rb[8];
        tld8(taddr + (uint32_t)(8 * g), r);
        if (SP == 1) {
          tld8(taddr + 8u, rc);
          tld8(taddr + 16u, rb);
        } else if (SP == 2) {
          tld8(taddr + (uint32_t)(16 + 8 * g), rc);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (live) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int n = 8 * g + i;
            if (n < Cout) {
              float val = __uint_as_float(r[i]);
              if (SP == 1) val = fmaf(__uint_as_float(rc[i]), 1.f / 2048.f, __fadd_rn(val, __uint_as_float(rb[i])));
              if (SP == 2) val = fmaf(__uint_as_float(rc[i]), 1.f / 2048.f, val);
              if (pass == 2)
                val += ob[(long long)n * N];      // partial sums (+ bias) of the first launch
              else
                val += s_bias[n];
              ob[(long long)n * N] = (act_out && pass != 1) ? lrelu01(val) : val;
              st_s[n] += val;
              st_q[n] = fmaf(val, val, st_q[n]);
            }
          }
        }
      }
    }
  };

  // instruction descriptor: fp32 accumulate; A and B formats bf16 (1) or, in SPLIT mode, fp16 (0); N = MMA_N, M = 128
  const uint32_t idesc = (1u << 4) | (SPLIT ? 0u : ((1u << 7) | (1u << 10))) | ((uint32_t)(MMA_N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t idesc32 = (idesc & ~(0x3Fu << 17)) | ((uint32_t)(32 >> 3) << 17);   // the same with N = 32
  const uint32_t ring_base = smem_u32(smem), b_base = smem_u32(smem + OFF_B);

  auto arrive = [&](int bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bars + bar)) : "memory"); };

  if (warp < 8) {
    // ---------------- workers: stage plane d+1, then drain plane d-1 while the MMAs of plane d run
    stage_plane(d0 - 1);
    stage_plane(d0);
    uint32_t ph[2] = {0u, 0u};
    for (int d = d0; d < d1; ++d) {
      const int par = (d - d0) & 1;
      stage_plane(d + 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
      arrive(3 + par);                                               // planes d-1, d, d+1 are in the ring
      if (d > d0) {
        const int pb = par ^ 1;
        if (SP != 0 && pass == 2) prefetch_partial(d - 1);
        mbar_wait(smem_u32(bars + 1 + pb), ph[pb]);                  // MMAs of plane d-1 complete
        ph[pb] ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        drain_plane(d - 1, pb);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        arrive(5 + pb);                                              // TMEM buffer pb may be overwritten
        if (((d - d0) & 3) == 0) flush_stats();
      }
    }
    {
      const int pb = (d1 - 1 - d0) & 1;
      if (SP != 0 && pass == 2) prefetch_partial(d1 - 1);
      mbar_wait(smem_u32(bars + 1 + pb), ph[pb]);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      drain_plane(d1 - 1, pb);
      flush_stats();
    }
  } else {
    // ---------------- issuer warp: one elected thread issues 4 x 27 (15, 30, 31) MMAs per plane.  The thread is chosen
    // with elect.sync (the compiler then knows the region runs on ONE thread and feeds UTCHMMA from uniform registers
    // without its generic per-lane ELECT / R2UR / branch loop), and a descriptor is a 32-bit add: only the 14-bit start
    // address (and, for paired taps, the leading-dimension offset) in its LOW word changes; the high word (stride-dimension
    // offset 128 B, descriptor version) is the constant 0x4008.
    uint32_t leader = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(leader));
    if (leader) {
      mbar_wait(smem_u32(bars), 0);   // weights
      uint32_t ph_s[2] = {0u, 0u}, ph_d[2] = {0u, 0u};
      auto desc64 = [](uint32_t lo) {
        uint64_t dsc;
        asm("mov.b64 %0, {%1, %2};" : "=l"(dsc) : "r"(lo), "r"(0x4008u));
        return dsc;
      };
      constexpr uint32_t BLK16 = 2 * MMA_N;                                  // one B block in 16-byte units
      const uint32_t db_lo = (b_base >> 4) | ((uint32_t)MMA_N << 16);        // LBO = MMA_N rows x 16 B between the K chunks
      constexpr uint32_t LBO_ROW = 1u << 16, LBO_WRAP = (uint32_t)(P - 2) << 16, LBO_CB = (uint32_t)PLANE_POS << 16;
#define SMILE_MMA(DCOL, DA, DB, IDESC, FIRST)                                                                              \
  if (FIRST)                                                                                                               \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" \
                 ::"r"(DCOL), "l"(desc64(DA)), "l"(desc64(DB)), "r"(IDESC) : "memory");                                    \
  else                                                                                                                     \
    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" \
                 ::"r"(DCOL), "l"(desc64(DA)), "l"(desc64(DB)), "r"(IDESC) : "memory");
      for (int d = d0; d < d1; ++d) {
        const int par = (d - d0) & 1;
        mbar_wait(smem_u32(bars + 3 + par), ph_s[par]);
        ph_s[par] ^= 1u;
        if (d - d0 >= 2) {            // the accumulator buffer was last used by plane d-2: drained?
          mbar_wait(smem_u32(bars + 5 + par), ph_d[par]);
          ph_d[par] ^= 1u;
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t slot16[3];
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) slot16[kd] = (ring_base + (uint32_t)(((d - 1 + kd + RING) % RING) * PB)) >> 4;
#pragma unroll 1
        for (int m = 0; m < MT; ++m) {
          const uint32_t dcol = tmem + (uint32_t)(par * (MT * MCOLS) + m * MCOLS);
          const uint32_t a0 = slot16[0] + (uint32_t)(m * 128), a1 = slot16[1] + (uint32_t)(m * 128),
                         a2 = slot16[2] + (uint32_t)(m * 128);
          if (CIN8) {
#pragma unroll
            for (int i = 0; i < 15; ++i) {
              if (CIN4 && i >= 9) break;
              // CIN4: MMA i = kernel row (kd, kh): K chunk 0 = taps kw 0, 1 at the row's first position, chunk 1 = taps
              // kw 2, - two positions further.  Otherwise MMA i = tap pair:
              const int kd = CIN4 ? i / 3 : i / 5, t0 = CIN4 ? (i % 3) * 3 : (i % 5) * 2;   // first tap inside the plane
              const int kh = t0 / 3, kw = t0 % 3;
              // second tap t0 + 1: next column (distance 1 position) unless t0 ends a row (distance P - 2); the lone ninth
              // tap has zero weights on its second K chunk
              const uint32_t lbo = CIN4 ? (2u << 16) : (kw == 2 && t0 != 8) ? LBO_WRAP : LBO_ROW;
              const uint32_t da = ((kd == 0) ? a0 : (kd == 1) ? a1 : a2) + (uint32_t)(kh * P + kw) + lbo;
              const uint32_t db = db_lo + (uint32_t)i * BLK16;
              if (SP == 1) {
                if (i == 0) {   // clear the 32 columns of this M tile: any A x the zero block (N = 32), accumulate off
                  SMILE_MMA(dcol, da, (b_base >> 4) + 30u * BLK16 + (32u << 16), idesc32, true)
                }
                // even pair: [W_hi | W_lo] -> hh_a, corr;  odd pair: [W_lo | W_hi] at column 8 -> corr, hh_b
                SMILE_MMA(dcol + ((i & 1) ? 8u : 0u), da, db, idesc, false)
                // A_lo (second block of the slot) x [0 | W_hi] (second set of B blocks) -> corr
                SMILE_MMA(dcol, da + (uint32_t)PLANE_POS, db + 15u * BLK16, idesc, false)
                continue;
              }
              if (SP == 2) {   // [W_hi | W_lo] -> hh, corr;  A_lo x [0 | W_hi] -> corr
                SMILE_MMA(dcol, da, db, idesc, i == 0)
                SMILE_MMA(dcol, da + (uint32_t)PLANE_POS, db + 15u * BLK16, idesc, false)
                continue;
              }
              SMILE_MMA(dcol, da, db, idesc, i == 0)
            }
          } else {
#pragma unroll
            for (int tap = 0; tap < 27; ++tap) {
              const int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
              const uint32_t da = ((kd == 0) ? a0 : (kd == 1) ? a1 : a2) + (uint32_t)(kh * P + kw) + LBO_CB;
              const uint32_t db = db_lo + (uint32_t)tap * BLK16;
              SMILE_MMA(dcol, da, db, idesc, tap == 0)
            }
          }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bars + 1 + par))
                     : "memory");
      }
#undef SMILE_MMA
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TCOLS));

  // ---- InstanceNorm statistics: per-warp fp64 sums (flush_stats) -> one fp64 atomic per channel and CTA
  if (out_stats != nullptr && pass != 1) {
    __syncthreads();
    if (tid < 2 * NT) {
      const int n = tid >> 1, which = tid & 1;
      if (n < Cout) {
        double tot = 0.0;
#pragma unroll
        for (int wv = 0; wv < 8; ++wv) tot += s_red[(wv * NT + n) * 2 + which];
        atomicAdd(out_stats + ((long long)b * Cout + n) * 2 + which, tot);
      }
    }
  }
}

}  // namespace

namespace {
// common launch path: SP = 0 bf16, 1 / 2 fp16 split (Lay<SP>); input channels ci0 .. ci0 + Cin - 1 of CinT; pass as in the kernel
template <int SP>
int launch_march(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                 double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps, cudaStream_t st,
                 int CinT, int ci0, int pass) {
  using L = Lay<SP>;
  const char* what = SP ? "conv3d(fp16-split march)" : "conv3d(bf16 march)";
  const int ntr = ceil_div(H, TR), ntc = ceil_div(W, TC);
  const long long tiles = (long long)B * ntr * ntc;
  // Two CTAs per SM (92-107 KB of shared memory, 128 / 256 TMEM columns each): while one stages / drains, the other's MMAs
  // run.  The depth is split so that the grid is just under two full waves of 2 x 148 CTAs (measured: 288-576 CTAs beat
  // 144 and anything that leaves a partial third wave; each split re-stages two halo planes, which is not what bounds it).
  // The fp16-split mode (twice the MMAs per plane, so the halo planes of a split weigh less against its tail balance) is
  // best with the FEWEST depth splits that fill whole waves of 2 x 148 CTAs: 8->8 @160x192x160 (144 tiles) 494 -> 488 us with
  // 2 splits instead of 4, 16->16 @80x96x80 284 -> 272, 8->16 103 -> 92; 160x192x224 (192 tiles) needs 3.
  static const int knob = [] { const char* e = getenv("SMILE_MARCH_CTAS"); return e ? atoi(e) : 0; }();
  int DS;
  if (knob > 0 || SP == 0) {
    DS = (int)((knob > 0 ? knob : 4 * kNumSMs) / tiles);
  } else {
    const long long wave = 2LL * kNumSMs;
    double best = -1.0;
    DS = 1;
    for (int cand = 1; cand <= 8; ++cand) {
      const long long ctas = tiles * cand;
      const double eff = (double)ctas / (double)(ceil_div_ll(ctas, wave) * wave);
      if (eff > best + 0.02) {      // a larger split count has to fill the waves noticeably better
        best = eff;
        DS = cand;
      }
    }
  }
  if (DS > D / 4) DS = D / 4;
  if (DS < 1) DS = 1;
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
  void* wprep = nullptr;
  cudaError_t e = cudaMallocAsync(&wprep, L::B_BYTES, st);
  if (e != cudaSuccess) {
    set_error("%s: cudaMallocAsync failed: %s", what, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  if (SP)
    conv_march_prep_split_kernel<SP><<<ceil_div(L::B_BYTES / 2, 256), 256, 0, st>>>(weight, reinterpret_cast<__half*>(wprep), Cout,
                                                                                  Cin, CinT, ci0);
  else
    conv_march_prep_kernel<<<ceil_div(27 * 2 * NT * 8, 256), 256, 0, st>>>(weight, reinterpret_cast<__nv_bfloat16*>(wprep), Cout,
                                                                         Cin);
  const unsigned grid = (unsigned)(tiles * DS);
  auto run = [&](auto kern) {
    cudaError_t e2 = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::SMEM);
    if (e2 != cudaSuccess) {
      set_error("%s: cannot reserve %d B of shared memory: %s", what, L::SMEM, cudaGetErrorString(e2));
      return SMILE_ERR_CUDA;
    }
    kern<<<grid, THREADS, L::SMEM, st>>>(in, wprep, bias, out, in_stats, out_stats, Cin, Cout, D, H, W, ntr, ntc, DS, act_out, eps,
                                         CinT, ci0, pass);
    return check_launch(what);
  };
  int rc;
  if (Cin <= 4)
    rc = in_stats ? run(conv_march_kernel<true, true, SP, true>) : run(conv_march_kernel<true, false, SP, true>);
  else if (SP != 0 || Cin <= 8)
    rc = in_stats ? run(conv_march_kernel<true, true, SP, false>) : run(conv_march_kernel<true, false, SP, false>);
  else
    rc = in_stats ? run(conv_march_kernel<false, true, 0, false>) : run(conv_march_kernel<false, false, 0, false>);
  cudaFreeAsync(wprep, st);
  return rc;
}
}  // namespace

// Depth-marching bf16 tensor-core conv for Cin <= 16 and Cout <= 16.  *handled = false otherwise.
int launch_conv3d_march_bf16(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                             double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                             cudaStream_t st, bool* handled) {
  *handled = false;
  if (Cin < 2 || Cin > 16 || Cout > 16 || D < 1 || H < 2 || W < 2) return SMILE_OK;
  if ((long long)Cin * D * H * W >= (1LL << 31)) return SMILE_OK;   // 32-bit element offsets inside the kernel
  *handled = true;
  return launch_march<0>(in, weight, bias, out, in_stats, out_stats, B, Cin, Cout, D, H, W, act_out, eps, st, Cin, 0, 0);
}

// Depth-marching fp16-split tensor-core conv (fp32-class accuracy) for the wide levels: at most 16 output channels and
// 2..16 input channels (more than 8 input channels = two launches over 8 + the rest, the second adds the first's partial
// sums).  By default only where it beats the SIMT kernels (measured, tools/conv_compare.py): volumes that fill the
// 16 x 30 tiles, 4, 6..8 or 13..16 input channels and enough output channels (the cost does not shrink with Cin below 8
// -- except for Cin <= 4, which packs a whole kernel row into one MMA -- nor with Cout; the SIMT kernels' does):
// 4->8 @160x192x160 432 -> 338 us, 8->8 755 -> 498, 8->16 @80x96x80 222 -> 103, 16->16 399 -> 318, 6->12 87 -> 78; 12->12 is
// behind (170 vs 148).  SMILE_CONV_SPLIT=0 keeps everything on the SIMT kernels, =2 takes every
// legal shape.  *handled = false otherwise.
int launch_conv3d_march_split(const float* in, const float* weight, const float* bias, float* out, const double* in_stats,
                              double* out_stats, int B, int Cin, int Cout, int D, int H, int W, int act_out, float eps,
                              cudaStream_t st, bool* handled) {
  *handled = false;
  static const int knob = [] { const char* e = getenv("SMILE_CONV_SPLIT"); return e ? atoi(e) : 1; }();
  if (knob == 0) return SMILE_OK;
  if (Cin < 2 || Cin > 16 || Cout > 16 || D < 1 || H < 2 || W < 2) return SMILE_OK;
  if ((long long)Cin * D * H * W >= (1LL << 30)) return SMILE_OK;   // 32-bit element offsets; |z| after InstanceNorm < 2^15
  if (knob != 2) {
    if (W < 60 || H < 32 || D < 8) return SMILE_OK;
    if (Cin <= 8 ? !((Cin == 4 || Cin >= 6) && Cout >= 6) : !(Cin >= 13 && Cout >= 8)) return SMILE_OK;
  }
  *handled = true;
  auto go = [&](int cin, int ci0, int pass) {
    if (Cout <= 8)
      return launch_march<1>(in, weight, bias, out, in_stats, out_stats, B, cin, Cout, D, H, W, act_out, eps, st, Cin, ci0, pass);
    return launch_march<2>(in, weight, bias, out, in_stats, out_stats, B, cin, Cout, D, H, W, act_out, eps, st, Cin, ci0, pass);
  };
  if (Cin <= 8) return go(Cin, 0, 0);
  const int rc = go(8, 0, 1);
  if (rc != SMILE_OK) return rc;
  return go(Cin - 8, 8, 2);
}

}  // namespace smile
