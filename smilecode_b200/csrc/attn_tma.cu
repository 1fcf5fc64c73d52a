// Fused heads==1 ModeT level, TMA-staged (the headline kernel).
//
//   w        = ModeTransformer(q, k)                       reference ModeT/models.py:308-334
//   flow_out = post * (SpatialTransformer(flow_in, w) + w)  models.py:403 / 408 (49-67)
//   moved    = SpatialTransformer(moving, flow_out)         models.py:410
//
// A CTA marches a column of TH rows x 32 voxels along D.  One producer warp streams, per plane,
// three TMA boxes into an mbarrier ring: the key plane with its 1-voxel halo (out-of-volume
// elements are zero-filled by TMA == the zero padding of models.py:319), the query plane and the
// three flow_in planes with halo.  TH consumer warps (lanes along W) keep the partial logits of
// the three voxels a key plane contributes to in registers, so each 24-byte key row is read
// from shared memory once (9 rows per output voxel instead of 27); dot products are packed
// fma.rn.f32x2.  |w| <= 1, so the compose sample lives in the same 3x3x3 window and is gathered
// from the flow ring (global-memory path when a corner leaves the window, i.e. |w| == 1).
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"

namespace smile {
namespace {

constexpr int TW = 32;        // voxels per tile row (one per lane)
// TMA (tiled mode) needs the innermost start coordinate to be a multiple of 16 bytes -- measured on
// B200: an unaligned start raises "illegal instruction" (tools/probe/tma_probe.cu) -- so the halo
// boxes start a little further left than the 1-voxel halo: 2 voxels (48 B) for keys, 4 floats for flow.
constexpr int KW = TW + 4;    // key row: voxels w0-2 .. w0+33
constexpr int KOFF = 1;       // tile column of voxel (w - 1) for lane 0
constexpr int FWP = 40;       // flow row: floats w0-4 .. w0+35 (160 B)
constexpr int FOFF = 3;       // tile column of voxel (w - 1) for lane 0
constexpr int HD = 6;         // head_dim of the reference configuration
constexpr float kLog2e = 1.4426950408889634f;

constexpr int MAXSEG = 16;    // depth segments one CTA may own (host caps units_per_cta accordingly)

struct Seg {
  int b, h0, w0, d_a, L, s_begin;  // s_begin: index of the segment's first stage in the CTA's stage sequence
};

template <int TH, int NS>
struct Cfg {
  static constexpr int NF = NS + 3;  // flow ring: planes p-2..p live + NS in flight
  static constexpr int KROWS = TH + 2;
  static constexpr int K_BYTES = KROWS * KW * HD * 4;
  static constexpr int K_STRIDE = (K_BYTES + 127) / 128 * 128;
  static constexpr int Q_BYTES = TH * TW * HD * 4;
  static constexpr int Q_STRIDE = (Q_BYTES + 127) / 128 * 128;
  static constexpr int F_PLANE = KROWS * FWP;  // floats per flow component
  static constexpr int F_BYTES = 3 * F_PLANE * 4;
  static constexpr int F_STRIDE = (F_BYTES + 127) / 128 * 128;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_Q = OFF_K + NS * K_STRIDE;
  static constexpr int OFF_F = OFF_Q + NS * Q_STRIDE;
  static constexpr int OFF_BAR = OFF_F + NF * F_STRIDE;   // NS full barriers
  static constexpr int OFF_CNT = OFF_BAR + 32;            // NS arrival counters
  static constexpr int OFF_RPB = OFF_CNT + 32;            // 28 floats
  static constexpr int OFF_SEG = OFF_RPB + 128;           // (MAXSEG + 1) segments
  static constexpr int SMEM = OFF_SEG + (MAXSEG + 1) * (int)sizeof(Seg) + 64;
  static constexpr int THREADS = TH * 32;
};

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
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
// A warp that runs ahead of its CTA must not burn issue slots polling (the arbiter favours it over
// the warps it is waiting for): back off with nanosleep between polls.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  do {
    __nanosleep(64);
  } while (!mbar_try_wait(bar, parity));
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ uint32_t atom_add_acq_rel(uint32_t addr, uint32_t v) {
  uint32_t old;
  asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(v) : "memory");
  return old;
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
__device__ __forceinline__ float2 add2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// <q, k> over the six channels: three packed FMAs + one add.
__device__ __forceinline__ float dot6(const float2 (&q)[3], float2 k0, float2 k1, float2 k2) {
  float2 a = mul2(q[0], k0);
  a = fma2(q[1], k1, a);
  a = fma2(q[2], k2, a);
  return a.x + a.y;
}

// Sampling coordinate of SpatialTransformer + grid_sample(align_corners=True): same value, bit for
// bit, as common.cuh:st_coord.  p/(S-1) is Markstein's correctly rounded division with the
// host-computed rc = RN(1/(S-1)); 2*(q-0.5) and (n+1)/2 are exact scalings, so
// x = RN(RN(RN(q - 0.5) + 0.5) * (S-1)).
__device__ __forceinline__ float st_coord_fast(float idx, float f, float sm1, float rc) {
  const float p = __fadd_rn(idx, f);
  const float q0 = __fmul_rn(p, rc);
  const float r = __fmaf_rn(-q0, sm1, p);
  const float q = __fmaf_rn(r, rc, q0);
  return __fmul_rn(__fadd_rn(__fsub_rn(q, 0.5f), 0.5f), sm1);
}

struct Dims {
  int B, D, H, W;
  int ncol_h, ncol_w;
  long long total_units;  // B * ncol_h * ncol_w * D plane-steps
  int units_per_cta;
  float dm1, hm1, wm1;    // S - 1
  float rd, rh, rw;       // RN(1 / (S - 1))
};

// Softmax over the 27 logits (log2 domain) and expectation of the tap offsets (models.py:328-332).
// L[] holds <q,k>; logit = qscale * <q,k> + rpb (rpb pre-scaled by log2 e) is formed here.
__device__ __forceinline__ void softmax_expect27(float (&L)[27], const float* __restrict__ s_rpb, float qscale,
                                                 float& od, float& oh, float& ow) {
  const float2 sc2 = make_float2(qscale, qscale);  // scale * log2(e)
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const float2 r = reinterpret_cast<const float2*>(s_rpb)[i];
    const float2 v = fma2(make_float2(L[2 * i], L[2 * i + 1]), sc2, r);
    L[2 * i] = v.x;
    L[2 * i + 1] = v.y;
  }
  L[26] = fmaf(L[26], qscale, s_rpb[26]);
  float m9[9];  // max as a depth-3 tree of 3-input max (13 instructions, no 13-deep dependent chain)
#pragma unroll
  for (int i = 0; i < 9; ++i) m9[i] = max3(L[3 * i], L[3 * i + 1], L[3 * i + 2]);
  const float m = max3(max3(m9[0], m9[1], m9[2]), max3(m9[3], m9[4], m9[5]), max3(m9[6], m9[7], m9[8]));
  const float2 nm = make_float2(-m, -m);
  float p[27];
#pragma unroll
  for (int i = 0; i < 13; ++i) {
    const float2 v = add2(make_float2(L[2 * i], L[2 * i + 1]), nm);
    p[2 * i] = ex2(v.x);
    p[2 * i + 1] = ex2(v.y);
  }
  p[26] = ex2(L[26] - m);
  float row[9], rw[9];
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    row[i] = (p[3 * i] + p[3 * i + 2]) + p[3 * i + 1];
    rw[i] = p[3 * i + 2] - p[3 * i];
  }
  const float c0 = (row[0] + row[1]) + row[2], c1 = (row[3] + row[4]) + row[5], c2 = (row[6] + row[7]) + row[8];
  const float sum = (c0 + c2) + c1;
  const float sd = c2 - c0;
  const float sh = ((row[2] - row[0]) + (row[5] - row[3])) + (row[8] - row[6]);
  const float sw = (((rw[0] + rw[1]) + (rw[2] + rw[3])) + ((rw[4] + rw[5]) + (rw[6] + rw[7]))) + rw[8];
  const float inv = rcp_approx(sum);
  od = sd * inv;
  oh = sh * inv;
  ow = sw * inv;
}

// One axis of a zeros-padded trilinear sample: corner indices clamped into the volume (so both
// loads are always legal) and the weight of an out-of-volume corner forced to 0, which adds +-0
// where torch's grid_sampler skips the corner (GridSampler.h:209-211) -- same sum.
__device__ __forceinline__ void axis_corners(float c, int S, int& i0, int& i1, float& w0, float& w1) {
  const float f = floorf(c);
  const int j0 = __float2int_rd(c), j1 = j0 + 1;
  w1 = __fsub_rn(c, f);
  w0 = __fsub_rn(__fadd_rn(f, 1.0f), c);
  w0 = ((unsigned)j0 < (unsigned)S) ? w0 : 0.f;
  w1 = ((unsigned)j1 < (unsigned)S) ? w1 : 0.f;
  i0 = min(max(j0, 0), S - 1);
  i1 = min(max(j1, 0), S - 1);
}

template <int TH, int NS, bool COMPOSE>
__device__ __forceinline__ void issue_stage(uint32_t sbase, const Seg* __restrict__ segs, int& pseg, int n,
                                            const CUtensorMap* tm_k, const CUtensorMap* tm_q, const CUtensorMap* tm_f) {
  using C = Cfg<TH, NS>;
  constexpr int NF = C::NF;
  while (n >= segs[pseg + 1].s_begin) ++pseg;
  const Seg sg = segs[pseg];
  const int p = sg.d_a - 1 + (n - sg.s_begin);
  const int slot = n % NS, fslot = n % NF;
  const uint32_t full = sbase + C::OFF_BAR + 8 * slot;
  mbar_expect_tx(full, C::K_BYTES + C::Q_BYTES + (COMPOSE ? C::F_BYTES : 0));
  tma_load_4d(sbase + C::OFF_K + slot * C::K_STRIDE, tm_k, full, (sg.w0 - 2) * HD, sg.h0 - 1, p, sg.b);
  tma_load_4d(sbase + C::OFF_Q + slot * C::Q_STRIDE, tm_q, full, sg.w0 * HD, sg.h0, p + 1, sg.b);
  if (COMPOSE) tma_load_4d(sbase + C::OFF_F + fslot * C::F_STRIDE, tm_f, full, sg.w0 - 4, sg.h0 - 1, p, sg.b * 3);
}

template <int TH, int NS, bool TWOPASS, bool COMPOSE, bool MOVED, int MINB>
__global__ void __launch_bounds__(TH * 32, MINB)
fused_march_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_q,
                   const __grid_constant__ CUtensorMap tm_f, const float* __restrict__ rpb,
                   const float* __restrict__ flow_in, const float* __restrict__ moving, float* __restrict__ out0,
                   float* __restrict__ moved, const Dims dm, float qscale, float post, int Cmov) {
  using C = Cfg<TH, NS>;
  constexpr int NF = C::NF;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + C::OFF_BAR;
  const uint32_t cnt_base = sbase + C::OFF_CNT;
  float* s_rpb = reinterpret_cast<float*>(smem + C::OFF_RPB);
  Seg* segs = reinterpret_cast<Seg*>(smem + C::OFF_SEG);
  int* s_nseg = reinterpret_cast<int*>(smem + C::OFF_SEG + (MAXSEG + 1) * sizeof(Seg));

  const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
  const int D = dm.D, H = dm.H, W = dm.W, HW = H * W;
  const int N = D * HW;  // < 2^31 / 6 (checked by the launcher)

  if (tid == 0) {
    const long long u_begin = (long long)blockIdx.x * dm.units_per_cta;
    long long u_end = u_begin + dm.units_per_cta;
    if (u_end > dm.total_units) u_end = dm.total_units;
    int ns = 0, stage = 0;
    long long u = u_begin;
    while (u < u_end && ns < MAXSEG) {
      const long long col = u / D;
      Seg sg;
      sg.d_a = (int)(u - col * D);
      sg.L = (int)((u_end - u) < (long long)(D - sg.d_a) ? (u_end - u) : (long long)(D - sg.d_a));
      sg.w0 = (int)(col % dm.ncol_w) * TW;
      const long long t2 = col / dm.ncol_w;
      sg.h0 = (int)(t2 % dm.ncol_h) * TH;
      sg.b = (int)(t2 / dm.ncol_h);
      sg.s_begin = stage;
      segs[ns++] = sg;
      stage += sg.L + 2;
      u += sg.L;
    }
    Seg sentinel = {0, 0, 0, 0, 0, stage};
    segs[ns] = sentinel;
    *s_nseg = ns;
    for (int i = 0; i < NS; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      reinterpret_cast<uint32_t*>(smem + C::OFF_CNT)[i] = 0;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (tid < 28) s_rpb[tid] = (tid < 27 && rpb != nullptr) ? rpb[tid] * kLog2e : 0.f;
  __syncthreads();
  const int nseg = *s_nseg;
  const int total_stages = segs[nseg].s_begin;
  int pseg = 0;  // segment cursor of the TMA issue path (per warp; only lane 0 uses it)
  if (tid == 0) {
    for (int n = 0; n < NS && n < total_stages; ++n) issue_stage<TH, NS, COMPOSE>(sbase, segs, pseg, n, &tm_k, &tm_q, &tm_f);
  }

  int slot = 0, fslot = 0, it = 0;
  uint32_t par = 0;
  bool ready = false;
  // per-thread shared-memory bases (bytes)
  const uint8_t* q_thr = smem + C::OFF_Q + (r * TW + lane) * (HD * 4);
  const uint8_t* k_thr = smem + C::OFF_K + (r * KW + lane + KOFF) * (HD * 4);
  const uint8_t* f_thr = smem + C::OFF_F + (r * FWP + lane + FOFF) * 4;

  for (int si = 0; si < nseg; ++si) {
    const Seg sg = segs[si];
    const int h = sg.h0 + r, wg = sg.w0 + lane;
    const bool valid = (h < H) && (wg < W);
    const float hf = (float)h, wf = (float)wg;
    float* ob = out0 + (long long)sg.b * 3 * N;
    const float* fb = COMPOSE ? flow_in + (long long)sg.b * 3 * N : nullptr;
    const float* mb = MOVED ? moving + (long long)sg.b * Cmov * N : nullptr;
    float* mvb = MOVED ? moved + (long long)sg.b * Cmov * N : nullptr;
    int vo = (sg.d_a - 2) * HW + h * W + wg;  // linear offset of the voxel that completes at the current stage
    int f_m2 = 0, f_m1 = 0;                  // flow ring byte offsets of planes p-2, p-1

    float2 q[3][3];
    float lg[3][27];
#pragma unroll
    for (int s = 0; s < 3; ++s)
#pragma unroll
      for (int i = 0; i < 3; ++i) q[s][i] = make_float2(0.f, 0.f);

    const int nsteps = sg.L + 2;
    for (int z0 = 0; z0 < nsteps; z0 += 3) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int z = z0 + j;
        if (z >= nsteps) break;
        const int sN = j, sM = (j + 2) % 3, sO = (j + 1) % 3;  // new (tap plane 0), middle (1), oldest (2: completes)
        if (!ready) mbar_wait(bar_full + 8 * slot, par);

        // query of the voxel that starts at this plane (depth p + 1), scaled into the log2 domain
        {
          const float2* qs = reinterpret_cast<const float2*>(q_thr + slot * C::Q_STRIDE);
          q[sN][0] = qs[0];
          q[sN][1] = qs[1];
          q[sN][2] = qs[2];
        }
        const float2* ks = reinterpret_cast<const float2*>(k_thr + slot * C::K_STRIDE);
        // pass 1: the nine taps that complete the oldest voxel (TWOPASS), or all 27 products of the plane
#pragma unroll
        for (int i = 0; i < 9; ++i) {
          const float2* kr = ks + ((i / 3) * KW + (i % 3)) * 3;
          const float2 k0 = kr[0], k1 = kr[1], k2 = kr[2];
          lg[sO][18 + i] = dot6(q[sO], k0, k1, k2);
          if (!TWOPASS) {
            lg[sN][i] = dot6(q[sN], k0, k1, k2);
            lg[sM][9 + i] = dot6(q[sM], k0, k1, k2);
          }
        }
        const int f_0 = fslot * C::F_STRIDE;

        float fo0 = 0.f, fo1 = 0.f, fo2 = 0.f;  // flow_out of the completed voxel (for the moved sample)
        if (z >= 2) {
          const int v = sg.d_a - 2 + z;  // completed voxel depth
          float w0, w1, w2;
          softmax_expect27(lg[sO], s_rpb, qscale, w0, w1, w2);
          if (!COMPOSE) {
            if (valid) {
              ob[vo] = w0;
              ob[vo + N] = w1;
              ob[vo + 2 * N] = w2;
            }
          } else {
            const float vf = (float)v;
            const float cz = st_coord_fast(vf, w0, dm.dm1, dm.rd);
            const float cy = st_coord_fast(hf, w1, dm.hm1, dm.rh);
            const float cx = st_coord_fast(wf, w2, dm.wm1, dm.rw);
            const float z0f = floorf(cz), y0f = floorf(cy), x0f = floorf(cx);
            const int lz = __float2int_rd(cz) - (v - 1), ly = __float2int_rd(cy) - (h - 1),
                      lx = __float2int_rd(cx) - (wg - 1);
            float f0, f1, f2;
            if (((unsigned)lz <= 1u) && ((unsigned)ly <= 1u) && ((unsigned)lx <= 1u)) {
              const float wz1 = __fsub_rn(cz, z0f), wy1 = __fsub_rn(cy, y0f), wx1 = __fsub_rn(cx, x0f);
              const float wz0 = __fsub_rn(__fadd_rn(z0f, 1.0f), cz), wy0 = __fsub_rn(__fadd_rn(y0f, 1.0f), cy),
                          wx0 = __fsub_rn(__fadd_rn(x0f, 1.0f), cx);
              const float w00 = __fmul_rn(wx0, wy0), w10 = __fmul_rn(wx1, wy0), w01 = __fmul_rn(wx0, wy1),
                          w11 = __fmul_rn(wx1, wy1);
              const float wa0 = __fmul_rn(w00, wz0), wa1 = __fmul_rn(w10, wz0), wa2 = __fmul_rn(w01, wz0),
                          wa3 = __fmul_rn(w11, wz0), wb0 = __fmul_rn(w00, wz1), wb1 = __fmul_rn(w10, wz1),
                          wb2 = __fmul_rn(w01, wz1), wb3 = __fmul_rn(w11, wz1);
              const int o = (ly * FWP + lx) * 4;
              const float* pa = reinterpret_cast<const float*>(f_thr + (lz ? f_m1 : f_m2) + o);
              const float* pb = reinterpret_cast<const float*>(f_thr + (lz ? f_0 : f_m1) + o);
              float acc[3];
#pragma unroll
              for (int c = 0; c < 3; ++c) {
                const float* a = pa + c * C::F_PLANE;
                const float* bq = pb + c * C::F_PLANE;
                float t = __fmul_rn(a[0], wa0);
                t = __fadd_rn(t, __fmul_rn(a[1], wa1));
                t = __fadd_rn(t, __fmul_rn(a[FWP], wa2));
                t = __fadd_rn(t, __fmul_rn(a[FWP + 1], wa3));
                t = __fadd_rn(t, __fmul_rn(bq[0], wb0));
                t = __fadd_rn(t, __fmul_rn(bq[1], wb1));
                t = __fadd_rn(t, __fmul_rn(bq[FWP], wb2));
                t = __fadd_rn(t, __fmul_rn(bq[FWP + 1], wb3));
                acc[c] = t;
              }
              f0 = acc[0];
              f1 = acc[1];
              f2 = acc[2];
            } else {
              // a corner left the staged window (|w| == 1 up to rounding): exact global-memory gather
              f0 = f1 = f2 = 0.f;
              if (valid) {
                TriSample s;
                tri_setup(s, cz, cy, cx, D, H, W);
                f0 = tri_gather(s, fb);
                f1 = tri_gather(s, fb + N);
                f2 = tri_gather(s, fb + 2 * N);
              }
            }
            f0 = __fmul_rn(post, __fadd_rn(f0, w0));
            f1 = __fmul_rn(post, __fadd_rn(f1, w1));
            f2 = __fmul_rn(post, __fadd_rn(f2, w2));
            fo0 = f0;
            fo1 = f1;
            fo2 = f2;
            if (valid) {
              ob[vo] = f0;
              ob[vo + N] = f1;
              ob[vo + 2 * N] = f2;
            }
          }
        }

        // moved = T(moving, flow_out): addresses + weights now, the eight loads fly during pass 2
        float mval[8], mwt[8];
        const bool do_moved = MOVED && COMPOSE && (z >= 2) && valid;
        if (MOVED) {
#pragma unroll
          for (int c = 0; c < 8; ++c) mval[c] = mwt[c] = 0.f;
        }
        if (do_moved) {
          int z0i, z1i, y0i, y1i, x0i, x1i;
          float uz0, uz1, uy0, uy1, ux0, ux1;
          axis_corners(st_coord_fast((float)(sg.d_a - 2 + z), fo0, dm.dm1, dm.rd), D, z0i, z1i, uz0, uz1);
          axis_corners(st_coord_fast(hf, fo1, dm.hm1, dm.rh), H, y0i, y1i, uy0, uy1);
          axis_corners(st_coord_fast(wf, fo2, dm.wm1, dm.rw), W, x0i, x1i, ux0, ux1);
          const float u00 = __fmul_rn(ux0, uy0), u10 = __fmul_rn(ux1, uy0), u01 = __fmul_rn(ux0, uy1),
                      u11 = __fmul_rn(ux1, uy1);
          mwt[0] = __fmul_rn(u00, uz0); mwt[1] = __fmul_rn(u10, uz0); mwt[2] = __fmul_rn(u01, uz0);
          mwt[3] = __fmul_rn(u11, uz0); mwt[4] = __fmul_rn(u00, uz1); mwt[5] = __fmul_rn(u10, uz1);
          mwt[6] = __fmul_rn(u01, uz1); mwt[7] = __fmul_rn(u11, uz1);
          const unsigned r00 = (unsigned)((z0i * H + y0i) * W), r01 = (unsigned)((z0i * H + y1i) * W),
                         r10 = (unsigned)((z1i * H + y0i) * W), r11 = (unsigned)((z1i * H + y1i) * W);
          const unsigned ux0i = (unsigned)x0i, ux1i = (unsigned)x1i;
          mval[0] = __ldg(mb + (r00 + ux0i)); mval[1] = __ldg(mb + (r00 + ux1i));
          mval[2] = __ldg(mb + (r01 + ux0i)); mval[3] = __ldg(mb + (r01 + ux1i));
          mval[4] = __ldg(mb + (r10 + ux0i)); mval[5] = __ldg(mb + (r10 + ux1i));
          mval[6] = __ldg(mb + (r11 + ux0i)); mval[7] = __ldg(mb + (r11 + ux1i));
        }

        // pass 2: first nine taps of the new voxel, middle nine of the previous one
#pragma unroll
        for (int i = 0; i < (TWOPASS ? 9 : 0); ++i) {
          const float2* kr = ks + ((i / 3) * KW + (i % 3)) * 3;
          const float2 k0 = kr[0], k1 = kr[1], k2 = kr[2];
          lg[sN][i] = dot6(q[sN], k0, k1, k2);
          lg[sM][9 + i] = dot6(q[sM], k0, k1, k2);
        }

        if (do_moved) {
          float t = __fmul_rn(mval[0], mwt[0]);
#pragma unroll
          for (int c = 1; c < 8; ++c) t = __fadd_rn(t, __fmul_rn(mval[c], mwt[c]));
          mvb[vo] = t;
        }

        // release the slot; the warp that arrives last re-arms it with the stage NS steps ahead
        __syncwarp();
        if (lane == 0) {
          const uint32_t prev = atom_add_acq_rel(cnt_base + 4 * slot, 1u);
          if (prev == TH - 1) {
            reinterpret_cast<volatile uint32_t*>(smem + C::OFF_CNT)[slot] = 0;
            const int n = it + NS;
            if (n < total_stages) {
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
              issue_stage<TH, NS, COMPOSE>(sbase, segs, pseg, n, &tm_k, &tm_q, &tm_f);
            }
          }
        }
        ++it;
        vo += HW;
        f_m2 = f_m1;
        f_m1 = f_0;
        if (++slot == NS) {
          slot = 0;
          par ^= 1u;
        }
        if (++fslot == NF) fslot = 0;
        // poll the next stage's barrier now so its ~90-cycle query latency hides behind this step's tail
        ready = (it < total_stages) && mbar_try_wait(bar_full + 8 * slot, par);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
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
    set_error("modet_fused(TMA): cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return false;
  }
  return true;
}

template <int TH, int NS, bool TWOPASS, bool COMPOSE, bool MOVED, int MINB>
int launch_cfg(const CUtensorMap& mk, const CUtensorMap& mq, const CUtensorMap& mf, const float* rpb, const float* flow_in,
               const float* moving, float* out0, float* moved, const Dims& dm, int grid, float qscale, float post, int Cmov,
               cudaStream_t st) {
  using C = Cfg<TH, NS>;
  auto kern = fused_march_kernel<TH, NS, TWOPASS, COMPOSE, MOVED, MINB>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) {
    set_error("modet_fused(TMA): cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  kern<<<grid, C::THREADS, C::SMEM, st>>>(mk, mq, mf, rpb, flow_in, moving, out0, moved, dm, qscale, post, Cmov);
  return check_launch("modet_fused(TMA)");
}

}  // namespace

// Handles head_dim == 6 levels whose rows meet TMA's 16-byte stride rule (W % 4 == 0); everything
// else stays on the generic CUDA kernels of attn.cu (*handled = false).
int launch_modet_attn_tma(const float* q, const float* k, const float* rpb, const float* flow_in, const float* moving,
                          float* w_out, float* flow_out, float* moved, int B, int D, int H, int W, float scale, float post,
                          int Cmov, cudaStream_t st, bool* handled) {
  *handled = false;
  constexpr int TH = 8;
  const bool compose = flow_in != nullptr;
  if (W % 4 != 0 || D < 2 || H < 2 || W < 2) return SMILE_OK;
  if (moved != nullptr && Cmov != 1) return SMILE_OK;  // the fused sampler handles the single-channel moving image
  if ((long long)D * H * W * HD >= (1LL << 31)) return SMILE_OK;
  if (get_encode() == nullptr) return SMILE_OK;
  *handled = true;

  Dims dm;
  dm.B = B; dm.D = D; dm.H = H; dm.W = W;
  dm.ncol_h = ceil_div(H, TH);
  dm.ncol_w = ceil_div(W, TW);
  dm.total_units = (long long)B * dm.ncol_h * dm.ncol_w * D;
  static const int variant_for_slots = [] { const char* e = getenv("SMILE_FUSED_VARIANT"); return e ? atoi(e) : 0; }();
  long long slots = (variant_for_slots == 4 ? 3LL : 2LL) * kNumSMs;
  long long per = ceil_div_ll(dm.total_units, slots);
  if (per < 4) per = 4;                                        // amortise the two halo planes of a segment
  if (per > (long long)(MAXSEG - 2) * D) per = (long long)(MAXSEG - 2) * D;  // bound the per-CTA segment table
  dm.units_per_cta = (int)per;
  const int grid = (int)ceil_div_ll(dm.total_units, per);
  dm.dm1 = (float)(D - 1); dm.hm1 = (float)(H - 1); dm.wm1 = (float)(W - 1);
  dm.rd = 1.0f / dm.dm1; dm.rh = 1.0f / dm.hm1; dm.rw = 1.0f / dm.wm1;

  CUtensorMap mk, mq, mf;
  const cuuint64_t qk_dims[4] = {(cuuint64_t)W * HD, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t qk_str[3] = {(cuuint64_t)W * HD * 4, (cuuint64_t)H * W * HD * 4, (cuuint64_t)D * H * W * HD * 4};
  const cuuint32_t k_box[4] = {KW * HD, TH + 2, 1, 1};
  const cuuint32_t q_box[4] = {TW * HD, TH, 1, 1};
  if (!encode4(&mk, k, qk_dims, qk_str, k_box) || !encode4(&mq, q, qk_dims, qk_str, q_box)) return SMILE_ERR_CUDA;
  if (compose) {
    const cuuint64_t f_dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * 3};
    const cuuint64_t f_str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
    const cuuint32_t f_box[4] = {FWP, TH + 2, 1, 3};
    if (!encode4(&mf, flow_in, f_dims, f_str, f_box)) return SMILE_ERR_CUDA;
  } else {
    mf = mq;
  }
  const float qscale = scale * kLog2e;
#define SMILE_LAUNCH(NSV, TP, MB)                                                                                            \
  do {                                                                                                                    \
    if (!compose)                                                                                                         \
      return launch_cfg<TH, NSV, TP, false, false, MB>(mk, mq, mf, rpb, nullptr, nullptr, w_out, nullptr, dm, grid, qscale,   \
                                                   1.0f, 0, st);                                                          \
    if (moved != nullptr)                                                                                                 \
      return launch_cfg<TH, NSV, TP, true, true, MB>(mk, mq, mf, rpb, flow_in, moving, flow_out, moved, dm, grid, qscale,     \
                                                 post, Cmov, st);                                                         \
    return launch_cfg<TH, NSV, TP, true, false, MB>(mk, mq, mf, rpb, flow_in, nullptr, flow_out, nullptr, dm, grid, qscale,   \
                                                post, 0, st);                                                             \
  } while (0)
  static const int variant = [] { const char* e = getenv("SMILE_FUSED_VARIANT"); return e ? atoi(e) : 0; }();
  switch (variant) {  // tuning knob for profiling runs; 0 is the production configuration
    case 1: SMILE_LAUNCH(4, false, 2);
    case 2: SMILE_LAUNCH(3, true, 2);
    case 4: SMILE_LAUNCH(3, false, 3);
    default: SMILE_LAUNCH(3, false, 2);
  }
#undef SMILE_LAUNCH
}

}  // namespace smile
