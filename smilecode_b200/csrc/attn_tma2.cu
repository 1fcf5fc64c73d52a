// Fused heads==1 ModeT level, second generation: two voxels per thread.
//
//   w        = ModeTransformer(q, k)                       reference ModeT/models.py:308-334
//   flow_out = post * (SpatialTransformer(flow_in, w) + w)  models.py:403 / 408 (49-67)
//   moved    = SpatialTransformer(moving, flow_out)         models.py:410
//
// Same TMA-staged marching scheme as attn_tma.cu (a CTA marches a column of rows x 32 voxels along D; per plane a key
// box with halo, a query box and a flow box land in an mbarrier ring; slots are re-armed by whichever warp first sees
// them released), but the FMA pipe -- which bounds the first kernel (r02n profile: 311 fma-pipe instructions per
// voxel at one issue per two cycles) -- now does about 200 per voxel:
//   * a thread owns the voxel PAIR (h, w), (h+1, w).  Every packed fp32x2 instruction works on the pair: lane x is
//     voxel A, lane y voxel B.  The two key rows the pair shares are multiplied against both queries with ONE scalar
//     broadcast operand (6 FFMA2 for two 6-channel dot products, no horizontal add), the two rows only one of them
//     needs are paired with each other.  81 FFMA2 per voxel for the 27 logits = the minimum for 162 FMAs.
//   * key rows are read from shared memory once per pair (18 LDS.64 per voxel instead of 27).
//   * relative position bias enters as the initial value of the dot-product accumulators (one broadcast LDS.64).
//   * softmax without a running maximum when the caller passes the LayerNorm affine parameters that produced q and k:
//     |logit| <= scale * (max|gamma| * sqrt(C) + |beta|_2)^2 + max|rpb| is evaluated in the kernel prologue and, when it
//     is far from the fp32 exponent range, exponentials are accumulated directly (no max tree, no rescale).  Otherwise
//     (no parameters given, or a large bound) the same loop runs with an online maximum -- safe for any input.
//   * coordinates, trilinear lerps, softmax sums: all packed over the pair; floor() of the moved-image coordinate by a
//     round-down add of 1.5 * 2^23 instead of F2I/I2F on the XU pipe that the 27 ex2 already load.
#include <cuda.h>
#include <cudaTypedefs.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "common.cuh"
#include "kernels.h"
#include "ptx_util.cuh"

namespace smile {
namespace {

constexpr int TW = 32;        // voxels per tile row (one per lane)
constexpr int KW = TW + 4;    // key row: voxels w0-2 .. w0+33 (TMA start must be 16-byte aligned: 2 voxels = 48 B)
constexpr int KOFF = 1;       // tile column of voxel (w - 1) for lane 0
constexpr int FWP = 40;       // flow row: floats w0-4 .. w0+35
constexpr int FOFF = 3;       // tile column of voxel (w - 1) for lane 0
constexpr int HD = 6;
constexpr float kLog2e = 1.4426950408889634f;
constexpr int MAXSEG = 16;
constexpr float kFastBoundLog2 = 100.0f;  // |logit * log2 e| below this: ex2 / 27-term sums cannot overflow or vanish

struct Seg {
  int b, h0, w0, d_a, L, s_begin;
};

template <int TH, int NS>
struct Cfg2 {
  static constexpr int NF = NS + 3;
  static constexpr int ROWS = 2 * TH;       // voxel rows of the tile
  static constexpr int KROWS = ROWS + 2;
  static constexpr int K_BYTES = KROWS * KW * HD * 4;
  static constexpr int K_STRIDE = (K_BYTES + 127) / 128 * 128;
  static constexpr int Q_BYTES = ROWS * TW * HD * 4;
  static constexpr int Q_STRIDE = (Q_BYTES + 127) / 128 * 128;
  static constexpr int F_PLANE = KROWS * FWP;
  static constexpr int F_BYTES = 3 * F_PLANE * 4;
  static constexpr int F_STRIDE = (F_BYTES + 127) / 128 * 128;
  static constexpr int OFF_K = 0;
  static constexpr int OFF_Q = OFF_K + NS * K_STRIDE;
  static constexpr int OFF_F = OFF_Q + NS * Q_STRIDE;
  static constexpr int OFF_BAR = OFF_F + NF * F_STRIDE;
  static constexpr int OFF_CNT = OFF_BAR + 32;
  static constexpr int OFF_NEXT = OFF_CNT + 32;
  static constexpr int OFF_BIAS = OFF_NEXT + 16;            // [3 phases][3 slots][3 row combos][3 dx] p2 (A, B), x log2 e
  static constexpr int OFF_SEG = OFF_BIAS + 81 * 8 + 8;     // + fast-path flag
  static constexpr int SMEM = OFF_SEG + (MAXSEG + 1) * (int)sizeof(Seg) + 64;
  static constexpr int THREADS = TH * 32;
};

struct Dims2 {
  int B, D, H, W;
  int ncol_h, ncol_w;
  long long total_units;   // (row band, plane) units: B * ncol_h * D
  int group;               // CTAs per group = ncol_w: the W-neighbour columns of a row band march in lockstep
  float dm1, hm1, wm1;
  float rd, rh, rw;
};

// Work partition, computed on the host: CTA group g marches the (row band, plane) units [u[g], u[g + 1]).  The ranges are
// balanced by STEPS, not units: a range that crosses the end of a band is two segments and every segment costs two ramp
// steps (its halo planes), so equal unit counts left the crossing groups 4 steps (9 %) behind the others.
struct UnitTable {
  int u[kNumSMs + 1];
};

// per-pair running softmax state, packed (lo: voxel A, hi: voxel B).  FAST: s, ah, aw are plain sums of exponentials, nd
// holds the first tap plane's sum until the last plane turns it into sum(last) - sum(first).  SAFE adds the running
// maximum m (log2 domain) and every update rescales.
struct Acc2 {
  p2 s, nd, ah, aw, m;
};

// Exact replay of common.cuh:st_coord for the voxel pair (see st_coord_fast above): packed, same roundings.
__device__ __forceinline__ p2 st_coord2(p2 idx, p2 f, float sm1, float nsm1, float rc) {
  const p2 p = padd(idx, f);
  const p2 q0 = pmuls(p, rc);
  const p2 r = pfmas(q0, nsm1, p);
  const p2 q = pfmas(r, rc, q0);
  return pmuls(padds(padds(q, -0.5f), 0.5f), sm1);
}

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}

// floor(c) for |c| < 2^22 without the XU pipe: t = RD(c + 1.5 * 2^23) carries floor(c) in its low mantissa bits.
__device__ __forceinline__ void floor_magic(float c, int& j, float& jf) {
  float t;
  asm("add.rm.f32 %0, %1, 0f4B400000;" : "=f"(t) : "f"(c));
  j = __float_as_int(t) - 0x4B400000;
  jf = __fsub_rn(t, 12582912.0f);
}

__device__ __forceinline__ void axis_corners2(float c, int S, int& i0, int& i1, float& fr, bool& in0, bool& in1) {
  const int j0 = __float2int_rd(c), j1 = j0 + 1;
  fr = (fabsf(c) < 1e9f) ? __fsub_rn(c, (float)j0) : 0.f;
  in0 = (unsigned)j0 < (unsigned)S;
  in1 = (unsigned)j1 < (unsigned)S;
  i0 = min(max(j0, 0), S - 1);
  i1 = min(max(j1, 0), S - 1);
}

struct F3b { float a, b, c; };
__device__ __noinline__ F3b compose_sample_global2(const float* __restrict__ fb, float cz, float cy, float cx, int D, int H,
                                                   int W) {
  TriSample s;
  tri_setup(s, cz, cy, cx, D, H, W);
  const int N = D * H * W;
  F3b r;
  r.a = tri_gather(s, fb);
  r.b = tri_gather(s, fb + N);
  r.c = tri_gather(s, fb + 2 * N);
  return r;
}

struct Corners8b { float v[8]; float fz, fy, fx; };
__device__ __forceinline__ Corners8b moved_corners_border2(const float* __restrict__ mb, float mz, float my, float mx, int D,
                                                        int H, int W, bool active) {
  int z0i, z1i, y0i, y1i, x0i, x1i;
  bool zi0, zi1, yi0, yi1, xi0, xi1;
  Corners8b c;
  axis_corners2(mz, D, z0i, z1i, c.fz, zi0, zi1);
  axis_corners2(my, H, y0i, y1i, c.fy, yi0, yi1);
  axis_corners2(mx, W, x0i, x1i, c.fx, xi0, xi1);
  const float* r00 = mb + (z0i * H + y0i) * W;
  const float* r01 = mb + (z0i * H + y1i) * W;
  const float* r10 = mb + (z1i * H + y0i) * W;
  const float* r11 = mb + (z1i * H + y1i) * W;
  zi0 = zi0 && active;
  zi1 = zi1 && active;
  c.v[0] = (zi0 && yi0 && xi0) ? __ldg(r00 + x0i) : 0.f;
  c.v[1] = (zi0 && yi0 && xi1) ? __ldg(r00 + x1i) : 0.f;
  c.v[2] = (zi0 && yi1 && xi0) ? __ldg(r01 + x0i) : 0.f;
  c.v[3] = (zi0 && yi1 && xi1) ? __ldg(r01 + x1i) : 0.f;
  c.v[4] = (zi1 && yi0 && xi0) ? __ldg(r10 + x0i) : 0.f;
  c.v[5] = (zi1 && yi0 && xi1) ? __ldg(r10 + x1i) : 0.f;
  c.v[6] = (zi1 && yi1 && xi0) ? __ldg(r11 + x0i) : 0.f;
  c.v[7] = (zi1 && yi1 && xi1) ? __ldg(r11 + x1i) : 0.f;
  return c;
}

template <int TH, int NS, bool COMPOSE>
__device__ __noinline__ void issue_stage2(uint32_t sbase, const Seg* __restrict__ segs, int n, const CUtensorMap* tm_k,
                                          const CUtensorMap* tm_q, const CUtensorMap* tm_f) {
  using C = Cfg2<TH, NS>;
  constexpr int NF = C::NF;
  int pseg = 0;
  while (n >= segs[pseg + 1].s_begin) ++pseg;
  const Seg sg = segs[pseg];
  const int p = sg.d_a - 1 + (n - sg.s_begin);
  const int slot = n % NS, fslot = n % NF;
  const uint32_t full = sbase + C::OFF_BAR + 8 * slot;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  mbar_expect_tx(full, C::K_BYTES + C::Q_BYTES + (COMPOSE ? C::F_BYTES : 0));
  tma_load_4d(sbase + C::OFF_K + slot * C::K_STRIDE, tm_k, full, (sg.w0 - 2) * HD, sg.h0 - 1, p, sg.b);
  tma_load_4d(sbase + C::OFF_Q + slot * C::Q_STRIDE, tm_q, full, sg.w0 * HD, sg.h0, p + 1, sg.b);
  if (COMPOSE) tma_load_4d(sbase + C::OFF_F + fslot * C::F_STRIDE, tm_f, full, sg.w0 - 4, sg.h0 - 1, p, sg.b * 3);
}

template <int TH, int NS, bool COMPOSE>
__device__ __forceinline__ void try_issue2(uint32_t sbase, const Seg* __restrict__ segs, int total_stages,
                                           const CUtensorMap* tm_k, const CUtensorMap* tm_q, const CUtensorMap* tm_f) {
  using C = Cfg2<TH, NS>;
  const uint32_t next_addr = sbase + C::OFF_NEXT;
  const uint32_t n = lds_volatile(next_addr);
  if ((int)n < total_stages) {
    const uint32_t k = n / NS;
    if (mbar_test(sbase + C::OFF_CNT + 8 * (n - k * NS), (k - 1) & 1u)) {
      if (atom_cas_relaxed(next_addr, n, n + 1) == n) issue_stage2<TH, NS, COMPOSE>(sbase, segs, (int)n, tm_k, tm_q, tm_f);
    }
  }
}

// Fold the three logit pairs of one key-row combination into the running state of a voxel pair.  The update is the same
// for every tap plane (the marching loop is NOT unrolled over the three in-flight slots: the unrolled body was 72 KB of
// code and `no_instruction` became the second largest stall, r03b profile): a slot's accumulators are reset when its
// pair completes, so "+=" also serves the first plane, and the depth numerator uses a per-slot coefficient
// cnd = -1 / 0 / +1 (first / middle / last tap plane) that rotates with the slot's role.
//   COMBO: 0 = rows (hA-1 | hB+1): dy = -1 for A, +1 for B;  1 = row hA: dy = 0 for A, -1 for B;
//          2 = row hB: dy = +1 for A, 0 for B
// SAFE: l0..l2 are logits (log2 domain); the running maximum m starts at -1e30, so the first rescale factor is 0.
// FAST: they are already the exponentials, and the plane total is accumulated in P (end_plane adds it to s / nd).
template <int COMBO, bool SAFE>
__device__ __forceinline__ void fold_row(Acc2& A, p2& P, p2 cnd, p2 l0, p2 l1, p2 l2) {
  // dy of (A, B) for this row combination as packed multipliers of the row sum
  const p2 kdy = (COMBO == 0) ? pk(-1.f, 1.f) : (COMBO == 1) ? pk(0.f, -1.f) : pk(1.f, 0.f);
  if (SAFE) {
    const p2 mn = pk(fmaxf(max3(lo(l0), lo(l1), lo(l2)), lo(A.m)), fmaxf(max3(hi(l0), hi(l1), hi(l2)), hi(A.m)));
    const p2 a = pex2(psub(A.m, mn));
    A.m = mn;
    const p2 e0 = pex2(psub(l0, mn)), e1 = pex2(psub(l1, mn)), e2 = pex2(psub(l2, mn));
    const p2 R = padd(padd(e0, e2), e1);
    A.s = pfma(A.s, a, R);
    A.aw = pfma(A.aw, a, psub(e2, e0));
    A.ah = pfma(A.ah, a, pmul(R, kdy));
    A.nd = pfma(A.nd, a, pmul(R, cnd));
  } else {
    const p2 R = padd(padd(l0, l2), l1);
    A.aw = padd(A.aw, psub(l2, l0));
    A.ah = pfma(R, kdy, A.ah);
    P = (COMBO == 0) ? R : padd(P, R);
  }
}

template <bool SAFE>
__device__ __forceinline__ void end_plane(Acc2& A, p2 P, p2 cnd) {
  if (SAFE) return;
  A.s = padd(A.s, P);
  A.nd = pfma(P, cnd, A.nd);
}

__device__ __forceinline__ void reset_acc(Acc2& A) {
  A.s = A.nd = A.ah = A.aw = 0;
  A.m = pk(-1e30f, -1e30f);
}

__device__ __forceinline__ float ldg_pred(const float* p, bool pr) {
  float v;
  asm("{\n\t.reg .pred p;\n\tsetp.ne.s32 p, %2, 0;\n\tmov.f32 %0, 0f00000000;\n\t@p ld.global.nc.f32 %0, [%1];\n\t}"
      : "=f"(v)
      : "l"(p), "r"((int)pr));
  return v;
}

// The marching loop of one CTA.  SAFE selects the online-maximum softmax.
template <int TH, int NS, bool COMPOSE, bool MOVED, bool SAFE, int VAR>
__device__ __forceinline__ void march(uint8_t* smem, const CUtensorMap* tm_k, const CUtensorMap* tm_q, const CUtensorMap* tm_f,
                                      const float* __restrict__ flow_in, const float* __restrict__ moving,
                                      float* __restrict__ out0, float* __restrict__ moved, const Dims2& dm, float qscale,
                                      float post, int Cmov, int nseg, int total_stages, int four) {
  using C = Cfg2<TH, NS>;
  constexpr int NF = C::NF;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + C::OFF_BAR;
  const uint32_t cnt_base = sbase + C::OFF_CNT;
  const Seg* segs = reinterpret_cast<const Seg*>(smem + C::OFF_SEG);
  const p2* s_bias = reinterpret_cast<const p2*>(smem + C::OFF_BIAS);

  const int tid = threadIdx.x, lane = tid & 31, r = tid >> 5;
  const int D = dm.D, H = dm.H, W = dm.W, HW = H * W;
  const int N = D * HW;
  const float ndm1 = -dm.dm1, nhm1 = -dm.hm1, nwm1 = -dm.wm1;

  if (tid == 0) {
    for (int n = 0; n < NS && n < total_stages; ++n) issue_stage2<TH, NS, COMPOSE>(sbase, segs, n, tm_k, tm_q, tm_f);
  }

  int slot = 0, fslot = 0;
  uint32_t par = 0;
  // wait for a "full" barrier; a warp that has to wait keeps offering to issue the next stage
  auto wait_full = [&](uint32_t bar, uint32_t parity) {
    if (!mbar_try_wait_hint(bar, parity, 200)) {
      do {
        if (lane == 0) try_issue2<TH, NS, COMPOSE>(sbase, segs, total_stages, tm_k, tm_q, tm_f);
        __syncwarp();
      } while (!mbar_try_wait_hint(bar, parity, 200));
    }
  };
  const uint8_t* q_thr = smem + C::OFF_Q + ((2 * r) * TW + lane) * (HD * 4);
  const uint8_t* k_thr = smem + C::OFF_K + ((2 * r) * KW + lane + KOFF) * (HD * 4);
  const uint8_t* f_thr = smem + C::OFF_F + ((2 * r) * FWP + lane + FOFF) * 4;

  for (int si = 0; si < nseg; ++si) {
    const Seg sg = segs[si];
    const int hA = sg.h0 + 2 * r, wg = sg.w0 + lane;
    const bool validA = (hA < H) && (wg < W), validB = (hA + 1 < H) && (wg < W);
    const float hfA = (float)hA, hfB = (float)(hA + 1), wf = (float)wg;
    const p2 hf2 = pk(hfA, hfB), wf2 = pk(wf, wf);
    // Element offsets from the tensor bases fit in 32 bits (launcher checks B * 3 * N < 2^32): an address is one
    // IMAD.WIDE.U32 on a kernel-parameter base instead of a 64-bit per-batch pointer held (or rebuilt) in registers.
    const float* fb = COMPOSE ? flow_in + (long long)sg.b * 3 * N : nullptr;    // rare out-of-window path only
    const unsigned boff1 = (unsigned)sg.b * (unsigned)N;                          // moving / moved (one channel)
    unsigned vo = (unsigned)sg.b * 3u * (unsigned)N + (unsigned)((sg.d_a - 3) * HW + hA * W + wg);   // voxel A in out0; B is + W
    unsigned vm = boff1 + (unsigned)((sg.d_a - 3) * HW + hA * W + wg);                               // voxel A in moved
    float vf = (float)(sg.d_a - 3);
    int f_m3 = 0, f_m2 = 0, f_m1 = 0;

    p2 qq[3][HD];          // [slot][channel] = (qA, qB) * qscale
    Acc2 acc[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
      for (int c = 0; c < HD; ++c) qq[s][c] = 0;
      reset_acc(acc[s]);
    }
    p2 w0 = 0, w1 = 0, w2 = 0;   // attention output of the pair completed by the previous iteration

    const int nsteps = sg.L + 2;
    int j = 0;                   // step phase: slot j takes the new pair, slot (j + 1) % 3 completes
#pragma unroll 1
    for (int z = 0; z <= nsteps; ++z) {
      {
        if (lane == 0) try_issue2<TH, NS, COMPOSE>(sbase, segs, total_stages, tm_k, tm_q, tm_f);
        __syncwarp();
        p2 mv[8];           // moved-image corners (A, B): z0y0x0 z0y0x1 z0y1x0 z0y1x1 z1y0x0 ...
        p2 mfx = 0, mfy = 0, mfz = 0;
        bool pend = false;

        // ---- (1) the pair completed by the previous iteration
        if (z >= 3) {
          if (!COMPOSE) {
            if (validA) {
              out0[vo] = lo(w0);
              out0[vo + (unsigned)N] = lo(w1);
              out0[vo + 2u * (unsigned)N] = lo(w2);
            }
            if (validB) {
              out0[vo + (unsigned)W] = hi(w0);
              out0[vo + (unsigned)W + (unsigned)N] = hi(w1);
              out0[vo + (unsigned)W + 2u * (unsigned)N] = hi(w2);
            }
          } else {
            const p2 cz = st_coord2(pk(vf, vf), w0, dm.dm1, ndm1, dm.rd);
            const p2 cy = st_coord2(hf2, w1, dm.hm1, nhm1, dm.rh);
            const p2 cx = st_coord2(wf2, w2, dm.wm1, nwm1, dm.rw);
            const float czA = lo(cz), czB = hi(cz), cyA = lo(cy), cyB = hi(cy), cxA = lo(cx), cxB = hi(cx);
            // floor(c) is idx - 1 or idx while the sample stays inside the staged window (exact float compares)
            const bool bzA = czA >= vf, byA = cyA >= hfA, bxA = cxA >= wf;
            const bool bzB = czB >= vf, byB = cyB >= hfB, bxB = cxB >= wf;
            const bool inA = (czA >= vf - 1.0f) && (czA < vf + 1.0f) && (cyA >= hfA - 1.0f) && (cyA < hfA + 1.0f) &&
                             (cxA >= wf - 1.0f) && (cxA < wf + 1.0f);
            const bool inB = (czB >= vf - 1.0f) && (czB < vf + 1.0f) && (cyB >= hfB - 1.0f) && (cyB < hfB + 1.0f) &&
                             (cxB >= wf - 1.0f) && (cxB < wf + 1.0f);
            const p2 fz = psub(cz, pk(bzA ? vf : vf - 1.0f, bzB ? vf : vf - 1.0f));
            const p2 fy = psub(cy, pk(byA ? hfA : hfA - 1.0f, byB ? hfB : hfB - 1.0f));
            const p2 fx = psub(cx, pk(bxA ? wf : wf - 1.0f, bxB ? wf : wf - 1.0f));
            const p2 one = pk(1.f, 1.f);
            const p2 gz = psub(one, fz), gy = psub(one, fy), gx = psub(one, fx);
            // window corner (z-1|z, y-1|y, x-1|x) of each voxel in the flow ring; B's window starts one row below A's
            const int oA = (byA ? FWP * 4 : 0) + (bxA ? 4 : 0);
            const int oB = (byB ? 2 * FWP * 4 : FWP * 4) + (bxB ? 4 : 0);
            const float* paA = reinterpret_cast<const float*>(f_thr + (bzA ? f_m2 : f_m3) + oA);
            const float* pbA = reinterpret_cast<const float*>(f_thr + (bzA ? f_m1 : f_m2) + oA);
            const float* paB = reinterpret_cast<const float*>(f_thr + (bzB ? f_m2 : f_m3) + oB);
            const float* pbB = reinterpret_cast<const float*>(f_thr + (bzB ? f_m1 : f_m2) + oB);
            p2 fo[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              const int o = c * C::F_PLANE;
              const p2 a00 = pk(paA[o], paB[o]), a01 = pk(paA[o + 1], paB[o + 1]);
              const p2 a10 = pk(paA[o + FWP], paB[o + FWP]), a11 = pk(paA[o + FWP + 1], paB[o + FWP + 1]);
              const p2 b00 = pk(pbA[o], pbB[o]), b01 = pk(pbA[o + 1], pbB[o + 1]);
              const p2 b10 = pk(pbA[o + FWP], pbB[o + FWP]), b11 = pk(pbA[o + FWP + 1], pbB[o + FWP + 1]);
              const p2 ra0 = pfma(a01, fx, pmul(a00, gx)), ra1 = pfma(a11, fx, pmul(a10, gx));
              const p2 rb0 = pfma(b01, fx, pmul(b00, gx)), rb1 = pfma(b11, fx, pmul(b10, gx));
              const p2 sa = pfma(ra1, fy, pmul(ra0, gy)), sb = pfma(rb1, fy, pmul(rb0, gy));
              fo[c] = pfma(sb, fz, pmul(sa, gz));
            }
            if (!(inA && inB)) {      // a corner left the staged window (|w| == 1 up to rounding): exact global gather
              float a[3] = {lo(fo[0]), lo(fo[1]), lo(fo[2])}, bq[3] = {hi(fo[0]), hi(fo[1]), hi(fo[2])};
              if (!inA) {
                a[0] = a[1] = a[2] = 0.f;
                if (validA) {
                  const F3b g = compose_sample_global2(fb, czA, cyA, cxA, D, H, W);
                  a[0] = g.a; a[1] = g.b; a[2] = g.c;
                }
              }
              if (!inB) {
                bq[0] = bq[1] = bq[2] = 0.f;
                if (validB) {
                  const F3b g = compose_sample_global2(fb, czB, cyB, cxB, D, H, W);
                  bq[0] = g.a; bq[1] = g.b; bq[2] = g.c;
                }
              }
#pragma unroll
              for (int c = 0; c < 3; ++c) fo[c] = pk(a[c], bq[c]);
            }
            const p2 f0 = pmuls(padd(fo[0], w0), post);
            const p2 f1 = pmuls(padd(fo[1], w1), post);
            const p2 f2v = pmuls(padd(fo[2], w2), post);
            if (validA) {
              out0[vo] = lo(f0);
              out0[vo + (unsigned)N] = lo(f1);
              out0[vo + 2u * (unsigned)N] = lo(f2v);
            }
            if (validB) {
              out0[vo + (unsigned)W] = hi(f0);
              out0[vo + (unsigned)W + (unsigned)N] = hi(f1);
              out0[vo + (unsigned)W + 2u * (unsigned)N] = hi(f2v);
            }
            if (MOVED) {
              const p2 mz = st_coord2(pk(vf, vf), f0, dm.dm1, ndm1, dm.rd);
              const p2 my = st_coord2(hf2, f1, dm.hm1, nhm1, dm.rh);
              const p2 mx = st_coord2(wf2, f2v, dm.wm1, nwm1, dm.rw);
              int jzA, jyA, jxA, jzB, jyB, jxB;
              float zfA, yfA, xfA, zfB, yfB, xfB;
              floor_magic(lo(mz), jzA, zfA);
              floor_magic(lo(my), jyA, yfA);
              floor_magic(lo(mx), jxA, xfA);
              floor_magic(hi(mz), jzB, zfB);
              floor_magic(hi(my), jyB, yfB);
              floor_magic(hi(mx), jxB, xfB);
              // VAR bit 2: zero padding is part of the straight path -- every corner is a predicated load whose predicate is
              // "inside the volume", whatever the sample position.  Before, a warp with ONE lane sampling across a face of
              // the volume took the slow path below with all its lanes; the CTAs that own border rows / columns / planes
              // were the last to finish (ncu: sm__cycles_active avg = 0.89 of max on border-crossing flows, 0.92
              // now).  Only coordinates outside floor_magic's exact range (|c| >= 2^22, NaN) still leave.
              constexpr bool XIN = (VAR & 4) != 0;
              const bool intA = validA && ((unsigned)jzA < (unsigned)(D - 1)) && ((unsigned)jyA < (unsigned)(H - 1)) &&
                                ((unsigned)jxA < (unsigned)(W - 1));
              const bool intB = validB && ((unsigned)jzB < (unsigned)(D - 1)) && ((unsigned)jyB < (unsigned)(H - 1)) &&
                                ((unsigned)jxB < (unsigned)(W - 1));
              const bool tame = (fabsf(lo(mz)) < 4.0e6f) && (fabsf(lo(my)) < 4.0e6f) && (fabsf(lo(mx)) < 4.0e6f) &&
                                (fabsf(hi(mz)) < 4.0e6f) && (fabsf(hi(my)) < 4.0e6f) && (fabsf(hi(mx)) < 4.0e6f);
              if (XIN && __all_sync(0xffffffffu, tame)) {
                mfz = psub(mz, pk(zfA, zfB));
                mfy = psub(my, pk(yfA, yfB));
                mfx = psub(mx, pk(xfA, xfB));
                const bool z0A = validA && ((unsigned)jzA < (unsigned)D), z1A = validA && ((unsigned)(jzA + 1) < (unsigned)D);
                const bool z0B = validB && ((unsigned)jzB < (unsigned)D), z1B = validB && ((unsigned)(jzB + 1) < (unsigned)D);
                const bool y0A = (unsigned)jyA < (unsigned)H, y1A = (unsigned)(jyA + 1) < (unsigned)H;
                const bool y0B = (unsigned)jyB < (unsigned)H, y1B = (unsigned)(jyB + 1) < (unsigned)H;
                const bool x0A = (unsigned)jxA < (unsigned)W, x1A = (unsigned)(jxA + 1) < (unsigned)W;
                const bool x0B = (unsigned)jxB < (unsigned)W, x1B = (unsigned)(jxB + 1) < (unsigned)W;
                // signed element offsets: corners outside the volume form addresses that are never dereferenced
                // (|offset| < 2^31 because |j| < 2^22 here and B * N < 2^31)
                const int iA = (int)(boff1 + (((unsigned)jzA * (unsigned)H + (unsigned)jyA) * (unsigned)W + (unsigned)jxA));
                const int iB = (int)(boff1 + (((unsigned)jzB * (unsigned)H + (unsigned)jyB) * (unsigned)W + (unsigned)jxB));
                const float *pA0 = moving + (ptrdiff_t)iA, *pA1 = moving + (ptrdiff_t)(iA + W),
                            *pA2 = moving + (ptrdiff_t)(iA + HW), *pA3 = moving + (ptrdiff_t)(iA + HW + W);
                const float *pB0 = moving + (ptrdiff_t)iB, *pB1 = moving + (ptrdiff_t)(iB + W),
                            *pB2 = moving + (ptrdiff_t)(iB + HW), *pB3 = moving + (ptrdiff_t)(iB + HW + W);
                mv[0] = pk(ldg_pred(pA0, z0A && y0A && x0A), ldg_pred(pB0, z0B && y0B && x0B));
                mv[1] = pk(ldg_pred(pA0 + 1, z0A && y0A && x1A), ldg_pred(pB0 + 1, z0B && y0B && x1B));
                mv[2] = pk(ldg_pred(pA1, z0A && y1A && x0A), ldg_pred(pB1, z0B && y1B && x0B));
                mv[3] = pk(ldg_pred(pA1 + 1, z0A && y1A && x1A), ldg_pred(pB1 + 1, z0B && y1B && x1B));
                mv[4] = pk(ldg_pred(pA2, z1A && y0A && x0A), ldg_pred(pB2, z1B && y0B && x0B));
                mv[5] = pk(ldg_pred(pA2 + 1, z1A && y0A && x1A), ldg_pred(pB2 + 1, z1B && y0B && x1B));
                mv[6] = pk(ldg_pred(pA3, z1A && y1A && x0A), ldg_pred(pB3, z1B && y1B && x0B));
                mv[7] = pk(ldg_pred(pA3 + 1, z1A && y1A && x1A), ldg_pred(pB3 + 1, z1B && y1B && x1B));
              } else if (!XIN && __all_sync(0xffffffffu, intA && intB)) {
                mfz = psub(mz, pk(zfA, zfB));
                mfy = psub(my, pk(yfA, yfB));
                mfx = psub(mx, pk(xfA, xfB));
                const unsigned iA = boff1 + (unsigned)((jzA * H + jyA) * W + jxA), iB = boff1 + (unsigned)((jzB * H + jyB) * W + jxB);
                const float *pA0 = moving + iA, *pA1 = moving + (iA + (unsigned)W), *pA2 = moving + (iA + (unsigned)HW),
                            *pA3 = moving + (iA + (unsigned)(HW + W));
                const float *pB0 = moving + iB, *pB1 = moving + (iB + (unsigned)W), *pB2 = moving + (iB + (unsigned)HW),
                            *pB3 = moving + (iB + (unsigned)(HW + W));
                mv[0] = pk(__ldg(pA0), __ldg(pB0));
                mv[1] = pk(__ldg(pA0 + 1), __ldg(pB0 + 1));
                mv[2] = pk(__ldg(pA1), __ldg(pB1));
                mv[3] = pk(__ldg(pA1 + 1), __ldg(pB1 + 1));
                mv[4] = pk(__ldg(pA2), __ldg(pB2));
                mv[5] = pk(__ldg(pA2 + 1), __ldg(pB2 + 1));
                mv[6] = pk(__ldg(pA3), __ldg(pB3));
                mv[7] = pk(__ldg(pA3 + 1), __ldg(pB3 + 1));
              } else {
                const Corners8b ca = moved_corners_border2(moving + boff1, lo(mz), lo(my), lo(mx), D, H, W, validA);
                const Corners8b cb = moved_corners_border2(moving + boff1, hi(mz), hi(my), hi(mx), D, H, W, validB);
#pragma unroll
                for (int i = 0; i < 8; ++i) mv[i] = pk(ca.v[i], cb.v[i]);
                mfz = pk(ca.fz, cb.fz);
                mfy = pk(ca.fy, cb.fy);
                mfx = pk(ca.fx, cb.fx);
              }
              pend = true;
            }
          }
        }

        // ---- (2) key plane of this iteration
        if (z < nsteps) {
          wait_full(bar_full + 8 * slot, par);
          {   // queries of the new pair into slot j (uniform branch on the phase: the loads target the slot's registers)
            const p2* qa = reinterpret_cast<const p2*>(q_thr + slot * C::Q_STRIDE);
            const p2* qb = reinterpret_cast<const p2*>(q_thr + slot * C::Q_STRIDE + TW * HD * 4);
            const p2 a0 = qa[0], a1 = qa[1], a2 = qa[2], b0 = qb[0], b1 = qb[1], b2 = qb[2];
#define SMILE_LOAD_Q(S)                                   \
  qq[S][0] = pmuls(pk(lo(a0), lo(b0)), qscale);           \
  qq[S][1] = pmuls(pk(hi(a0), hi(b0)), qscale);           \
  qq[S][2] = pmuls(pk(lo(a1), lo(b1)), qscale);           \
  qq[S][3] = pmuls(pk(hi(a1), hi(b1)), qscale);           \
  qq[S][4] = pmuls(pk(lo(a2), lo(b2)), qscale);           \
  qq[S][5] = pmuls(pk(hi(a2), hi(b2)), qscale);
            if (j == 0) {
              SMILE_LOAD_Q(0)
            } else if (j == 1) {
              SMILE_LOAD_Q(1)
            } else {
              SMILE_LOAD_Q(2)
            }
#undef SMILE_LOAD_Q
          }
          // depth coefficient of each slot at this phase: tap plane (j - s) mod 3 = 0 / 1 / 2  ->  -1 / 0 / +1
          const p2 cm = pk(-1.f, -1.f), cz0 = 0, cp = pk(1.f, 1.f);
          p2 cnd[3];
          if (j == 0) {
            cnd[0] = cm; cnd[1] = cp; cnd[2] = cz0;
          } else if (j == 1) {
            cnd[0] = cz0; cnd[1] = cm; cnd[2] = cp;
          } else {
            cnd[0] = cp; cnd[1] = cz0; cnd[2] = cm;
          }
          const p2* ks = reinterpret_cast<const p2*>(k_thr + slot * C::K_STRIDE);
          const p2* bias = s_bias + j * 27;     // [slot][combo][dx] at this phase
          p2 P[3] = {0, 0, 0};
#pragma unroll
          for (int combo = 0; combo < 3; ++combo) {
            p2 E[3][3];     // [slot][dx]: exponentials (FAST) / logits (SAFE) of the pair (A, B)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) {
              p2 a0 = bias[(0 * 3 + combo) * 3 + dx], a1 = bias[(1 * 3 + combo) * 3 + dx], a2 = bias[(2 * 3 + combo) * 3 + dx];
              if (combo == 0) {
                // Rows only one voxel of the pair needs: hA-1 (tile row 2r, dy = -1) feeds A, hB+1 (tile row 2r+3, dy = +1)
                // feeds B.  Each is multiplied in broadcast form against the (A, B) query pairs and only one lane of the
                // result is used (A's from the first chain, B's from the second): twice the FFMA2 of a packed-packed
                // product, but the FMA pipe has the room (23 % busy), whereas interleaving the two rows into register
                // pairs costs ~170 moves per step with LDS.64 and doubles the shared-memory wavefronts with LDS.32
                // (both measured, r03b / r03c).
                const p2* ka = ks + (0 * KW + dx) * 3;
                const p2* kb = ks + (3 * KW + dx) * 3;
                p2 c0 = a0, c1 = a1, c2 = a2;      // second chains (row hB+1); lane hi carries B's logit
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                  const p2 va = ka[i], vb = kb[i];
                  const float ax = lo(va), ay = hi(va), bx = lo(vb), by = hi(vb);
                  a0 = pfmas(qq[0][2 * i], ax, a0);
                  a1 = pfmas(qq[1][2 * i], ax, a1);
                  a2 = pfmas(qq[2][2 * i], ax, a2);
                  c0 = pfmas(qq[0][2 * i], bx, c0);
                  c1 = pfmas(qq[1][2 * i], bx, c1);
                  c2 = pfmas(qq[2][2 * i], bx, c2);
                  a0 = pfmas(qq[0][2 * i + 1], ay, a0);
                  a1 = pfmas(qq[1][2 * i + 1], ay, a1);
                  a2 = pfmas(qq[2][2 * i + 1], ay, a2);
                  c0 = pfmas(qq[0][2 * i + 1], by, c0);
                  c1 = pfmas(qq[1][2 * i + 1], by, c1);
                  c2 = pfmas(qq[2][2 * i + 1], by, c2);
                }
                if (SAFE) {
                  E[0][dx] = pk(lo(a0), hi(c0));
                  E[1][dx] = pk(lo(a1), hi(c1));
                  E[2][dx] = pk(lo(a2), hi(c2));
                } else {     // MUFU reads and writes single registers: regrouping the lanes is free
                  E[0][dx] = pk(ex2(lo(a0)), ex2(hi(c0)));
                  E[1][dx] = pk(ex2(lo(a1)), ex2(hi(c1)));
                  E[2][dx] = pk(ex2(lo(a2)), ex2(hi(c2)));
                }
                continue;
              } else {
                const p2* kr = ks + (combo * KW + dx) * 3;   // tile row 2r+1 (= hA) or 2r+2 (= hB)
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                  const p2 v = kr[i];
                  const float vx = lo(v), vy = hi(v);
                  a0 = pfmas(qq[0][2 * i], vx, a0);
                  a1 = pfmas(qq[1][2 * i], vx, a1);
                  a2 = pfmas(qq[2][2 * i], vx, a2);
                  a0 = pfmas(qq[0][2 * i + 1], vy, a0);
                  a1 = pfmas(qq[1][2 * i + 1], vy, a1);
                  a2 = pfmas(qq[2][2 * i + 1], vy, a2);
                }
              }
              E[0][dx] = SAFE ? a0 : pex2(a0);
              E[1][dx] = SAFE ? a1 : pex2(a1);
              E[2][dx] = SAFE ? a2 : pex2(a2);
            }
            if (combo == 2) {   // all shared-memory reads of the key / query slot are issued: release it
              __syncwarp();
              mbar_arrive_lane0(cnt_base + 8 * slot, lane);
            }
#pragma unroll
            for (int sl = 0; sl < 3; ++sl) {
              if (combo == 0) fold_row<0, SAFE>(acc[sl], P[sl], cnd[sl], E[sl][0], E[sl][1], E[sl][2]);
              if (combo == 1) fold_row<1, SAFE>(acc[sl], P[sl], cnd[sl], E[sl][0], E[sl][1], E[sl][2]);
              if (combo == 2) fold_row<2, SAFE>(acc[sl], P[sl], cnd[sl], E[sl][0], E[sl][1], E[sl][2]);
            }
          }
#pragma unroll
          for (int sl = 0; sl < 3; ++sl) end_plane<SAFE>(acc[sl], P[sl], cnd[sl]);
          // the pair in slot (j + 1) % 3 has seen its last tap plane: expected offsets, then reset the slot
#define SMILE_FINISH(S)                                                              \
  {                                                                                  \
    const p2 inv = pk(rcp_approx(lo(acc[S].s)), rcp_approx(hi(acc[S].s)));           \
    w0 = pmul(acc[S].nd, inv);                                                       \
    w1 = pmul(acc[S].ah, inv);                                                       \
    w2 = pmul(acc[S].aw, inv);                                                       \
    reset_acc(acc[S]);                                                               \
  }
          if (j == 0) SMILE_FINISH(1) else if (j == 1) SMILE_FINISH(2) else SMILE_FINISH(0)
#undef SMILE_FINISH
          j = (j == 2) ? 0 : j + 1;
          f_m3 = f_m2;
          f_m2 = f_m1;
          f_m1 = fslot * C::F_STRIDE;
          if (++slot == NS) {
            slot = 0;
            par ^= 1u;
          }
          if (++fslot == NF) fslot = 0;
        }

        // ---- (3) finish the moved samples of part (1)
        if (MOVED && pend) {
          const p2 one = pk(1.f, 1.f);
          const p2 gx = psub(one, mfx), gy = psub(one, mfy), gz = psub(one, mfz);
          const p2 r00 = pfma(mv[1], mfx, pmul(mv[0], gx)), r01 = pfma(mv[3], mfx, pmul(mv[2], gx));
          const p2 r10 = pfma(mv[5], mfx, pmul(mv[4], gx)), r11 = pfma(mv[7], mfx, pmul(mv[6], gx));
          const p2 s0 = pfma(r01, mfy, pmul(r00, gy)), s1 = pfma(r11, mfy, pmul(r10, gy));
          const p2 t = pfma(s1, mfz, pmul(s0, gz));
          if (validA) moved[vm] = lo(t);
          if (validB) moved[vm + (unsigned)W] = hi(t);
        }
        vo += (unsigned)HW;
        vm += (unsigned)HW;
        vf += 1.0f;
      }
    }
  }
}


template <int TH, int NS, bool COMPOSE, bool MOVED, int VAR>
__global__ void __launch_bounds__(TH * 32, 1)
fused_march2_kernel(const __grid_constant__ CUtensorMap tm_k, const __grid_constant__ CUtensorMap tm_q,
                    const __grid_constant__ CUtensorMap tm_f, const __grid_constant__ UnitTable tab,
                    const float* __restrict__ rpb,
                    const float* __restrict__ ln_gamma, const float* __restrict__ ln_beta,
                    const float* __restrict__ flow_in, const float* __restrict__ moving, float* __restrict__ out0,
                    float* __restrict__ moved, const Dims2 dm, float scale, float post, int Cmov, int force_safe, int four) {
  using C = Cfg2<TH, NS>;
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  Seg* segs = reinterpret_cast<Seg*>(smem + C::OFF_SEG);
  int* s_nseg = reinterpret_cast<int*>(smem + C::OFF_SEG + (MAXSEG + 1) * sizeof(Seg));
  float2* s_bias = reinterpret_cast<float2*>(smem + C::OFF_BIAS);
  int* s_fast = reinterpret_cast<int*>(smem + C::OFF_BIAS + 81 * 8);
  const int tid = threadIdx.x;
  const int D = dm.D;

  if (tid == 0) {
    // CTA (g, m): group g owns a contiguous range of (row band, plane) units, member m takes column m of every band in
    // it.  The ncol_w members of a group therefore read W-adjacent boxes of the same planes at the same time, and the
    // 128-byte lines their halos share (a 160-byte flow row spans three) come from DRAM once and from L2 afterwards
    // (measured: 492 -> see profiles/r03*_fused_v2_full.txt MB read per launch).
    const int g = blockIdx.x / dm.group, m = blockIdx.x - g * dm.group;
    const long long u_begin = tab.u[g], u_end = tab.u[g + 1];
    int ns = 0, stage = 0;
    long long u = u_begin;
    while (u < u_end && ns < MAXSEG) {
      const long long band = u / D;
      Seg sg;
      sg.d_a = (int)(u - band * D);
      sg.L = (int)((u_end - u) < (long long)(D - sg.d_a) ? (u_end - u) : (long long)(D - sg.d_a));
      sg.w0 = m * TW;
      sg.h0 = (int)(band % dm.ncol_h) * C::ROWS;
      sg.b = (int)(band / dm.ncol_h);
      sg.s_begin = stage;
      segs[ns++] = sg;
      stage += sg.L + 2;
      u += sg.L;
    }
    Seg sentinel = {0, 0, 0, 0, 0, stage};
    segs[ns] = sentinel;
    *s_nseg = ns;
    for (int i = 0; i < NS; ++i) {
      mbar_init(sbase + C::OFF_BAR + 8 * i, 1);
      mbar_init(sbase + C::OFF_CNT + 8 * i, TH);
    }
    *reinterpret_cast<volatile uint32_t*>(smem + C::OFF_NEXT) = NS;
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // rigorous logit bound from the LayerNorm affine parameters (|LN(x)|_2 <= sqrt(C) exactly):
    // fast path if small
    int fast = 0;
    if (ln_gamma != nullptr && ln_beta != nullptr && !force_safe) {
      float gmax = 0.f, b2 = 0.f, rmax = 0.f;
      // NaN-propagating maxima (fmaxf would drop a NaN parameter; it must reach the comparison below)
      for (int c = 0; c < HD; ++c) {
        const float g = fabsf(ln_gamma[c]);
        gmax = (g > gmax || g != g) ? g : gmax;
        b2 = fmaf(ln_beta[c], ln_beta[c], b2);
      }
      if (rpb != nullptr)
        for (int t = 0; t < 27; ++t) {
          const float v = fabsf(rpb[t]);
          rmax = (v > rmax || v != v) ? v : rmax;
        }
      const float qn = gmax * sqrtf((float)HD) + sqrtf(b2);
      const float bound = (fabsf(scale) * qn * qn + rmax) * kLog2e;
      fast = (bound <= kFastBoundLog2) ? 1 : 0;   // NaN parameters compare false -> safe path
    }
    *s_fast = fast;
  }
  if (tid < 81) {
    // entry (phase j, slot, combo, dx): biases of voxel A and voxel B for that key row; the slot's tap plane at phase j
    // is (j - slot) mod 3 (slot j holds the newest pair)
    const int j = tid / 27, sl = (tid / 9) % 3, combo = (tid / 3) % 3, dx = tid % 3;
    const int role = (j - sl + 3) % 3;
    const int dyA = (combo == 0) ? 0 : combo, dyB = (combo == 0) ? 2 : combo - 1;   // row index dy + 1
    const float bA = rpb != nullptr ? rpb[role * 9 + dyA * 3 + dx] * kLog2e : 0.f;
    const float bB = rpb != nullptr ? rpb[role * 9 + dyB * 3 + dx] * kLog2e : 0.f;
    s_bias[tid] = make_float2(bA, bB);
  }
  __syncthreads();
  const int nseg = *s_nseg;
  const int total_stages = segs[nseg].s_begin;
  const float qscale = scale * kLog2e;
  if (*s_fast)
    march<TH, NS, COMPOSE, MOVED, false, VAR>(smem, &tm_k, &tm_q, &tm_f, flow_in, moving, out0, moved, dm, qscale, post, Cmov, nseg,
                                              total_stages, four);
  else
    march<TH, NS, COMPOSE, MOVED, true, VAR>(smem, &tm_k, &tm_q, &tm_f, flow_in, moving, out0, moved, dm, qscale, post, Cmov, nseg,
                                             total_stages, four);
}

// ---------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------
PFN_cuTensorMapEncodeTiled get_encode2() {
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

// L2 promotion of the halo boxes: their rows start 16 B (flow) / 48 B (keys) before a 128-byte boundary, so with 128-byte
// promotion a 160-byte flow row pulls three 128-byte lines (measured: 492 MB read for 315 MB algorithmic).  SMILE_TMA_PROMO
// = 0 none / 1 64 B / 2 128 B / 3 256 B overrides the default for experiments.
bool encode4b(CUtensorMap* map, const void* base, const cuuint64_t (&dims)[4], const cuuint64_t (&strides)[3],
              const cuuint32_t (&box)[4], CUtensorMapL2promotion promo) {
  static const int knob = [] { const char* e = getenv("SMILE_TMA_PROMO"); return e ? atoi(e) : -1; }();
  if (knob >= 0 && knob <= 3) promo = (CUtensorMapL2promotion)knob;
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult rc = get_encode2()(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<void*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("modet_fused(TMA v2): cuTensorMapEncodeTiled failed with CUresult %d", (int)rc);
    return false;
  }
  return true;
}

// Step-balanced partition of the (row band, plane) units over at most `maxg` CTA groups: every group gets a budget of T
// steps, a segment of L planes costs L + 2.  Returns the number of groups used (maxg + 1: does not fit).
int partition_units(long long total, int D, int T, int maxg, int* u) {
  long long pos = 0;
  int g = 0;
  u[0] = 0;
  while (pos < total) {
    if (g == maxg) return maxg + 1;
    int budget = T, nseg = 0;
    while (pos < total && budget >= 3 && nseg < MAXSEG - 1) {
      const long long band_end = (pos / D + 1) * D;
      long long L = band_end - pos;
      if (L > budget - 2) L = budget - 2;
      if (L > total - pos) L = total - pos;
      pos += L;
      budget -= (int)L + 2;
      ++nseg;
    }
    u[++g] = (int)pos;
  }
  return g;
}

template <int TH, int NS, bool COMPOSE, bool MOVED, int VAR>
int launch_cfg2(const CUtensorMap& mk, const CUtensorMap& mq, const CUtensorMap& mf, const UnitTable& tab, const float* rpb,
                const float* g, const float* bta, const float* flow_in, const float* moving, float* out0, float* moved,
                const Dims2& dm, int grid, float scale, float post, int Cmov, int force_safe, cudaStream_t st) {
  using C = Cfg2<TH, NS>;
  auto kern = fused_march2_kernel<TH, NS, COMPOSE, MOVED, VAR>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) {
    set_error("modet_fused(TMA v2): cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  kern<<<grid, C::THREADS, C::SMEM, st>>>(mk, mq, mf, tab, rpb, g, bta, flow_in, moving, out0, moved, dm, scale, post, Cmov,
                                          force_safe, 4);
  return check_launch("modet_fused(TMA v2)");
}

template <int TH, int NS, int VAR>
int launch_tiles2(const float* q, const float* k, const float* rpb, const float* g, const float* bta, const float* flow_in,
                  const float* moving, float* w_out, float* flow_out, float* moved, int B, int D, int H, int W, float scale,
                  float post, int Cmov, int force_safe, cudaStream_t st, bool* handled) {
  using C = Cfg2<TH, NS>;
  const bool compose = flow_in != nullptr;
  Dims2 dm;
  dm.B = B; dm.D = D; dm.H = H; dm.W = W;
  dm.ncol_h = ceil_div(H, C::ROWS);
  dm.ncol_w = ceil_div(W, TW);
  dm.group = dm.ncol_w;
  dm.total_units = (long long)B * dm.ncol_h * D;
  if (dm.total_units >= (1LL << 31)) return SMILE_OK;   // not handled: generic kernel
  // One CTA per SM and all groups co-resident when the volume allows it; wider / longer volumes take more groups (more
  // than one wave, still correct) up to the size of the table.
  int groups = kNumSMs / dm.group;
  if (groups < 1) groups = 1;
  UnitTable tab;
  const int t_cap = (MAXSEG - 2) * D;
  int T = (int)ceil_div_ll(dm.total_units, (long long)groups) + 2, used = 0;
  if (T < 3) T = 3;
  for (;; ++T) {
    if (T > t_cap) {
      T = t_cap;
      groups = kNumSMs;
    }
    used = partition_units(dm.total_units, D, T, groups, tab.u);
    if (used <= groups) break;
    if (T == t_cap) return SMILE_OK;                     // does not fit the table: generic kernel
  }
  for (int i = used + 1; i <= kNumSMs; ++i) tab.u[i] = tab.u[used];
  const int grid = used * dm.group;
  *handled = true;
  dm.dm1 = (float)(D - 1); dm.hm1 = (float)(H - 1); dm.wm1 = (float)(W - 1);
  dm.rd = 1.0f / dm.dm1; dm.rh = 1.0f / dm.hm1; dm.rw = 1.0f / dm.wm1;

  CUtensorMap mk, mq, mf;
  const cuuint64_t qk_dims[4] = {(cuuint64_t)W * HD, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
  const cuuint64_t qk_str[3] = {(cuuint64_t)W * HD * 4, (cuuint64_t)H * W * HD * 4, (cuuint64_t)D * H * W * HD * 4};
  const cuuint32_t k_box[4] = {KW * HD, (cuuint32_t)C::KROWS, 1, 1};
  const cuuint32_t q_box[4] = {TW * HD, (cuuint32_t)C::ROWS, 1, 1};
  if (!encode4b(&mk, k, qk_dims, qk_str, k_box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B) ||
      !encode4b(&mq, q, qk_dims, qk_str, q_box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B))
    return SMILE_ERR_CUDA;
  if (compose) {
    const cuuint64_t f_dims[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B * 3};
    const cuuint64_t f_str[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)D * H * W * 4};
    const cuuint32_t f_box[4] = {FWP, (cuuint32_t)C::KROWS, 1, 3};
    if (!encode4b(&mf, flow_in, f_dims, f_str, f_box, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return SMILE_ERR_CUDA;
  } else {
    mf = mq;
  }
  if (!compose)
    return launch_cfg2<TH, NS, false, false, VAR>(mk, mq, mf, tab, rpb, g, bta, nullptr, nullptr, w_out, nullptr, dm, grid, scale,
                                                  1.0f, 0, force_safe, st);
  if (moved != nullptr)
    return launch_cfg2<TH, NS, true, true, VAR>(mk, mq, mf, tab, rpb, g, bta, flow_in, moving, flow_out, moved, dm, grid, scale,
                                                post, Cmov, force_safe, st);
  return launch_cfg2<TH, NS, true, false, VAR>(mk, mq, mf, tab, rpb, g, bta, flow_in, nullptr, flow_out, nullptr, dm, grid, scale,
                                               post, 0, force_safe, st);
}

}  // namespace

// Second-generation fused kernel (two voxels per thread).  Same contract as launch_modet_attn_tma; `ln_gamma` /
// `ln_beta` (device pointers to the HD LayerNorm affine parameters that produced BOTH q and k, or null) enable the
// maximum-free softmax when the logit bound they imply is small.
int launch_modet_attn_tma2(const float* q, const float* k, const float* rpb, const float* ln_gamma, const float* ln_beta,
                           const float* flow_in, const float* moving, float* w_out, float* flow_out, float* moved, int B,
                           int D, int H, int W, float scale, float post, int Cmov, cudaStream_t st, bool* handled) {
  *handled = false;
  if (W % 4 != 0 || D < 2 || H < 2 || W < 2) return SMILE_OK;
  if (moved != nullptr && Cmov != 1) return SMILE_OK;
  if ((long long)D * H * W * HD >= (1LL << 31)) return SMILE_OK;
  if ((long long)B * 3 * D * H * W >= (1LL << 32)) return SMILE_OK;   // 32-bit element offsets inside the kernel
  if (get_encode2() == nullptr) return SMILE_OK;
  // SMILE_FUSED_V2: bit 0 = online-maximum softmax even when the bound allows the fast loop (profiling / test knob);
  // bit 2 = the round-3 border handling (whole-warp slow path at the faces of the volume) for A/B runs
  static const int variant = [] { const char* e = getenv("SMILE_FUSED_V2"); return e ? atoi(e) : 0; }();
  const int force_safe = (variant & 1);
  const int var = (variant & 4) ? 0 : 4;
#define SMILE_GO(V)                                                                                                       \
  return launch_tiles2<12, 3, V>(q, k, rpb, ln_gamma, ln_beta, flow_in, moving, w_out, flow_out, moved, B, D, H, W, scale, \
                                 post, Cmov, force_safe, st, handled)
  if (var == 4) SMILE_GO(4);
  SMILE_GO(0);
#undef SMILE_GO
}

}  // namespace smile
