// Fused heads==1 ModeT level, TMA-staged (the headline kernel).
//
//   w        = ModeTransformer(q, k)                       reference ModeT/models.py:308-334
//   flow_out = post * (SpatialTransformer(flow_in, w) + w)  models.py:403 / 408 (49-67)
//   moved    = SpatialTransformer(moving, flow_out)         models.py:410
//
// A CTA of TH warps marches a column of TH rows x 32 voxels along D.  Per plane three TMA boxes land in an mbarrier
// ring: the key plane with its 1-voxel halo (out-of-volume elements are zero-filled by TMA == the zero padding of
// models.py:319), the query plane and the three flow_in planes with halo.  There is no producer warp: every warp
// releases a ring slot with an mbarrier arrive, and whichever warp first sees a slot released by all claims the next
// stage (CAS) and issues its TMA loads.  Lanes run along W.  A key plane holds tap plane 2 of the oldest in-flight
// voxel, 1 of the middle one and 0 of the newest: every 24-byte key row is read from shared memory once and its nine
// logits per voxel are folded into a running (online) softmax state, so no logits are kept across planes.  Dot
// products and lerps are packed fma.rn.f32x2.  |w| <= 1, so the compose sample lives in the same 3x3x3 window and is
// gathered from the flow ring (global-memory path when a corner leaves the window, i.e. |w| == 1).
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
// TMA (tiled mode) needs the innermost start coordinate to be a multiple of 16 bytes -- measured on
// B200: an unaligned start raises "illegal instruction" (tools/probe/tma_probe.cu) -- so the halo
// boxes start a little further left than the 1-voxel halo: 2 voxels (48 B) for keys, 4 floats for flow.
constexpr int KW = TW + 4;    // key row: voxels w0-2 .. w0+33
constexpr int KOFF = 1;       // tile column of voxel (w - 1) for lane 0
// Flow row pitch: 40 floats (w0-4 .. w0+35) are needed.  Wider pitches change the bank pattern of the compose gather
// (with 64 the bank depends on the lane only); measured on the B200 at 160x192x160: 40 -> 155.1 us, 48 -> 153.8 us,
// 64 -> 155 us, i.e. no effect beyond noise -- the gather conflicts are not what limits the kernel.  48 is kept.
constexpr int FWP = 48;
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
  static constexpr int OFF_CNT = OFF_BAR + 32;            // NS "slot empty" barriers (one arrival per warp)
  static constexpr int OFF_NEXT = OFF_CNT + 32;          // next stage to be issued (claimed with a CAS)
  static constexpr int OFF_RPB = OFF_NEXT + 16;            // 3 tap planes x 12 floats (9 used; rows 16-byte aligned)
  static constexpr int OFF_SEG = OFF_RPB + 192;           // (MAXSEG + 1) segments
  static constexpr int SMEM = OFF_SEG + (MAXSEG + 1) * (int)sizeof(Seg) + 64;
  static constexpr int THREADS = TH * 32;
};

struct Dims {
  int B, D, H, W;
  int ncol_h, ncol_w;
  long long total_units;  // B * ncol_h * ncol_w * D plane-steps
  int units_per_cta;
  float dm1, hm1, wm1;    // S - 1
  float rd, rh, rw;       // RN(1 / (S - 1))
};

// Running softmax state of one in-flight voxel (log2 domain): maximum so far, sum of exponentials, and the
// exponential-weighted sums that give the expected tap offset (models.py:328-332).  nd holds the sum of the FIRST tap
// plane (offset -1 along depth) until the last plane turns it into  sum(last plane) - sum(first plane).
struct Acc {
  float m, s, nd, ah, aw;
};

// Fold the nine logits of one tap plane (index i = 3*(dy+1) + (dx+1)) into the running state.
// L[] holds <q,k>; logit = qscale * <q,k> + rpb (rpb pre-scaled by log2 e) is formed here.
// ROLE 0: first plane of the voxel (initialises A), 1: middle plane, 2: last plane (A.nd becomes the depth numerator).
template <int ROLE>
__device__ __forceinline__ void fold9(Acc& A, const float (&L)[9], const float* __restrict__ rpb9, float qscale) {
  float l[9];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float4 r = reinterpret_cast<const float4*>(rpb9)[i];
    const float2 v0 = fma2s(make_float2(L[4 * i], L[4 * i + 1]), qscale, make_float2(r.x, r.y));
    const float2 v1 = fma2s(make_float2(L[4 * i + 2], L[4 * i + 3]), qscale, make_float2(r.z, r.w));
    l[4 * i] = v0.x;
    l[4 * i + 1] = v0.y;
    l[4 * i + 2] = v1.x;
    l[4 * i + 3] = v1.y;
  }
  l[8] = fmaf(L[8], qscale, rpb9[8]);
  const float pm = max3(max3(l[0], l[1], l[2]), max3(l[3], l[4], l[5]), max3(l[6], l[7], l[8]));
  float mn = pm, a = 0.f;
  if (ROLE != 0) {
    mn = fmaxf(A.m, pm);
    a = ex2(A.m - mn);
  }
  const float nm = -mn;
  float e[9];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 v = add2s(make_float2(l[2 * i], l[2 * i + 1]), nm);
    e[2 * i] = ex2(v.x);
    e[2 * i + 1] = ex2(v.y);
  }
  e[8] = ex2(l[8] + nm);
  // rows 0 and 1 ride in the two lanes of the packed adds, row 2 is scalar
  const float2 X = make_float2(e[0], e[3]), Y = make_float2(e[2], e[5]), Z = make_float2(e[1], e[4]);
  const float2 R01 = add2(add2(X, Y), Z);
  const float2 D01 = sub2(Y, X);
  const float r2 = (e[6] + e[8]) + e[7], d2 = e[8] - e[6];
  const float S = (R01.x + R01.y) + r2;
  const float SH = r2 - R01.x;
  const float SW = (D01.x + D01.y) + d2;
  if (ROLE == 0) {
    A.s = S;
    A.nd = S;
    A.ah = SH;
    A.aw = SW;
  } else {
    A.s = fmaf(A.s, a, S);
    A.nd = (ROLE == 1) ? A.nd * a : fmaf(-A.nd, a, S);
    A.ah = fmaf(A.ah, a, SH);
    A.aw = fmaf(A.aw, a, SW);
  }
  A.m = mn;
}

// One axis of a zeros-padded trilinear sample from global memory (the moved image): floor index, fraction,
// corner indices clamped into the volume (so both loads are always legal) and in-bounds flags; an out-of-volume
// corner contributes 0 like torch's grid_sampler (GridSampler.h:209-211).
__device__ __forceinline__ void axis_corners(float c, int S, int& i0, int& i1, float& fr, bool& in0, bool& in1) {
  const int j0 = __float2int_rd(c), j1 = j0 + 1;
  fr = (fabsf(c) < 1e9f) ? __fsub_rn(c, (float)j0) : 0.f;
  in0 = (unsigned)j0 < (unsigned)S;
  in1 = (unsigned)j1 < (unsigned)S;
  i0 = min(max(j0, 0), S - 1);
  i1 = min(max(j1, 0), S - 1);
}

template <int TH, int NS, bool COMPOSE>
__device__ __noinline__ int issue_stage(uint32_t sbase, const Seg* __restrict__ segs, int pseg, int n,
                                        const CUtensorMap* tm_k, const CUtensorMap* tm_q, const CUtensorMap* tm_f) {
  using C = Cfg<TH, NS>;
  constexpr int NF = C::NF;
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
  return pseg;
}

// trilinear combination of eight corner values held as (z0, z1) lane pairs: x, then y, then z
__device__ __forceinline__ float tri_combine(float2 y0x0, float2 y0x1, float2 y1x0, float2 y1x1, float fx, float fy,
                                             float fz) {
  const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
  const float2 r0 = fma2s(y0x1, fx, mul2s(y0x0, gx));
  const float2 r1 = fma2s(y1x1, fx, mul2s(y1x0, gx));
  const float2 s = fma2s(r1, fy, mul2s(r0, gy));
  return fmaf(s.y, fz, s.x * gz);
}

// Rare paths, kept out of line so the marching loop stays small in the instruction cache.
struct F3 { float a, b, c; };
// compose sample whose corner left the staged window (|w| == 1 up to rounding): exact global-memory gather
__device__ __noinline__ F3 compose_sample_global(const float* __restrict__ fb, float cz, float cy, float cx, int D, int H,
                                                 int W) {
  TriSample s;
  tri_setup(s, cz, cy, cx, D, H, W);
  const int N = D * H * W;
  F3 r;
  r.a = tri_gather(s, fb);
  r.b = tri_gather(s, fb + N);
  r.c = tri_gather(s, fb + 2 * N);
  return r;
}
struct Corners8 { float v[8]; float fz, fy, fx; };  // v: z0y0x0 z0y0x1 z0y1x0 z0y1x1 z1y0x0 ...
// moved-image corners near the volume border (or of an inactive lane): clamped addresses, zeros outside.
// Inlined: about a fifth of the warps of a 160-wide volume touch a border, and the loads must stay in flight
// across the dot products like those of the interior path.
__device__ __forceinline__ Corners8 moved_corners_border(const float* __restrict__ mb, float mz, float my, float mx, int D,
                                                      int H, int W, bool active) {
  int z0i, z1i, y0i, y1i, x0i, x1i;
  bool zi0, zi1, yi0, yi1, xi0, xi1;
  Corners8 c;
  axis_corners(mz, D, z0i, z1i, c.fz, zi0, zi1);
  axis_corners(my, H, y0i, y1i, c.fy, yi0, yi1);
  axis_corners(mx, W, x0i, x1i, c.fx, xi0, xi1);
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

// Re-arming a ring slot is not tied to a particular warp: whoever notices first that every warp has released the slot
// of stage n - NS claims stage n with a CAS on a shared counter and issues its TMA loads.  The check sits where warps
// have time to spare (top of an iteration, and inside the wait for a stage that is not there yet), so the work lands
// on warps that run ahead and not on the slowest one.  Called by lane 0 only.
template <int TH, int NS, bool COMPOSE>
__device__ __forceinline__ int try_issue(uint32_t sbase, const Seg* __restrict__ segs, int pseg, int total_stages,
                                         const CUtensorMap* tm_k, const CUtensorMap* tm_q, const CUtensorMap* tm_f) {
  using C = Cfg<TH, NS>;
  const uint32_t next_addr = sbase + C::OFF_NEXT;
  const uint32_t n = lds_volatile(next_addr);
  if ((int)n < total_stages) {
    const uint32_t k = n / NS;  // the slot (n % NS) is in its k-th use; its previous use was released in phase k - 1
    if (mbar_test(sbase + C::OFF_CNT + 8 * (n - k * NS), (k - 1) & 1u)) {
      if (atom_cas_relaxed(next_addr, n, n + 1) == n) pseg = issue_stage<TH, NS, COMPOSE>(sbase, segs, pseg, (int)n, tm_k, tm_q, tm_f);
    }
  }
  return pseg;
}

// One iteration of the marching loop handles two things that are independent of each other:
//   (1) the voxel whose 27 logits the PREVIOUS iteration completed: softmax, compose, stores, and the issue of the
//       eight moved-image gathers;
//   (2) the key plane of THIS iteration: wait for its TMA stage, 27 dot products, release the stage;
//   (3) combine the gathers issued in (1) -- their latency is covered by (2), and (1) covers the TMA wait of (2).
template <int TH, int NS, bool COMPOSE, bool MOVED, int MINB, bool DBG = false>
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
      mbar_init(cnt_base + 8 * i, TH);
      if (i == 0) *reinterpret_cast<volatile uint32_t*>(smem + C::OFF_NEXT) = NS;
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 36) s_rpb[tid] = (tid % 12 < 9 && rpb != nullptr) ? rpb[(tid / 12) * 9 + tid % 12] * kLog2e : 0.f;
  __syncthreads();
  const int nseg = *s_nseg;
  const int total_stages = segs[nseg].s_begin;
  int pseg = 0;  // segment cursor of the TMA issue path (per warp; only lane 0 uses it)
  if (tid == 0) {
    for (int n = 0; n < NS && n < total_stages; ++n) pseg = issue_stage<TH, NS, COMPOSE>(sbase, segs, pseg, n, &tm_k, &tm_q, &tm_f);
  }

  int slot = 0, fslot = 0, it = 0;
  uint32_t par = 0;
  long long dbg_wait = 0, dbg_t0 = DBG ? clock64() : 0;
  int dbg_fail = 0;
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
    // linear offset of the voxel handled by part (1) of the current iteration; unsigned so that an address is one
    // IMAD.WIDE.U32 (wraps harmlessly for the not-yet-valid voxels in front of the segment, which are never stored)
    unsigned vo = (unsigned)((sg.d_a - 3) * HW + h * W + wg);
    float vf = (float)(sg.d_a - 3);            // its depth
    int f_m3 = 0, f_m2 = 0, f_m1 = 0;          // flow ring byte offsets of the stages it-3, it-2, it-1

    float2 q[3][3];
    Acc acc[3];
#pragma unroll
    for (int s = 0; s < 3; ++s) {
#pragma unroll
      for (int i = 0; i < 3; ++i) q[s][i] = make_float2(0.f, 0.f);
      acc[s].m = acc[s].s = acc[s].nd = acc[s].ah = acc[s].aw = 0.f;
    }
    float w0 = 0.f, w1 = 0.f, w2 = 0.f;  // attention output of the voxel completed by the previous iteration

    const int nsteps = sg.L + 2;  // key planes of this segment; iteration nsteps only finishes the last voxel
    for (int z0 = 0; z0 <= nsteps; z0 += 3) {
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        const int z = z0 + j;
        if (z > nsteps) break;
        // storage roles of the three in-flight voxels: new (tap plane 0), middle (1), oldest (2: completes)
        const int sN = j, sM = (j + 2) % 3, sO = (j + 1) % 3;
        // every warp looks once per step, here and nowhere else: looking right after releasing a slot puts the work
        // on the last arriver (measured 162 vs 155 us), looking with two warps of eight re-arms too late (161 us)
        if (lane == 0) pseg = try_issue<TH, NS, COMPOSE>(sbase, segs, pseg, total_stages, &tm_k, &tm_q, &tm_f);
        __syncwarp();
        float2 mv[4];            // moved-image corners as (z0, z1) pairs: y0x0, y0x1, y1x0, y1x1
        float mfx = 0.f, mfy = 0.f, mfz = 0.f;
        bool pend = false;

        // ---- (1) voxel completed by the previous iteration
        if (z >= 3) {
          if (!COMPOSE) {
            if (valid) {
              ob[vo] = w0;
              ob[vo + (unsigned)N] = w1;
              ob[vo + 2u * (unsigned)N] = w2;
            }
          } else {
            const float cz = st_coord_fast(vf, w0, dm.dm1, dm.rd);
            const float cy = st_coord_fast(hf, w1, dm.hm1, dm.rh);
            const float cx = st_coord_fast(wf, w2, dm.wm1, dm.rw);
            // floor(c) is idx - 1 or idx when the sample stays in the staged window (exact float compares)
            const bool bz = cz >= vf, by = cy >= hf, bx = cx >= wf;
            const bool inwin = (cz >= vf - 1.0f) && (cz < vf + 1.0f) && (cy >= hf - 1.0f) && (cy < hf + 1.0f) &&
                               (cx >= wf - 1.0f) && (cx < wf + 1.0f);
            float f0, f1, f2;
            {
              // shared-memory addresses are inside the staged window whatever the coordinates are
              const float fz = __fsub_rn(cz, bz ? vf : vf - 1.0f), fy = __fsub_rn(cy, by ? hf : hf - 1.0f),
                          fx = __fsub_rn(cx, bx ? wf : wf - 1.0f);
              const int o = (by ? FWP * 4 : 0) + (bx ? 4 : 0);
              const float* pa = reinterpret_cast<const float*>(f_thr + (bz ? f_m2 : f_m3) + o);
              const float* pb = reinterpret_cast<const float*>(f_thr + (bz ? f_m1 : f_m2) + o);
              const float gx = 1.0f - fx, gy = 1.0f - fy, gz = 1.0f - fz;
              // channels 0 and 1 ride in the two lanes of the packed ops
              const float* pa1 = pa + C::F_PLANE;
              const float* pb1 = pb + C::F_PLANE;
              const float2 a00 = make_float2(pa[0], pa1[0]), a01 = make_float2(pa[1], pa1[1]),
                           a10 = make_float2(pa[FWP], pa1[FWP]), a11 = make_float2(pa[FWP + 1], pa1[FWP + 1]),
                           b00 = make_float2(pb[0], pb1[0]), b01 = make_float2(pb[1], pb1[1]),
                           b10 = make_float2(pb[FWP], pb1[FWP]), b11 = make_float2(pb[FWP + 1], pb1[FWP + 1]);
              const float2 ra0 = fma2s(a01, fx, mul2s(a00, gx)), ra1 = fma2s(a11, fx, mul2s(a10, gx)),
                           rb0 = fma2s(b01, fx, mul2s(b00, gx)), rb1 = fma2s(b11, fx, mul2s(b10, gx));
              const float2 sa = fma2s(ra1, fy, mul2s(ra0, gy)), sb = fma2s(rb1, fy, mul2s(rb0, gy));
              const float2 t01 = fma2s(sb, fz, mul2s(sa, gz));
              f0 = t01.x;
              f1 = t01.y;
              // channel 2: lanes = (plane a, plane b)
              const float* pa2 = pa + 2 * C::F_PLANE;
              const float* pb2 = pb + 2 * C::F_PLANE;
              f2 = tri_combine(make_float2(pa2[0], pb2[0]), make_float2(pa2[1], pb2[1]),
                               make_float2(pa2[FWP], pb2[FWP]), make_float2(pa2[FWP + 1], pb2[FWP + 1]), fx, fy, fz);
            }
            if (!inwin) {
              f0 = f1 = f2 = 0.f;
              if (valid) {
                const F3 g = compose_sample_global(fb, cz, cy, cx, D, H, W);
                f0 = g.a;
                f1 = g.b;
                f2 = g.c;
              }
            }
            f0 = __fmul_rn(post, __fadd_rn(f0, w0));
            f1 = __fmul_rn(post, __fadd_rn(f1, w1));
            f2 = __fmul_rn(post, __fadd_rn(f2, w2));
            if (valid) {
              ob[vo] = f0;
              ob[vo + (unsigned)N] = f1;
              ob[vo + 2u * (unsigned)N] = f2;
            }
            if (MOVED) {
              // moved = T(moving, flow_out): issue the eight gathers now, combine them after the dot products
              const float mz = st_coord_fast(vf, f0, dm.dm1, dm.rd);
              const float my = st_coord_fast(hf, f1, dm.hm1, dm.rh);
              const float mx = st_coord_fast(wf, f2, dm.wm1, dm.rw);
              const int jz = __float2int_rd(mz), jy = __float2int_rd(my), jx = __float2int_rd(mx);
              const bool interior = valid && ((unsigned)jz < (unsigned)(D - 1)) && ((unsigned)jy < (unsigned)(H - 1)) &&
                                    ((unsigned)jx < (unsigned)(W - 1));
              if (__all_sync(0xffffffffu, interior)) {
                mfz = __fsub_rn(mz, (float)jz);
                mfy = __fsub_rn(my, (float)jy);
                mfx = __fsub_rn(mx, (float)jx);
                const unsigned i00 = (unsigned)((jz * H + jy) * W + jx), i01 = i00 + (unsigned)W,
                               i10 = i00 + (unsigned)HW, i11 = i10 + (unsigned)W;
                const float *p00 = mb + i00, *p01 = mb + i01, *p10 = mb + i10, *p11 = mb + i11;
                mv[0] = make_float2(__ldg(p00), __ldg(p10));
                mv[1] = make_float2(__ldg(p00 + 1), __ldg(p10 + 1));
                mv[2] = make_float2(__ldg(p01), __ldg(p11));
                mv[3] = make_float2(__ldg(p01 + 1), __ldg(p11 + 1));
              } else {
                const Corners8 c = moved_corners_border(mb, mz, my, mx, D, H, W, valid);
                mv[0] = make_float2(c.v[0], c.v[4]);
                mv[1] = make_float2(c.v[1], c.v[5]);
                mv[2] = make_float2(c.v[2], c.v[6]);
                mv[3] = make_float2(c.v[3], c.v[7]);
                mfz = c.fz;
                mfy = c.fy;
                mfx = c.fx;
              }
              pend = true;
            }
          }
        }

        // ---- (2) key plane of this iteration
        if (z < nsteps) {
          {
            const long long t0 = DBG ? clock64() : 0;
            if (!mbar_try_wait_hint(bar_full + 8 * slot, par, 200)) {
              do {  // the stage is late: use the time to re-arm slots other warps have released
                if (lane == 0) pseg = try_issue<TH, NS, COMPOSE>(sbase, segs, pseg, total_stages, &tm_k, &tm_q, &tm_f);
                __syncwarp();
              } while (!mbar_try_wait_hint(bar_full + 8 * slot, par, 200));
              if (DBG) ++dbg_fail;
            }
            if (DBG) dbg_wait += clock64() - t0;
          }
          {
            const float2* qs = reinterpret_cast<const float2*>(q_thr + slot * C::Q_STRIDE);
            q[sN][0] = qs[0];
            q[sN][1] = qs[1];
            q[sN][2] = qs[2];
          }
          const float2* ks = reinterpret_cast<const float2*>(k_thr + slot * C::K_STRIDE);
          // tap plane 2 of the oldest voxel (completes it), 1 of the middle one, 0 of the new one
          float LO[9], LM[9], LN[9];
#pragma unroll
          for (int i = 0; i < 9; ++i) {
            const float2* kr = ks + ((i / 3) * KW + (i % 3)) * 3;
            const float2 k0 = kr[0], k1 = kr[1], k2 = kr[2];
            LO[i] = dot6(q[sO], k0, k1, k2);
            LN[i] = dot6(q[sN], k0, k1, k2);
            LM[i] = dot6(q[sM], k0, k1, k2);
          }
          // this warp is done with the key/query slot
          __syncwarp();
          mbar_arrive_lane0(cnt_base + 8 * slot, lane);
          // online softmax: the three folds are independent of each other
          fold9<0>(acc[sN], LN, s_rpb, qscale);
          fold9<1>(acc[sM], LM, s_rpb + 12, qscale);
          fold9<2>(acc[sO], LO, s_rpb + 24, qscale);
          {
            const float inv = rcp_approx(acc[sO].s);
            w0 = acc[sO].nd * inv;
            w1 = acc[sO].ah * inv;
            w2 = acc[sO].aw * inv;
          }
          ++it;
          f_m3 = f_m2;
          f_m2 = f_m1;
          f_m1 = fslot * C::F_STRIDE;
          if (++slot == NS) {
            slot = 0;
            par ^= 1u;
          }
          if (++fslot == NF) fslot = 0;
        }

        // ---- (3) finish the moved sample of part (1)
        if (MOVED && pend) {
          const float t = tri_combine(mv[0], mv[1], mv[2], mv[3], mfx, mfy, mfz);
          if (valid) mvb[vo] = t;
        }
        vo += (unsigned)HW;
        vf += 1.0f;
      }
    }
  }
  if (DBG && lane == 0 && (blockIdx.x % 59 == 0))
    printf("cta %3d warp %d sm %2d: stages %3d total %7lld cyc, barrier wait %7lld cyc (%4.1f%%), first try failed %3d\n",
           (int)blockIdx.x, r, (int)__smid(), total_stages, clock64() - dbg_t0, dbg_wait,
           100.0 * (double)dbg_wait / (double)(clock64() - dbg_t0), dbg_fail);
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

template <int TH, int NS, bool COMPOSE, bool MOVED, int MINB, bool DBG = false>
int launch_cfg(const CUtensorMap& mk, const CUtensorMap& mq, const CUtensorMap& mf, const float* rpb, const float* flow_in,
               const float* moving, float* out0, float* moved, const Dims& dm, int grid, float qscale, float post, int Cmov,
               cudaStream_t st) {
  using C = Cfg<TH, NS>;
  auto kern = fused_march_kernel<TH, NS, COMPOSE, MOVED, MINB, DBG>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
  if (e != cudaSuccess) {
    set_error("modet_fused(TMA): cannot reserve %d B of shared memory: %s", C::SMEM, cudaGetErrorString(e));
    return SMILE_ERR_CUDA;
  }
  kern<<<grid, C::THREADS, C::SMEM, st>>>(mk, mq, mf, rpb, flow_in, moving, out0, moved, dm, qscale, post, Cmov);
  return check_launch("modet_fused(TMA)");
}

}  // namespace

namespace {
template <int TH, int CTAS>
int launch_tiles(const float* q, const float* k, const float* rpb, const float* flow_in, const float* moving, float* w_out,
                 float* flow_out, float* moved, int B, int D, int H, int W, float scale, float post, int Cmov,
                 cudaStream_t st, int variant) {
  static_assert((TH & (TH - 1)) == 0, "the slot arrival logic assumes a power-of-two warp count");
  const bool compose = flow_in != nullptr;
  Dims dm;
  dm.B = B; dm.D = D; dm.H = H; dm.W = W;
  dm.ncol_h = ceil_div(H, TH);
  dm.ncol_w = ceil_div(W, TW);
  dm.total_units = (long long)B * dm.ncol_h * dm.ncol_w * D;
  long long slots = (long long)(variant == 3 ? 3 : CTAS) * kNumSMs;
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
#define SMILE_LAUNCH(NSV, MB)                                                                                                \
  do {                                                                                                                    \
    if (!compose)                                                                                                         \
      return launch_cfg<TH, NSV, false, false, MB>(mk, mq, mf, rpb, nullptr, nullptr, w_out, nullptr, dm, grid, qscale,   \
                                                   1.0f, 0, st);                                                          \
    if (moved != nullptr)                                                                                                 \
      return launch_cfg<TH, NSV, true, true, MB>(mk, mq, mf, rpb, flow_in, moving, flow_out, moved, dm, grid, qscale,     \
                                                 post, Cmov, st);                                                         \
    return launch_cfg<TH, NSV, true, false, MB>(mk, mq, mf, rpb, flow_in, nullptr, flow_out, nullptr, dm, grid, qscale,   \
                                                post, 0, st);                                                             \
  } while (0)
  switch (variant) {  // tuning knob for profiling runs; 0 is the production configuration
    case 1: SMILE_LAUNCH(4, 2);  // 4-deep ring: 159.7 us vs 155 us (measured)
    case 3: SMILE_LAUNCH(3, 3);  // 3 CTAs/SM at 80 registers: spills, 200 us (measured)
    case 9:  // per-warp barrier-wait timing printed from the kernel (profiling aid)
      if (compose && moved != nullptr)
        return launch_cfg<TH, 3, true, true, 2, true>(mk, mq, mf, rpb, flow_in, moving, flow_out, moved, dm, grid, qscale, post,
                                                      Cmov, st);
      SMILE_LAUNCH(3, 2);
    default: SMILE_LAUNCH(3, CTAS);
  }
#undef SMILE_LAUNCH
}
}  // namespace

// Handles head_dim == 6 levels whose rows meet TMA's 16-byte stride rule (W % 4 == 0); everything
// else stays on the generic CUDA kernels of attn.cu (*handled = false).
int launch_modet_attn_tma(const float* q, const float* k, const float* rpb, const float* flow_in, const float* moving,
                          float* w_out, float* flow_out, float* moved, int B, int D, int H, int W, float scale, float post,
                          int Cmov, cudaStream_t st, bool* handled) {
  *handled = false;
  if (W % 4 != 0 || D < 2 || H < 2 || W < 2) return SMILE_OK;
  if (moved != nullptr && Cmov != 1) return SMILE_OK;  // the fused sampler handles the single-channel moving image
  if ((long long)D * H * W * HD >= (1LL << 31)) return SMILE_OK;
  if (get_encode() == nullptr) return SMILE_OK;
  *handled = true;
  static const int variant = [] { const char* e = getenv("SMILE_FUSED_VARIANT"); return e ? atoi(e) : 0; }();
  if (variant == 4)  // 4-warp CTAs, four per SM (four independent rings): 164 us vs 155.7 us (measured), not used
    return launch_tiles<4, 4>(q, k, rpb, flow_in, moving, w_out, flow_out, moved, B, D, H, W, scale, post, Cmov, st, 0);
  return launch_tiles<8, 2>(q, k, rpb, flow_in, moving, w_out, flow_out, moved, B, D, H, W, scale, post, Cmov, st, variant);
}


}  // namespace smile
