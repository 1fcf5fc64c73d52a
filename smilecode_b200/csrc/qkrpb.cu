// a3: C-ABI twins of the reference's only native interface, the `modet` extension
// (ModeT-cu/modet/modet.cpp:4-37, modet_kernel.cu:17-381; called from ModeT-cu/functional.py:5-28):
//
//   modet_fw(query, key, rpb)             -> attn = q . kpad(window) + rpb        (logits, pre-softmax)
//   modet_bw(d_attn, query, key, bias)    -> d_query, d_key (padded shape), d_rpb
//
// Same tensors and layouts as the reference: q [B,heads,H,W,T,hd] (already scaled, ModeT-cu/models.py:304),
// kpad [B,heads,H+2,W+2,T+2,hd] (zero padded by the caller, models.py:305-306), rpb [heads,3,3,3],
// attn / d_attn [B,heads,H,W,T,27] with tap t = (ti*3 + tj)*3 + tk.  Outputs are caller allocated.
//
// B200 notes: all four kernels are HBM/L2-bound gathers.  One thread owns one (b, head, voxel)
// [or one padded key position for dK]: the hd-float rows it touches are contiguous, adjacent lanes
// touch adjacent rows, so every load instruction covers one contiguous segment per warp; the 27
// logits of a voxel stay in registers and leave as one contiguous 108-byte run per thread.  dK is a
// gather over the (<= 27) query positions whose window covers the key, and dRPB is reduced warp ->
// CTA -> one atomicAdd per (head, tap) and CTA -- no atomics on dK, unlike a scatter formulation.
#include "common.cuh"
#include "kernels.h"

namespace smile {

namespace {

constexpr int kThreads = 128;

__global__ void __launch_bounds__(kThreads) qkrpb_fwd_kernel(const float* __restrict__ q, const float* __restrict__ kp,
                                                             const float* __restrict__ rpb, float* __restrict__ attn,
                                                             int heads, int H, int W, int T, int hd, long long total) {
  const int PW = W + 2, PT = T + 2;
  const long long PHWT = (long long)(H + 2) * PW * PT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    // i = ((bh * H + y) * W + x) * T + z
    const int z = (int)(i % T);
    long long r = i / T;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const long long bh = r / H;
    const int head = (int)(bh % heads);
    const float* qr = q + i * hd;
    const float* kb = kp + (bh * PHWT + ((long long)y * PW + x) * PT + z) * hd;  // tap (0,0,0) of the padded volume
    float* out = attn + i * 27;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const float* kr = kb + (((long long)(t / 9) * PW + (t / 3) % 3) * PT + t % 3) * hd;
      float acc = 0.f;
      for (int d = 0; d < hd; ++d) acc = fmaf(__ldg(qr + d), __ldg(kr + d), acc);
      out[t] = acc + (rpb ? __ldg(rpb + head * 27 + t) : 0.f);
    }
  }
}

// head_dim == 6 (the reference configuration): the same computation with the memory system in mind.  A thread's 27
// logits are 108 contiguous bytes, so a plain `out[t] = ...` is 27 store instructions of 32 scattered 4-byte words each
// (measured 684 GB/s, 1.45x the reference's modet_fw).  Here a CTA of 128 consecutive voxels stages its [128][27] logits
// in shared memory (stride 27 words: conflict-free) and writes the 13.8 KB run with coalesced 16-byte stores; key rows are
// read as three 8-byte words per tap (adjacent lanes -> adjacent rows, L1 serves the 27-fold reuse).
__global__ void __launch_bounds__(kThreads) qkrpb_fwd6_kernel(const float* __restrict__ q, const float* __restrict__ kp,
                                                              const float* __restrict__ rpb, float* __restrict__ attn,
                                                              int heads, int H, int W, int T, long long total) {
  __shared__ __align__(16) float s_out[kThreads * 27];
  const int PW = W + 2, PT = T + 2;
  const long long PHWT = (long long)(H + 2) * PW * PT;
  const long long ntiles = ceil_div_ll(total, kThreads);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * kThreads, i = i0 + threadIdx.x;
    if (i < total) {
      const int z = (int)(i % T);
      long long r = i / T;
      const int x = (int)(r % W);
      r /= W;
      const int y = (int)(r % H);
      const long long bh = r / H;
      const int head = (int)(bh % heads);
      const float2* qr = reinterpret_cast<const float2*>(q + i * 6);
      const float2 q0 = __ldg(qr), q1 = __ldg(qr + 1), q2 = __ldg(qr + 2);
      const float* kb = kp + (bh * PHWT + ((long long)y * PW + x) * PT + z) * 6;
      const float* rb = rpb ? rpb + head * 27 : nullptr;
#pragma unroll
      for (int t = 0; t < 27; ++t) {
        const float2* kr = reinterpret_cast<const float2*>(kb + (((long long)(t / 9) * PW + (t / 3) % 3) * PT + t % 3) * 6);
        const float2 k0 = __ldg(kr), k1 = __ldg(kr + 1), k2 = __ldg(kr + 2);
        // same summation order as the generic kernel (and the reference's loop over d): sequential fma from 0
        float acc = 0.f;
        acc = fmaf(q0.x, k0.x, acc);
        acc = fmaf(q0.y, k0.y, acc);
        acc = fmaf(q1.x, k1.x, acc);
        acc = fmaf(q1.y, k1.y, acc);
        acc = fmaf(q2.x, k2.x, acc);
        acc = fmaf(q2.y, k2.y, acc);
        s_out[threadIdx.x * 27 + t] = acc + (rb ? __ldg(rb + t) : 0.f);
      }
    }
    __syncthreads();
    const long long left = total - i0;
    const int nv = (int)(left < kThreads ? left : kThreads);
    float4* dst = reinterpret_cast<float4*>(attn + i0 * 27);       // i0 * 27 * 4 B is a multiple of 16 (i0 = k * 128)
    const int n4 = nv * 27 / 4;
    for (int j = threadIdx.x; j < n4; j += kThreads) dst[j] = reinterpret_cast<const float4*>(s_out)[j];
    for (int j = n4 * 4 + threadIdx.x; j < nv * 27; j += kThreads) attn[i0 * 27 + j] = s_out[j];
    __syncthreads();
  }
}

// d_query for head_dim == 6: d_attn rows staged through shared memory with coalesced 16-byte loads
__global__ void __launch_bounds__(kThreads) qkrpb_dq6_kernel(const float* __restrict__ g, const float* __restrict__ kp,
                                                             float* __restrict__ dq, int H, int W, int T, long long total) {
  __shared__ __align__(16) float s_g[kThreads * 27];
  const int PW = W + 2, PT = T + 2;
  const long long PHWT = (long long)(H + 2) * PW * PT;
  const long long ntiles = ceil_div_ll(total, kThreads);
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long i0 = tile * kThreads, i = i0 + threadIdx.x;
    const long long left = total - i0;
    const int nv = (int)(left < kThreads ? left : kThreads);
    const float4* src = reinterpret_cast<const float4*>(g + i0 * 27);
    const int n4 = nv * 27 / 4;
    for (int j = threadIdx.x; j < n4; j += kThreads) reinterpret_cast<float4*>(s_g)[j] = __ldg(src + j);
    for (int j = n4 * 4 + threadIdx.x; j < nv * 27; j += kThreads) s_g[j] = __ldg(g + i0 * 27 + j);
    __syncthreads();
    if (i < total) {
      const int z = (int)(i % T);
      long long r = i / T;
      const int x = (int)(r % W);
      r /= W;
      const int y = (int)(r % H);
      const long long bh = r / H;
      const float* kb = kp + (bh * PHWT + ((long long)y * PW + x) * PT + z) * 6;
      float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int t = 0; t < 27; ++t) {      // tap order per channel as in the generic kernel: t ascending
        const float gt = s_g[threadIdx.x * 27 + t];
        const float2* kr = reinterpret_cast<const float2*>(kb + (((long long)(t / 9) * PW + (t / 3) % 3) * PT + t % 3) * 6);
        const float2 k0 = __ldg(kr), k1 = __ldg(kr + 1), k2 = __ldg(kr + 2);
        acc[0] = fmaf(gt, k0.x, acc[0]);
        acc[1] = fmaf(gt, k0.y, acc[1]);
        acc[2] = fmaf(gt, k1.x, acc[2]);
        acc[3] = fmaf(gt, k1.y, acc[3]);
        acc[4] = fmaf(gt, k2.x, acc[4]);
        acc[5] = fmaf(gt, k2.y, acc[5]);
      }
      float2* o = reinterpret_cast<float2*>(dq + i * 6);
      o[0] = make_float2(acc[0], acc[1]);
      o[1] = make_float2(acc[2], acc[3]);
      o[2] = make_float2(acc[4], acc[5]);
    }
    __syncthreads();
  }
}

// d_query[b,h,n,d] = sum_t d_attn[b,h,n,t] * kpad[b,h,n+off(t),d]       (modet_kernel.cu:156-207)
__global__ void __launch_bounds__(kThreads) qkrpb_dq_kernel(const float* __restrict__ g, const float* __restrict__ kp,
                                                            float* __restrict__ dq, int H, int W, int T, int hd,
                                                            long long total) {
  const int PW = W + 2, PT = T + 2;
  const long long PHWT = (long long)(H + 2) * PW * PT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int z = (int)(i % T);
    long long r = i / T;
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const long long bh = r / H;
    const float* kb = kp + (bh * PHWT + ((long long)y * PW + x) * PT + z) * hd;
    const float* gr = g + i * 27;
    float gt[27];
#pragma unroll
    for (int t = 0; t < 27; ++t) gt[t] = __ldg(gr + t);
    for (int d = 0; d < hd; ++d) {
      float acc = 0.f;
#pragma unroll
      for (int t = 0; t < 27; ++t)
        acc = fmaf(gt[t], __ldg(kb + (((long long)(t / 9) * PW + (t / 3) % 3) * PT + t % 3) * hd + d), acc);
      dq[i * hd + d] = acc;
    }
  }
}

// d_key[b,h,m,d] over the PADDED key volume (the caller's pad-backward crops it, as autograd does for
// the reference): sum over taps t of d_attn[b,h,m-off(t),t] * q[b,h,m-off(t),d] for query positions
// inside the volume                                                        (modet_kernel.cu:209-267)
__global__ void __launch_bounds__(kThreads) qkrpb_dk_kernel(const float* __restrict__ g, const float* __restrict__ q,
                                                            float* __restrict__ dk, int H, int W, int T, int hd,
                                                            long long total) {
  const int PH = H + 2, PW = W + 2, PT = T + 2;
  const long long HWT = (long long)H * W * T;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int pz = (int)(i % PT);
    long long r = i / PT;
    const int px = (int)(r % PW);
    r /= PW;
    const int py = (int)(r % PH);
    const long long bh = r / PH;
    float* out = dk + i * hd;
    for (int d = 0; d < hd; ++d) out[d] = 0.f;
#pragma unroll
    for (int t = 0; t < 27; ++t) {
      const int y = py - t / 9, x = px - (t / 3) % 3, z = pz - t % 3;  // query position whose tap t lands here
      if (y >= 0 && y < H && x >= 0 && x < W && z >= 0 && z < T) {
        const long long n = bh * HWT + ((long long)y * W + x) * T + z;
        const float gv = __ldg(g + n * 27 + t);
        const float* qr = q + n * hd;
        for (int d = 0; d < hd; ++d) out[d] = fmaf(gv, __ldg(qr + d), out[d]);
      }
    }
  }
}

// d_rpb[h,t] = sum_{b,n} d_attn[b,h,n,t]                                   (modet_kernel.cu:269-317)
__global__ void __launch_bounds__(256) qkrpb_drpb_kernel(const float* __restrict__ g, float* __restrict__ drpb, int B,
                                                         int heads, long long HWT) {
  __shared__ float s_part[8][27];
  const int head = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float acc[27];
#pragma unroll
  for (int t = 0; t < 27; ++t) acc[t] = 0.f;
  for (int b = 0; b < B; ++b) {
    const float* gb = g + ((long long)b * heads + head) * HWT * 27;
    for (long long n = (long long)blockIdx.x * blockDim.x + threadIdx.x; n < HWT; n += (long long)gridDim.x * blockDim.x) {
#pragma unroll
      for (int t = 0; t < 27; ++t) acc[t] += __ldg(gb + n * 27 + t);
    }
  }
#pragma unroll
  for (int t = 0; t < 27; ++t) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[t] += __shfl_xor_sync(0xffffffffu, acc[t], o);
    if (lane == 0) s_part[warp][t] = acc[t];
  }
  __syncthreads();
  if (threadIdx.x < 27) {
    float tot = 0.f;
    for (int w = 0; w < 8; ++w) tot += s_part[w][threadIdx.x];
    atomicAdd(drpb + head * 27 + threadIdx.x, tot);
  }
}

inline int grid_for(long long n, int block, int per_sm) {
  long long g = ceil_div_ll(n, block);
  const long long cap = (long long)kNumSMs * per_sm;
  return (int)(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int launch_qkrpb_fwd(const float* q, const float* kpad, const float* rpb, float* attn, int B, int heads, int H, int W,
                     int T, int hd, cudaStream_t st) {
  const long long total = (long long)B * heads * H * W * T;
  if (hd == 6)
    qkrpb_fwd6_kernel<<<grid_for(total, kThreads, 16), kThreads, 0, st>>>(q, kpad, rpb, attn, heads, H, W, T, total);
  else
    qkrpb_fwd_kernel<<<grid_for(total, kThreads, 32), kThreads, 0, st>>>(q, kpad, rpb, attn, heads, H, W, T, hd, total);
  return check_launch("modet_qkrpb_fwd");
}

int launch_qkrpb_bwd(const float* d_attn, const float* q, const float* kpad, float* dq, float* dk, float* drpb, int B,
                     int heads, int H, int W, int T, int hd, cudaStream_t st) {
  const long long total = (long long)B * heads * H * W * T;
  const long long ptotal = (long long)B * heads * (H + 2) * (W + 2) * (T + 2);
  if (hd == 6)
    qkrpb_dq6_kernel<<<grid_for(total, kThreads, 16), kThreads, 0, st>>>(d_attn, kpad, dq, H, W, T, total);
  else
    qkrpb_dq_kernel<<<grid_for(total, kThreads, 32), kThreads, 0, st>>>(d_attn, kpad, dq, H, W, T, hd, total);
  int rc = check_launch("modet_qkrpb_bwd(dq)");
  if (rc) return rc;
  qkrpb_dk_kernel<<<grid_for(ptotal, kThreads, 32), kThreads, 0, st>>>(d_attn, q, dk, H, W, T, hd, ptotal);
  rc = check_launch("modet_qkrpb_bwd(dk)");
  if (rc) return rc;
  if (drpb != nullptr) {
    cudaError_t e = cudaMemsetAsync(drpb, 0, (size_t)heads * 27 * sizeof(float), st);
    if (e != cudaSuccess) {
      set_error("modet_qkrpb_bwd: memset of d_rpb failed: %s", cudaGetErrorString(e));
      return SMILE_ERR_CUDA;
    }
    const long long HWT = (long long)H * W * T;
    dim3 grid(grid_for(HWT, 256, 4), heads);
    qkrpb_drpb_kernel<<<grid, 256, 0, st>>>(d_attn, drpb, B, heads, HWT);
    rc = check_launch("modet_qkrpb_bwd(drpb)");
  }
  return rc;
}

}  // namespace smile
