// PTX helpers shared by the TMA-staged fused attention kernels (attn_tma.cu, attn_tma2.cu): mbarrier / TMA wrappers,
// packed fp32x2 arithmetic (SASS FFMA2 / FMUL2 / FADD2) and the exact sampling-coordinate replay.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace smile {
namespace {

// ---------------------------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
// arrival by one lane of a converged warp, as a predicated instruction (no divergent branch)
__device__ __forceinline__ void mbar_arrive_lane0(uint32_t bar, int lane) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, 0;\n\t@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t}" ::"r"(bar), "r"(lane)
               : "memory");
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
__device__ __forceinline__ bool mbar_try_wait_hint(uint32_t bar, uint32_t parity, uint32_t ns) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(ns)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ uint32_t lds_volatile(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared::cta.u32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t atom_cas_relaxed(uint32_t addr, uint32_t cmp, uint32_t val) {
  uint32_t old;
  asm volatile("atom.relaxed.cta.shared::cta.cas.b32 %0, [%1], %2, %3;" : "=r"(old) : "r"(addr), "r"(cmp), "r"(val) : "memory");
  return old;
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
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {  // non-blocking probe
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// Packed fp32x2 arithmetic (SASS FFMA2 / FMUL2 / FADD2).  The "s" forms take a scalar that the hardware broadcasts
// to both lanes (operand modifier R.F32), so no duplicated register pair is needed.
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
__device__ __forceinline__ float2 fma2s(float2 a, float b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\tmov.b64 rc, {%5, %6};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b), "f"(c.x), "f"(c.y));
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
__device__ __forceinline__ float2 mul2s(float2 a, float b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
      "mul.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b));
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
__device__ __forceinline__ float2 add2s(float2 a, float b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %4};\n\t"
      "add.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b));
  return d;
}
__device__ __forceinline__ float2 sub2(float2 a, float2 b) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\t"
      "sub.rn.f32x2 rd, ra, rb;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return d;
}
__device__ __forceinline__ unsigned __smid() {
  unsigned v;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(v));
  return v;
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


}  // namespace
}  // namespace smile

// ---------------------------------------------------------------------------------------------
// Packed pairs kept in ONE 64-bit register (attn_tma2.cu).  With the float2 helpers above every call re-packs its
// operands inside its own asm block, and ptxas materialises those packs as MOV / IMAD.MOV whenever the two halves do
// not already sit in an aligned register pair (measured: 173 moves per thread-step in the first two-voxel kernel).
// A p2 value is packed once (pk) and stays packed across uses.
// ---------------------------------------------------------------------------------------------
namespace smile {
namespace {
typedef unsigned long long p2;

__device__ __forceinline__ p2 pk(float x, float y) {
  p2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
__device__ __forceinline__ float lo(p2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
  return x;
}
__device__ __forceinline__ float hi(p2 a) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(a));
  return y;
}
__device__ __forceinline__ p2 pfma(p2 a, p2 b, p2 c) {
  p2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ p2 pfmas(p2 a, float b, p2 c) {   // scalar b broadcast to both lanes (SASS operand form R.F32)
  p2 d;
  asm("{\n\t.reg .b64 rb;\n\tmov.b64 rb, {%2, %2};\n\tfma.rn.f32x2 %0, %1, rb, %3;\n\t}" : "=l"(d) : "l"(a), "f"(b), "l"(c));
  return d;
}
__device__ __forceinline__ p2 pmul(p2 a, p2 b) {
  p2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ p2 pmuls(p2 a, float b) {
  p2 d;
  asm("{\n\t.reg .b64 rb;\n\tmov.b64 rb, {%2, %2};\n\tmul.rn.f32x2 %0, %1, rb;\n\t}" : "=l"(d) : "l"(a), "f"(b));
  return d;
}
__device__ __forceinline__ p2 padd(p2 a, p2 b) {
  p2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ p2 padds(p2 a, float b) {
  p2 d;
  asm("{\n\t.reg .b64 rb;\n\tmov.b64 rb, {%2, %2};\n\tadd.rn.f32x2 %0, %1, rb;\n\t}" : "=l"(d) : "l"(a), "f"(b));
  return d;
}
__device__ __forceinline__ p2 psub(p2 a, p2 b) {
  p2 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ p2 pex2(p2 a) { return pk(ex2(lo(a)), ex2(hi(a))); }
}  // namespace
}  // namespace smile
