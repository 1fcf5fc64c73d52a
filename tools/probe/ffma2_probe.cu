// Peak issue rate of FFMA vs FFMA2 (packed fp32x2) per SM sub-partition on B200.
#include <cuda_runtime.h>
#include <cstdio>
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 d;
  asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\tmov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
      "fma.rn.f32x2 rd, ra, rb, rc;\n\tmov.b64 {%0, %1}, rd;\n\t}"
      : "=f"(d.x), "=f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, i);
  float2 b = make_float2(s, s * 0.5f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, b.y); a[i].y = fmaf(a[i].y, b.x, b.y); }
        else if (MODE == 1) a[i] = fma2(a[i], b, b);
        else a[i] = fma2(make_float2(b.x, b.x), a[i], b);   // broadcast-operand form
      }
  }
  float r = 0; for (int i = 0; i < 8; ++i) r += a[i].x + a[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  for (int mode = 0; mode < 3; ++mode) for (int warps = 4; warps <= 32; warps *= 2) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&] { if (mode == 0) k<0><<<148, warps * 32>>>(d, iters, 1.0001f); else if (mode == 1) k<1><<<148, warps * 32>>>(d, iters, 1.0001f); else k<2><<<148, warps * 32>>>(d, iters, 1.0001f); };
    run(); cudaDeviceSynchronize();
    cudaEventRecord(e0); run(); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma_per_thread = (double)iters * 4 * 8 * 2;   // scalar FMAs
    double tf = fma_per_thread * warps * 32 * 148 * 2 / (ms * 1e-3) / 1e12;
    double inst = (double)iters * 4 * 8 * (mode == 0 ? 2 : 1) * warps * 148;  // warp instructions
    printf("mode %d (%s) warps/SM %2d: %.3f ms  %.1f TFLOP/s  %.3f warp-inst/ns/SM\n", mode, mode == 0 ? "FFMA" : mode == 1 ? "FFMA2" : "FFMA2 bcast", warps, ms, tf, inst / (ms * 1e6) / 148);
  }
  return 0;
}
