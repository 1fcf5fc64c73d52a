// Probe: accuracy of tcgen05 kind::tf32 accumulation in TMEM over a long K (the 27*Cin reduction of a conv3d),
// single-pass TF32 vs the 3-pass split (hi*hi + lo*hi + hi*lo), against fp64 and against a sequential fp32 sum.
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, N = 32, KB = 128;  // K elements staged per pass (32 chunks of 4)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// mode 0: single TF32 pass on the raw fp32 bits; mode 1: 3xTF32
__global__ void __launch_bounds__(128) probe(const float* A, const float* B, float* D, int K, int mode) {
  extern __shared__ __align__(1024) uint8_t smem[];
  float* sAh = reinterpret_cast<float*>(smem);                 // [KB/4][M][4]
  float* sAl = sAh + KB * M;
  float* sBh = sAl + KB * M;                                   // [KB/4][N][4]
  float* sBl = sBh + KB * N;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sBl + KB * N);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
  uint32_t phase = 0;
  for (int k0 = 0; k0 < K; k0 += KB) {
    for (int i = tid; i < KB * M; i += 128) {  // A[m][k] row-major in global
      const int kk = i % KB, m = i / KB;
      const float v = A[(size_t)m * K + k0 + kk];
      const float h = mode ? tf32_hi(v) : v;
      sAh[((kk >> 2) * M + m) * 4 + (kk & 3)] = h;
      sAl[((kk >> 2) * M + m) * 4 + (kk & 3)] = v - h;
    }
    for (int i = tid; i < KB * N; i += 128) {
      const int kk = i % KB, n = i / KB;
      const float v = B[(size_t)n * K + k0 + kk];
      const float h = mode ? tf32_hi(v) : v;
      sBh[((kk >> 2) * N + n) * 4 + (kk & 3)] = h;
      sBl[((kk >> 2) * N + n) * 4 + (kk & 3)] = v - h;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      for (int c = 0; c < KB / 8; ++c) {
        const int npass = mode ? 3 : 1;
        for (int p = 0; p < npass; ++p) {
          const float* a = (p == 1) ? sAl : sAh;
          const float* b = (p == 2) ? sBl : sBh;
          const uint64_t da = make_desc(smem_u32(a) + (2 * c) * M * 16, M * 16, 128);
          const uint64_t db = make_desc(smem_u32(b) + (2 * c) * N * 16, N * 16, 128);
          const uint32_t acc = (k0 > 0 || c > 0 || p > 0) ? 1u : 0u;
          asm volatile(
              "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
              "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
              ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
              : "memory");
        }
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(phase) : "memory");
    }
    phase ^= 1u;
    __syncthreads();
  }
  asm volatile("tcgen05.fence::after_thread_sync;");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t v[8];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main() {
  const int Ks[3] = {216, 1728, 3456};
  const int smem = (2 * KB * M + 2 * KB * N) * 4 + 64;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int K : Ks) {
    const int Kp = (K + KB - 1) / KB * KB;
    std::vector<float> hA((size_t)M * Kp, 0.f), hB((size_t)N * Kp, 0.f);
    srand(7);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) hA[(size_t)m * Kp + k] = (float)rand() / RAND_MAX * 2.f - 1.f;
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) hB[(size_t)n * Kp + k] = ((float)rand() / RAND_MAX * 2.f - 1.f) * 0.1f;
    float *dA, *dB, *dD;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dB, hB.size() * 4); cudaMalloc(&dD, M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 4, cudaMemcpyHostToDevice);
    std::vector<double> ref(M * N);
    std::vector<float> seq(M * N);
    double scale = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
      double s = 0; float f = 0.f;
      for (int k = 0; k < K; ++k) { s += (double)hA[(size_t)m * Kp + k] * hB[(size_t)n * Kp + k]; f = fmaf(hA[(size_t)m * Kp + k], hB[(size_t)n * Kp + k], f); }
      ref[m * N + n] = s; seq[m * N + n] = f; scale = fmax(scale, fabs(s));
    }
    double e_seq = 0;
    for (int i = 0; i < M * N; ++i) e_seq = fmax(e_seq, fabs(seq[i] - ref[i]));
    for (int mode = 0; mode < 2; ++mode) {
      probe<<<1, 128, smem>>>(dA, dB, dD, Kp, mode);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<float> hD(M * N);
      cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
      double err = 0, bias = 0;
      for (int i = 0; i < M * N; ++i) { err = fmax(err, fabs(hD[i] - ref[i])); bias += hD[i] - ref[i]; }
      printf("K=%4d %-6s max|D-fp64| = %.3e (rel to max|D| %.3e)  mean signed err %.3e   [fp32 sequential FMA: %.3e]\n", K,
             mode ? "3xTF32" : "TF32", err, err / scale, bias / (M * N), e_seq);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dD);
  }
  return 0;
}
