// Probe for the tcgen05 path the tensor-core conv3d uses (tools/probe: hardware facts, not product code).
// Checks, against a CPU result, on one CTA:
//   * K-major / no-swizzle shared-memory descriptors with a 16-byte row pitch (8 bf16 channels per voxel position),
//   * an arbitrary 16-byte-aligned START address (a convolution tap is a shift of the A rows),
//   * LBO used as "distance to the second 8-wide K chunk" (second channel block OR the next tap),
//   * accumulation over several tcgen05.mma into TMEM, commit -> mbarrier, tcgen05.ld 32x32b epilogue.
//   D[m][n] = sum over (chunk c) sum_{j<8} A[m + shift_c][j] * B[c][n][j]
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int M = 128, PA = 512;  // PA positions staged
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

template <int N>
__global__ void __launch_bounds__(128) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* D, int4 shifts,
                                             int lbo_a_positions) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __nv_bfloat16* sA = reinterpret_cast<__nv_bfloat16*>(smem);                    // [PA][8]
  __nv_bfloat16* sB = reinterpret_cast<__nv_bfloat16*>(smem + PA * 16);          // [4 chunks][N][8]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + PA * 16 + 4 * N * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  for (int i = tid; i < PA * 8; i += 128) sA[i] = A[i];
  for (int i = tid; i < 4 * N * 8; i += 128) sB[i] = B[i];
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(N < 32 ? 32 : N));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = *tmem_slot;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    // MMA 0: chunks 0,1 ; A chunk0 starts at position shifts.x, chunk1 lbo_a_positions further
    // MMA 1: chunks 2,3 ; A chunk0 starts at position shifts.y, chunk1 lbo_a_positions further
    for (int i = 0; i < 2; ++i) {
      const int s0 = i == 0 ? shifts.x : shifts.y;
      const uint64_t da = make_desc(smem_u32(sA) + s0 * 16, lbo_a_positions * 16, 128);
      const uint64_t db = make_desc(smem_u32(sB) + (2 * i) * N * 16, N * 16, 128);
      const uint32_t acc = i;
      asm volatile(
          "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
          "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
          ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc)
          : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                   : "=r"(ok) : "r"(smem_u32(bar)), "r"(0u) : "memory");
    }
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
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(N < 32 ? 32 : N));
}

template <int N>
int run(int sx, int sy, int lbo) {
  std::vector<__nv_bfloat16> hA(PA * 8), hB(4 * N * 8);
  std::vector<float> fA(PA * 8), fB(4 * N * 8);
  srand(1);
  for (size_t i = 0; i < hA.size(); ++i) { float v = (rand() % 17 - 8) / 8.0f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
  for (size_t i = 0; i < hB.size(); ++i) { float v = (rand() % 13 - 6) / 4.0f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
  __nv_bfloat16 *dA, *dB; float* dD;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dD, M * N * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaMemset(dD, 0xff, M * N * 4);
  const int smem = PA * 16 + 4 * N * 16 + 64;
  cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<N><<<1, 128, smem>>>(dA, dB, dD, make_int4(sx, sy, 0, 0), lbo);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("N=%d shifts (%d,%d) lbo %d: CUDA error %s\n", N, sx, sy, lbo, cudaGetErrorString(e)); return 1; }
  std::vector<float> hD(M * N);
  cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double ref = 0;
      for (int i = 0; i < 2; ++i) {
        const int s0 = i == 0 ? sx : sy;
        for (int c = 0; c < 2; ++c)
          for (int j = 0; j < 8; ++j)
            ref += (double)fA[(m + s0 + c * lbo) * 8 + j] * (double)fB[((2 * i + c) * N + n) * 8 + j];
      }
      const double err = fabs(ref - hD[m * N + n]);
      if (err > maxerr) maxerr = err;
    }
  printf("N=%3d shifts (%3d,%3d) lbo %3d positions: max |D - ref| = %.3e  %s\n", N, sx, sy, lbo, maxerr, maxerr < 1e-3 ? "OK" : "MISMATCH");
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
  return maxerr < 1e-3 ? 0 : 1;
}

int main() {
  int bad = 0;
  bad += run<32>(0, 128, 256);   // aligned starts, chunk 1 = a separate plane
  bad += run<32>(3, 77, 1);      // arbitrary 16-byte-aligned starts, chunk 1 = the next position (adjacent tap)
  bad += run<16>(5, 9, 83);      // N = 16, chunk 1 = the row below (tap + Wp)
  bad += run<64>(1, 2, 200);
  bad += run<128>(7, 300, 37);
  printf(bad ? "umma probe: FAILED\n" : "umma probe: all OK\n");
  return bad;
}
