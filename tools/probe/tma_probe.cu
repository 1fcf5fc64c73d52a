// Bisecting probe for the TMA path: ./tma_probe <variant>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void probe(const __grid_constant__ CUtensorMap tm, float* out, int n, int c0, int c1, int c2, int c3, int bytes) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) unsigned long long bar;
  uint32_t b = smem_u32(&bar), dst = smem_u32(smem);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 32) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                   ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1) : "memory");
    else
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                   ::"r"(dst), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(b), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
  }
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(b), "r"(0) : "memory");
  } while (!ok);
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = reinterpret_cast<float*>(smem)[i];
}
int main(int argc, char** argv) {
  int variant = argc > 1 ? atoi(argv[1]) : 0;
  void* p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  const int B = 1, D = 8, H = 16, W = 64, C = 6;
  size_t n = (size_t)B * D * H * W * C;
  std::vector<float> h(n); for (size_t i = 0; i < n; ++i) h[i] = (float)i;
  float *d, *o; cudaMalloc(&d, n * 4); cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  CUtensorMap tm; cuuint32_t es[4] = {1, 1, 1, 1}; CUresult rc; int boxn = 0; int rank = 4; int c[4] = {0, 0, 0, 0};
  if (variant == 0) {         // 2D, small box
    rank = 2; cuuint64_t dims[2] = {(cuuint64_t)W * C, (cuuint64_t)B * D * H}; cuuint64_t str[1] = {(cuuint64_t)W * C * 4}; cuuint32_t box[2] = {64, 4};
    rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE); boxn = 64 * 4;
  } else {
    cuuint64_t dims[4] = {(cuuint64_t)W * C, (cuuint64_t)H, (cuuint64_t)D, (cuuint64_t)B};
    cuuint64_t str[3] = {(cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4, (cuuint64_t)D * H * W * C * 4};
    cuuint32_t box[4] = {64, 4, 1, 1};
    if (variant >= 2) { box[0] = 204; box[1] = 10; }
    if (variant >= 3) { c[0] = -6; c[1] = -1; c[2] = -1; }
    if (variant == 4) { box[0] = 192; box[1] = 8; c[0] = 0; c[1] = 0; c[2] = 1; }
    if (variant == 5) { c[0] = -8; c[1] = -1; c[2] = -1; }
    if (variant == 6) { c[0] = 2; c[1] = 0; c[2] = 0; }
    if (variant == 7) { c[0] = 4; c[1] = -1; c[2] = 8; }
    if (variant == 8) { c[0] = 380; c[1] = 12; c[2] = 7; }
    rc = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, str, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    boxn = box[0] * box[1];
  }
  printf("variant %d encode rc=%d boxn=%d\n", variant, (int)rc, boxn);
  cudaMalloc(&o, boxn * 4);
  if (rank == 2) { cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); probe<2><<<1, 64, 65536>>>(tm, o, boxn, c[0], c[1], c[2], c[3], boxn * 4); }
  else { cudaFuncSetAttribute(probe<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536); probe<4><<<1, 64, 65536>>>(tm, o, boxn, c[0], c[1], c[2], c[3], boxn * 4); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("sync: %s\n", cudaGetErrorString(e));
  if (e == cudaSuccess) { std::vector<float> r(boxn); cudaMemcpy(r.data(), o, boxn * 4, cudaMemcpyDeviceToHost); printf("first: %g %g %g ... [boxw] %g last %g\n", r[0], r[1], r[2], r[boxn / (variant>=2&&variant!=4?10:(variant==4?8:4))], r[boxn - 1]); }
  return 0;
}
