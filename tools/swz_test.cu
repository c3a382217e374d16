// Development aid: prints the shared-memory placement of a TMA box {32 floats, 32 rows} under SWIZZLE_128B and
// SWIZZLE_128B_ATOM_32B (the layouts the mma.sync tile's fragment loads must follow).  nvcc -arch=sm_100a -o swz_test swz_test.cu
#include <cstdio>
#include <vector>
#include "../ilswiss_b200/csrc/ilsw_tmap.h"
using namespace ilsw;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const __grid_constant__ CUtensorMap tm, float* out) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(&bar)), "r"(4096));
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(sm)), "l"(&tm), "r"(smem_u32(&bar)), "r"(0), "r"(0) : "memory");
  }
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) out[i] = reinterpret_cast<float*>(sm)[i];
}
int main() {
  std::vector<float> h(64 * 64);
  for (int i = 0; i < 64 * 64; ++i) h[i] = (float)((i / 64) * 32 + (i % 64));   // value = r*32 + c for c < 32
  float *d, *o; cudaMalloc(&d, h.size() * 4); cudaMalloc(&o, 4096);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  for (int mode = 0; mode < 2; ++mode) {
    CUtensorMap tm;
    if (make_tmap_2d(&tm, d, 64, 64, 64, 32, mode == 1)) { printf("encode failed\n"); return 1; }
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    k<<<1, 128, 8192>>>(tm, o);
    std::vector<float> r(1024);
    cudaMemcpy(r.data(), o, 4096, cudaMemcpyDeviceToHost);
    printf("mode %s: %s\n", mode ? "SWIZZLE_128B_ATOM_32B" : "SWIZZLE_128B", cudaGetErrorString(cudaGetLastError()));
    int bad16 = 0, bad32a = 0, bad32b = 0;
    for (int w = 0; w < 1024; ++w) {
      const int v = (int)r[w], rr = v / 32, c = v % 32;
      if (w != rr * 32 + (((c / 4) ^ (rr & 7)) * 4) + c % 4) ++bad16;
      if (w != rr * 32 + (((c / 8) ^ (rr & 3)) * 8) + c % 8) ++bad32a;
      if (w != rr * 32 + (((c / 8) ^ ((rr >> 1) & 3)) * 8) + c % 8) ++bad32b;
    }
    printf("  mismatches: chunk16^(r&7) %d | atom32^(r&3) %d | atom32^((r>>1)&3) %d\n", bad16, bad32a, bad32b);
    for (int rr = 0; rr < 9; ++rr) { printf("  row %d first words of each 16B chunk:", rr); for (int c = 0; c < 8; ++c) printf(" %4d", (int)r[rr * 32 + c * 4]); printf("\n"); }
  }
  return 0;
}
