// Standalone check + timing of the tcgen05/TMA GEMM tile (ilswiss_b200/csrc/ilsw_tc5.cuh) against a double-precision
// host reference: the three operand-layout combinations the step programs use (forward, backward-data, weight
// gradient + bias gradient), ragged extents, the epilogue variants, and a probe of how the tensor core reads fp32
// words (truncation to TF32 -- the lo-panel split relies on it).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o /tmp/tc5_test tools/tc5_test.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define ILSW_TC5_DEBUG 1
#include "../ilswiss_b200/csrc/ilsw_tc5.cuh"
#include "../ilswiss_b200/csrc/ilsw_tmap.h"

using namespace ilsw;
constexpr int BN = 64;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__global__ void __launch_bounds__(256, 1) tc5_kernel(GemmOp o, int ntiles, int* fail, unsigned long long* stamps, int reps) {
  extern __shared__ unsigned char dyn_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_raw) + 1023) & ~uintptr_t(1023));
  __shared__ tc5::Sync sy;
  const bool stamp = stamps && blockIdx.x == 0 && threadIdx.x == 0;
  if (stamp) stamps[0] = gtime();
  tc5::setup<BN>(sy, smem);
  if (stamp) stamps[1] = gtime();
  tc5::State st{0u, 0u};
  bool good = true;
  for (int rep = 0; rep < reps && good; ++rep) {       // warm instruction cache / steady state: the LAST repetition is stamped
    int nt = 0;
    if (stamp) stamps[7] = gtime();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      if (!tc5::gemm_tile<BN>(o, tile, smem, sy, st)) { if (threadIdx.x == 0) atomicExch(fail, 1); good = false; break; }
      if (stamp && nt < 4) stamps[2 + nt] = gtime();
      ++nt;
    }
  }
  tc5::teardown<BN>(sy);
  if (stamp) stamps[6] = gtime();
}

static float frand() { return (float)rand() / (float)RAND_MAX * 2.f - 1.f; }

struct Case { const char* name; int M, N, K, a_mc, b_nc, aug, act, mask, bias, ksplit; };

static double run_case(const Case& c, bool timing) {
  const int M = c.M, N = c.N, K = c.K;
  // leading dimensions: multiples of 4 floats
  const int lda = c.a_mc ? (M + 3) / 4 * 4 : (K + 3) / 4 * 4;
  const int ldb = c.b_nc ? (N + 3) / 4 * 4 : (K + 3) / 4 * 4;
  const int ldc = strstr(c.name, "unaligned") ? N : (N + 3) / 4 * 4;   // packed rows exercise the scalar store path
  const int S = c.ksplit > 1 ? c.ksplit : 1;
  const size_t nA = (size_t)(c.a_mc ? K : M) * lda, nB = (size_t)(c.b_nc ? K : N) * ldb, nC1 = (size_t)M * ldc + M, nC = nC1 * S;
  std::vector<float> hA(nA), hB(nB), hH(nC), hbias(N), hC(nC, -777.f), hbo(M, -777.f);
  // split-K layout: partial s of C at dC + s * nC1, of bias_out at dC + s * nC1 + M * ldc
  for (auto& v : hA) v = frand();
  for (auto& v : hB) v = frand();
  for (auto& v : hH) v = frand();
  for (auto& v : hbias) v = frand();
  float *dA, *dB, *dC, *dH, *dbias, *dbo; int* dfail;
  CK(cudaMalloc(&dA, nA * 4)); CK(cudaMalloc(&dB, nB * 4)); CK(cudaMalloc(&dC, nC * 4)); CK(cudaMalloc(&dH, nC * 4));
  CK(cudaMalloc(&dbias, N * 4)); CK(cudaMalloc(&dbo, M * 4)); CK(cudaMalloc(&dfail, 4));
  CK(cudaMemcpy(dA, hA.data(), nA * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), nB * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dH, hH.data(), nC * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, hC.data(), nC * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dbo, hbo.data(), M * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dfail, 0, 4));
  CUtensorMap hm[2]; CUtensorMap* dm;
  int r1 = c.a_mc ? make_tmap_2d(&hm[0], dA, lda, M, K, 32, true) : make_tmap_2d(&hm[0], dA, lda, K, M, 128, false);
  int r2 = c.b_nc ? make_tmap_2d(&hm[1], dB, ldb, N, K, 32, true) : make_tmap_2d(&hm[1], dB, ldb, K, N, BN, false);
  if (r1 || r2) { printf("%s: tensor map encode failed (%d, %d)\n", c.name, r1, r2); exit(3); }
  CK(cudaMalloc(&dm, sizeof(hm))); CK(cudaMemcpy(dm, hm, sizeof(hm), cudaMemcpyHostToDevice));
  GemmOp o; memset(&o, 0, sizeof(o));
  o.A = dA; o.lda = lda; o.a_mc = c.a_mc; o.B = dB; o.ldb = ldb; o.b_nc = c.b_nc; o.M = M; o.N = N; o.K = K;
  o.aug_ones = c.aug; o.C = dC; o.ldc = ldc; o.bias_out = c.aug ? (S > 1 ? dC + (size_t)M * ldc : dbo) : nullptr; o.bias = c.bias ? dbias : nullptr;
  o.ksplit = S; o.split_stride = (int)nC1;
  o.H = c.mask ? dH : nullptr; o.ldh = ldc; o.act = c.act; o.mask = c.mask;
  o.tiles_m = (M + 127) / 128; o.tiles_n = (N + BN - 1) / BN; o.tc5 = 1; o.tmapA = dm; o.tmapB = dm + 1;
  const int ntiles = S * o.tiles_m * o.tiles_n;
  const size_t smem = tc5::Geom<BN>::kSmemBytes + 1024;
  CK(cudaFuncSetAttribute(tc5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int grid = ntiles < 148 ? ntiles : 148;
  unsigned long long* dst; CK(cudaMalloc(&dst, 64)); CK(cudaMemset(dst, 0, 64));
  tc5_kernel<<<grid, 256, smem>>>(o, ntiles, dfail, nullptr, 1);
  CK(cudaDeviceSynchronize());
  int fail = 0; CK(cudaMemcpy(&fail, dfail, 4, cudaMemcpyDeviceToHost));
  if (fail) { printf("%s: PIPELINE TIMEOUT\n", c.name); return 1.0; }
  CK(cudaMemcpy(hC.data(), dC, nC * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(hbo.data(), dbo, M * 4, cudaMemcpyDeviceToHost));
  if (S > 1) {     // sum the partial arenas in split order (what the flat Adam job does)
    for (size_t i = 0; i < nC1; ++i) { float a = hC[i]; for (int sp = 1; sp < S; ++sp) a += hC[sp * nC1 + i]; hC[i] = a; }
    for (int m = 0; m < M; ++m) hbo[m] = hC[(size_t)M * ldc + m];
  }
  double worst = 0; int nbad = 0;
  for (int m = 0; m < M; ++m) {
    for (int n = 0; n < N; ++n) {
      double s = 0, sa = 0;
      for (int k = 0; k < K; ++k) {
        const double a = c.a_mc ? hA[(size_t)k * lda + m] : hA[(size_t)m * lda + k];
        const double b = c.b_nc ? hB[(size_t)k * ldb + n] : hB[(size_t)n * ldb + k];
        s += a * b; sa += fabs(a * b);
      }
      double v = s;
      if (c.bias) v += hbias[n];
      if (c.act == ACT_RELU) v = v > 0 ? v : 0;
      if (c.mask == ACT_RELU) v = hH[(size_t)m * ldc + n] > 0 ? v : 0;
      const double err = fabs((double)hC[(size_t)m * ldc + n] - v) / (sa + 1e-30);
      if (err > worst) worst = err;
      if (!(err < 1e-3) && nbad++ < 3) printf("%s: C[%d,%d] = %g expected %g\n", c.name, m, n, hC[(size_t)m * ldc + n], v);
    }
    if (c.aug) {
      double s = 0, sa = 0;
      for (int k = 0; k < K; ++k) { const double a = hA[(size_t)k * lda + m]; s += a; sa += fabs(a); }
      const double err = fabs((double)hbo[m] - s) / (sa + 1e-30);
      if (err > worst) worst = err;
      if (!(err < 1e-3) && nbad++ < 3) printf("%s: bias_out[%d] = %g expected %g\n", c.name, m, hbo[m], s);
    }
  }
  // untouched padding columns stay untouched
  for (int m = 0; m < M; ++m) for (int n = N; n < ldc; ++n) if (S == 1 && hC[(size_t)m * ldc + n] != -777.f) { printf("%s: wrote padding\n", c.name); worst = 1.0; m = M; break; }
  double us = 0;
  if (timing) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) tc5_kernel<<<grid, 256, smem>>>(o, ntiles, dfail, nullptr, 1);
    cudaEventRecord(e0);
    const int reps = 20;
    for (int i = 0; i < reps; ++i) tc5_kernel<<<grid, 256, smem>>>(o, ntiles, dfail, nullptr, 1);
    cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
    tc5_kernel<<<grid, 256, smem>>>(o, ntiles, dfail, dst, 8);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1); us = ms * 1e3 / reps;
  }
  printf("%-28s M=%4d N=%4d K=%4d tiles=%3d  max err / sum|ab| = %.3e%s", c.name, M, N, K, ntiles, worst, worst < 5e-6 ? "  OK" : "  TOO LARGE");
  if (timing) {
    unsigned long long hs[8]; CK(cudaMemcpy(hs, dst, 64, cudaMemcpyDeviceToHost));
    printf("   %.2f us/launch; CTA0 (8th in-kernel repetition): setup %.2f us, tiles", us, (hs[1] - hs[0]) * 1e-3);
    for (int i = 0; i < 4 && hs[2 + i]; ++i) printf(" %.2f", (hs[2 + i] - (i ? hs[1 + i] : hs[7])) * 1e-3);
    printf(" us");
    unsigned long long ht[4][16]; CK(cudaMemcpyFromSymbol(ht, tc5::g_tc5_t, sizeof(ht)));
    const unsigned long long t0 = hs[7];
    const char* names[4] = {"producer", "mma(wait,commit)", "splitter(w2: landed,arrived)", "epilogue(start,acc_full,staged,-,done,end)"};
    for (int r = 0; r < 4; ++r) { printf("\n      %s:", names[r]); for (int i = 0; i < 16; ++i) if (ht[r][i] >= t0) printf(" %.2f", (ht[r][i] - t0) * 1e-3); }
  }
  cudaFree(dst);
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dH); cudaFree(dbias); cudaFree(dbo); cudaFree(dfail); cudaFree(dm);
  return worst;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s, %d SMs, tile 128x%d, %d stages, %d B shared\n", p.name, p.multiProcessorCount, BN, tc5::kStages, tc5::Geom<BN>::kSmemBytes);
  srand(1);
  const Case cases[] = {
      {"fwd small", 128, 64, 32, 0, 0, 0, ACT_NONE, ACT_NONE, 0, 1},
      {"fwd bias+relu", 1024, 256, 393, 0, 0, 0, ACT_RELU, ACT_NONE, 1, 1},
      {"dx relu-mask", 1024, 256, 256, 0, 1, 0, ACT_NONE, ACT_RELU, 0, 1},
      {"dw +bias-grad", 256, 393, 1024, 1, 1, 1, ACT_NONE, ACT_NONE, 0, 1},
      {"dw +bias-grad split-K 3", 256, 393, 1024, 1, 1, 1, ACT_NONE, ACT_NONE, 0, 3},
      {"dw K=4096 (HER) split-K 4", 300, 300, 4096, 1, 1, 1, ACT_NONE, ACT_NONE, 0, 4},
      {"fwd ragged", 300, 300, 100, 0, 0, 0, ACT_RELU, ACT_NONE, 1, 1},
      {"dx ragged", 1000, 300, 300, 0, 1, 0, ACT_NONE, ACT_RELU, 0, 1},
      {"dw ragged", 300, 28, 1000, 1, 1, 1, ACT_NONE, ACT_NONE, 0, 1},
      {"fwd unaligned ldc (N=393)", 256, 393, 256, 0, 0, 0, ACT_NONE, ACT_NONE, 1, 1},
  };
  double worst = 0;
  for (const Case& c : cases) {
    double w = run_case(c, true);
    if (w > worst) worst = w;
  }
  printf("worst %.3e -> %s\n", worst, worst < 5e-6 ? "PASS" : "FAIL");
  return worst < 5e-6 ? 0 : 1;
}
