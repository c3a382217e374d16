// Micro-benchmarks that calibrate the latency floor of the persistent engine on this GPU:
// grid-barrier variants, L2 round trip, launch overhead.   nvcc -arch=sm_100a -O3 -rdc=false
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) { unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) { unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned atom_add_acq_rel(unsigned* p, unsigned v) { unsigned r; asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory"); return r; }
__device__ __forceinline__ unsigned atom_add_relaxed(unsigned* p, unsigned v) { unsigned r; asm volatile("atom.add.relaxed.gpu.global.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory"); return r; }
__device__ __forceinline__ void st_release(unsigned* p, unsigned v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed(unsigned* p, unsigned v) { asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ void red_add_release(unsigned* p, unsigned v) { asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

struct Bar { unsigned count; unsigned gen; unsigned pad[30]; unsigned flags[256 * 32]; };

// variant 0: engine barrier (atom acq_rel + ld.acquire poll)
__device__ __forceinline__ void bar_v0(Bar* b, unsigned n, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned target = gen + 1;
    if (atom_add_acq_rel(&b->count, 1) == n - 1) { st_relaxed(&b->count, 0); st_release(&b->gen, target); }
    else while (ld_acquire_gpu(&b->gen) != target) {}
  }
  __syncthreads(); gen++;
}
// variant 1: monotonically increasing counter, no reset, no return value needed: red.release + poll on count
__device__ __forceinline__ void bar_v1(Bar* b, unsigned n, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned target = (gen + 1) * n;
    red_add_release(&b->count, 1);
    while ((int)(ld_acquire_gpu(&b->count) - target) < 0) {}
  }
  __syncthreads(); gen++;
}
// variant 2: like v1 but relaxed polling + one fence at the end
__device__ __forceinline__ void bar_v2(Bar* b, unsigned n, unsigned& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned target = (gen + 1) * n;
    __threadfence();
    atom_add_relaxed(&b->count, 1);
    while ((int)(ld_relaxed_gpu(&b->count) - target) < 0) {}
    __threadfence();
  }
  __syncthreads(); gen++;
}
// variant 3: per-CTA flags, master CTA gathers with 148 threads then releases
__device__ __forceinline__ void bar_v3(Bar* b, unsigned n, unsigned& gen) {
  __syncthreads();
  unsigned target = gen + 1;
  if (blockIdx.x == 0) {
    if (threadIdx.x > 0 && threadIdx.x < n) while (ld_acquire_gpu(&b->flags[threadIdx.x * 32]) != target) {}
    __syncthreads();
    if (threadIdx.x == 0) st_release(&b->gen, target);
  } else if (threadIdx.x == 0) {
    st_release(&b->flags[blockIdx.x * 32], target);
    while (ld_acquire_gpu(&b->gen) != target) {}
  }
  __syncthreads(); gen++;
}

// variant 5/6: NPOLL warps poll the counter (their requests are naturally staggered, so the completion is seen after a
// fraction of a round trip on average), whoever sees it first releases the CTA through a shared-memory flag; no closing
// __syncthreads
template <int NPOLL>
__device__ __forceinline__ void bar_multi(Bar* b, unsigned n, unsigned& gen) {
  __shared__ volatile unsigned s_flag;
  __syncthreads();
  const unsigned target = (gen + 1) * n, want = gen + 1;
  if (threadIdx.x == 0) red_add_release(&b->count, 1);
  if ((threadIdx.x >> 5) < NPOLL) {
    while (s_flag != want) {
      if ((int)(ld_acquire_gpu(&b->count) - target) >= 0) { s_flag = want; break; }
    }
  } else {
    while (s_flag != want) {}
  }
  gen++;
}

template <int V>
__global__ void __launch_bounds__(256, 1) k_bar(Bar* b, int iters, unsigned gen0, float* sink, int work) {
  unsigned gen = gen0;
  float acc = 0.f;
  for (int i = 0; i < iters; ++i) {
    if (work) { sink[blockIdx.x * 256 + threadIdx.x] = acc + i; }   // one store per thread before the barrier
    if (V == 0) bar_v0(b, gridDim.x, gen);
    else if (V == 1) bar_v1(b, gridDim.x, gen);
    else if (V == 2) bar_v2(b, gridDim.x, gen);
    else if (V == 3) bar_v3(b, gridDim.x, gen);
    else if (V == 5) bar_multi<2>(b, gridDim.x, gen);
    else if (V == 6) bar_multi<4>(b, gridDim.x, gen);
    else if (V == 7) bar_multi<1>(b, gridDim.x, gen);
    else cg::this_grid().sync();
  }
  if (acc < 0) sink[0] = acc;
}

__global__ void k_chase(const unsigned* next, int iters, unsigned* out) {   // dependent L2 loads
  unsigned p = threadIdx.x + blockIdx.x * 1024;
  for (int i = 0; i < iters; ++i) p = __ldcg(next + p);
  if (p == 0xffffffffu) out[0] = p;
}

template <int V>
float run_bar(Bar* b, int grid, int iters, float* sink, int work, unsigned& gen) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  void* args[] = {&b, &iters, &gen, &sink, &work};
  cudaMemset(b, 0, sizeof(Bar)); gen = 0;
  for (int rep = 0; rep < 2; ++rep) {
    cudaMemset(b, 0, sizeof(Bar)); gen = 0;
    cudaEventRecord(e0);
    cudaLaunchCooperativeKernel((void*)k_bar<V>, dim3(grid), dim3(256), args, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
  }
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error: %s\n", cudaGetErrorString(e));
  return ms * 1e6f / iters;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int grid = p.multiProcessorCount;
  Bar* b; cudaMalloc(&b, sizeof(Bar));
  float* sink; cudaMalloc(&sink, 256 * 1024 * 4);
  unsigned gen;
  const int iters = 20000;
  printf("device %s, %d SMs, %d iters\n", p.name, grid, iters);
  for (int work = 0; work < 2; ++work) {
    printf("work=%d  v0(atom acq_rel + ld.acquire): %.0f ns/barrier\n", work, run_bar<0>(b, grid, iters, sink, work, gen));
    printf("work=%d  v1(red.release + ld.acquire count): %.0f ns\n", work, run_bar<1>(b, grid, iters, sink, work, gen));
    printf("work=%d  v2(fence + relaxed atom/poll + fence): %.0f ns\n", work, run_bar<2>(b, grid, iters, sink, work, gen));
    printf("work=%d  v3(per-CTA flags + master): %.0f ns\n", work, run_bar<3>(b, grid, iters, sink, work, gen));
    printf("work=%d  v4(cooperative_groups grid.sync): %.0f ns\n", work, run_bar<4>(b, grid, iters, sink, work, gen));
    printf("work=%d  v7(1 polling warp + shared flag, no closing sync): %.0f ns\n", work, run_bar<7>(b, grid, iters, sink, work, gen));
    printf("work=%d  v5(2 polling warps + shared flag): %.0f ns\n", work, run_bar<5>(b, grid, iters, sink, work, gen));
    printf("work=%d  v6(4 polling warps + shared flag): %.0f ns\n", work, run_bar<6>(b, grid, iters, sink, work, gen));
  }
  for (int g : {8, 32, 74}) printf("grid=%d v1: %.0f ns\n", g, run_bar<1>(b, g, iters, sink, 1, gen));
  // L2 pointer chase (working set 4 MB, resident in L2)
  {
    const int N = 1 << 20;
    unsigned* h = new unsigned[N];
    for (int i = 0; i < N; ++i) h[i] = (unsigned)(((long long)i * 40503 + 12345) % N);
    unsigned *d, *out; cudaMalloc(&d, N * 4); cudaMalloc(&out, 4);
    cudaMemcpy(d, h, N * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_chase<<<1, 32>>>(d, 20000, out); cudaEventRecord(e1); cudaEventSynchronize(e1); }
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("L2 dependent-load round trip (ld.global.cg, 1 warp): %.0f ns\n", ms * 1e6f / 20000);
  }
  return 0;
}
