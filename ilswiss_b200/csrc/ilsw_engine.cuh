// ilswiss_b200 -- the persistent, cooperative step-engine kernel for sm_100a.
//
// ONE launch executes n_steps full gradient steps (gather -> forwards -> backwards -> Adam ->
// Polyak -> stats) by walking the phase program of ilsw_program.h.  One CTA per SM; phases are
// separated by a grid-wide barrier; all state that persists between steps (parameters, Adam
// moments, replay ring) stays in HBM/L2, activations and gradients live in an L2-resident
// scratch arena and never leave the chip during a train call.
#pragma once
#include <cuda_runtime.h>

#include "ilsw_ops.cuh"
#include "ilsw_rows_fast.cuh"
#include "ilsw_tc5.cuh"

// experiment switches of the tensor-core tile (see DESIGN.md 4.4 for the measurements behind the defaults)
#ifndef ILSW_ADAM_PREFETCH
#define ILSW_ADAM_PREFETCH 0
#endif
#ifndef ILSW_EIN_FIRST
#define ILSW_EIN_FIRST 0
#endif
#ifndef ILSW_TC128_SIMPLE
#define ILSW_TC128_SIMPLE 1
#endif
#ifndef ILSW_VECRAG
#define ILSW_VECRAG 1
#endif
// epilogue inputs (bias / mask source / previous value) of a tile are prefetched with cp.async into a shared scratch
// instead of registers: at the 255-register cap the twelve values were spilled to local memory right after their loads,
// and every spill store waits for its load -- four serialised L2 round trips (~1.3 us) in the prologue of EVERY tile
// (ncu source view, profiles/r2_ncu_engine_sac_hopper_prefix.txt: STL ... stall_long_sb)
// cross-proxy fence in front of the TMA panel loads of the mma.sync tile.  The operands were written with generic stores by
// OTHER CTAs in an earlier phase and published through the grid barrier (release / acquire at gpu scope); the TMA unit reads
// them from L2, where those stores are already performed.  The tcgen05 tile has always issued its loads without the fence
// (parity green on every TD3 / HER case), so it is off by default; -DILSW_TMA_GLOBAL_FENCE=1 restores it.
#ifndef ILSW_TMA_GLOBAL_FENCE
#define ILSW_TMA_GLOBAL_FENCE 0
#endif
#ifndef ILSW_EIN_CPASYNC
#define ILSW_EIN_CPASYNC 1
#endif

namespace ilsw {

struct BarrierState { unsigned count; unsigned gen; unsigned pad[30]; };

// cross-replica exchange state (NVLink peer memory mapped with CUDA IPC), see ilsw_abi.cu
struct Replica {
  int world, rank;
  int n;                      // policy parameter count
  int nstride;                // slot stride (n rounded up to 4 floats: 16-byte aligned slots)
  const float* grad;          // local policy gradient arena
  int g_splits, g_split_stride;   // > 1: the local gradient is the sum of split-K partial arenas (see AdamOp)
  float* recv_local;          // [2][world][n]  (parity, source rank)
  float* recv_peer[8];        // recv_local of every rank (self included)
  unsigned* flags_local;      // [8] sequence numbers written by the peers
  unsigned* flags_peer[8];
  unsigned seq0;              // exchanges completed before this launch
  unsigned long long* cnt_local;      // arrival counter of the epilogue-push exchange: every REPLICA adds 1 per policy update, when the
  unsigned long long* cnt_peer[8];    // last of its CTAs has arrived on `arrive_local` (a system fence behind its pushes); never reset
  unsigned* arrive_local;             // this replica's own CTA arrival counter (local memory, monotonic)
  // NVLS multicast mapping of the same exchange buffer (ilsw_replica_connect_symm): a store / reduction on these addresses is
  // replicated to every replica by the NVSwitch -- the epilogue push costs ONE store per element instead of `world`
  float* recv_mc;                     // nullptr: no multicast mapping, per-peer stores
  unsigned long long* cnt_mc;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_add_release_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("red.release.sys.global.add.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

constexpr long long kSpinLimit = 4000000000LL;  // ~2 s at 2 GHz: a hung peer/CTA aborts the launch

// Sense-reversal grid barrier (all CTAs are co-resident: cooperative launch, 1 CTA / SM).
// Arrival is ONE atom.add.acq_rel.gpu (release of this CTA's phase writes, cumulative over the
// preceding bar.sync), departure ONE ld.acquire.gpu poll loop: two L2 round trips, no membar.sc.
// Returns false if the launch was aborted (timeout); every thread of every CTA then exits.
__device__ __forceinline__ unsigned atom_add_acq_rel_gpu(unsigned* p, unsigned v) {
  unsigned r;
  asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], %2;" : "=r"(r) : "l"(p), "r"(v) : "memory");
  return r;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_gpu(unsigned* p, unsigned v) {
  asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void red_add_release_gpu(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// `gen` counts the barriers passed; bar->count is monotonic (never reset): barrier k is complete
// when count >= (k+1)*nblocks.  Arrival = one fire-and-forget red.release (no return value to wait
// for), departure = ld.acquire polling: measured 1.25 us on 148 SMs vs 1.9 us for the
// atom+generation variant (tools/microbench.cu, profiles/r1_microbench.txt).
__device__ __forceinline__ bool grid_barrier(BarrierState* bar, unsigned nblocks, unsigned& gen, int* abort_flag) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    int ok = 1;
    const unsigned target = (gen + 1u) * nblocks;
    red_add_release_gpu(&bar->count, 1u);
    long long t0 = 0;
    unsigned spins = 0;
#ifndef ILSW_BARRIER_RELAXED
#define ILSW_BARRIER_RELAXED 0
#endif
#if ILSW_BARRIER_RELAXED
    while ((int)(ld_relaxed_gpu(&bar->count) - target) < 0) {
#else
    while ((int)(ld_acquire_gpu(&bar->count) - target) < 0) {
#endif
      if ((++spins & 1023u) == 0u) {
        if (t0 == 0) t0 = clock64();
        long long dt = clock64() - t0;
        if (dt > kSpinLimit) {
          if (*(volatile int*)abort_flag) { ok = 0; break; }
          if (dt > 2 * kSpinLimit) { atomicExch(abort_flag, 1); ok = 0; break; }
        }
      }
    }
    s_ok = ok;
  }
  __syncthreads();
  gen += 1;
  return s_ok != 0;
}

// ------------------------------------------------------------------------------------------
// fp32 SIMT GEMM tile: 32x32 outputs per CTA job, K swept in chunks of 32 through a
// double-buffered shared-memory stage; the 256 threads are 4 K-groups x (8x8 threads x 4x4
// register micro-tiles); partial sums of the K-groups are reduced through shared memory in a
// fixed order (deterministic).  Operands are read with 128-bit ld.global.cg when aligned.
// ------------------------------------------------------------------------------------------
constexpr int kTM = 32, kTN = 32, kTK = 32, kLd = 36;   // smem row stride (floats), 16B aligned
constexpr int kStageFloats = kTK * kLd;                  // one operand stage
constexpr int kGemmSmemFloats = 4 * kStageFloats;        // 2 stages x (A,B) = 18432 B

struct TileLoader {
  // each thread moves 4 elements of A and 4 of B per K chunk
  float ra[4], rb[4];
};

__device__ __forceinline__ void tile_load_A(const GemmOp& o, int m0, int k0, float* r, bool fast, const L0FuseOp* fz) {
  const int tid = threadIdx.x;
  if (fz) {       // fused first layer (GemmOp::a0): produced element by element in the exact mode
    const int m = m0 + (tid >> 3), k = k0 + ((tid & 7) << 2);
#pragma unroll
    for (int i = 0; i < 4; ++i) r[i] = (m < o.M && k + i < o.K) ? gemm_A_fused(*fz, m, k + i) : 0.f;
    return;
  }
  if (!o.a_mc) {  // k contiguous: thread -> (m = tid/8, k = (tid%8)*4 .. +3)
    int m = m0 + (tid >> 3), k = k0 + ((tid & 7) << 2);
    if (m < o.M && fast && k + 3 < o.K) {
      float4 v = __ldcg(reinterpret_cast<const float4*>(o.A + (size_t)m * o.lda + k));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (m < o.M && k + i < o.K) ? __ldcg(o.A + (size_t)m * o.lda + k + i) : 0.f;
    }
  } else {        // m contiguous: thread -> (k = tid/8, m = (tid%8)*4 .. +3)
    int k = k0 + (tid >> 3), m = m0 + ((tid & 7) << 2);
    if (k < o.K && fast && m + 3 < o.M) {
      float4 v = __ldcg(reinterpret_cast<const float4*>(o.A + (size_t)k * o.lda + m));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) r[i] = (k < o.K && m + i < o.M) ? __ldcg(o.A + (size_t)k * o.lda + m + i) : 0.f;
    }
  }
}
__device__ __forceinline__ void tile_store_A(const GemmOp& o, float* As, const float* r) {
  const int tid = threadIdx.x;
  if (!o.a_mc) {
    int m = tid >> 3, k = (tid & 7) << 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) As[(k + i) * kLd + m] = r[i];
  } else {
    int k = tid >> 3, m = (tid & 7) << 2;
    *reinterpret_cast<float4*>(As + k * kLd + m) = make_float4(r[0], r[1], r[2], r[3]);
  }
}
__device__ __forceinline__ void tile_load_B(const GemmOp& o, int n0, int k0, float* r, bool fast) {
  const int tid = threadIdx.x;
  const int Nt = o.N + o.aug_ones;
  if (!o.b_nc) {  // k contiguous: thread -> (n = tid/8, k = (tid%8)*4 .. +3)
    int n = n0 + (tid >> 3), k = k0 + ((tid & 7) << 2);
    if (n < o.N && fast && k + 3 < o.K) {
      float4 v = __ldcg(reinterpret_cast<const float4*>(o.B + (size_t)n * o.ldb + k));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (k + i < o.K) {
          if (n < o.N) v = __ldcg(o.B + (size_t)n * o.ldb + k + i);
          else if (n < Nt) v = 1.0f;
        }
        r[i] = v;
      }
    }
  } else {        // n contiguous: thread -> (k = tid/8, n = (tid%8)*4 .. +3)
    int k = k0 + (tid >> 3), n = n0 + ((tid & 7) << 2);
    if (k < o.K && fast && n + 3 < o.N) {
      float4 v = __ldcg(reinterpret_cast<const float4*>(o.B + (size_t)k * o.ldb + n));
      r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float v = 0.f;
        if (k < o.K) {
          if (n + i < o.N) v = __ldcg(o.B + (size_t)k * o.ldb + n + i);
          else if (n + i < Nt) v = 1.0f;
        }
        r[i] = v;
      }
    }
  }
}
__device__ __forceinline__ void tile_store_B(const GemmOp& o, float* Bs, const float* r) {
  const int tid = threadIdx.x;
  if (!o.b_nc) {
    int n = tid >> 3, k = (tid & 7) << 2;
#pragma unroll
    for (int i = 0; i < 4; ++i) Bs[(k + i) * kLd + n] = r[i];
  } else {
    int k = tid >> 3, n = (tid & 7) << 2;
    *reinterpret_cast<float4*>(Bs + k * kLd + n) = make_float4(r[0], r[1], r[2], r[3]);
  }
}

// fused optimiser step on gradient element gi (value g) -- the tile epilogues' Adam(+Polyak)
__device__ __forceinline__ void adam_fused_elem(const AdamOp& ad, const AdamCoef& cf, int gi, float g) {
  adam_elem_g(ad, cf, gi, g, adam_has_shadow(ad));
}

__device__ __noinline__ void gemm_tile_device(const GemmOp& o, int tile, float* smem, const AdamOp* ad, const AdamCoef* cf, const PushCtx* push, const L0FuseOp* fz) {
  const int tid = threadIdx.x;
  const int tm = tile / o.tiles_n, tn = tile - tm * o.tiles_n;
  const int m0 = tm * kTM, n0 = tn * kTN;
  const int kg = tid >> 6, t = tid & 63, ty = t >> 3, tx = t & 7;
  const int nchunks = (o.K + kTK - 1) / kTK;
  const bool fastA = ((o.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.A) & 15) == 0);
  const bool fastB = ((o.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.B) & 15) == 0);
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float ra[4], rb[4];
  tile_load_A(o, m0, 0, ra, fastA, fz);
  tile_load_B(o, n0, 0, rb, fastB);
  tile_store_A(o, smem, ra);
  tile_store_B(o, smem + kStageFloats, rb);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    float* As = smem + (c & 1) * 2 * kStageFloats;
    float* Bs = As + kStageFloats;
    if (c + 1 < nchunks) {
      tile_load_A(o, m0, (c + 1) * kTK, ra, fastA, fz);
      tile_load_B(o, n0, (c + 1) * kTK, rb, fastB);
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      const int k = kg * 8 + kk;
      float4 a4 = *reinterpret_cast<const float4*>(As + k * kLd + ty * 4);
      float4 b4 = *reinterpret_cast<const float4*>(Bs + k * kLd + tx * 4);
      float av[4] = {a4.x, a4.y, a4.z, a4.w}, bv[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (c + 1 < nchunks) {
      float* An = smem + ((c + 1) & 1) * 2 * kStageFloats;
      tile_store_A(o, An, ra);
      tile_store_B(o, An + kStageFloats, rb);
    }
    __syncthreads();
  }
  // cross K-group reduction through smem: red[(kg*64 + t)*17 + i*4 + j]
  float* red = smem;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) red[(kg * 64 + t) * 17 + i * 4 + j] = acc[i][j];
  __syncthreads();
  const int Nt = o.N;      // the aug column (n == N) is produced below from column sums of A
  const int m = m0 + ty * 4 + kg;
  EpiIn ein[4];
  float vout[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (m < o.M && n < Nt) ein[j] = epi_load(o, m, n);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) v += red[(g * 64 + t) * 17 + kg * 4 + j];
    vout[j] = v;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int n = n0 + tx * 4 + j;
    if (m < o.M && n < Nt) {
      epi_store(o, m, n, vout[j], ein[j], push);
      if (ad) adam_fused_elem(*ad, *cf, gemm_grad_index(o, *ad, m, n), o.accumulate ? vout[j] + ein[j].prev : vout[j]);
    }
  }
  if (o.aug_ones && tn == 0 && tid < 32 && m0 + tid < o.M) {     // bias gradient: bias_out[m] = sum_k A(m,k)
    const int mm = m0 + tid;
    float ssum = 0.f;
    for (int k = 0; k < o.K; ++k) ssum += gemm_A(o, mm, k);
    const EpiIn e = epi_load(o, mm, o.N);
    epi_store(o, mm, o.N, ssum, e, push);
    if (ad) adam_fused_elem(*ad, *cf, gemm_grad_index(o, *ad, mm, o.N), o.accumulate ? ssum + e.prev : ssum);
  }
  if (fz && tn == 0) {         // materialise the fused first layer of this row block (backward pass)
    for (int e = tid; e < kTM * o.K; e += kThreads) {
      const int r = e / o.K, k = e - r * o.K;
      if (m0 + r < o.M) fz->out[(size_t)(m0 + r) * fz->ldo + k] = gemm_A_fused(*fz, m0 + r, k);
    }
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// TF32 tensor-core GEMM tile (speed mode; mma.sync.m16n8k8, fp32 accumulate).
// 32x32 outputs per CTA job; K panels of 128 of both operands are brought from L2 into shared
// memory with 16-byte cp.async.cg (one L2 round trip per 128-deep stage instead of one per
// 32-wide chunk), two stages in flight.  Each of the 8 warps
// owns a 16x8 accumulator fragment over the full K.  Shared layouts are chosen per operand so
// that BOTH the 128-bit fill and the fragment reads are bank-conflict free:
//   k-contiguous operand  -> [32 rows][260]   (fragment bank = 4*row + k  = lane)
//   m/n-contiguous operand -> [256 k][40]     (fragment bank = 8*k  + row = lane-permutation)
// mode 1: single-pass TF32 (operands rounded with cvt.rna);  mode 3: 3xTF32 split
// (hi/lo, small terms first) which recovers fp32-level accuracy on the tensor cores.
// ------------------------------------------------------------------------------------------
// K per stage is a template parameter: 256 (one CTA per SM, 160 KB staging: a K=256 GEMM is a single
// stage) or 128 (two CTAs per SM, 80 KB staging each).
constexpr int kMS = 40;                       // row stride of an m/n-contiguous panel
template <int KC> struct TcGeom {
  static constexpr int kKS = KC + 4;                    // row stride of a k-contiguous panel
  static constexpr int kOperandFloats = KC * kMS;       // >= 32 * kKS
  static constexpr int kStageFloats = 2 * kOperandFloats;
  static constexpr int kSmemFloats = 2 * kStageFloats;  // 2 stages x (A,B): 160 KB (KC=256) / 80 KB (KC=128)
};
// K per stage of the one-CTA/SM variant.  128 (two stages of 40 KB: a K = 256 GEMM has both stages in flight at once) keeps the
// kernel's dynamic shared memory near 105 KB, so the SM's unified L1/shared carve-out leaves 124 KB of L1 instead of 60 KB
// (256: 190 KB of shared memory) -- the engine spills to a 1.8 KB stack frame and every phase starts with a cold L1, so the
// L1 size is worth more than the single-stage K = 256 panel (measured: growing shared memory past the next carve-out step,
// 28 KB of L1, cost 16 us per SAC step).
#ifndef ILSW_KC1
#define ILSW_KC1 128
#endif
constexpr int kKC1 = ILSW_KC1;
constexpr int tc_smem_floats(int ctas) { return ctas == 2 ? TcGeom<128>::kSmemFloats : TcGeom<kKC1>::kSmemFloats; }
// bytes of the GEMM staging area at the start of dynamic shared memory; the tcgen05 engine variant (one CTA per SM)
// also fits the TMA stages of ilsw_tc5.cuh (+ 1 KB to align them to the 1024-byte swizzle atom)
constexpr int kEpiScratchFloats = 12 * kThreads;      // [12 slots][256 threads]: bias, mask source, previous value x 4 columns
constexpr size_t engine_staging_bytes(int ctas, bool tc5) {
  // + 1 KB: the tile's staging area starts on a 1024-byte boundary (TMA boxes with the 128-byte swizzle)
  return tc5 ? (size_t)tc5::Geom<kTc5BN>::kSmemBytes + 1024 : (size_t)(tc_smem_floats(ctas) + kEpiScratchFloats) * sizeof(float) + 1024;
}
// The engine is compiled in two occupancy variants: CTAS=1 (255 registers/thread, lowest single-job
// latency: B=256 workloads) and CTAS=2 (128 registers, twice the tile parallelism per SM: B=1024).

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc, int src_bytes) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16s(unsigned smem_dst, const float* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
// fp32 -> tf32 with round-to-nearest (ties away) done in INTEGER arithmetic (IADD + LOP3 on the fast
// pipes).  cvt.rna.tf32.f32 issues on the 16-lane/clk conversion pipe: with 6-12 conversions per
// MMA it, not the tensor core, paced the tile (measured 2.9 us for 32 k-steps).  Same result as
// cvt.rna for finite values.
__device__ __forceinline__ uint32_t cvt_tf32(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// profiling (opt-in, ilsw_trainer_set_profiling): thread 0 of CTA 0 stamps the stages of its LAST tile of
// every phase (read with ilsw_read_tile_ns): [phase][0..4] = tile start, panels issued, panels landed, MMA
// done, epilogue done.  `prof_phase` < 0 (the production setting) compiles to one predictable branch per
// stamp: no global traffic on the critical path of CTA 0.
#define ILSW_TSTAMP(i) do { if (prof_phase >= 0 && threadIdx.x == 0) g_tile_ns[prof_phase][i] = globaltimer_ns(); } while (0)

// Fragment words of ONE k-step (k8) for the whole 32x32 tile: A = 2 m16 tiles x 4 words, B = 4 n8 tiles
// x 2 words (m16n8k8 TF32 fragment layout: a0 (g,q) a1 (g+8,q) a2 (g,q+4) a3 (g+8,q+4); b0 (q,g) b1 (q+4,g)).
// Explicit 32-bit shared addresses, the 16 loads issued back to back.
struct FragSet { float a[8]; float b[8]; };
__device__ __forceinline__ void fragset_load(FragSet& f, unsigned ap, unsigned a_row8, unsigned a_k4, unsigned a_mt,
                                             unsigned bp, unsigned b_k4, unsigned b_nt, unsigned b_nt2) {
  asm volatile(
      "ld.shared.f32 %0, [%8];\n\t"
      "ld.shared.f32 %1, [%9];\n\t"
      "ld.shared.f32 %2, [%10];\n\t"
      "ld.shared.f32 %3, [%11];\n\t"
      "ld.shared.f32 %4, [%12];\n\t"
      "ld.shared.f32 %5, [%13];\n\t"
      "ld.shared.f32 %6, [%14];\n\t"
      "ld.shared.f32 %7, [%15];"
      : "=f"(f.a[0]), "=f"(f.a[1]), "=f"(f.a[2]), "=f"(f.a[3]), "=f"(f.a[4]), "=f"(f.a[5]), "=f"(f.a[6]), "=f"(f.a[7])
      : "r"(ap), "r"(ap + a_row8), "r"(ap + a_k4), "r"(ap + a_row8 + a_k4),
        "r"(ap + a_mt), "r"(ap + a_mt + a_row8), "r"(ap + a_mt + a_k4), "r"(ap + a_mt + a_row8 + a_k4)
      : "memory");
  asm volatile(
      "ld.shared.f32 %0, [%8];\n\t"
      "ld.shared.f32 %1, [%9];\n\t"
      "ld.shared.f32 %2, [%10];\n\t"
      "ld.shared.f32 %3, [%11];\n\t"
      "ld.shared.f32 %4, [%12];\n\t"
      "ld.shared.f32 %5, [%13];\n\t"
      "ld.shared.f32 %6, [%14];\n\t"
      "ld.shared.f32 %7, [%15];"
      : "=f"(f.b[0]), "=f"(f.b[1]), "=f"(f.b[2]), "=f"(f.b[3]), "=f"(f.b[4]), "=f"(f.b[5]), "=f"(f.b[6]), "=f"(f.b[7])
      : "r"(bp), "r"(bp + b_k4), "r"(bp + b_nt), "r"(bp + b_nt + b_k4),
        "r"(bp + b_nt2), "r"(bp + b_nt2 + b_k4), "r"(bp + b_nt2 + b_nt), "r"(bp + b_nt2 + b_nt + b_k4)
      : "memory");
}
__device__ __forceinline__ void mma_tf32p(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// all MMAs of one k-step: 8 independent accumulator tiles, so consecutive MMAs never depend on each
// other (mode 3: the 8 lo*hi products, then the 8 hi*lo, then the 8 hi*hi -- small terms first per
// accumulator).  n_mt / n_nt: live m16 / n8 tiles of a ragged output tile (warp-uniform).
template <bool FULL>
__device__ __forceinline__ void fragset_mma(float (&acc)[8][4], const FragSet& f, int mode, int n_mt, int n_nt) {
  uint32_t ah[8], bh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ah[i] = cvt_tf32(f.a[i]); bh[i] = cvt_tf32(f.b[i]); }
  if (mode == 3) {
    uint32_t al[8], bl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      al[i] = cvt_tf32(f.a[i] - __uint_as_float(ah[i]));
      bl[i] = cvt_tf32(f.b[i] - __uint_as_float(bh[i]));
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        if (FULL || (mt < n_mt && nt < n_nt)) mma_tf32p(acc[mt * 4 + nt], al + 4 * mt, bh + 2 * nt);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
        if (FULL || (mt < n_mt && nt < n_nt)) mma_tf32p(acc[mt * 4 + nt], ah + 4 * mt, bl + 2 * nt);
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
      if (FULL || (mt < n_mt && nt < n_nt)) mma_tf32p(acc[mt * 4 + nt], ah + 4 * mt, bh + 2 * nt);
}
__device__ __forceinline__ void cp_async4(unsigned smem_dst, const float* gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void sts_u32(unsigned addr, float v) {
  asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}

// Fill one K stage of BOTH operands into shared memory (single, compact routine: code size
// matters -- the engine's instruction working set must stay cache resident).
//   k-contiguous operand: element(row,k) = base[row*ld + k] -> smem[row*kKS + k]
//   otherwise           : element(row,k) = base[k*ld + row] -> smem[k*kMS + row]
// vec   : 16-byte cp.async.cg (aligned base, ld % 4 == 0; ragged extents copy partial vectors, zero filled)
// !vec  : 4-byte cp.async.ca per element (unaligned rows such as W0[H x 14], ragged edges, the
//         ones column of the bias gradient).  Nothing is staged through registers, so all copies
//         of a stage are in flight together.  (.ca allocates in L1: safe because every grid
//         barrier's ld.acquire.gpu invalidates the L1 -- CCTL.IVALL -- before a new phase reads.)
template <int KC>
__device__ __noinline__ void tc_fill_stage(const GemmOp& o, float* stage, int m0, int n0, int k0, int klen, bool vecA, bool vecB, int skip_mask) {
  constexpr int kKS = TcGeom<KC>::kKS, kOperandFloats = TcGeom<KC>::kOperandFloats;
  constexpr int kVecPerRow = KC / 4, kVecShift = (KC == 256 ? 6 : 5), kVecIters = 32 * KC / 4 / kThreads;
  const int tid = threadIdx.x;
  const int kpad = (klen + 15) & ~15;    // zero padded to TWO MMA k-steps (the k loop is unrolled by 2, unguarded)
#pragma unroll 1
  for (int op = o.a0 ? 1 : 0; op < 2; ++op) {      // a fused first layer produces its own A panel (tc_produce_l0)
    if ((skip_mask >> op) & 1) continue;             // this panel arrives through the TMA unit (tc_tma_stage)
    const bool isB = op != 0;
    const bool contig_k = isB ? !o.b_nc : !o.a_mc;
    const float* base = isB ? o.B : o.A;
    const int ld = isB ? o.ldb : o.lda;
    const int R = isB ? o.N : o.M;
    const int row0 = isB ? n0 : m0;
    const bool vec = isB ? vecB : vecA;
    float* sm = stage + op * kOperandFloats;
    const int sr = contig_k ? ld : 1, sk = contig_k ? 1 : ld;        // source strides
    const int dr = contig_k ? kKS : 1, dk = contig_k ? 1 : kMS;      // shared strides
    if (vec) {
      // thread -> one 16-byte vector per iteration; (row, k) advance by constant steps, so the loop body is
      // predicate + cp.async + 4 adds (the step engine is instruction-latency bound: every instruction
      // removed from a tile's prologue is ~5 cycles off the critical path)
      int r = contig_k ? (tid >> kVecShift) : ((tid & 7) << 2);
      int k = contig_k ? ((tid & (kVecPerRow - 1)) << 2) : (tid >> 3);
      const int rstep = contig_k ? (kThreads >> kVecShift) : 0, kstep = contig_k ? 0 : (kThreads >> 3);
      const int rlim = R - row0;
      const float* src = base + (size_t)(row0 + r) * sr + (size_t)(k0 + k) * sk;
      const size_t sstep = (size_t)rstep * sr + (size_t)kstep * sk;
      unsigned dst = (unsigned)__cvta_generic_to_shared(sm + r * dr + k * dk);
      const unsigned dstep = 4u * (unsigned)(rstep * dr + kstep * dk);
      // a vector at a ragged edge (N = 11, 14 ... or K % 4 != 0) copies only its live floats: cp.async zero-fills the
      // rest.  The live count along the vector direction is loop invariant (k is fixed per thread in the k-contiguous
      // layout, r in the other), so the loop body stays predicate + cp.async + 4 adds.
#if ILSW_VECRAG
      const int vbytes = 4 * min(max(contig_k ? klen - k : rlim - r, 0), 4);
#endif
#pragma unroll 4
      for (int i = 0; i < kVecIters; ++i) {                    // 32 x KC floats = 8*KC vectors
#if ILSW_VECRAG
        const int nb = (contig_k ? (r < rlim) : (k < klen)) ? vbytes : 0;
        cp_async16s(dst, nb ? src : base, nb);
#else
        const bool in = (r < rlim) && (k < klen);
        cp_async16s(dst, in ? src : base, in ? 16 : 0);
#endif
        src += sstep; dst += dstep; r += rstep; k += kstep;
      }
    } else {
      const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm);
      // only rows that exist (plus the ones column) and kpad k's are touched; the rest of the
      // 32-row fragment range is zeroed so the MMA sees exact zeros.  Element e = tid + j*256 maps to
      // (r, k) = (e / kpad, e % kpad) [k contiguous] or (e % 32, e / 32); both advance incrementally.
      const int nelem = 32 * kpad;
      int r, k, rq, kq;
      if (contig_k) { r = tid / kpad; k = tid - r * kpad; rq = kThreads / kpad; kq = kThreads - rq * kpad; }
      else { r = tid & 31; k = tid >> 5; rq = 0; kq = kThreads >> 5; }
      const int rlim = R - row0;
#pragma unroll 2
      for (int e = tid; e < nelem; e += kThreads) {
        const unsigned dst = sbase + 4u * (unsigned)(r * dr + k * dk);
        if (r < rlim && k < klen) cp_async4(dst, base + (size_t)(row0 + r) * sr + (size_t)(k0 + k) * sk);
        else sts_u32(dst, 0.f);
        r += rq; k += kq;
        if (contig_k && k >= kpad) { k -= kpad; r += 1; }
      }
    }
  }
}

// Fused first layer (GemmOp::a0): the k-contiguous A panel [32 rows][kKS] of one stage is PRODUCED here instead of
// copied -- on the tensor cores, like the layer it feeds: warp w computes panel columns [32 w, 32 w + 32) (output features
// of the first layer) for the 32 rows of the tile as 2 x 4 m16n8k8 tiles over the K0 <= 32 inputs, fragments loaded
// straight from global memory (one L2 round trip, no staging, no CTA barrier), same operand split as the main loop
// (mode 3: lo*hi, hi*lo, hi*hi).  Bias + activation are applied on the accumulators, which go to the panel and -- in the
// tn == 0 tile -- to L0FuseOp::out for the backward pass.  (A first version on the SIMT pipes, thread per column with W0^T
// staged in shared memory, took 8 us per tile: profiles/r2_phase_profile_l0_simt.txt.)
template <int KC>
__device__ __noinline__ void tc_produce_l0(const L0FuseOp& f, int M, float* panel, int m0, int k0, int klen, bool store_out, int mode) {
  constexpr int kKS = TcGeom<KC>::kKS;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, q = lane & 3;
  const int kpad = (klen + 15) & ~15;
  const int n_base = 32 * warp;                       // first panel column of this warp
  if (n_base >= kpad) return;
  const int K0 = f.K0, nks = (K0 + 7) >> 3;         // <= kFuseL0MaxK / 8 k-steps
  float acc[8][4];
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[u][i] = 0.f;
  constexpr int kMaxKs = kFuseL0MaxK / 8;
  float af[kMaxKs][8], bf[kMaxKs][8];
  // all fragment loads first
#pragma unroll
  for (int ks = 0; ks < kMaxKs; ++ks) {
    if (ks < nks) {
      const int c0 = 8 * ks + q, c1 = c0 + 4;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int r0 = m0 + 16 * mt + g, r1 = r0 + 8;
        const float* x0 = f.X + (size_t)r0 * f.ldx;
        const float* x1 = f.X + (size_t)r1 * f.ldx;
        af[ks][4 * mt + 0] = (r0 < M && c0 < K0) ? __ldcg(x0 + c0) : 0.f;
        af[ks][4 * mt + 1] = (r1 < M && c0 < K0) ? __ldcg(x1 + c0) : 0.f;
        af[ks][4 * mt + 2] = (r0 < M && c1 < K0) ? __ldcg(x0 + c1) : 0.f;
        af[ks][4 * mt + 3] = (r1 < M && c1 < K0) ? __ldcg(x1 + c1) : 0.f;
      }
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = n_base + 8 * nt + g;            // output feature within the stage
        const float* w = f.W + (size_t)(k0 + n) * K0;
        bf[ks][2 * nt + 0] = (n < klen && c0 < K0) ? __ldcg(w + c0) : 0.f;
        bf[ks][2 * nt + 1] = (n < klen && c1 < K0) ? __ldcg(w + c1) : 0.f;
      }
    }
  }
#pragma unroll
  for (int ks = 0; ks < kMaxKs; ++ks) {
    if (ks < nks) {
      uint32_t ah[8], bh[8], al[8], bl[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ah[i] = cvt_tf32(af[ks][i]); bh[i] = cvt_tf32(bf[ks][i]);
        al[i] = cvt_tf32(af[ks][i] - __uint_as_float(ah[i])); bl[i] = cvt_tf32(bf[ks][i] - __uint_as_float(bh[i]));
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          if (mode == 3) {
            mma_tf32p(acc[mt * 4 + nt], al + 4 * mt, bh + 2 * nt);
            mma_tf32p(acc[mt * 4 + nt], ah + 4 * mt, bl + 2 * nt);
          }
          mma_tf32p(acc[mt * 4 + nt], ah + 4 * mt, bh + 2 * nt);
        }
    }
  }
  // accumulator layout of m16n8: c0,c1 -> (row g, cols 2q, 2q+1); c2,c3 -> (row g+8, same cols)
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    const int n = n_base + 8 * nt + 2 * q;
    const bool live0 = n < klen, live1 = n + 1 < klen;
    const float b0v = live0 ? __ldcg(f.b + k0 + n) : 0.f, b1v = live1 ? __ldcg(f.b + k0 + n + 1) : 0.f;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int r = 16 * mt + g + 8 * h;
        const float v0 = live0 ? act_apply(acc[mt * 4 + nt][2 * h] + b0v, f.act) : 0.f;      // zero padding beyond klen
        const float v1 = live1 ? act_apply(acc[mt * 4 + nt][2 * h + 1] + b1v, f.act) : 0.f;
        *reinterpret_cast<float2*>(panel + r * kKS + n) = make_float2(v0, v1);
        if (store_out && m0 + r < M) {
          float* dst = f.out + (size_t)(m0 + r) * f.ldo + k0 + n;
          if (live1) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);
          else if (live0) *dst = v0;
        }
      }
    }
  }
}

// TMA panels (GemmOp::tma): the panel of a stage is a row of boxes {32 floats, 32 rows} of the operand's tensor map, 4 KB
// each, box j = k range [32 j, 32 j + 32) of the stage.  Warp w issues box w of both operands (one elected lane: two
// instructions per warp instead of sixteen cp.async per thread -- the panel issue was 1.7 us of a 6.3 us tile,
// profiles/r2_phase_profile.txt); out-of-range rows / k's are zero-filled by the TMA unit, ragged extents need no code.
//   k-contiguous operand (element(r,k) = base[r*ld + k]): box = [32 r][32 k], SWIZZLE_128B:
//       byte(r,k) = (k>>5)*4096 + r*128 + ((((k&31)>>2) ^ (r&7)) << 4) + (k&3)*4
//   m/n-contiguous operand (element(r,k) = base[k*ld + r]): box = [32 k][32 r], SWIZZLE_128B_ATOM_32B:
//       byte(r,k) = (k>>5)*4096 + (k&31)*128 + (((r>>3) ^ (k&3)) << 5) + (r&7)*4
// Both keep the m16n8k8 fragment loads bank-conflict free, and because warp w owns k-steps w, w + 8, ... the swizzle
// terms are loop invariant per lane (see the address set-up in gemm_tile_tc).
struct TmaState { unsigned long long* bar; unsigned par; };     // two stage barriers; bit s of par = phase parity of bar[s]
__device__ __forceinline__ void tc_tma_stage(const GemmOp& o, unsigned stage_addr, unsigned operand_bytes, int m0, int n0, int k0,
                                             int klen, unsigned long long* bar) {
  const int warp = threadIdx.x >> 5;
  const int nb = (klen + 31) >> 5;                 // boxes per operand
  const int nops = (o.tma & 1) + ((o.tma >> 1) & 1);
  if (threadIdx.x == 0) tc5::mbar_arrive_expect_tx(bar, (unsigned)(nb * nops) * 4096u);
  if (warp < nb && tc5::elect_one()) {
#if ILSW_TMA_GLOBAL_FENCE
    asm volatile("fence.proxy.async.global;" ::: "memory");     // operands written by generic stores of the previous phase
#endif
    const int kk = k0 + 32 * warp;
    if (o.tma & 1) {
      if (o.a_mc) tc5::tma_load_2d(stage_addr + 4096u * warp, o.tmapA, m0, kk, bar);
      else tc5::tma_load_2d(stage_addr + 4096u * warp, o.tmapA, kk, m0, bar);
    }
    if (o.tma & 2) {
      if (o.b_nc) tc5::tma_load_2d(stage_addr + operand_bytes + 4096u * warp, o.tmapB, n0, kk, bar);
      else tc5::tma_load_2d(stage_addr + operand_bytes + 4096u * warp, o.tmapB, kk, n0, bar);
    }
  }
  __syncwarp();
}

constexpr int kRedLd = 36;                    // row stride of a per-warp partial tile (floats; 16-byte aligned rows)
constexpr int kRedFloats = 8 * 32 * kRedLd;   // 36 KB: fits one staging stage of either occupancy variant

// 32x32 output tile.  The K range of every stage is split over the 8 warps (warp w owns k-steps w, w+8,
// ...): each warp accumulates the WHOLE 32x32 tile for its k-steps in 8 independent m16n8 accumulators
// (24 independent MMAs per k-step in 3xTF32 mode -> the loop is issue bound, not MMA-latency bound; one
// fragment word feeds 2-4 MMAs), then the 8 partial tiles are summed through shared memory in warp order
// (fixed order: bit-reproducible).  Measured predecessor (one 16x8 accumulator per warp over the full K,
// 96 dependent MMAs): 3.3 us of a 5.9 us tile.
template <int KC, int CT>      // CT: CTAs per SM of the engine variant (register budget 255 / 128)
__device__ __noinline__ void gemm_tile_tc(const GemmOp& og, int tile, float* smem, int mode, int prof_phase, const AdamOp* ad, const AdamCoef* cf, const PushCtx* push, const L0FuseOp* fz, TmaState& tma) {
  constexpr int kKC = KC, kKS = TcGeom<KC>::kKS, kOperandFloats = TcGeom<KC>::kOperandFloats, kTcStageFloats = TcGeom<KC>::kStageFloats;
  static_assert(kRedFloats <= TcGeom<KC>::kStageFloats, "partial tiles must fit one stage");
  const GemmOp o = og;                       // registers / local copy: the op descriptor lives in shared memory (reading it in place measured 4 us/step slower;
                                             // a NON-const copy, patched in place for K splits, cost 25 us/step: every field then lives in local memory)
  const int kbase = o.kbase;                 // first K index of this job's K range in the TMA maps (split-K jobs, see gemm_tile_tc_split)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tm = tile / o.tiles_n, tn = tile - tm * o.tiles_n;
  const int m0 = tm * 32, n0 = tn * 32;
  const int g = lane >> 2, q = lane & 3;
  const bool a_kc = !o.a_mc, b_kc = !o.b_nc;
  // 16-byte panel copies need 16-byte aligned source vectors: aligned base and leading dimension (ragged extents are fine)
#if ILSW_VECRAG
  const bool vecA = ((o.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.A) & 15) == 0);
  const bool vecB = ((o.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.B) & 15) == 0);
#else
  const bool vecA = ((o.lda & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.A) & 15) == 0) && ((o.K & 3) == 0) && ((o.M & 3) == 0);
  const bool vecB = ((o.ldb & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.B) & 15) == 0) && ((o.K & 3) == 0) && ((o.N & 3) == 0);
#endif
  const int nstages = (o.K + kKC - 1) / kKC;
  const int Nt = o.N;
  // bias gradient (aug_ones): bias_out[m] = sum_k A(m,k), taken from the A panels (m-contiguous layout) by the tn == 0 tile
  const bool do_aug = o.aug_ones && tn == 0;
  float colsum = 0.f;
  // live fragment tiles of a ragged output tile (N = 3, 14, 1 ...): warp-uniform
  const int n_mt = min(2, (o.M - m0 + 15) >> 4), n_nt = min(4, (Nt - n0 + 7) >> 3);
  const bool full = (n_mt == 2) && (n_nt == 4);
  // epilogue mapping: thread -> (row = tid / 8, 4 consecutive columns)
  const int er = m0 + (tid >> 3), ec = n0 + ((tid & 7) << 2);
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(smem);
  // Stage 0 and the epilogue inputs are requested FIRST: the accumulator / fragment-address set-up below (~100 dependent
  // instructions at 8 warps per SM) then runs while the panels are in flight instead of in front of them
  ILSW_TSTAMP(0);
  {
    const int klen0 = min(kKC, o.K);
    if (o.tma) tc_tma_stage(o, sbase, 4u * (unsigned)kOperandFloats, m0, n0, kbase, klen0, &tma.bar[0]);
    if (o.tma != 3) tc_fill_stage<KC>(og, smem, m0, n0, 0, klen0, vecA, vecB, o.tma);
    cp_async_commit();
    if (fz)           // the B panel is in flight; produce the A panel meanwhile
      tc_produce_l0<KC>(*fz, o.M, smem, m0, 0, klen0, tn == 0, mode);
#if ILSW_EIN_CPASYNC
    const unsigned sc = sbase + 4u * (unsigned)(TcGeom<KC>::kSmemFloats + tid);
    if (er < o.M) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (ec + i < Nt) {
          if (o.bias) cp_async4(sc + 4u * (unsigned)(i * kThreads), o.bias + ec + i);
          if (o.mask != ACT_NONE) cp_async4(sc + 4u * (unsigned)((4 + i) * kThreads), o.H + (size_t)er * o.ldh + ec + i);
          if (o.accumulate) cp_async4(sc + 4u * (unsigned)((8 + i) * kThreads), o.C + (size_t)er * o.ldc + ec + i);
        }
      }
    }
    cp_async_commit();
#endif
  }
  ILSW_TSTAMP(1);
  float acc[8][4];
#pragma unroll
  for (int u = 0; u < 8; ++u)
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[u][i] = 0.f;
  EpiIn ein[4];
  // fused-optimiser tiles: the Adam state of this thread's 4 elements (and of the bias element of a tn == 0 tile) is
  // fetched while the panels are in flight -- no L2 round trip left in the epilogue
#if ILSW_ADAM_PREFETCH
  int gi[4] = {-1, -1, -1, -1}, gib = -1;
  float am[4], av[4], ap[4], at[4], bm = 0.f, bv = 0.f, bp = 0.f, bt = 0.f;
#endif
  // epilogue inputs by register loads (builds without ILSW_EIN_CPASYNC)
#if ILSW_EIN_FIRST || !ILSW_EIN_CPASYNC
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (er < o.M && ec + i < Nt) ein[i] = epi_load(o, er, ec + i);
#endif
#define ILSW_ADAM_STATE_LOADS() \
  if (ad) { \
    _Pragma("unroll") \
    for (int i = 0; i < 4; ++i) \
      if (er < o.M && ec + i < Nt) { \
        gi[i] = gemm_grad_index(o, *ad, er, ec + i); \
        am[i] = __ldcg(ad->m + gi[i]); av[i] = __ldcg(ad->v + gi[i]); ap[i] = __ldcg(ad->p + gi[i]); \
        at[i] = ad->target ? __ldcg(ad->target + gi[i]) : 0.f; \
      } \
    if (do_aug && tid < 32 && m0 + tid < o.M) { \
      gib = gemm_grad_index(o, *ad, m0 + tid, o.N); \
      bm = __ldcg(ad->m + gib); bv = __ldcg(ad->v + gib); bp = __ldcg(ad->p + gib); \
      bt = ad->target ? __ldcg(ad->target + gib) : 0.f; \
    } \
  }
#if ILSW_ADAM_PREFETCH
  ILSW_ADAM_STATE_LOADS();
#endif
  // fragment addressing (32-bit shared addresses, bytes)
  const bool a_tma = (o.tma & 1) != 0, b_tma = (o.tma & 2) != 0;
  // per operand: offset of this lane's first fragment word for k-step `warp` (x_w), advance per 8 k-steps (x_it), and the
  // offsets between the fragment words of one k-step.  The TMA layouts are swizzled; all offsets wrap in 32 bits.
  unsigned a_w, a_it, a_row8, a_k4, a_mt, b_w, b_it, b_k4, b_nt, b_nt2;
  if (a_tma) {
    a_it = 8192u;
    if (a_kc) { a_w = (unsigned)((warp >> 2) * 4096 + g * 128 + (((2 * (warp & 3)) ^ g) << 4) + q * 4); a_row8 = 1024u; a_mt = 2048u; a_k4 = (g & 1) ? 0u - 16u : 16u; }
    else { a_w = (unsigned)((warp >> 2) * 4096 + (8 * (warp & 3) + q) * 128 + (q << 5) + g * 4); a_k4 = 512u;
           // rows g, g + 8, g + 16, g + 24 = atoms 0..3, XORed with k & 3 = q (the same for k and k + 4)
           a_row8 = (q & 1) ? 0u - 32u : 32u; a_mt = (q & 2) ? 0u - 64u : 64u; }
  } else {
    const unsigned a_k8 = 8u * (a_kc ? 4 : 4 * kMS);
    a_w = 4u * (a_kc ? g * kKS + q : q * kMS + g) + (unsigned)warp * a_k8; a_it = 8u * a_k8;
    a_row8 = 4u * (a_kc ? 8 * kKS : 8); a_k4 = 4u * (a_kc ? 4 : 4 * kMS); a_mt = 2u * a_row8;
  }
  if (b_tma) {
    b_it = 8192u;
    if (b_kc) { b_w = (unsigned)((warp >> 2) * 4096 + g * 128 + (((2 * (warp & 3)) ^ g) << 4) + q * 4); b_nt = 1024u; b_nt2 = 2048u; b_k4 = (g & 1) ? 0u - 16u : 16u; }
    else { b_w = (unsigned)((warp >> 2) * 4096 + (8 * (warp & 3) + q) * 128 + (q << 5) + g * 4); b_k4 = 512u;
           b_nt = (q & 1) ? 0u - 32u : 32u; b_nt2 = (q & 2) ? 0u - 64u : 64u; }
  } else {
    const unsigned b_k8 = 8u * (b_kc ? 4 : 4 * kMS);
    b_w = 4u * (b_kc ? g * kKS + q : q * kMS + g) + (unsigned)warp * b_k8; b_it = 8u * b_k8;
    b_k4 = 4u * (b_kc ? 4 : 4 * kMS); b_nt = 4u * (b_kc ? 8 * kKS : 8); b_nt2 = 2u * b_nt;
  }
#pragma unroll 1
  for (int st = 0; st < nstages; ++st) {
    if (st + 1 < nstages) {
      const int k0 = (st + 1) * kKC;
      if (o.tma) tc_tma_stage(o, sbase + 4u * (unsigned)(((st + 1) & 1) * kTcStageFloats), 4u * (unsigned)kOperandFloats, m0, n0, kbase + k0,
                              min(kKC, o.K - k0), &tma.bar[(st + 1) & 1]);
      if (o.tma != 3) tc_fill_stage<KC>(og, smem + ((st + 1) & 1) * kTcStageFloats, m0, n0, k0, min(kKC, o.K - k0), vecA, vecB, o.tma);
      cp_async_commit();
      if (fz)           // the B panel is in flight; produce the A panel meanwhile
        tc_produce_l0<KC>(*fz, o.M, smem + ((st + 1) & 1) * kTcStageFloats, m0, k0, min(kKC, o.K - k0), tn == 0, mode);
    }
    if (st + 1 < nstages) cp_async_wait<1>(); else cp_async_wait<0>();
    if (o.tma) {       // every consuming thread observes the completion of the stage's boxes
      if (!tc5::mbar_wait(&tma.bar[st & 1], (tma.par >> (st & 1)) & 1u)) __trap();
      tma.par ^= 1u << (st & 1);
    }
    __syncthreads();
    if (st == 0) ILSW_TSTAMP(2);
    {
      const int klen = min(kKC, o.K - st * kKC);
      const int ksteps = (klen + 7) >> 3;            // panels are zero padded up to a multiple of 16
      unsigned pa = sbase + 4u * (unsigned)((st & 1) * kTcStageFloats) + a_w;
      unsigned pb = sbase + 4u * (unsigned)((st & 1) * kTcStageFloats + kOperandFloats) + b_w;
      // software pipeline over this warp's k-steps: the fragment loads of the next k-step are issued
      // before the MMAs of the current one (ping-pong fragment registers)
      FragSet f0, f1;
      int ks = warp;
      if (ks < ksteps) fragset_load(f0, pa, a_row8, a_k4, a_mt, pb, b_k4, b_nt, b_nt2);
      if (full && CT == 2 && ILSW_TC128_SIMPLE) {
        // two CTAs per SM (128-register cap): one fragment set, no ping-pong -- four warps per scheduler hide the shared
        // memory latency, and the second fragment set is what pushed this variant into local-memory spills
#pragma unroll 1
        for (; ks < ksteps; ks += 8) {
          fragset_mma<true>(acc, f0, mode, 2, 4);
          pa += a_it; pb += b_it;
          if (ks + 8 < ksteps) fragset_load(f0, pa, a_row8, a_k4, a_mt, pb, b_k4, b_nt, b_nt2);
        }
      } else if (full) {          // all 8 fragment tiles live: straight-line MMAs, no guards
#pragma unroll 1
        while (ks < ksteps) {
          int kn = ks + 8;
          pa += a_it; pb += b_it;
          if (kn < ksteps) fragset_load(f1, pa, a_row8, a_k4, a_mt, pb, b_k4, b_nt, b_nt2);
          fragset_mma<true>(acc, f0, mode, 2, 4);
          ks = kn;
          if (ks >= ksteps) break;
          kn = ks + 8;
          pa += a_it; pb += b_it;
          if (kn < ksteps) fragset_load(f0, pa, a_row8, a_k4, a_mt, pb, b_k4, b_nt, b_nt2);
          fragset_mma<true>(acc, f1, mode, 2, 4);
          ks = kn;
        }
      } else {             // ragged output tile (N = 3, 14, 1 ...): skip the dead fragment tiles
#pragma unroll 1
        for (; ks < ksteps; ks += 8) {
          fragset_mma<false>(acc, f0, mode, n_mt, n_nt);
          pa += a_it; pb += b_it;
          if (ks + 8 < ksteps) fragset_load(f0, pa, a_row8, a_k4, a_mt, pb, b_k4, b_nt, b_nt2);
        }
      }
    }
    if (do_aug) {      // thread -> (m = tid % 32, k = tid / 32 + 8 i): conflict-free reads of the m-contiguous A panel
      const int klen = min(kKC, o.K - st * kKC);
      float sacc = 0.f;
      if (a_tma) {     // swizzled boxes [32 k][32 m]: word(m,k) = (k>>5)*1024 + (k&31)*32 + (((m>>3) ^ (k&3)) << 3) + (m&7)
        const float* As = smem + (st & 1) * kTcStageFloats;
        const int m = tid & 31;
        for (int k = tid >> 5; k < klen; k += 8) sacc += As[(k >> 5) * 1024 + (k & 31) * 32 + (((m >> 3) ^ (k & 3)) << 3) + (m & 7)];
      } else {
        const float* As = smem + (st & 1) * kTcStageFloats + (tid >> 5) * kMS + (tid & 31);
        for (int k = tid >> 5; k < klen; k += 8, As += 8 * kMS) sacc += *As;
      }
      colsum += sacc;
    }
    __syncthreads();   // stage buffer may be refilled by the next iteration's prefetch / reused for the partials
  }
  // cross-warp reduction of the partial tiles (stage 0 area; all warps are past their MMAs)
  const int nred = min(8, (min(kKC, o.K) + 7) >> 3);      // warps that own at least one k-step (stage 0 is the longest)
  if (warp < nred) {
    float* red = smem + warp * (32 * kRedLd);
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        float* r0p = red + (mt * 16 + g) * kRedLd + nt * 8 + 2 * q;
        *reinterpret_cast<float2*>(r0p) = make_float2(acc[mt * 4 + nt][0], acc[mt * 4 + nt][1]);
        *reinterpret_cast<float2*>(r0p + 8 * kRedLd) = make_float2(acc[mt * 4 + nt][2], acc[mt * 4 + nt][3]);
      }
  }
  if (do_aug) smem[kRedFloats + tid] = colsum;
  __syncthreads();
  float4 sum = *reinterpret_cast<const float4*>(smem + (tid >> 3) * kRedLd + ((tid & 7) << 2));
  for (int w = 1; w < nred; ++w) {
    const float4 v = *reinterpret_cast<const float4*>(smem + w * (32 * kRedLd) + (tid >> 3) * kRedLd + ((tid & 7) << 2));
    sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
  }
  ILSW_TSTAMP(3);
  const float outv[4] = {sum.x, sum.y, sum.z, sum.w};
#if ILSW_EIN_CPASYNC
  {
    const float* sc = smem + TcGeom<KC>::kSmemFloats + tid;       // landed with the last stage's cp.async.wait_group 0
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      ein[i].bias = o.bias ? sc[i * kThreads] : 0.f;
      ein[i].h = (o.mask != ACT_NONE) ? sc[(4 + i) * kThreads] : 0.f;
      ein[i].prev = o.accumulate ? sc[(8 + i) * kThreads] : 0.f;
    }
  }
#endif
  const bool ad_sh = ad && adam_has_shadow(*ad);      // aligned W0 copies to maintain (tcgen05 programs only)
#if !ILSW_ADAM_PREFETCH
  int gi[4] = {-1, -1, -1, -1}, gib = -1;
  float am[4], av[4], ap[4], at[4], bm = 0.f, bv = 0.f, bp = 0.f, bt = 0.f;
  ILSW_ADAM_STATE_LOADS();
#endif
  if (ad) {
    // weight-gradient tile with the optimiser fused (Adam state prefetched in the prologue)
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (gi[i] >= 0) {
        const float gv = o.accumulate ? outv[i] + ein[i].prev : outv[i];
        o.C[(size_t)er * o.ldc + ec + i] = gv;
        adam_math_store(*ad, *cf, gi[i], gv, am[i], av[i], ap[i], at[i], ad_sh);
      }
  } else {
    if (push) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (er < o.M && ec + i < Nt) epi_store(o, er, ec + i, outv[i], ein[i], push);
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (er < o.M && ec + i < Nt) epi_store(o, er, ec + i, outv[i], ein[i]);
    }
  }
  if (do_aug && tid < 32 && m0 + tid < o.M) {
    float bsum = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) bsum += smem[kRedFloats + w * 32 + tid];
    const EpiIn e = epi_load(o, m0 + tid, o.N);
    epi_store(o, m0 + tid, o.N, bsum, e, push);
    if (ad) adam_math_store(*ad, *cf, gib, o.accumulate ? bsum + e.prev : bsum, bm, bv, bp, bt, ad_sh);
  }
  tc5::fence_proxy_async();   // this tile's generic accesses to the staging area precede the next tile's TMA writes
  __syncthreads();     // the partial tiles are read before the next job's panels overwrite them
  ILSW_TSTAMP(4);
}

// Weight-gradient GEMM split along K (= batch) over partial gradient arenas (GemmOp::ksplit on a plain tile; operands k-major:
// a_mc && b_nc): job (ks, tile) sums the 32-deep K blocks [b0, b1) into arena ks.  Tiny outputs (the discriminator's W1: 128 x 23)
// would otherwise walk K = 512 on 4 CTAs while 144 SMs wait (GAIL phase 5: 15 us).  A shifted copy of the descriptor in shared
// memory, then the same tile.
template <int KC, int CT>
__device__ __noinline__ void gemm_tile_tc_split(const GemmOp& og, int tile_in, float* smem, int mode, int prof_phase, TmaState& tma) {
  __shared__ GemmOp s_tsplit;
  const int per = og.tiles_m * og.tiles_n, ks = tile_in / per, tile = tile_in - ks * per;
  __syncthreads();
  if (threadIdx.x == 0) {
    GemmOp o = og;
    const int nkb = (o.K + 31) >> 5;
    const int kb = 32 * ((ks * nkb) / o.ksplit), ke = min(o.K, 32 * (((ks + 1) * nkb) / o.ksplit));
    o.A += (size_t)kb * o.lda; o.B += (size_t)kb * o.ldb; o.K = ke - kb; o.kbase = kb;
    o.C += (size_t)ks * o.split_stride;
    if (o.bias_out) o.bias_out += (size_t)ks * o.split_stride;
    s_tsplit = o;
  }
  __syncthreads();
  gemm_tile_tc<KC, CT>(s_tsplit, tile, smem, mode, prof_phase, nullptr, nullptr, nullptr, nullptr, tma);
}

// Skinny weight-gradient tile: M <= 8 output rows (critic output layer M = 1, policy heads M = A), a_mc && b_nc.
//   C[m, n0+col] = sum_k A[k*lda + m] * B[k*ldb + n0 + col]   (+ bias_out[m] = sum_k A[k*lda + m])
// Exact fp32 on the SIMT pipes, straight from L2 (no panels, no MMA): thread -> (col = tid % 32, k = tid / 32 + 8 i), 8 k's
// of loads in flight per thread; the 8 k-groups are summed through shared memory in a fixed order.  A ragged MMA tile
// for these shapes costs a full tile (5 us); this is ~1.5 us.
// Split along K (tcgen05 programs, K = batch >= 512): job (ks, tile) sums k in [ks * K / S, (ks + 1) * K / S) into the
// partial gradient arena ks (C + ks * split_stride); one CTA walking K = 1024 alone took 23 us.
__device__ __noinline__ void gemm_tile_skinny(const GemmOp& og, int tile, float* smem, const AdamOp* ad, const AdamCoef* cf, const PushCtx* push) {
  const GemmOp o = og;
  const int tid = threadIdx.x, col = tid & 31, kg = tid >> 5;
  const int n = tile * 32 + col;
  const int M = o.M;
  const bool nv = n < o.N;
  float acc[8], asum[8];
#pragma unroll
  for (int m = 0; m < 8; ++m) { acc[m] = 0.f; asum[m] = 0.f; }
  const float* Bp = o.B + n;
#pragma unroll 1
  for (int k0 = kg; k0 < o.K; k0 += 64) {
    float bv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int k = k0 + 8 * u;
      bv[u] = (k < o.K && nv) ? __ldcg(Bp + (size_t)k * o.ldb) : 0.f;
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      if (m < M) {
        float av[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k0 + 8 * u;
          av[u] = k < o.K ? __ldcg(o.A + (size_t)k * o.lda + m) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) { acc[m] = fmaf(av[u], bv[u], acc[m]); asum[m] += av[u]; }
      }
    }
  }
#pragma unroll
  for (int m = 0; m < 8; ++m)
    if (m < M) {
      smem[(kg * 8 + m) * 32 + col] = acc[m];
      if (col == 0) smem[2048 + kg * 8 + m] = asum[m];
    }
  __syncthreads();
  {
    const int m = tid >> 5;                       // thread -> output (m, col)
    if (m < M && nv) {
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += smem[(w * 8 + m) * 32 + col];
      const EpiIn e = epi_load(o, m, n);
      epi_store(o, m, n, v, e, push);
      if (ad) adam_fused_elem(*ad, *cf, gemm_grad_index(o, *ad, m, n), o.accumulate ? v + e.prev : v);
    }
    if (o.aug_ones && tile == 0 && tid < M) {     // bias gradient
      float v = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) v += smem[2048 + w * 8 + tid];
      const EpiIn e = epi_load(o, tid, o.N);
      epi_store(o, tid, o.N, v, e, push);
      if (ad) adam_fused_elem(*ad, *cf, gemm_grad_index(o, *ad, tid, o.N), o.accumulate ? v + e.prev : v);
    }
  }
  __syncthreads();
}
// split form (tcgen05 programs only): a shifted copy of the descriptor per K range, then the same tile
__device__ __noinline__ void gemm_tile_skinny_split(const GemmOp& og, int tile_in, float* smem) {
  __shared__ GemmOp s_split;
  const int nsplit = og.ksplit, ks = tile_in / og.tiles_n, tile = tile_in - ks * og.tiles_n;
  const int kbeg = (ks * og.K) / nsplit, kend = ((ks + 1) * og.K) / nsplit;
  __syncthreads();
  if (threadIdx.x == 0) {
    GemmOp o = og;
    o.A += (size_t)kbeg * o.lda; o.B += (size_t)kbeg * o.ldb; o.K = kend - kbeg;
    o.C += (size_t)ks * o.split_stride;
    if (o.bias_out) o.bias_out += (size_t)ks * o.split_stride;
    s_split = o;
  }
  __syncthreads();
  gemm_tile_skinny(s_split, tile, smem, nullptr, nullptr, nullptr);
}
ILSW_HD bool gemm_is_skinny(const GemmOp& o) { return o.M <= 8 && o.a_mc && o.b_nc && o.tiles_m == 1; }


// ------------------------------------------------------------------------------------------
// replica exchange: every rank PUSHES its policy gradient into every rank's receive slot over
// NVLink peer stores, then publishes a sequence number; the Adam job sums the slots in rank
// order (bit-identical on all ranks) -- no NCCL call, no host round trip, fused into the step.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool replica_exchange(const Replica& rp, unsigned seq, BarrierState* bar, unsigned& gen, int* abort_flag) {
  const int parity = (int)(seq & 1u);
  const size_t slot = ((size_t)parity * rp.world + rp.rank) * (size_t)rp.nstride;
  const int n4 = rp.n >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(rp.grad) + i);
    for (int sp = 1; sp < rp.g_splits; ++sp) {
      const float4 w = __ldcg(reinterpret_cast<const float4*>(rp.grad + (size_t)sp * rp.g_split_stride) + i);
      v.x += w.x; v.y += w.y; v.z += w.z; v.w += w.w;
    }
    for (int r = 0; r < rp.world; ++r) reinterpret_cast<float4*>(rp.recv_peer[r] + slot)[i] = v;
  }
  for (int i = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < rp.n; i += gridDim.x * blockDim.x) {
    float v = __ldcg(rp.grad + i);
    for (int sp = 1; sp < rp.g_splits; ++sp) v += __ldcg(rp.grad + (size_t)sp * rp.g_split_stride + i);
    for (int r = 0; r < rp.world; ++r) rp.recv_peer[r][slot + i] = v;
  }
  __threadfence_system();
  if (!grid_barrier(bar, gridDim.x, gen, abort_flag)) return false;
  if (blockIdx.x == 0 && threadIdx.x < rp.world) st_release_sys(rp.flags_peer[threadIdx.x] + rp.rank, seq);
  __shared__ int s_ok2;
  if (threadIdx.x == 0) s_ok2 = 1;
  __syncthreads();
  if (threadIdx.x < rp.world) {
    long long t0 = clock64();
    while ((int)(ld_acquire_sys(rp.flags_local + threadIdx.x) - seq) < 0) {
      long long dt = clock64() - t0;
      // peers may start their launch seconds later (module load, host jitter): be patient (~20 s)
      if (dt > kSpinLimit && (*(volatile int*)abort_flag || dt > 10 * kSpinLimit)) {
        atomicExch(abort_flag, 1);
        s_ok2 = 0;
        break;
      }
    }
  }
  __syncthreads();
  return s_ok2 != 0;
}

// epilogue-push exchange, consumer side: wait until every replica has signalled update `seq` (1-based)
__device__ __forceinline__ bool replica_wait_pushes(const Replica& rp, unsigned seq, int* abort_flag) {
  __shared__ int s_ok3;
  if (threadIdx.x == 0) {
    const unsigned long long want = (unsigned long long)seq * (unsigned long long)rp.world;
    int ok = 1;
    const long long t0 = clock64();
    while (ld_acquire_sys_u64(rp.cnt_local) < want) {
      const long long dt = clock64() - t0;
      if (dt > kSpinLimit && (*(volatile int*)abort_flag || dt > 10 * kSpinLimit)) { atomicExch(abort_flag, 1); ok = 0; break; }
    }
    s_ok3 = ok;
  }
  __syncthreads();
  return s_ok3 != 0;
}
// producer side, two levels: every CTA fences its pushes system-wide and arrives on the replica's LOCAL counter; the CTA that
// completes the replica's arrivals for update `seq` signals the peers ONCE (one multimem.red through the NVSwitch, or `world`
// peer reductions).  The flat form -- every CTA of every replica adding to every replica's counter -- put world x 148
// serialised system-scope atomics on one address per step: 296 at N = 2, 1 184 at N = 8, the part of the exchange cost that
// grew with N (per-replica step 159.7 -> 175.4 us from N = 2 to N = 8 with per-GPU step times equal to 0.6 %).
__device__ __forceinline__ void replica_signal_pushes(const Replica& rp, unsigned seq) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned old = atom_add_acq_rel_gpu(rp.arrive_local, 1u);
    if (old + 1u == seq * gridDim.x) {          // modular arithmetic: consistent across a wrap of the 32-bit counter
      __threadfence_system();
      if (rp.cnt_mc) asm volatile("multimem.red.release.sys.global.add.u64 [%0], %1;" ::"l"(rp.cnt_mc), "l"(1ull) : "memory");
      else
        for (int r = 0; r < rp.world; ++r) red_add_release_sys_u64(rp.cnt_peer[r], 1ull);
    }
  }
}

__device__ __forceinline__ float replica_reduced_grad(const Replica& rp, unsigned seq, int i) {
  const size_t base = (size_t)(seq & 1u) * rp.world * (size_t)rp.nstride;
  float g = 0.f;
  for (int r = 0; r < rp.world; ++r) g += __ldcv(rp.recv_local + base + (size_t)r * rp.nstride + i);
  return g;
}

// ------------------------------------------------------------------------------------------
// the engine
// ------------------------------------------------------------------------------------------
// Dynamic shared memory: [GEMM staging (kTcSmemFloats)] [program copy: phases, ops, ctx].
// The program is immutable during a launch; keeping it in shared memory removes every L2 round
// trip from job dispatch (the L1 is invalidated at each grid barrier).
struct SmemProgram { const Phase* phases; const Op* ops; const Ctx* ctx; int n_phases; };

__device__ __forceinline__ size_t align16(size_t x) { return (x + 15) & ~size_t(15); }

// one flat Adam(+Polyak) chunk of kAdamChunk elements: ALL loads first, then the arithmetic, then the stores.
// The load half is BRANCH FREE (out-of-range lanes read a valid address and are masked at the stores): with a branch per
// element the compiler kept the state arrays in local memory and stored every value right after loading it, so each
// element waited for its own L2 round trips in turn -- 14 us per job on the B200 (ncu: STL ... stall_long_sb).
__device__ __noinline__ void adam_job(const AdamOp& aos, const AdamCoef& cfs, int j, bool reduced, const Replica& rp, unsigned xseq, int world) {
  const AdamOp ao = aos;
  const AdamCoef cf = cfs;
  const float gscale = (ao.grad_scale_world && world > 1) ? 1.0f / (float)world : 1.0f;
  const int beg = ao.begin + j * kAdamChunk, end = min(ao.n, beg + kAdamChunk);
  constexpr int E = kAdamChunk / kThreads;
  float g[E], m[E], v[E], p[E], tg[E];
  int idx[E];
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = beg + threadIdx.x + u * kThreads;
    idx[u] = i < end ? i : beg;
  }
  if (reduced) {                // rank-ordered sum of the replicas' receive slots
    const size_t base = (size_t)(xseq & 1u) * rp.world * (size_t)rp.nstride;
#pragma unroll
    for (int u = 0; u < E; ++u) {
      float part[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float x = __ldcv(rp.recv_local + base + (size_t)(r < rp.world ? r : 0) * rp.nstride + idx[u]);
        part[r] = r < rp.world ? x : 0.f;
      }
      float acc = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) acc += part[r];
      g[u] = acc;
    }
  } else {                      // split-K partial arenas in split order (one arena without splits)
#pragma unroll
    for (int u = 0; u < E; ++u) {
      float part[kMaxGradSplits];
#pragma unroll
      for (int sp = 0; sp < kMaxGradSplits; ++sp) {
        const bool live = sp == 0 || sp < ao.g_splits;
        const float x = __ldcg(ao.g + (size_t)(live ? sp : 0) * ao.g_split_stride + idx[u]);
        part[sp] = live ? x : 0.f;
      }
      float acc = part[0];
#pragma unroll
      for (int sp = 1; sp < kMaxGradSplits; ++sp) acc += part[sp];
      g[u] = acc;
    }
  }
  const float* tptr = ao.target ? ao.target : ao.p;
#pragma unroll
  for (int u = 0; u < E; ++u) {
    // __fmul_rn: the scaled gradient must be ROUNDED before Adam consumes it -- a plain `* gscale` gets contracted into
    // the first FMA of adam_math_store ((s * gscale) - m), which made R identical replicas differ from one replica by an
    // ulp (tools/replica_check.py, test A: the single-replica program applies Adam in the weight-gradient epilogues)
    g[u] = __fmul_rn(g[u], gscale);
    m[u] = ao.m[idx[u]]; v[u] = ao.v[idx[u]]; p[u] = ao.p[idx[u]];
    tg[u] = tptr[idx[u]];
  }
  const bool sh = adam_has_shadow(ao);
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = beg + threadIdx.x + u * kThreads;
    if (i < end) adam_math_store(ao, cf, i, g[u], m[u], v[u], p[u], tg[u], sh);
  }
}

// flat Polyak chunk, same structure
__device__ __noinline__ void polyak_job(const PolyakOp& pos, int j) {
  const PolyakOp po = pos;
  const int beg = j * kAdamChunk, end = min(po.n, beg + kAdamChunk);
  constexpr int E = kAdamChunk / kThreads;
  float tv[E], sv[E];
  const float om = (float)(1.0 - (double)po.tau);
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = beg + threadIdx.x + u * kThreads, ii = i < end ? i : beg;
    tv[u] = po.target[ii]; sv[u] = __ldcg(po.src + ii);
  }
#pragma unroll
  for (int u = 0; u < E; ++u) {
    const int i = beg + threadIdx.x + u * kThreads;
    if (i < end) {
      const float tn = tv[u] * om + sv[u] * po.tau;
      po.target[i] = tn;
      shadow_store(po.sh_t, i, tn);
    }
  }
}

template <int CTAS, bool TC5>
__global__ void __launch_bounds__(kThreads, CTAS)
ilsw_engine_kernel(const Program* __restrict__ prog, RunArgs a_param, BarrierState* bar, Replica rp_param) {
  static_assert(!TC5 || CTAS == 1, "the tcgen05 variant runs one CTA per SM");
  static_assert(!TC5 || engine_staging_bytes(1, true) >= (size_t)(tc_smem_floats(1) + kEpiScratchFloats) * sizeof(float), "staging area also serves the mma.sync tile");
  static_assert(!TC5 || (size_t)(tc_smem_floats(1) + kEpiScratchFloats) * sizeof(float) + 1024 <= (size_t)tc5::kStages * tc5::Geom<kTc5BN>::kStageBytes, "the all-ones tile of the tcgen05 path lies beyond the mma.sync tile's scratch");
  unsigned char* dyn_smem = reinterpret_cast<unsigned char*>(ilsw_dyn_smem_f);
  float* smem = ilsw_dyn_smem_f;
  // launch arguments live in shared memory: row kernels take them by reference and the replica tables are indexed
  // dynamically -- from the parameter bank either would force a per-thread local-memory copy
  __shared__ RunArgs s_a;
  __shared__ Replica s_rp;
  __shared__ AdamCoef s_coefs[kMaxNets];       // this step's Adam coefficients per optimiser slot
  __shared__ int s_adam_op[kMaxNets];
  __shared__ unsigned s_gen;
  __shared__ double s_p1[kMaxNets], s_p2[kMaxNets], s_b1[kMaxNets], s_b2[kMaxNets];   // running beta^t per Adam slot
  __shared__ int s_pt[kMaxNets];
  __shared__ tc5::Sync s_tc5;
  __shared__ PushCtx s_push;
  __shared__ unsigned long long s_tma_bar[2];  // stage barriers of the mma.sync tile's TMA panels (tc_tma_stage)
  __shared__ int s_last_ph;                    // last active phase of the current step (hosts Ctx::tail_op1)
  __shared__ int s_job0[kMaxPhases];           // (op index << 16 | job within the op) of this CTA's FIRST job of every phase, -1: none
  __shared__ unsigned char s_pinfo[kMaxPhases];  // per phase, resolved once per launch: 1 = can be active in this launch, 2 = also has a
                                                 // per-step condition (first step / TD3 policy / statistics step), 4 = exchange, 8 = push
  const int n_phases = prog->n_phases, n_ops = prog->n_ops;
  constexpr int KC = CTAS == 2 ? 128 : kKC1;
  unsigned char* pbase = dyn_smem + engine_staging_bytes(CTAS, TC5);
  // tcgen05 variant only (the B = 256 variant keeps its loop state minimal: everything live across a grid barrier that
  // does not fit the register file is reloaded from local memory through L2 -- the barrier's acquire invalidates L1)
  [[maybe_unused]] unsigned char* tc5_smem = nullptr;
  [[maybe_unused]] tc5::State tc5_state{0u, 0u};
  if constexpr (TC5) {
    tc5_smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));
    tc5::setup<kTc5BN>(s_tc5, tc5_smem);
  }
// a failed barrier / exchange / tile wait ends the launch for every thread of the CTA (uniform); the tcgen05 variant leaves
// through the common exit because tensor memory must be released before the CTA retires
#define ILSW_ALIVE (!TC5 || alive)
#define ILSW_DIE() { if constexpr (TC5) { alive = false; break; } else { return; } }
  Phase* s_phases = reinterpret_cast<Phase*>(pbase);
  Op* s_ops = reinterpret_cast<Op*>(pbase + align16(sizeof(Phase) * (size_t)n_phases));
  Ctx* s_ctx = reinterpret_cast<Ctx*>(reinterpret_cast<unsigned char*>(s_ops) + align16(sizeof(Op) * (size_t)n_ops));
  {
    const int* src; int* dst; int n;
    src = reinterpret_cast<const int*>(prog->phases); dst = reinterpret_cast<int*>(s_phases); n = (int)(sizeof(Phase) * n_phases / 4);
    for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = src[i];
    src = reinterpret_cast<const int*>(prog->ops); dst = reinterpret_cast<int*>(s_ops); n = (int)(sizeof(Op) * n_ops / 4);
    for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = src[i];
    src = reinterpret_cast<const int*>(&prog->ctx); dst = reinterpret_cast<int*>(s_ctx); n = (int)(sizeof(Ctx) / 4);
    for (int i = threadIdx.x; i < n; i += kThreads) dst[i] = src[i];
  }
  if (threadIdx.x == 0) {
    s_gen = 0u; s_a = a_param; s_rp = rp_param;   // the host zeroes the barrier state before every launch
    if (blockIdx.x == 0 && a_param.bar_other) *a_param.bar_other = 0u;     // the previous launch's barrier state: idle, reused by the next
    tc5::mbar_init(&s_tma_bar[0], 1); tc5::mbar_init(&s_tma_bar[1], 1);
    tc5::fence_barrier_init();
  }
  if (threadIdx.x < kMaxNets) { s_pt[threadIdx.x] = -1; s_adam_op[threadIdx.x] = -1; }
  __syncthreads();
  const RunArgs& a = s_a;
  const Replica& rp = s_rp;
  if (threadIdx.x == 0) {   // one pow() per optimiser per LAUNCH; afterwards beta^t is a running product
    for (int i = 0; i < n_ops; ++i)
      if (s_ops[i].kind == OP_ADAM) {
        const AdamOp& ao = s_ops[i].adam;
        if (s_adam_op[ao.slot] < 0) s_adam_op[ao.slot] = i;
        s_b1[ao.slot] = ao.beta1; s_b2[ao.slot] = ao.beta2; s_pt[ao.slot] = a.t0[ao.slot];
        s_p1[ao.slot] = pow(ao.beta1, (double)a.t0[ao.slot]); s_p2[ao.slot] = pow(ao.beta2, (double)a.t0[ao.slot]);
      }
  }
  __syncthreads();
  // job -> (op, local job) of this CTA's first job per phase, resolved once per launch: the dispatch of a phase is then one
  // shared-memory load instead of a dependent walk over the phase's op list behind every grid barrier
  for (int ph = threadIdx.x; ph < n_phases; ph += kThreads) {
    const Phase& P = s_phases[ph];
    int j = (int)blockIdx.x, oi = P.op_begin, v = -1;
    if (j < P.total_jobs) {
      while (j >= s_ops[oi].n_jobs) { j -= s_ops[oi].n_jobs; ++oi; }
      v = (oi << 16) | j;
    }
    s_job0[ph] = v;
    // launch-invariant half of phase_active(): replica count and update mode are fixed for the launch
    const RunArgs& a0 = s_a;
    bool on = true;
    if ((P.cond & COND_DISC_PART) && a0.update_mode == UPDATE_POLICY_ONLY) on = false;
    if ((P.cond & COND_POLICY_PART) && a0.update_mode == UPDATE_DISC_ONLY) on = false;
    if ((P.cond & COND_WORLD_1) && a0.world > 1) on = false;
    if ((P.cond & COND_WORLD_N) && a0.world <= 1) on = false;
    const bool dyn = (P.cond & (COND_FIRST_STEP | COND_TD3_POLICY | COND_TD3_POLICY_OR_STATS)) != 0;
    s_pinfo[ph] = (unsigned char)((on ? 1 : 0) | (dyn ? 2 : 0) | ((P.collective && s_rp.world > 1) ? 4 : 0) | ((P.push && s_rp.world > 1) ? 8 : 0));
  }
  __syncthreads();
  const Ctx& c = *s_ctx;
  int* abort_flag = &c.dyn->abort_flag;
  unsigned gen = s_gen;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int prec = c.hp.gemm_precision;
  const bool fast_rows = fast_rows_ok(c);

  [[maybe_unused]] bool alive = true;
  // staging area of the mma.sync tile on a 1024-byte boundary (its TMA boxes are swizzled on absolute address bits)
  float* tile_smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(dyn_smem) + 1023) & ~uintptr_t(1023));
  TmaState tma_state{s_tma_bar, 0u};
  for (int s = 0; s < a.n_steps && ILSW_ALIVE; ++s) {
    const bool stamp = (s == a.n_steps - 1) && blockIdx.x == 0 && threadIdx.x == 0;
    if (stamp) c.phase_ns[0] = globaltimer_ns();
    if (threadIdx.x < kMaxNets && s_pt[threadIdx.x] >= 0) {
      const int tn = adam_t(a, c.hp, threadIdx.x, s);
      while (s_pt[threadIdx.x] < tn) { s_p1[threadIdx.x] *= s_b1[threadIdx.x]; s_p2[threadIdx.x] *= s_b2[threadIdx.x]; s_pt[threadIdx.x]++; }
      s_coefs[threadIdx.x] = adam_coef_pw(s_ops[s_adam_op[threadIdx.x]].adam, s_p1[threadIdx.x], s_p2[threadIdx.x], 1);
    }
    if (c.tail_op1 && threadIdx.x == 32) {      // which phase carries the next-step row op this step
      int last = -1;
      for (int ph = n_phases - 1; ph >= 0 && last < 0; --ph)
        if (phase_active(s_phases[ph], c.hp, a, s)) last = ph;
      s_last_ph = last;
    }
    __syncthreads();
    for (int ph = 0; ph < n_phases && ILSW_ALIVE; ++ph) {
      const Phase& P = s_phases[ph];
      const unsigned pinfo = s_pinfo[ph];
      if (!(pinfo & 1u) || ((pinfo & 2u) && !phase_active(P, c.hp, a, s))) { if (stamp) c.phase_ns[ph + 1] = c.phase_ns[ph]; continue; }
      const bool exchange = (pinfo & 4u) != 0, pushing = (pinfo & 8u) != 0;
      // exchange sequence number = number of policy updates so far (parity double-buffers the slots)
      unsigned xseq = 0u;
      if (pinfo & 12u) xseq = rp.seq0 + (unsigned)(adam_t(a, c.hp, SLOT_POLICY, s) - a.t0[SLOT_POLICY]);
      if (exchange && P.collective == 1 && !replica_exchange(rp, xseq, bar, gen, abort_flag)) ILSW_DIE()
      if (exchange && P.collective == 2 && !replica_wait_pushes(rp, xseq, abort_flag)) ILSW_DIE()
      if (pushing) {       // this rank's receive slot (parity of this update) on every replica
        if (threadIdx.x < rp.world)
          s_push.peer[threadIdx.x] = rp.recv_peer[threadIdx.x] + ((size_t)(xseq & 1u) * rp.world + rp.rank) * (size_t)rp.nstride;
        if (threadIdx.x == 0) {
          s_push.grad_base = rp.grad; s_push.world = rp.world;
          s_push.mc = rp.recv_mc ? rp.recv_mc + ((size_t)(xseq & 1u) * rp.world + rp.rank) * (size_t)rp.nstride : nullptr;
        }
        __syncthreads();
      }
      const PushCtx* push = pushing ? &s_push : nullptr;
      const int tail_jobs = (c.tail_op1 && ph == s_last_ph) ? s_ops[c.tail_op1 - 1].n_jobs : 0;
      for (int job = blockIdx.x; job < P.total_jobs + tail_jobs && ILSW_ALIVE; job += gridDim.x) {
        int j = job, oi = P.op_begin;
        if (job >= P.total_jobs) { j = job - P.total_jobs; oi = c.tail_op1 - 1; }
        else if (job == (int)blockIdx.x) { const int v = s_job0[ph]; oi = v >> 16; j = v & 0xffff; }
        else { while (j >= s_ops[oi].n_jobs) { j -= s_ops[oi].n_jobs; ++oi; } }
        const Op& o = s_ops[oi];
        if (o.kind == OP_GEMM) {
          const AdamOp* ad = o.gemm.adam ? &s_ops[o.gemm.adam - 1].adam : nullptr;
          const AdamCoef* cf = ad ? &s_coefs[ad->slot] : nullptr;
          const L0FuseOp* fz = o.gemm.a0 ? &s_ops[o.gemm.a0 - 1].l0 : nullptr;
          const int gk = o.gemm.kind;
          if (gk == GK_TILE && prec != 0) {       // the common case first
            gemm_tile_tc<KC, CTAS>(o.gemm, j, tile_smem, prec, (a.profile && blockIdx.x == 0) ? ph : -1, ad, cf, push, fz, tma_state);
          } else if (TC5 && gk == GK_TC5) {
            if constexpr (TC5) {
              if (!tc5::gemm_tile<kTc5BN>(o.gemm, j, tc5_smem, s_tc5, tc5_state)) {
                if (threadIdx.x == 0) atomicExch(abort_flag, 1);     // the other CTAs leave their barrier wait
                alive = false;
              }
            }
          } else if (TC5 && gk == GK_SKINNY_SPLIT) gemm_tile_skinny_split(o.gemm, j, smem);
          else if (gk == GK_SKINNY || gk == GK_SKINNY_SPLIT) gemm_tile_skinny(o.gemm, j, smem, ad, cf, push);
          else if (prec == 0) gemm_tile_device(o.gemm, j, smem, ad, cf, push, fz);
          else gemm_tile_tc_split<KC, CTAS>(o.gemm, j, tile_smem, prec, (a.profile && blockIdx.x == 0) ? ph : -1, tma_state);
        } else if (o.kind == OP_ROW) {
          RowEnv env; env.lane = lane; env.nl = 32; env.warp = warp; env.sm = smem;
          env.prof = (a.profile && blockIdx.x == 0) ? ph : -1;
          const int s_row = s + o.row.arg0;          // arg0 = 1: prefetch job for the next step
          if (s_row < a.n_steps && !(fast_rows && run_row_job_fast(c, a, o.row.kind, o.row.rows, s_row, j + o.row.arg1 / kRowsPerJob, env))) {
            const int row = o.row.arg1 + j * kRowsPerJob + warp;     // arg1: first row of the job range
            if (row < o.row.rows) run_row(c, a, o.row.kind, s_row, row, lane, 32);
          }
        } else if (o.kind == OP_ADAM) {
          adam_job(o.adam, s_coefs[o.adam.slot], j, exchange && o.adam.grad_scale_world, rp, xseq, a.world);
        } else if (o.kind == OP_POLYAK) {
          polyak_job(o.polyak, j);
        } else if (o.kind == OP_SHADOW) {
          const int beg = j * kAdamChunk, end = min(o.shadow.dst.n, beg + kAdamChunk);
          for (int i = beg + threadIdx.x; i < end; i += kThreads) shadow_refresh_elem(o.shadow, i);
        }
      }
      if constexpr (TC5) { if (!alive) break; }
      if (pushing) replica_signal_pushes(rp, xseq);
      if (stamp) c.phase_ns[kMaxPhases + 1 + ph] = globaltimer_ns();
      if (a.profile && s == a.n_steps - 1 && threadIdx.x == 0 && blockIdx.x < kMaxGrid) c.cta_ns[ph * kMaxGrid + blockIdx.x] = globaltimer_ns();
      // descriptor prefetch for this CTA's first tile of the next phase (a wrong guess -- inactive phase -- is harmless)
      if (threadIdx.x == 0) {
        const int nph = ph + 1 < n_phases ? ph + 1 : 0, v = s_job0[nph];
        if (v >= 0) {
          const Op& no = s_ops[v >> 16];
          if (no.kind == OP_GEMM && (no.gemm.tma || no.gemm.tc5)) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(no.gemm.tmapA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(no.gemm.tmapB) : "memory");
          }
        }
      }
      if (!grid_barrier(bar, gridDim.x, gen, abort_flag)) ILSW_DIE()
      if (stamp) c.phase_ns[ph + 1] = globaltimer_ns();
    }
  }
  // the launch's losses go to the host mailbox (every loss row was written before the last grid barrier of its step)
  if (ILSW_ALIVE && blockIdx.x == 0 && a.mail_losses) {
    const int nl = a.n_steps * kLossSlots;
    const float* src = c.loss_log + (size_t)a.loss_log_offset * kLossSlots;
    for (int i = threadIdx.x; i < nl; i += kThreads) a.mail_losses[i] = __ldcg(src + i);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(a.mail_done), "l"(a.mail_seq) : "memory");
  }
  if constexpr (TC5) tc5::teardown<kTc5BN>(s_tc5);
#undef ILSW_ALIVE
#undef ILSW_DIE
}

// ------------------------------------------------------------------------------------------
// A1: sampler-side policy inference (policies.py:245-246): one CTA per env row.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ilsw_policy_act_kernel(MlpPtrs P, int algo, float max_act, float noise_std,
                                                               float noise_clip, const float* obs, int n, int deterministic,
                                                               uint64_t seed, float* act_out) {
  extern __shared__ float sh[];  // [in_dim] + 2*[hid]
  const int row = blockIdx.x;
  if (row >= n) return;
  const int O = P.in_dim, Hd = P.hid, A = P.out_dim;
  float* x = sh; float* h0 = sh + O; float* h1 = h0 + Hd;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int k = threadIdx.x; k < O; k += blockDim.x) x[k] = obs[(size_t)row * O + k];
  __syncthreads();
  for (int o = warp; o < Hd; o += nw) {
    float sacc = 0.f;
    for (int k = lane; k < O; k += 32) sacc += x[k] * P.p[P.oW0 + (size_t)o * O + k];
    sacc = wsum(sacc) + P.p[P.ob0 + o];
    if (lane == 0) h0[o] = sacc > 0.f ? sacc : 0.f;
  }
  __syncthreads();
  for (int o = warp; o < Hd; o += nw) {
    float sacc = 0.f;
    for (int k = lane; k < Hd; k += 32) sacc += h0[k] * P.p[P.oW1 + (size_t)o * Hd + k];
    sacc = wsum(sacc) + P.p[P.ob1 + o];
    if (lane == 0) h1[o] = sacc > 0.f ? sacc : 0.f;
  }
  __syncthreads();
  for (int j = warp; j < A; j += nw) {
    float mu = 0.f, ls = 0.f;
    for (int k = lane; k < Hd; k += 32) {
      mu += h1[k] * P.p[P.oW2 + (size_t)j * Hd + k];
      if (P.heads == 2) ls += h1[k] * P.p[P.oW3 + (size_t)j * Hd + k];
    }
    mu = wsum(mu) + P.p[P.ob2 + j];
    ls = wsum(ls);
    if (lane == 0) {
      float out;
      if (algo == 2) {  // TD3: MlpGaussianNoisePolicy (policies.py:166-188)
        out = max_act * tanhf(mu);
        if (!deterministic) {
          float nz = noise_std * philox_normal(seed, 0u, (uint32_t)row, (uint32_t)j, 7u);
          out += fminf(fmaxf(nz, -noise_clip), noise_clip);
        }
      } else {          // tanh-Gaussian (policies.py:248-283)
        float z = mu;
        if (!deterministic) {
          float lsv = fminf(fmaxf(ls + P.p[P.ob3 + j], -20.f), 2.f);
          z = mu + philox_normal(seed, 0u, (uint32_t)row, (uint32_t)j, 7u) * expf(lsv);
        }
        out = tanhf(z);
      }
      act_out[(size_t)row * A + j] = out;
    }
  }
}

}  // namespace ilsw
