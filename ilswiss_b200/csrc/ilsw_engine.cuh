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

namespace ilsw {

struct BarrierState { unsigned count; unsigned gen; unsigned pad[30]; };

// cross-replica exchange state (NVLink peer memory mapped with CUDA IPC), see ilsw_abi.cu
struct Replica {
  int world, rank;
  int n;                      // policy parameter count
  const float* grad;          // local policy gradient arena
  float* recv_local;          // [2][world][n]  (parity, source rank)
  float* recv_peer[8];        // recv_local of every rank (self included)
  unsigned* flags_local;      // [8] sequence numbers written by the peers
  unsigned* flags_peer[8];
  unsigned seq0;              // exchanges completed before this launch
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
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

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr long long kSpinLimit = 4000000000LL;  // ~2 s at 2 GHz: a hung peer/CTA aborts the launch

// Sense-reversal grid barrier (all CTAs are co-resident: cooperative launch, 1 CTA / SM).
// Returns false if the launch was aborted (timeout); every thread of every CTA then exits.
__device__ __forceinline__ bool grid_barrier(BarrierState* bar, unsigned nblocks, unsigned& gen, int* abort_flag) {
  __shared__ int s_ok;
  __syncthreads();
  if (threadIdx.x == 0) {
    int ok = 1;
    const unsigned target = gen + 1;
    __threadfence();  // release this CTA's writes (also invalidates L1 on sm_100)
    unsigned prev = atomicAdd(&bar->count, 1u);
    if (prev == nblocks - 1) {
      atomicExch(&bar->count, 0u);
      __threadfence();
      atomicExch(&bar->gen, target);
    } else {
      long long t0 = clock64();
      while (ld_acquire_gpu(&bar->gen) != target) {
        long long dt = clock64() - t0;
        if (dt > kSpinLimit) {
          if (*(volatile int*)abort_flag) { ok = 0; break; }
          if (dt > 2 * kSpinLimit) { atomicExch(abort_flag, 1); ok = 0; break; }
        }
      }
    }
    __threadfence();
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

__device__ __forceinline__ void tile_load_A(const GemmOp& o, int m0, int k0, float* r, bool fast) {
  const int tid = threadIdx.x;
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

__device__ __noinline__ void gemm_tile_device(const GemmOp& o, int tile, float* smem) {
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
  tile_load_A(o, m0, 0, ra, fastA);
  tile_load_B(o, n0, 0, rb, fastB);
  tile_store_A(o, smem, ra);
  tile_store_B(o, smem + kStageFloats, rb);
  __syncthreads();
  for (int c = 0; c < nchunks; ++c) {
    float* As = smem + (c & 1) * 2 * kStageFloats;
    float* Bs = As + kStageFloats;
    if (c + 1 < nchunks) {
      tile_load_A(o, m0, (c + 1) * kTK, ra, fastA);
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
  const int Nt = o.N + o.aug_ones;
  const int m = m0 + ty * 4 + kg;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float v = 0.f;
#pragma unroll
    for (int g = 0; g < 4; ++g) v += red[(g * 64 + t) * 17 + kg * 4 + j];
    const int n = n0 + tx * 4 + j;
    if (m < o.M && n < Nt) gemm_epilogue(o, m, n, v);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// replica exchange: every rank PUSHES its policy gradient into every rank's receive slot over
// NVLink peer stores, then publishes a sequence number; the Adam job sums the slots in rank
// order (bit-identical on all ranks) -- no NCCL call, no host round trip, fused into the step.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ bool replica_exchange(const Replica& rp, unsigned seq, BarrierState* bar, unsigned& gen, int* abort_flag) {
  const int parity = (int)(seq & 1u);
  const size_t slot = ((size_t)parity * rp.world + rp.rank) * (size_t)rp.n;
  const int n4 = rp.n >> 2;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += gridDim.x * blockDim.x) {
    float4 v = __ldcg(reinterpret_cast<const float4*>(rp.grad) + i);
    for (int r = 0; r < rp.world; ++r) reinterpret_cast<float4*>(rp.recv_peer[r] + slot)[i] = v;
  }
  for (int i = (n4 << 2) + blockIdx.x * blockDim.x + threadIdx.x; i < rp.n; i += gridDim.x * blockDim.x) {
    float v = __ldcg(rp.grad + i);
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
      if (dt > kSpinLimit && (*(volatile int*)abort_flag || dt > 2 * kSpinLimit)) {
        atomicExch(abort_flag, 1);
        s_ok2 = 0;
        break;
      }
    }
  }
  __syncthreads();
  return s_ok2 != 0;
}

__device__ __forceinline__ float replica_reduced_grad(const Replica& rp, unsigned seq, int i) {
  const size_t base = (size_t)(seq & 1u) * rp.world * (size_t)rp.n;
  float g = 0.f;
  for (int r = 0; r < rp.world; ++r) g += __ldcv(rp.recv_local + base + (size_t)r * rp.n + i);
  return g;
}

// ------------------------------------------------------------------------------------------
// the engine
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1)
ilsw_engine_kernel(const Program* __restrict__ prog, RunArgs a, BarrierState* bar, Replica rp) {
  __shared__ __align__(16) float smem[kGemmSmemFloats];
  __shared__ AdamCoef s_coef;
  __shared__ unsigned s_gen;
  const Ctx& c = prog->ctx;
  int* abort_flag = &c.dyn->abort_flag;
  if (threadIdx.x == 0) s_gen = ld_acquire_gpu(&bar->gen);
  __syncthreads();
  unsigned gen = s_gen;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  for (int s = 0; s < a.n_steps; ++s) {
    const bool stamp = (s == a.n_steps - 1) && blockIdx.x == 0 && threadIdx.x == 0;
    if (stamp) c.phase_ns[0] = globaltimer_ns();
    for (int ph = 0; ph < prog->n_phases; ++ph) {
      const Phase& P = prog->phases[ph];
      if (!phase_active(P, c.hp, a, s)) { if (stamp) c.phase_ns[ph + 1] = c.phase_ns[ph]; continue; }
      const bool exchange = P.collective && rp.world > 1;
      // exchange sequence number = number of policy updates so far (parity double-buffers the slots)
      const unsigned xseq = rp.seq0 + (unsigned)(adam_t(a, c.hp, SLOT_POLICY, s) - a.t0[SLOT_POLICY]);
      if (exchange && !replica_exchange(rp, xseq, bar, gen, abort_flag)) return;
      for (int job = blockIdx.x; job < P.total_jobs; job += gridDim.x) {
        int j = job, oi = P.op_begin;
        while (j >= prog->ops[oi].n_jobs) { j -= prog->ops[oi].n_jobs; ++oi; }
        const Op& o = prog->ops[oi];
        if (o.kind == OP_GEMM) {
          gemm_tile_device(o.gemm, j, smem);
        } else if (o.kind == OP_ROW) {
          const int row = j * kRowsPerJob + warp;
          if (row < o.row.rows) run_row(c, a, o.row.kind, s, row, lane, 32);
        } else if (o.kind == OP_ADAM) {
          __syncthreads();
          if (threadIdx.x == 0) s_coef = adam_coef(o.adam, adam_t(a, c.hp, o.adam.slot, s), a.world);
          __syncthreads();
          const AdamCoef cf = s_coef;
          const int beg = j * kAdamChunk, end = min(o.adam.n, beg + kAdamChunk);
          const bool reduced = exchange && o.adam.grad_scale_world;
          for (int i = beg + threadIdx.x; i < end; i += kThreads) {
            float g = reduced ? replica_reduced_grad(rp, xseq, i) : __ldcg(o.adam.g + i);
            adam_elem_g(o.adam, cf, i, g * cf.gscale);
          }
        } else if (o.kind == OP_POLYAK) {
          const int beg = j * kAdamChunk, end = min(o.polyak.n, beg + kAdamChunk);
          for (int i = beg + threadIdx.x; i < end; i += kThreads) polyak_elem(o.polyak, i);
        }
      }
      if (!grid_barrier(bar, gridDim.x, gen, abort_flag)) return;
      if (stamp) c.phase_ns[ph + 1] = globaltimer_ns();
    }
  }
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
