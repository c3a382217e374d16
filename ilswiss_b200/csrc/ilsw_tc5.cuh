// ilswiss_b200 -- Blackwell-native GEMM tile of the step engine: TMA (cp.async.bulk.tensor, 128-byte swizzle) stages
// the fp32 operand panels, tcgen05.mma (kind::tf32, one issuing thread) accumulates a 128 x BN tile in tensor memory,
// tcgen05.ld brings it back for the fused epilogue (bias / activation / mask / Adam + Polyak).
//
// Used for the phases where the layer really is a dense GEMM (batch >= 512: TD3-Humanoid B = 1024, HER B = 4096); the
// B = 256 programs keep the latency-optimised 32 x 32 tile of ilsw_engine.cuh.
//
// Precision ("3xTF32", same contract as gemm_precision 3 of the mma.sync tile): the tensor core reads the upper 19 bits
// of an fp32 operand word (hi = truncation to TF32); six helper warps derive lo = x - hi from the landed panel into a
// second shared buffer of identical (swizzled) layout, and every k-step issues lo*hi, hi*lo, hi*hi (small terms first)
// into the same fp32 accumulator: products carry ~2^-21 relative error, accumulation is fp32.
//
// Roles inside the 256-thread CTA during the main loop:
//   warp 0 / lane 0 : TMA producer        (waits `empty[s]`, arms `full_raw[s]`, issues the boxes of a K block of 32)
//   warps 2..7      : splitters           (wait `full_raw[s]`, write the lo panels, fence.proxy.async, arrive `full_lo[s]`)
//   warp 1 / lane 0 : MMA issuer          (waits `full_lo[s]`, 4 k-steps x 3 tcgen05.mma, tcgen05.commit -> `empty[s]`)
//   all 8 warps     : epilogue            (wait `acc_full`, tcgen05.ld 32 lanes x 32 columns per warp, transpose through
//                                          shared memory so that global traffic is row-coalesced)
// Operand layouts (both TMA boxes and UMMA descriptors use SWIZZLE_128B, panels are 1024-byte aligned):
//   K-major  (element(r,k) = base[r*ld + k]) : one box {32 k, R rows}  -> [R][32] rows of 128 B, SBO = 1024
//   MN-major (element(r,k) = base[k*ld + r]) : R/32 boxes {32 r, 32 k} -> [32 k][32 r] blocks of 4 KB, LBO = 4096
//            (32-bit MN-major operands use the 128B swizzle with 32-byte atoms on both the TMA and the UMMA side)
// The bias gradient of a weight-gradient GEMM (aug_ones: column sums of the MN-major A operand) is one more pair of
// MMAs per k-step against a constant all-ones B tile (N = 16) into 16 spare TMEM columns.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "ilsw_ops.cuh"

namespace ilsw {
namespace tc5 {

constexpr int kBM = 128;          // tile rows = TMEM lanes (cta_group::1, M = 128)
constexpr int kBK = 32;           // K per stage: one 128-byte swizzle row of fp32
// 3 stages of 48 KB (147 KB + program copy: the 164 KB carve-out step, 92 KB of L1) instead of 4 (196 KB: 228 KB step, 28 KB of
// L1).  The K loop is paced by shared-memory bandwidth at 0.65 us per block, TMA latency is ~1.2 us: three blocks in flight
// cover it, and the engine's spill traffic gets three times the L1 (see ILSW_KC1 in ilsw_engine.cuh).
#ifndef ILSW_TC5_STAGES
#define ILSW_TC5_STAGES 3
#endif
constexpr int kStages = ILSW_TC5_STAGES;
constexpr int kSplitWarps = 6;    // warps 2..7
constexpr int kSplitGroupWarps = 3;   // two groups, alternate K blocks
constexpr int kOnesBytes = 2048;  // 16 rows x 128 B of 1.0f (B operand of the bias-gradient MMAs)

template <int BN> struct Geom {
  static constexpr int kABytes = kBM * kBK * 4;                    // 16 KB
  static constexpr int kBBytes = BN * kBK * 4;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;    // [A raw][A lo][B raw][B lo]
  static constexpr int kSmemBytes = kStages * kStageBytes + kOnesBytes;
  static constexpr int kTmemCols = (BN + 16 <= 128) ? 128 : (BN + 16 <= 256 ? 256 : 512);
  static constexpr int kChunks = (kBM + BN) * 8;                   // 16-byte chunks per stage (A then B)
  static_assert(kChunks % (2 * kSplitGroupWarps * 32) == 0, "splitter threads own an even number of chunks");
};

struct Sync {                    // lives in static shared memory of the engine kernel
  unsigned long long full_raw[kStages], full_lo[kStages], empty[kStages], acc_full;
  uint32_t tmem_base;
  uint32_t pad;
};
struct State { uint32_t it; uint32_t ntile; };
#ifdef ILSW_TC5_DEBUG
__device__ int g_tc5_dbg[4];      // tools/tc5_test.cu: MN-major descriptor overrides {sbo, lbo, layout type} (0 = default)
__device__ unsigned long long g_tc5_t[4][16];   // CTA 0 stamps: [0] producer, [1] mma, [2] splitter warp 2, [3] epilogue
__device__ __forceinline__ unsigned long long dbg_time() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
#define TC5_STAMP(role, i) do { if (blockIdx.x == 0 && (i) < 16) g_tc5_t[role][i] = dbg_time(); } while (0)
#else
#define TC5_STAMP(role, i) do { } while (0)
#endif   // per-thread running counters (k blocks / tiles done by this CTA)

// ---- PTX wrappers --------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* b, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
// bounded wait: a broken pipeline must not wedge the GPU -- after ~1 s the wait gives up and reports failure
__device__ __forceinline__ uint32_t mbar_try_wait(uint32_t a, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(a), "r"(parity) : "memory");
  return ok;
}
__device__ __noinline__ bool mbar_wait_slow(uint32_t a, uint32_t parity) {
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 22); ++spin)
    if (mbar_try_wait(a, parity)) return true;
  return false;
}
__device__ __forceinline__ bool mbar_wait(unsigned long long* b, uint32_t parity) {
  const uint32_t a = smem_u32(b);
  if (mbar_try_wait(a, parity)) return true;
  return mbar_wait_slow(a, parity);
}
// one lane of a converged warp (the compiler keeps the operands of the instructions issued under it in uniform registers;
// a plain `lane == 0` branch makes every tcgen05.mma / TMA issue a lane-serialising R2UR loop: measured 130 cycles per MMA)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred)::"memory");
  return pred != 0;
}
__device__ __forceinline__ void prefetch_tmap(const void* tmap) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(tmap) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, int c0, int c1, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
               ::"r"(dst), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// 1-D bulk copies of the TMA unit (no tensor map): global -> shared with mbarrier completion, shared -> global in a bulk group
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst, uint32_t src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {     // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void mma_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ uint32_t tmem_ld1(uint32_t taddr) {
  uint32_t v;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(taddr) : "memory");
  return v;
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor (sm_100 "version 1").  Offsets in bytes, encoded >> 4.
//   K-major panels : SWIZZLE_128B (layout type 2): rows of 128 B, 16-byte chunks XOR (row % 8); SBO = 1024 (8 rows)
//   MN-major panels: 32-bit operands only exist as SWIZZLE_128B_BASE32B (layout type 1): rows of 128 B (one k each),
//                    32-byte chunks XOR (row % 4); K atoms of 4 rows -> SBO = 512, LBO = stride of the 32-element MN blocks
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout_type) {
  return (uint64_t)((addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo >> 4) << 16) | ((uint64_t)(sbo >> 4) << 32) | (1ull << 46) |
         ((uint64_t)layout_type << 61);
}
// instruction descriptor: D = F32, A = B = TF32, majors, N >> 3 at bit 17, M >> 4 at bit 24
__device__ __forceinline__ uint32_t instr_desc(int a_mn, int b_mn, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

// once per kernel: barriers + tensor memory (call from all threads; `ones` = the constant B tile of the bias MMAs)
template <int BN>
__device__ __forceinline__ void setup(Sync& sy, unsigned char* smem) {
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&sy.full_raw[s], 1); mbar_init(&sy.full_lo[s], kSplitGroupWarps); mbar_init(&sy.empty[s], 1); }
    mbar_init(&sy.acc_full, 1);
    fence_barrier_init();
  }
  float* ones = reinterpret_cast<float*>(smem + kStages * Geom<BN>::kStageBytes);
  for (int i = threadIdx.x; i < kOnesBytes / 4; i += blockDim.x) ones[i] = 1.0f;
  fence_proxy_async();
  if (threadIdx.x < 32) tmem_alloc(&sy.tmem_base, Geom<BN>::kTmemCols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
}
template <int BN>
__device__ __forceinline__ void teardown(Sync& sy) {
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(sy.tmem_base, Geom<BN>::kTmemCols);
}

// One 128 x BN output tile of GEMM `og` (tiles_m = ceil(M/128), tiles_n = ceil(N/BN)).  Returns false when a pipeline
// wait timed out (the engine then aborts the launch).
// Weight-gradient GEMMs (K = batch) can be split along K over `ksplit` tiles: split ks writes its partial tile to
// C + ks * split_stride (bias_out likewise) and the flat Adam job sums the partial gradient arenas in a fixed order.
// The optimiser is never fused into this tile: 128 x 64 outputs on 256 threads would serialise 32 Adam updates per thread
// on the few CTAs that own a weight-gradient tile, the flat Adam phase spreads them over the whole grid.
template <int BN>
__device__ __noinline__ bool gemm_tile(const GemmOp& og, int tile, unsigned char* smem, Sync& sy, State& st) {
  using G = Geom<BN>;
  const GemmOp o = og;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int per_split = o.tiles_m * o.tiles_n;
  const int ks = tile / per_split, t2 = tile - ks * per_split;
  const int tm = t2 / o.tiles_n, tn = t2 - tm * o.tiles_n;
  const int m0 = tm * kBM, n0 = tn * BN;
  const int nkb_all = (o.K + kBK - 1) / kBK, nsplit = o.ksplit > 1 ? o.ksplit : 1;
  const int kb_first = (ks * nkb_all) / nsplit, nkb = ((ks + 1) * nkb_all) / nsplit - kb_first;
  const bool do_aug = o.aug_ones && tn == 0;
  const uint32_t sb = smem_u32(smem);
  const uint32_t it0 = st.it;
  bool ok = true;
  if (warp == 0) {                         // ---- TMA producer (whole warp in uniform control flow, one elected lane issues)
    TC5_STAMP(0, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t it = it0 + kb, s = it % kStages, ph = (it / kStages) & 1u;
      if (!mbar_wait(&sy.empty[s], ph ^ 1u)) { ok = false; break; }
      const uint32_t a_raw = sb + s * G::kStageBytes, b_raw = a_raw + 2 * G::kABytes;
      const int k0 = (kb_first + kb) * kBK;
      if (elect_one()) {
        mbar_arrive_expect_tx(&sy.full_raw[s], G::kABytes + G::kBBytes);
        if (!o.a_mc) tma_load_2d(a_raw, o.tmapA, k0, m0, &sy.full_raw[s]);
        else {
#pragma unroll
          for (int j = 0; j < kBM / 32; ++j) tma_load_2d(a_raw + j * 4096, o.tmapA, m0 + 32 * j, k0, &sy.full_raw[s]);
        }
        if (!o.b_nc) tma_load_2d(b_raw, o.tmapB, k0, n0, &sy.full_raw[s]);
        else {
#pragma unroll
          for (int j = 0; j < BN / 32; ++j) tma_load_2d(b_raw + j * 4096, o.tmapB, n0 + 32 * j, k0, &sy.full_raw[s]);
        }
      }
      __syncwarp();
      TC5_STAMP(0, 1 + kb);
    }
  } else if (warp == 1) {                  // ---- MMA issuer (uniform control flow, one elected lane issues)
    const uint32_t idesc = instr_desc(o.a_mc, o.b_nc, BN), idesc1 = instr_desc(o.a_mc, 0, 16);
    const uint32_t tmem_d = sy.tmem_base;
    const uint32_t a_lbo = o.a_mc ? 4096u : 16u, b_lbo = o.b_nc ? 4096u : 16u;
    const uint32_t a_sbo = o.a_mc ? 512u : 1024u, b_sbo = o.b_nc ? 512u : 1024u;
    const uint32_t a_lt = o.a_mc ? 1u : 2u, b_lt = o.b_nc ? 1u : 2u;
    const uint32_t a_adv = o.a_mc ? 64u : 2u, b_adv = o.b_nc ? 64u : 2u;      // descriptor units (16 B) per k-step of 8
    const uint64_t ones_desc = smem_desc(sb + kStages * G::kStageBytes, 16u, 1024u, 2u);
    TC5_STAMP(1, 0);
    for (int kb = 0; kb < nkb; ++kb) {
      const uint32_t it = it0 + kb, s = it % kStages, ph = (it / kStages) & 1u;
      if (!mbar_wait(&sy.full_raw[s], ph) || !mbar_wait(&sy.full_lo[s], ph)) { ok = false; break; }
      TC5_STAMP(1, 1 + 2 * kb);
      tc_fence_after();
      const uint32_t a_raw = sb + s * G::kStageBytes, a_lo = a_raw + G::kABytes, b_raw = a_raw + 2 * G::kABytes, b_lo = b_raw + G::kBBytes;
      const uint64_t da_hi = smem_desc(a_raw, a_lbo, a_sbo, a_lt), da_lo = smem_desc(a_lo, a_lbo, a_sbo, a_lt);
      const uint64_t db_hi = smem_desc(b_raw, b_lbo, b_sbo, b_lt), db_lo = smem_desc(b_lo, b_lbo, b_sbo, b_lt);
      const int ksteps = min(kBK / 8, (o.K - (kb_first + kb) * kBK + 7) >> 3);
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < kBK / 8; ++kk) {
          if (kk < ksteps) {
            const uint32_t acc = (kb | kk) ? 1u : 0u;
            const uint64_t ao = (uint64_t)(kk * a_adv), bo = (uint64_t)(kk * b_adv);
            mma_tf32(tmem_d, da_lo + ao, db_hi + bo, idesc, acc);
            mma_tf32(tmem_d, da_hi + ao, db_lo + bo, idesc, 1u);
            mma_tf32(tmem_d, da_hi + ao, db_hi + bo, idesc, 1u);
            if (do_aug) {
              mma_tf32(tmem_d + BN, da_lo + ao, ones_desc, idesc1, acc);
              mma_tf32(tmem_d + BN, da_hi + ao, ones_desc, idesc1, 1u);
            }
          }
        }
        mma_commit(&sy.empty[s]);          // frees the stage when these MMAs have read it
        if (kb == nkb - 1) mma_commit(&sy.acc_full);
      }
      __syncwarp();
      TC5_STAMP(1, 2 + 2 * kb);
    }
  } else {                                 // ---- splitters: lo = x - trunc_tf32(x), same (swizzled) offsets as the raw panel
    // two groups of three warps work on alternate K blocks, so that the shared-memory latency of one block overlaps
    // the other's (one group alone: 0.3 us per block, the pace of the whole pipeline)
    const int grp = (warp - 2) / kSplitGroupWarps, gtid = tid - 64 - grp * (kSplitGroupWarps * 32);
    constexpr int kPer = G::kChunks / (kSplitGroupWarps * 32);
    for (int kb = grp; kb < nkb; kb += 2) {
      const uint32_t it = it0 + kb, s = it % kStages, ph = (it / kStages) & 1u;
      if (!mbar_wait(&sy.full_raw[s], ph)) { ok = false; break; }
      if (tid == 64) TC5_STAMP(2, kb);
      const uint32_t a_raw = sb + s * G::kStageBytes;
#pragma unroll
      for (int h = 0; h < 2; ++h) {        // two half batches bound the registers held across the loads
        float4 x[kPer / 2];
        uint32_t src[kPer / 2];
#pragma unroll
        for (int u = 0; u < kPer / 2; ++u) {
          const int c = gtid + (h * (kPer / 2) + u) * (kSplitGroupWarps * 32);
          // chunk c of [A raw | B raw]: A occupies kBM*8 chunks; the lo panel sits kABytes (A) / kBBytes (B) behind
          src[u] = c < kBM * 8 ? a_raw + 16u * c : a_raw + 2 * G::kABytes + 16u * (c - kBM * 8);
          asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x[u].x), "=f"(x[u].y), "=f"(x[u].z), "=f"(x[u].w) : "r"(src[u]) : "memory");
        }
#pragma unroll
        for (int u = 0; u < kPer / 2; ++u) {
          const int c = gtid + (h * (kPer / 2) + u) * (kSplitGroupWarps * 32);
          const uint32_t dst = src[u] + (c < kBM * 8 ? G::kABytes : G::kBBytes);
          const float lx = x[u].x - __uint_as_float(__float_as_uint(x[u].x) & 0xffffe000u);
          const float ly = x[u].y - __uint_as_float(__float_as_uint(x[u].y) & 0xffffe000u);
          const float lz = x[u].z - __uint_as_float(__float_as_uint(x[u].z) & 0xffffe000u);
          const float lw = x[u].w - __uint_as_float(__float_as_uint(x[u].w) & 0xffffe000u);
          asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(dst), "f"(lx), "f"(ly), "f"(lz), "f"(lw) : "memory");
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sy.full_lo[s]);
      if (tid == 64) TC5_STAMP(2, kb + 1);
    }
  }
  // ---- epilogue inputs (bias, mask source) are requested BEFORE the wait for the accumulator: their L2 round trip overlaps
  // the tail of the MMA pipeline instead of following it (thread (rb = tid / 16, cg = tid % 16) owns columns 4 cg .. 4 cg + 3
  // of rows rb, rb + 16, ... -- the mapping of the store loop below)
  constexpr int kER = kBM / 16;            // 8 rows per thread
  const int e_cg = tid & 15, e_rb = tid >> 4, e_n = n0 + 4 * e_cg;
  const bool vec_h = o.mask != ACT_NONE && ((o.ldh & 3) == 0) && ((reinterpret_cast<uintptr_t>(o.H) & 15) == 0) && (e_n + 3 < o.N);
  float b4[4] = {0.f, 0.f, 0.f, 0.f};
  float h[kER][4];
  if (e_n < o.N) {
    if (o.bias) {
#pragma unroll
      for (int j = 0; j < 4; ++j) if (e_n + j < o.N) b4[j] = __ldcg(o.bias + e_n + j);
    }
    if (o.mask != ACT_NONE) {
#pragma unroll
      for (int i = 0; i < kER; ++i) {
        const int m = m0 + i * 16 + e_rb;
        if (m < o.M) {
          const float* hp = o.H + (size_t)m * o.ldh + e_n;
          if (vec_h) { const float4 t4 = __ldcg(reinterpret_cast<const float4*>(hp)); h[i][0] = t4.x; h[i][1] = t4.y; h[i][2] = t4.z; h[i][3] = t4.w; }
          else {
#pragma unroll
            for (int j = 0; j < 4; ++j) h[i][j] = (e_n + j < o.N) ? __ldcg(hp + j) : 0.f;
          }
        }
      }
    }
  }
  // ---- epilogue: accumulator ready when every MMA of this tile has retired
  if (tid == 128) TC5_STAMP(3, 0);
  if (!mbar_wait(&sy.acc_full, st.ntile & 1u)) ok = false;
  if (tid == 128) TC5_STAMP(3, 1);
  ok = __syncthreads_and(ok);              // also: every TMA box has been consumed -> the stage area is free for the staging tile
  if (!ok) return false;
  tc_fence_after();
  // TMEM -> registers (thread = row, 32 columns) -> shared staging tile T[128][BN + 4] (16-byte rows, conflict-free both
  // ways) -> row-coalesced, 128-bit global traffic: thread (rb = tid / 16, cg = tid % 16) owns columns 4 cg .. 4 cg + 3 of
  // rows rb, rb + 16, ...  The epilogue kind is decided ONCE per tile (the per-element generic epilogue of the 32 x 32
  // tiles costs ~40 dependent instructions per element: 0.8 us per 4 rows with 2 warps per scheduler).
  constexpr int kTLd = BN + 4;
  float* T = reinterpret_cast<float*>(smem);
  {
    const int q = warp & 3, half = warp >> 2;
    static_assert(BN == 64, "two 32-column halves per lane quarter");
    uint32_t v[32];
    tmem_ld32(sy.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 32), v);
    tmem_wait_ld();
    float* trow = T + (q * 32 + lane) * kTLd + half * 32;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      *reinterpret_cast<float4*>(trow + 4 * j) = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                                              __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
    if (do_aug && half == 0) {             // bias gradient: column BN of the accumulator (all 16 spare columns are equal)
      const float bsum = __uint_as_float(tmem_ld1(sy.tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)BN));
      tmem_wait_ld();
      const int m = m0 + q * 32 + lane;
      if (m < o.M) o.bias_out[(size_t)ks * o.split_stride + m] = bsum;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 128) TC5_STAMP(3, 2);
  {
    const int cg = e_cg, rb = e_rb, n = e_n;
    float* Cb = o.C + (size_t)ks * o.split_stride;
    const bool vec_c = ((o.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(Cb) & 15) == 0) && (n + 3 < o.N);
    // the epilogue kind is decided ONCE per tile; the three kinds the step programs use are straight-line loops
    const int kind = (o.act == ACT_NONE && o.mask == ACT_NONE) ? 0 : (o.act == ACT_RELU && o.mask == ACT_NONE) ? 1
                   : (o.act == ACT_NONE && o.mask == ACT_RELU) ? 2 : 3;
    if (n < o.N) {
#define ILSW_TC5_STORE_ROWS(EXPR) \
      _Pragma("unroll") \
      for (int i = 0; i < kER; ++i) { \
        const int row = i * 16 + rb, m = m0 + row; \
        if (m < o.M) { \
          const float4 a4 = *reinterpret_cast<const float4*>(T + row * kTLd + 4 * cg); \
          float a[4] = {a4.x + b4[0], a4.y + b4[1], a4.z + b4[2], a4.w + b4[3]}; \
          _Pragma("unroll") \
          for (int j = 0; j < 4; ++j) { EXPR; } \
          float* cp = Cb + (size_t)m * o.ldc + n; \
          if (vec_c) *reinterpret_cast<float4*>(cp) = make_float4(a[0], a[1], a[2], a[3]); \
          else { \
            _Pragma("unroll") \
            for (int j = 0; j < 4; ++j) if (n + j < o.N) cp[j] = a[j]; \
          } \
        } \
      }
      if (kind == 0) { ILSW_TC5_STORE_ROWS((void)0) }
      else if (kind == 1) { ILSW_TC5_STORE_ROWS(a[j] = fmaxf(a[j], 0.f)) }
      else if (kind == 2) { ILSW_TC5_STORE_ROWS(a[j] = h[i][j] > 0.f ? a[j] : 0.f) }
      else {
        ILSW_TC5_STORE_ROWS(
          if (o.act == ACT_RELU) a[j] = fmaxf(a[j], 0.f);
          else if (o.act == ACT_TANH) a[j] = tanhf(a[j]);
          if (o.mask == ACT_RELU) a[j] = h[i][j] > 0.f ? a[j] : 0.f;
          else if (o.mask == ACT_TANH) a[j] *= (1.0f - h[i][j] * h[i][j]))
      }
#undef ILSW_TC5_STORE_ROWS
    }
  }
  if (tid == 128) TC5_STAMP(3, 4);
  st.it = it0 + nkb;
  st.ntile += 1;
  __syncthreads();                         // the staging tile is read before the next job's TMA boxes land in it
  if (tid == 128) TC5_STAMP(3, 5);
  return true;
}

}  // namespace tc5
}  // namespace ilsw
