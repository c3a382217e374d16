// ilswiss_b200 -- latency-optimised row jobs (hidden width <= 256, action dim <= 64).
//
// The step is latency bound: a row job's cost is the number of DEPENDENT L2 round trips
// (~0.4 us each), not its bytes or flops.  These versions of the hot row kernels
//   * stage the small weight matrices they need (policy heads, critic output rows) into shared
//     memory once per CTA job (one cooperative, coalesced L2 read shared by the 8 row-warps),
//   * issue ALL global loads of a row up front into registers (no store sits between two loads, so
//     nothing serialises them), compute, and only then store,
// which turns 10-40 dependent round trips into 2-3.  Semantics are identical to the generic
// per-row kernels in ilsw_ops.cuh (kept as fallback for other shapes and as an independent
// cross-check: the host simulator runs both and compares).
#pragma once
#include "ilsw_ops.cuh"

namespace ilsw {

#ifdef __CUDA_ARCH__
#define ILSW_VL 8        // elements per lane of a <=256-wide row vector (32 lanes)
#else
#define ILSW_VL 256      // host simulator: one lane owns the whole vector
#endif
constexpr int kFastMaxHid = 256;
constexpr int kFastMaxAct = 32;
constexpr int kRowStageFloats = 2 * kFastMaxAct * kFastMaxHid + 4 * kFastMaxHid;   // staged weights (68 KB + scratch <= the 80 KB GEMM staging area of the 2-CTA/SM variant)
constexpr int kRowScratchPerWarp = 4 * kFastMaxAct;                                // per-warp exchange area

// profiling (opt-in): thread 0 of CTA 0 stamps the stages of its last row job / GEMM tile of every phase
#if defined(__CUDACC__)
__device__ unsigned long long g_tile_ns[kMaxPhases][8];
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#endif
#ifdef __CUDA_ARCH__
#define ILSW_RSTAMP(e, i) do { if ((e).prof >= 0 && threadIdx.x == 0) g_tile_ns[(e).prof][i] = globaltimer_ns(); } while (0)
#else
#define ILSW_RSTAMP(e, i) do { } while (0)
#endif

struct RowEnv {
  int lane, nl, warp;
  int prof;             // phase index when CTA 0 profiles this job, else -1
  float* sm;            // host simulator: scratch buffer (>= kRowStageFloats + 8*kRowScratchPerWarp floats); device: the
                        // dynamic shared memory base (staging always starts at offset 0)
};

struct Vec { float v[ILSW_VL]; };

// Staged data lives in the CTA's dynamic SHARED memory on the device.  SPtr is a "shared pointer":
// on the device it is a float OFFSET into the typed extern shared array (plain LDS/STS with an
// immediate/register offset: no generic-address conversion per access, ordinary memory semantics so
// the compiler batches loads but never moves them across __syncthreads); on the host simulator it
// is an ordinary pointer.
#if defined(__CUDACC__)
extern __shared__ __align__(16) float ilsw_dyn_smem_f[];
#endif
struct SPtr {
  const float* p;   // host
  int off;          // device
  ILSW_HD SPtr operator+(size_t i) const { SPtr r; r.p = p ? p + i : p; r.off = off + (int)i; return r; }
};
ILSW_HD SPtr sptr_null() { SPtr r; r.p = nullptr; r.off = 0; return r; }
ILSW_HD float lds(const SPtr& s) {
#ifdef __CUDA_ARCH__
  return ilsw_dyn_smem_f[s.off];
#else
  return *s.p;
#endif
}
ILSW_HD void sts(const SPtr& s, float v) {
#ifdef __CUDA_ARCH__
  ilsw_dyn_smem_f[s.off] = v;
#else
  *const_cast<float*>(s.p) = v;
#endif
}

// Element k of a row vector owned by (lane, slot x).  Device: lane owns two groups of 4 CONSECUTIVE floats
// (k = 128 c + 4 lane + i): every global / shared access of a row is a fully coalesced 128-bit instruction, a
// quarter of the instructions of the element-strided mapping (the step is bound by the instruction count of the
// one warp that owns a row).  Needs width % 4 == 0 (fast_rows_ok).  Host simulator: one lane owns the whole vector.
ILSW_HD int vk(int lane, int x, int nl) {
#ifdef __CUDA_ARCH__
  (void)nl;
  return ((x >> 2) << 7) + (lane << 2) + (x & 3);
#else
  return lane + x * nl;
#endif
}
ILSW_HD void vzero(Vec& x) {
#pragma unroll
  for (int i = 0; i < ILSW_VL; ++i) x.v[i] = 0.f;
}
ILSW_HD void vload(Vec& x, const float* p, int n, int lane, int nl) {     // data written by other CTAs: through L2
#ifdef __CUDA_ARCH__
  (void)nl;
#pragma unroll
  for (int c = 0; c < ILSW_VL / 4; ++c) {
    const int k = (c << 7) + (lane << 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < n) v = __ldcg(reinterpret_cast<const float4*>(p + k));
    x.v[4 * c] = v.x; x.v[4 * c + 1] = v.y; x.v[4 * c + 2] = v.z; x.v[4 * c + 3] = v.w;
  }
#else
  for (int i = 0; i < ILSW_VL; ++i) { const int k = lane + i * nl; x.v[i] = k < n ? p[k] : 0.f; }
#endif
}
ILSW_HD void vstore(float* p, const Vec& x, int n, int lane, int nl) {
#ifdef __CUDA_ARCH__
  (void)nl;
#pragma unroll
  for (int c = 0; c < ILSW_VL / 4; ++c) {
    const int k = (c << 7) + (lane << 2);
    if (k < n) *reinterpret_cast<float4*>(p + k) = make_float4(x.v[4 * c], x.v[4 * c + 1], x.v[4 * c + 2], x.v[4 * c + 3]);
  }
#else
  for (int i = 0; i < ILSW_VL; ++i) { const int k = lane + i * nl; if (k < n) p[k] = x.v[i]; }
#endif
}
ILSW_HD void vload_s(Vec& x, const SPtr& w, int n, int lane, int nl) {    // staged (shared) data; offsets are multiples of 4
#ifdef __CUDA_ARCH__
  (void)nl;
#pragma unroll
  for (int c = 0; c < ILSW_VL / 4; ++c) {
    const int k = (c << 7) + (lane << 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < n) v = *reinterpret_cast<const float4*>(ilsw_dyn_smem_f + w.off + k);
    x.v[4 * c] = v.x; x.v[4 * c + 1] = v.y; x.v[4 * c + 2] = v.z; x.v[4 * c + 3] = v.w;
  }
#else
  for (int i = 0; i < ILSW_VL; ++i) { const int k = lane + i * nl; x.v[i] = k < n ? w.p[k] : 0.f; }
#endif
}
ILSW_HD float vdot_part(const Vec& a, const Vec& b) {      // this lane's share of <a, b> (padding slots are zero)
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILSW_VL; ++i) s += a.v[i] * b.v[i];
  return s;
}
ILSW_HD float vdot_s(const Vec& h, const SPtr& w, int n, int lane, int nl) {   // w staged
  Vec wv;
  vload_s(wv, w, n, lane, nl);
  return wsum(vdot_part(h, wv));
}

// CTA-cooperative staging of up to N small blocks in ONE batch: all global loads of a thread are issued before its
// first shared store (one L2 round trip for the whole set; separate loops would serialise them).  Element i of a
// block comes from src[(i % inner) * s_inner + (i / inner)] (inner == n: a plain copy; the strided form stages
// W0[:, O:O+A] of a critic transposed).  Device: all threads of the CTA must call.  Host: passthrough pointers
// (strided blocks are read in place by w0a()).
struct StageReq { const float* src; int n; int off; int inner; int s_inner; };
ILSW_HD StageReq stage_plain(const float* src, int n, int off) { StageReq r; r.src = src; r.n = n; r.off = off; r.inner = n > 0 ? n : 1; r.s_inner = 1; return r; }
// dst[j*Hd + nn] = W0[nn*K0 + O + j]
ILSW_HD StageReq stage_w0a(const MlpPtrs& Q, int O, int A, int Hd, int off) {
  StageReq r; r.src = Q.p + Q.oW0 + O; r.n = A * Hd; r.off = off; r.inner = Hd; r.s_inner = O + A; return r;
}
template <int N, int STRIDED_MASK = 0>      // bit r of STRIDED_MASK: request r uses the strided source form
ILSW_HD void cta_stage_multi(const RowEnv& e, const StageReq (&rq)[N], SPtr (&out)[N]) {
  (void)e;
#ifdef __CUDA_ARCH__
  constexpr int U = 8, T = kThreads;          // first pass: up to U elements per thread and request, all in flight together
  const int tid = (int)threadIdx.x;
  float v[N][U];
#pragma unroll
  for (int r = 0; r < N; ++r) {
    out[r].p = nullptr; out[r].off = rq[r].off;
    const float* src = rq[r].src;
    const int n = rq[r].n;
    if (((STRIDED_MASK >> r) & 1) && rq[r].inner <= T && n <= U * rq[r].inner) {
      // common strided shape (W0[:, O:O+A] of a critic: inner = hidden <= 256 threads, A <= 8 blocks): thread -> nn = tid,
      // slot u -> j = u.  No index division / carry chain, and a thread's A loads are adjacent words of ONE row of W0
      // (same 32-byte sector) instead of A different rows: the generic form below cost ~4 us of the 9 us staging stage of
      // the policy-backward row job (profiles/r2_phase_profile.txt, ROW(k29))
      const int inner = rq[r].inner, si = rq[r].s_inner;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u * inner < n) v[r][u] = (tid < inner) ? __ldcg(src + (size_t)tid * si + u) : 0.f;
    } else if ((STRIDED_MASK >> r) & 1) {
      const int inner = rq[r].inner, si = rq[r].s_inner;
      int j = tid / inner, nn = tid - j * inner;
      const int dj = T / inner, dn = T - dj * inner;
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (u * T < n) {          // warp-uniform: dead batches cost one branch
          v[r][u] = (tid + u * T < n) ? __ldcg(src + (size_t)nn * si + j) : 0.f;
          nn += dn; j += dj;
          if (nn >= inner) { nn -= inner; ++j; }
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u * T < n) v[r][u] = (tid + u * T < n) ? __ldcg(src + tid + u * T) : 0.f;
    }
  }
#pragma unroll
  for (int r = 0; r < N; ++r) {
    if (((STRIDED_MASK >> r) & 1) && rq[r].inner <= T && rq[r].n <= U * rq[r].inner) {
      const int inner = rq[r].inner;
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u * inner < rq[r].n && tid < inner) ilsw_dyn_smem_f[rq[r].off + u * inner + tid] = v[r][u];
    } else {
#pragma unroll
      for (int u = 0; u < U; ++u)
        if (u * T < rq[r].n && tid + u * T < rq[r].n) ilsw_dyn_smem_f[rq[r].off + tid + u * T] = v[r][u];
    }
  }
  // blocks larger than U*T elements (wide action spaces): the rest in plain loops
#pragma unroll
  for (int r = 0; r < N; ++r) {
    if (((STRIDED_MASK >> r) & 1) && rq[r].inner <= T && rq[r].n <= U * rq[r].inner) continue;
    for (int i = U * T + tid; i < rq[r].n; i += T) {
      const int inner = rq[r].inner;
      const size_t idx = ((STRIDED_MASK >> r) & 1) ? (size_t)(i % inner) * rq[r].s_inner + i / inner : (size_t)i;
      ilsw_dyn_smem_f[rq[r].off + i] = __ldcg(rq[r].src + idx);
    }
  }
#else
  for (int r = 0; r < N; ++r) { out[r].p = ((STRIDED_MASK >> r) & 1) ? nullptr : rq[r].src; out[r].off = 0; }
#endif
}
// single-block convenience form
ILSW_HD SPtr cta_stage(const RowEnv& e, const float* src, int n, int off) {
  StageReq rq[1] = {stage_plain(src, n, off)};
  SPtr out[1];
  cta_stage_multi(e, rq, out);
  return out[0];
}
// this lane's share of sum_n e0[n] * W0[n, O+j]  (staged transposed on the device, read in place on the host)
ILSW_HD float vdot_w0a_part(const Vec& e0, const SPtr& staged, const MlpPtrs& Q, int O, int A, int Hd, int j, int lane, int nl) {
#ifdef __CUDA_ARCH__
  (void)Q; (void)O; (void)A;
  Vec wv;
  vload_s(wv, staged + (size_t)j * Hd, Hd, lane, nl);
  return vdot_part(e0, wv);
#else
  (void)staged;
  float s = 0.f;
  for (int x = 0; x < ILSW_VL; ++x) { const int n = lane + x * nl; if (n < Hd) s += e0.v[x] * Q.p[Q.oW0 + (size_t)n * (O + A) + O + j]; }
  return s;
#endif
}
ILSW_HD void cta_sync() {
#ifdef __CUDA_ARCH__
  __syncthreads();
#endif
}
ILSW_HD SPtr warp_scratch(const RowEnv& e) {
  SPtr r;
  r.p = e.sm + kRowStageFloats + e.warp * kRowScratchPerWarp;   // host: scratch buffer; device: unused
  r.off = kRowStageFloats + e.warp * kRowScratchPerWarp;
  return r;
}

ILSW_HD bool fast_rows_ok(const Ctx& c) {
  if (c.hp.her) return false;     // HER-TD3 rows exist in the generic form only (ilsw_ops.cuh)
  return c.s.Hd <= kFastMaxHid && (c.s.Hd & 3) == 0 && c.s.A <= kFastMaxAct && 4 * c.s.A * c.s.Hd <= kRowStageFloats &&
         (!c.hp.has_disc || (c.d.Hd <= kFastMaxHid && (c.d.Hd & 3) == 0));
}

// per-lane slots of a per-action (A-wide) vector: the device handles A <= 32 with one slot per lane, the host
// simulator's single lane owns all of them
#ifdef __CUDA_ARCH__
#define ILSW_AL 1
#else
#define ILSW_AL kFastMaxAct
#endif

// ---------------------------------------------------------------------------------------------
// SAC policy heads (N2) on 2B rows
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_heads(const Ctx& c, int job, const RowEnv& e) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int A = S.A, Hd = S.Hd, B = S.B, O = S.O, lane = e.lane, nl = e.nl;
  ILSW_RSTAMP(e, 0);
  const int r = job * kRowsPerJob + e.warp;
  Vec h;
  float ep[ILSW_AL];
  if (r < 2 * B) {
    vload(h, S.h1p + (size_t)r * Hd, Hd, lane, nl);
    for (int j = lane, q = 0; j < A; j += nl, ++q) ep[q] = ldg(S.eps + (size_t)r * A + j);
  }
  const StageReq rq[4] = {stage_plain(P.p + P.oW2, A * Hd, 0), stage_plain(P.p + P.oW3, A * Hd, A * Hd),
                          stage_plain(P.p + P.ob2, A, 2 * A * Hd), stage_plain(P.p + P.ob3, A, 2 * A * Hd + A)};
  SPtr sp[4];
  cta_stage_multi(e, rq, sp);
  const SPtr Wm = sp[0], Ws = sp[1], bm = sp[2], bs = sp[3];
  cta_sync();
  ILSW_RSTAMP(e, 1);
  if (r < 2 * B) {
    const SPtr sc = warp_scratch(e);
#pragma unroll 1
    for (int j = 0; j < A; ++j) {
      float mu = vdot_s(h, Wm + (size_t)j * Hd, Hd, lane, nl) + lds(bm + j);
      float lr = vdot_s(h, Ws + (size_t)j * Hd, Hd, lane, nl) + lds(bs + j);
      if (lane == 0) { sts(sc + j, mu); sts(sc + A + j, lr); }
    }
    wsync();
    ILSW_RSTAMP(e, 2);
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int j = lane, q = 0; j < A; j += nl, ++q) {
      const float mu = lds(sc + j), lraw = lds(sc + A + j);
      const float ls = fminf(fmaxf(lraw, -20.0f), 2.0f);
      const float sig = expf(ls), cov = expf(2.0f * ls);
      const float z = ep[q] * sig + mu;
      const float t = tanhf(z);
      const float d = mu - z;
      s1 += d * d / cov; s2 += ls; s3 += logf(1.0f - t * t + 1e-6f);
      S.mean[(size_t)r * A + j] = mu; S.lraw[(size_t)r * A + j] = lraw;
      S.lstd[(size_t)r * A + j] = ls; S.act[(size_t)r * A + j] = t;
      if (r < B) S.Xna[(size_t)r * S.ld_oa + O + j] = t;
      else S.Xon[(size_t)(r - B) * S.ld_oa + O + j] = t;
    }
    s1 = wsum(s1); s2 = wsum(s2); s3 = wsum(s3);
    if (lane == 0) {
      float lp = -0.5f * s1;
      lp -= (s2 + 0.5f * 1.8378770664093453f);
      lp -= s3;
      S.logpi[r] = lp;
    }
  }
  ILSW_RSTAMP(e, 3);
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// critic target + output-layer backward (SAC and TD3)
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_critic_target(const Ctx& c, int job, const RowEnv& e, bool use_entropy, float loss_grad_factor) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, lane = e.lane, nl = e.nl;
  ILSW_RSTAMP(e, 0);
  const int b = job * kRowsPerJob + e.warp;
  Vec ht[2], hq[2];
  float rew = 0.f, term = 0.f, lp = 0.f, alpha = 0.f, tb[2] = {0.f, 0.f}, qb[2] = {0.f, 0.f};
  if (b < S.B) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      vload(ht[i], S.h1t[i] + (size_t)b * Hd, Hd, lane, nl);
      vload(hq[i], S.h1q[i] + (size_t)b * Hd, Hd, lane, nl);
      tb[i] = ldg(c.tqf[i].p + c.tqf[i].ob2);
      qb[i] = ldg(c.qf[i].p + c.qf[i].ob2);
    }
    rew = ldg(S.rew + b); term = ldg(S.term + b);
    if (use_entropy) { lp = ldg(S.logpi + b); alpha = ldg(&c.dyn->alpha); }
  }
  const StageReq rq[4] = {stage_plain(c.tqf[0].p + c.tqf[0].oW2, Hd, 0), stage_plain(c.tqf[1].p + c.tqf[1].oW2, Hd, Hd),
                          stage_plain(c.qf[0].p + c.qf[0].oW2, Hd, 2 * Hd), stage_plain(c.qf[1].p + c.qf[1].oW2, Hd, 3 * Hd)};
  SPtr sp[4];
  cta_stage_multi(e, rq, sp);
  cta_sync();
  ILSW_RSTAMP(e, 1);
  if (b < S.B) {
    const float tq0 = vdot_s(ht[0], sp[0], Hd, lane, nl) + tb[0];
    const float tq1 = vdot_s(ht[1], sp[1], Hd, lane, nl) + tb[1];
    const float tmin = fminf(tq0, tq1);
    const float rs = c.hp.reward_scale * rew;
    const float inner = use_entropy ? tmin - alpha * lp : tmin;
    const float y = rs + (1.0f - term) * c.hp.discount * inner;
    const float invB = 1.0f / (float)S.B;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      Vec w;
      vload_s(w, sp[2 + i], Hd, lane, nl);
      const float q = wsum(vdot_part(hq[i], w)) + qb[i];
      const float diff = q - y;
      const float dq = loss_grad_factor * diff * invB;
      if (lane == 0) { S.qp[i][b] = q; S.dq[i][b] = dq; S.lossterm[i][b] = diff * diff; }
#pragma unroll
      for (int x = 0; x < ILSW_VL; ++x) w.v[x] = hq[i].v[x] > 0.f ? dq * w.v[x] : 0.f;
      vstore(S.d1q[i] + (size_t)b * Hd, w, Hd, lane, nl);
    }
    if (lane == 0) { S.tq[0][b] = tq0; S.tq[1][b] = tq1; S.y[b] = y; }
  }
  ILSW_RSTAMP(e, 2);
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// SAC policy-loss terms + output-layer backward through min(Q1,Q2)(obs, a~)
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_ploss(const Ctx& c, int job, const RowEnv& e) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, A = S.A, B = S.B, lane = e.lane, nl = e.nl;
  ILSW_RSTAMP(e, 0);
  const int b = job * kRowsPerJob + e.warp;
  Vec h[2];
  float qb[2] = {0.f, 0.f}, lp = 0.f, alpha = 0.f, smu = 0.f, sls = 0.f;
  if (b < B) {
#pragma unroll
    for (int i = 0; i < 2; ++i) { vload(h[i], S.h1n[i] + (size_t)b * Hd, Hd, lane, nl); qb[i] = ldg(c.qf[i].p + c.qf[i].ob2); }
    lp = ldg(S.logpi + B + b); alpha = ldg(&c.dyn->alpha);
    for (int j = lane; j < A; j += nl) {
      const float mu = ldg(S.mean + (size_t)(B + b) * A + j), ls = ldg(S.lstd + (size_t)(B + b) * A + j);
      smu += mu * mu; sls += ls * ls;
    }
  }
  const StageReq rq[2] = {stage_plain(c.qf[0].p + c.qf[0].oW2, Hd, 0), stage_plain(c.qf[1].p + c.qf[1].oW2, Hd, Hd)};
  SPtr qw[2];
  cta_stage_multi(e, rq, qw);
  cta_sync();
  ILSW_RSTAMP(e, 1);
  if (b < B) {
    Vec w[2];
    float q[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) { vload_s(w[i], qw[i], Hd, lane, nl); q[i] = wsum(vdot_part(h[i], w[i])) + qb[i]; }
    smu = wsum(smu); sls = wsum(sls);
    const float invB = 1.0f / (float)B;
    const float w0 = q[0] < q[1] ? 1.f : (q[0] == q[1] ? 0.5f : 0.f);
    const float wq[2] = {w0, 1.f - w0};
    if (lane == 0) {
      S.qn[0][b] = q[0]; S.qn[1][b] = q[1];
      S.plterm[b] = alpha * lp - fminf(q[0], q[1]);
      S.regmu[b] = smu; S.regls[b] = sls;
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const float dq = -invB * wq[i];
#pragma unroll
      for (int x = 0; x < ILSW_VL; ++x) w[i].v[x] = h[i].v[x] > 0.f ? dq * w[i].v[x] : 0.f;
      vstore(S.e1[i] + (size_t)b * Hd, w[i], Hd, lane, nl);
    }
  }
  ILSW_RSTAMP(e, 2);
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// backward through the tanh-Gaussian head
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_pibwd(const Ctx& c, int job, const RowEnv& e, bool with_da) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int Hd = S.Hd, A = S.A, B = S.B, lane = e.lane, nl = e.nl;
  ILSW_RSTAMP(e, 0);
  const int b = job * kRowsPerJob + e.warp;
  const int r = B + b;
  // every global load of the row first (one L2 round trip), then the staging batch
  Vec h, e0[2];
  float alpha = 0.f, lpi = 0.f;
  float t_[ILSW_AL], mu_[ILSW_AL], ls_[ILSW_AL], lr_[ILSW_AL], ep_[ILSW_AL], ga_[ILSW_AL];
  if (b < B) {
    if (with_da) {
#pragma unroll
      for (int i = 0; i < 2; ++i) vload(e0[i], S.e0[i] + (size_t)b * Hd, Hd, lane, nl);
    }
    vload(h, S.h1p + (size_t)r * Hd, Hd, lane, nl);
    alpha = ldg(&c.dyn->alpha);
    lpi = ldg(S.logpi + r);
    for (int j = lane, q = 0; j < A; j += nl, ++q) {
      t_[q] = ldg(S.act + (size_t)r * A + j);
      mu_[q] = ldg(S.mean + (size_t)r * A + j); ls_[q] = ldg(S.lstd + (size_t)r * A + j); lr_[q] = ldg(S.lraw + (size_t)r * A + j);
      ep_[q] = ldg(S.eps + (size_t)r * A + j);
      ga_[q] = with_da ? 0.f : ldg(S.dA[0] + (size_t)b * A + j) + ldg(S.dA[1] + (size_t)b * A + j);
    }
  }
  const StageReq rq[4] = {stage_plain(P.p + P.oW2, A * Hd, 0), stage_plain(P.p + P.oW3, A * Hd, A * Hd),
                          with_da ? stage_w0a(c.qf[0], S.O, A, Hd, 2 * A * Hd) : stage_plain(nullptr, 0, 0),
                          with_da ? stage_w0a(c.qf[1], S.O, A, Hd, 3 * A * Hd) : stage_plain(nullptr, 0, 0)};
  SPtr sp[4];
  cta_stage_multi<4, 0xC>(e, rq, sp);
  const SPtr Wm = sp[0], Ws = sp[1];
  const SPtr sc = warp_scratch(e);
  cta_sync();
  ILSW_RSTAMP(e, 1);
  if (b < B) {
    if (with_da) {
      // dA[b,j] = sum_n e0_i[b,n] * W0_i[n, O+j], both critics (formerly a GEMM phase of its own)
#pragma unroll 1
      for (int j = 0; j < A; ++j) {
        const float acc = wsum(vdot_w0a_part(e0[0], sp[2], c.qf[0], S.O, A, Hd, j, lane, nl) +
                               vdot_w0a_part(e0[1], sp[3], c.qf[1], S.O, A, Hd, j, lane, nl));
        if (lane == 0) sts(sc + 2 * A + j, acc);
      }
      wsync();
    }
    ILSW_RSTAMP(e, 2);
    const float invB = 1.0f / (float)B, invBA = 1.0f / (float)(B * A);
    for (int j = lane, q = 0; j < A; j += nl, ++q) {
      const float gA = with_da ? lds(sc + 2 * A + j) : ga_[q];
      const float t = t_[q], mu = mu_[q], ls = ls_[q], lr = lr_[q];
      const float om = 1.0f - t * t;
      const float J = 2.0f * t * om / (om + 1e-6f);
      const float dz = gA * om + alpha * invB * J;
      const float dmu = dz + 2.0f * c.hp.mean_reg * mu * invBA;
      const float dl = dz * ep_[q] * expf(ls) - alpha * invB + 2.0f * c.hp.std_reg * ls * invBA;
      const float dlr = (lr >= -20.0f && lr <= 2.0f) ? dl : 0.f;
      sts(sc + j, dmu); sts(sc + A + j, dlr);
      S.dmean[(size_t)b * A + j] = dmu; S.dlraw[(size_t)b * A + j] = dlr;
    }
    wsync();
    ILSW_RSTAMP(e, 3);
    Vec acc;
    vzero(acc);
#pragma unroll 1
    for (int j = 0; j < A; ++j) {
      const float dm = lds(sc + j), dl = lds(sc + A + j);
      Vec wm, ws;
      vload_s(wm, Wm + (size_t)j * Hd, Hd, lane, nl);
      vload_s(ws, Ws + (size_t)j * Hd, Hd, lane, nl);
#pragma unroll
      for (int x = 0; x < ILSW_VL; ++x) acc.v[x] += dm * wm.v[x] + dl * ws.v[x];
    }
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) acc.v[x] = h.v[x] > 0.f ? acc.v[x] : 0.f;
    vstore(S.d1p + (size_t)b * Hd, acc, Hd, lane, nl);
    if (lane == 0) S.aterm[b] = lpi + c.hp.target_entropy;
  }
  ILSW_RSTAMP(e, 4);
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// TD3 heads / losses
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_td3_head(const Ctx& c, int job, const RowEnv& e, bool target) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = target ? c.tpolicy : c.policy;
  const int A = S.A, Hd = S.Hd, O = S.O, lane = e.lane, nl = e.nl;
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  float nz_[ILSW_AL];
  if (b < S.B) {
    vload(h, (target ? S.h1tp : S.h1p) + (size_t)b * Hd, Hd, lane, nl);
    if (target) for (int j = lane, q = 0; j < A; j += nl, ++q) nz_[q] = ldg(S.noise + (size_t)b * A + j);
  }
  const StageReq rq[2] = {stage_plain(P.p + P.oW2, A * Hd, 0), stage_plain(P.p + P.ob2, A, A * Hd)};
  SPtr sp[2];
  cta_stage_multi(e, rq, sp);
  const SPtr W = sp[0], bb = sp[1];
  cta_sync();
  if (b < S.B) {
    const SPtr sc = warp_scratch(e);
#pragma unroll 1
    for (int j = 0; j < A; ++j) {
      const float pre = vdot_s(h, W + (size_t)j * Hd, Hd, lane, nl) + lds(bb + j);
      if (lane == 0) sts(sc + j, pre);
    }
    wsync();
    for (int j = lane, q = 0; j < A; j += nl, ++q) {
      const float t = tanhf(lds(sc + j));
      if (target) {
        float nz = c.hp.policy_noise * nz_[q];
        nz = fminf(fmaxf(nz, -c.hp.noise_clip), c.hp.noise_clip);
        S.Xna[(size_t)b * S.ld_oa + O + j] = c.hp.max_act * t + nz;
      } else {
        S.act[(size_t)b * A + j] = t;
        S.Xon[(size_t)b * S.ld_oa + O + j] = c.hp.max_act * t;
      }
    }
  }
  cta_sync();
}

ILSW_HDN void job_td3_ploss(const Ctx& c, int job, const RowEnv& e) {
  const SacBufs& S = c.s;
  const MlpPtrs& Q = c.qf[0];
  const int Hd = S.Hd, lane = e.lane, nl = e.nl;
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  float qb = 0.f;
  if (b < S.B) { vload(h, S.h1n[0] + (size_t)b * Hd, Hd, lane, nl); qb = ldg(Q.p + Q.ob2); }
  const SPtr qw = cta_stage(e, Q.p + Q.oW2, Hd, 0);
  cta_sync();
  if (b < S.B) {
    Vec w;
    vload_s(w, qw, Hd, lane, nl);
    const float q = wsum(vdot_part(h, w)) + qb;
    const float dq = -1.0f / (float)S.B;
    if (lane == 0) { S.qn[0][b] = q; S.plterm[b] = -q; }
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) w.v[x] = h.v[x] > 0.f ? dq * w.v[x] : 0.f;
    vstore(S.e1[0] + (size_t)b * Hd, w, Hd, lane, nl);
  }
  cta_sync();
}

ILSW_HDN void job_td3_pibwd(const Ctx& c, int job, const RowEnv& e, bool with_da) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int Hd = S.Hd, A = S.A, lane = e.lane, nl = e.nl;
  const int b = job * kRowsPerJob + e.warp;
  Vec h, e0;
  float t_[ILSW_AL], ga_[ILSW_AL];
  if (b < S.B) {
    if (with_da) vload(e0, S.e0[0] + (size_t)b * Hd, Hd, lane, nl);
    vload(h, S.h1p + (size_t)b * Hd, Hd, lane, nl);
    for (int j = lane, q = 0; j < A; j += nl, ++q) {
      t_[q] = ldg(S.act + (size_t)b * A + j);
      ga_[q] = with_da ? 0.f : ldg(S.dA[0] + (size_t)b * A + j);
    }
  }
  const StageReq rq[2] = {stage_plain(P.p + P.oW2, A * Hd, 0),
                          with_da ? stage_w0a(c.qf[0], S.O, A, Hd, A * Hd) : stage_plain(nullptr, 0, 0)};
  SPtr sp[2];
  cta_stage_multi<2, 0x2>(e, rq, sp);
  const SPtr W = sp[0];
  const SPtr sc = warp_scratch(e);
  cta_sync();
  if (b < S.B) {
    if (with_da) {
#pragma unroll 1
      for (int j = 0; j < A; ++j) {
        const float acc = wsum(vdot_w0a_part(e0, sp[1], c.qf[0], S.O, A, Hd, j, lane, nl));
        if (lane == 0) sts(sc + A + j, acc);
      }
      wsync();
    }
    for (int j = lane, q = 0; j < A; j += nl, ++q) {
      const float t = t_[q];
      const float gA = with_da ? lds(sc + A + j) : ga_[q];
      const float dm = gA * c.hp.max_act * (1.0f - t * t);
      sts(sc + j, dm);
      S.dmean[(size_t)b * A + j] = dm;
    }
    wsync();
    Vec acc;
    vzero(acc);
#pragma unroll 1
    for (int j = 0; j < A; ++j) {
      const float dm = lds(sc + j);
      Vec w;
      vload_s(w, W + (size_t)j * Hd, Hd, lane, nl);
#pragma unroll
      for (int x = 0; x < ILSW_VL; ++x) acc.v[x] += dm * w.v[x];
    }
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) acc.v[x] = h.v[x] > 0.f ? acc.v[x] : 0.f;
    vstore(S.d1p + (size_t)b * Hd, acc, Hd, lane, nl);
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// discriminator rows
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_disc_head(const Ctx& c, int job, const RowEnv& e, int rows) {
  const DiscBufs& Dd = c.d;
  const MlpPtrs& N = c.disc;
  const int B = Dd.B, Hd = Dd.Hd, lane = e.lane, nl = e.nl, act = c.hp.disc_act;
  const int r = job * kRowsPerJob + e.warp;
  Vec h2;
  float b3 = 0.f;
  if (r < rows) { vload(h2, Dd.h2 + (size_t)r * Hd, Hd, lane, nl); b3 = ldg(N.p + N.ob2); }
  const SPtr w3 = cta_stage(e, N.p + N.oW2, Hd, 0);
  cta_sync();
  if (r < rows) {
    Vec w;
    vload_s(w, w3, Hd, lane, nl);
    const float y = wsum(vdot_part(h2, w)) + b3;
    const float cm = c.hp.disc_clamp;
    const float pass = (y >= -cm && y <= cm) ? 1.f : 0.f;
    if (r < 2 * B) {
      const float x = fminf(fmaxf(y, -cm), cm);
      const float t = r < B ? 1.f : 0.f;
      const float mv = fmaxf(-x, 0.f);
      const float ce = (1.0f - t) * x + mv + logf(expf(-mv) + expf(-x - mv));
      const float sg = 1.0f / (1.0f + expf(-x));
      const float dl = (sg - t) / (float)(2 * B) * pass;
      if (lane == 0) {
        Dd.y[r] = x; Dd.dlogit[r] = dl; Dd.ceterm[r] = ce;
        Dd.accterm[r] = ((x > 0.f ? 1.f : 0.f) == t) ? 1.f : 0.f;
      }
#pragma unroll
      for (int xx = 0; xx < ILSW_VL; ++xx) w.v[xx] = dl * w.v[xx] * disc_dact(act, h2.v[xx]);
      vstore(Dd.d2 + (size_t)r * Hd, w, Hd, lane, nl);
    } else {
      const int b = r - 2 * B;
      if (lane == 0) Dd.cmask[b] = pass;
#pragma unroll
      for (int xx = 0; xx < ILSW_VL; ++xx) w.v[xx] = pass * w.v[xx] * disc_dact(act, h2.v[xx]);
      vstore(Dd.dl2 + (size_t)b * Hd, w, Hd, lane, nl);
    }
  }
  cta_sync();
}

ILSW_HDN void job_disc_ew(const Ctx& c, int kind, int job, const RowEnv& e) {
  const DiscBufs& Dd = c.d;
  const int Hd = Dd.Hd, B = Dd.B, lane = e.lane, nl = e.nl, act = c.hp.disc_act;
  const int b = job * kRowsPerJob + e.warp;
  if (b >= B) return;
  const size_t ro = (size_t)b * Hd, ri = (size_t)(2 * B + b) * Hd;
  if (kind == ROW_DISC_EW1) {
    Vec h1, db, u1;
    vload(h1, Dd.h1 + ri, Hd, lane, nl); vload(db, Dd.db1 + ro, Hd, lane, nl); vload(u1, Dd.u1 + ro, Hd, lane, nl);
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) { h1.v[x] = db.v[x] * disc_dact(act, h1.v[x]); u1.v[x] = db.v[x] * u1.v[x]; }
    vstore(Dd.ub1 + ro, h1, Hd, lane, nl);
    vstore(Dd.sb1 + ro, u1, Hd, lane, nl);
  } else if (kind == ROW_DISC_EW2) {
    Vec h2, db, w3;
    vload(h2, Dd.h2 + ri, Hd, lane, nl); vload(db, Dd.db2 + ro, Hd, lane, nl); vload(w3, c.disc.p + c.disc.oW2, Hd, lane, nl);
    const float cmk = ldg(Dd.cmask + b);
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const float s2 = disc_dact(act, h2.v[x]);
      const float sb2 = db.v[x] * (cmk * w3.v[x]);
      w3.v[x] = (disc_curv(act, h2.v[x]) * sb2) * s2;
      db.v[x] = db.v[x] * s2;
    }
    vstore(Dd.t3 + ro, db, Hd, lane, nl);
    vstore(Dd.zb2 + ro, w3, Hd, lane, nl);
  } else {
    Vec h1, hb, sb;
    vload(h1, Dd.h1 + ri, Hd, lane, nl); vload(hb, Dd.hb1 + ro, Hd, lane, nl); vload(sb, Dd.sb1 + ro, Hd, lane, nl);
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) hb.v[x] = (hb.v[x] + disc_curv(act, h1.v[x]) * sb.v[x]) * disc_dact(act, h1.v[x]);
    vstore(Dd.zb1 + ro, hb, Hd, lane, nl);
  }
}

ILSW_HDN void job_disc_reward(const Ctx& c, int job, const RowEnv& e) {
  const DiscBufs& Dd = c.d;
  const MlpPtrs& N = c.disc;
  const int Hd = Dd.Hd, lane = e.lane, nl = e.nl;
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  float b3 = 0.f;
  if (b < Dd.B) { vload(h, Dd.rh2 + (size_t)b * Hd, Hd, lane, nl); b3 = ldg(N.p + N.ob2); }
  const SPtr w3 = cta_stage(e, N.p + N.oW2, Hd, 0);
  cta_sync();
  if (b < Dd.B) {
    const float y = vdot_s(h, w3, Hd, lane, nl) + b3;
    const float x = fminf(fmaxf(y, -c.hp.disc_clamp), c.hp.disc_clamp);
    float r;
    switch (c.hp.disc_mode) {
      case 0: r = x; break;
      case 1: r = (x > 20.f) ? x : log1pf(expf(x)); break;
      case 2: r = (-x > 20.f) ? x : -log1pf(expf(-x)); break;
      default: r = expf(x) * (-1.0f * x); break;
    }
    if (c.hp.clip_max_on) r = fminf(r, c.hp.rew_clip_max);
    if (c.hp.clip_min_on) r = fmaxf(r, c.hp.rew_clip_min);
    if (lane == 0) { c.s.rew[b] = r; Dd.rewraw[b] = r; }
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// SAC-alpha step epilogue: the 9 batch means are reduced by the 8 warps in parallel (one L2 round trip each
// instead of 9 in sequence on one warp), then lane 0 of warp 0 does the scalar work (loss log, float64 Adam of
// log alpha).  Same arithmetic as row_sac_final (ilsw_ops.cuh), which remains the generic fallback.
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_final(const Ctx& c, const RunArgs& a, int s, const RowEnv& e) {
  const SacBufs& S = c.s;
  const int B = S.B, A = S.A, lane = e.lane, nl = e.nl;
#ifdef __CUDA_ARCH__
  const int nw = 8;
#else
  const int nw = 1;                 // host simulator: warps run one after the other -> warp 0 does everything
  if (e.warp != 0) return;
#endif
  const SPtr sc = warp_scratch(e) + (size_t)0;
  SPtr res; res.p = e.sm + kRowStageFloats; res.off = kRowStageFloats;      // 9 means (scratch of warp 0)
  (void)sc;
  for (int v = e.warp; v < 9; v += nw) {
    const float* x = v == 0 ? S.lossterm[0] : v == 1 ? S.lossterm[1] : v == 2 ? S.plterm : v == 3 ? S.regmu : v == 4 ? S.regls
                   : v == 5 ? S.aterm : v == 6 ? S.qp[0] : v == 7 ? S.logpi + B : S.y;
    float acc = 0.f;
    for (int i0 = lane; i0 < B; i0 += 8 * nl) {
      float t[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { const int i = i0 + u * nl; t[u] = i < B ? ldg(x + i) : 0.f; }
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += t[u];
    }
    acc = wsum(acc) / (float)B;
    if (lane == 0) sts(res + v, acc);
  }
  cta_sync();
  if (e.warp == 0) {
    DynState* d = c.dyn;
    if (lane == 0) {
      const float l1 = 0.5f * lds(res + 0), l2 = 0.5f * lds(res + 1);
      const float rm = lds(res + 3) / (float)A, rl = lds(res + 4) / (float)A;
      const float pl = lds(res + 2) + (c.hp.mean_reg * rm + c.hp.std_reg * rl);
      const float am = lds(res + 5);
      float alpha_loss = 0.f;
      float* L = c.loss_log + (size_t)(a.loss_log_offset + s) * kLossSlots;
      L[L_QF1] = l1; L[L_QF2] = l2; L[L_POLICY] = pl;
      L[L_Q1_MEAN] = lds(res + 6); L[L_LOGPI_MEAN] = lds(res + 7); L[L_QT_MEAN] = lds(res + 8);
      if (c.hp.train_alpha) {
        alpha_loss = -((float)d->log_alpha * am);
        double g = (double)(-am);
        double b1 = c.hp.beta1, b2 = c.hp.beta2;
        d->alpha_t += 1;
        double w = 1.0 - b1;
        d->alpha_m = (w < 0.5) ? d->alpha_m + w * (g - d->alpha_m) : g - (g - d->alpha_m) * (1.0 - w);
        d->alpha_v = d->alpha_v * b2 + (1.0 - b2) * g * g;
        d->alpha_p1 *= b1; d->alpha_p2 *= b2;
        double bc1 = 1.0 - d->alpha_p1, bc2 = 1.0 - d->alpha_p2;
        double denom = sqrt(d->alpha_v) / sqrt(bc2) + c.hp.adam_eps;
        d->log_alpha = d->log_alpha + (-(c.hp.alpha_lr / bc1)) * d->alpha_m / denom;
        d->alpha = (float)exp(d->log_alpha);
      }
      L[L_ALPHA_LOSS] = alpha_loss;
      L[L_ALPHA] = d->alpha;
    }
    if (s == a.stats_step) snapshot_sac(c, lane, nl);
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// job-level dispatcher: returns true if a fast implementation handled the job
// (called by ALL threads of the CTA on the device; by each emulated warp on the host)
// ---------------------------------------------------------------------------------------------
ILSW_HD bool run_row_job_fast(const Ctx& c, const RunArgs& a, int kind, int rows, int s, int job, const RowEnv& e) {
  switch (kind) {
    case ROW_SAC_FINAL: job_sac_final(c, a, s, e); return true;
    case ROW_SAC_HEADS: job_sac_heads(c, job, e); return true;
    case ROW_SAC_TARGET: job_critic_target(c, job, e, true, 1.0f); return true;
    case ROW_TD3_TARGET: job_critic_target(c, job, e, false, 2.0f); return true;
    case ROW_SAC_PLOSS: job_sac_ploss(c, job, e); return true;
    case ROW_SAC_PIBWD: job_sac_pibwd(c, job, e, false); return true;
    case ROW_SAC_PIBWD_DA: job_sac_pibwd(c, job, e, true); return true;
    case ROW_TD3_THEAD: job_td3_head(c, job, e, true); return true;
    case ROW_TD3_PHEAD: job_td3_head(c, job, e, false); return true;
    case ROW_TD3_PLOSS: job_td3_ploss(c, job, e); return true;
    case ROW_TD3_PIBWD: job_td3_pibwd(c, job, e, false); return true;
    case ROW_TD3_PIBWD_DA: job_td3_pibwd(c, job, e, true); return true;
    case ROW_DISC_HEAD: job_disc_head(c, job, e, rows); return true;
    case ROW_DISC_EW1: case ROW_DISC_EW2: case ROW_DISC_EW3: job_disc_ew(c, kind, job, e); return true;
    case ROW_DISC_REWARD: job_disc_reward(c, job, e); return true;
    default: return false;
  }
}

}  // namespace ilsw
