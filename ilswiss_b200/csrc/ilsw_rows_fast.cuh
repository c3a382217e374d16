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

struct RowEnv {
  int lane, nl, warp;
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

ILSW_HD void vload(Vec& x, const float* p, int n, int lane, int nl) {
#pragma unroll
  for (int i = 0; i < ILSW_VL; ++i) { const int k = lane + i * nl; x.v[i] = k < n ? ldg(p + k) : 0.f; }
}
ILSW_HD void vload_plain(Vec& x, const float* p, int n, int lane, int nl) {   // staged (shared) or immutable data
#pragma unroll
  for (int i = 0; i < ILSW_VL; ++i) { const int k = lane + i * nl; x.v[i] = k < n ? p[k] : 0.f; }
}
ILSW_HD float vdot_s(const Vec& h, const SPtr& w, int n, int lane, int nl) {   // w staged/plain
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILSW_VL; ++i) { const int k = lane + i * nl; if (k < n) s += h.v[i] * lds(w + k); }
  return wsum(s);
}

// CTA-cooperative staging (device) / passthrough (host).  All threads of the CTA must call.
ILSW_HD SPtr cta_stage(const RowEnv& e, const float* src, int n, int off) {
  SPtr r;
#ifdef __CUDA_ARCH__
  (void)e;
  r.p = nullptr; r.off = off;
  // batches of 4 loads per thread are issued together (a store between two loads would
  // serialise them into separate L2 round trips), then stored
  for (int base = 0; base < n; base += 4 * (int)blockDim.x) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = base + u * (int)blockDim.x + (int)threadIdx.x; v[u] = i < n ? __ldcg(src + i) : 0.f; }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = base + u * (int)blockDim.x + (int)threadIdx.x; if (i < n) ilsw_dyn_smem_f[off + i] = v[u]; }
  }
#else
  (void)e; (void)n; (void)off;
  r.p = src; r.off = 0;
#endif
  return r;
}
// stages W0[:, O:O+A] of a critic TRANSPOSED: dst[j*Hd + n] = W0[n*K0 + O + j]
ILSW_HD SPtr cta_stage_w0a(const RowEnv& e, const MlpPtrs& Q, int O, int A, int Hd, int off) {
  SPtr r;
#ifdef __CUDA_ARCH__
  (void)e;
  r.p = nullptr; r.off = off;
  const int K0 = O + A, n = A * Hd;
  for (int base = 0; base < n; base += 4 * (int)blockDim.x) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = base + u * (int)blockDim.x + (int)threadIdx.x;      // i = j*Hd + nn
      v[u] = i < n ? __ldcg(Q.p + Q.oW0 + (size_t)(i % Hd) * K0 + O + i / Hd) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) { const int i = base + u * (int)blockDim.x + (int)threadIdx.x; if (i < n) ilsw_dyn_smem_f[off + i] = v[u]; }
  }
#else
  (void)e; (void)Q; (void)O; (void)A; (void)Hd; (void)off;
  r.p = nullptr; r.off = 0;   // host: read the strided weights directly (see w0a())
#endif
  return r;
}
ILSW_HD float w0a(const SPtr& staged, const MlpPtrs& Q, int O, int A, int Hd, int j, int n) {
#ifdef __CUDA_ARCH__
  (void)Q; (void)O; (void)A;
  return ilsw_dyn_smem_f[staged.off + j * Hd + n];
#else
  (void)staged; (void)Hd;
  return Q.p[Q.oW0 + (size_t)n * (O + A) + O + j];
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
  return c.s.Hd <= kFastMaxHid && c.s.A <= kFastMaxAct && 4 * c.s.A * c.s.Hd <= kRowStageFloats &&
         (!c.hp.has_disc || c.d.Hd <= kFastMaxHid);
}

// ---------------------------------------------------------------------------------------------
// SAC policy heads (N2) on 2B rows
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_heads(const Ctx& c, int job, const RowEnv& e) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int A = S.A, Hd = S.Hd, B = S.B, O = S.O, lane = e.lane, nl = e.nl;
  const SPtr Wm = cta_stage(e, P.p + P.oW2, A * Hd, 0);
  const SPtr Ws = cta_stage(e, P.p + P.oW3, A * Hd, A * Hd);
  const SPtr bm = cta_stage(e, P.p + P.ob2, A, 2 * A * Hd);
  const SPtr bs = cta_stage(e, P.p + P.ob3, A, 2 * A * Hd + A);
  const int r = job * kRowsPerJob + e.warp;
  Vec h;
  if (r < 2 * B) vload(h, S.h1p + (size_t)r * Hd, Hd, lane, nl);
  cta_sync();
  if (r < 2 * B) {
    const SPtr sc = warp_scratch(e);
#pragma unroll 1
    for (int j = 0; j < A; ++j) {
      float mu = vdot_s(h, Wm + (size_t)j * Hd, Hd, lane, nl) + lds(bm + j);
      float lr = vdot_s(h, Ws + (size_t)j * Hd, Hd, lane, nl) + lds(bs + j);
      if (lane == 0) { sts(sc + j, mu); sts(sc + A + j, lr); }
    }
    wsync();
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int j = lane; j < A; j += nl) {
      const float mu = lds(sc + j), lraw = lds(sc + A + j);
      const float ls = fminf(fmaxf(lraw, -20.0f), 2.0f);
      const float sig = expf(ls), cov = expf(2.0f * ls);
      const float z = ldg(S.eps + (size_t)r * A + j) * sig + mu;
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
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// critic target + output-layer backward (SAC and TD3)
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_critic_target(const Ctx& c, int job, const RowEnv& e, bool use_entropy, float loss_grad_factor) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, lane = e.lane, nl = e.nl;
  SPtr tw[2], qw[2];
  for (int i = 0; i < 2; ++i) {
    tw[i] = cta_stage(e, c.tqf[i].p + c.tqf[i].oW2, Hd, i * Hd);
    qw[i] = cta_stage(e, c.qf[i].p + c.qf[i].oW2, Hd, (2 + i) * Hd);
  }
  const int b = job * kRowsPerJob + e.warp;
  Vec ht[2], hq[2];
  float rew = 0.f, term = 0.f, lp = 0.f, alpha = 0.f, tb[2] = {0.f, 0.f}, qb[2] = {0.f, 0.f};
  if (b < S.B) {
    for (int i = 0; i < 2; ++i) {
      vload(ht[i], S.h1t[i] + (size_t)b * Hd, Hd, lane, nl);
      vload(hq[i], S.h1q[i] + (size_t)b * Hd, Hd, lane, nl);
      tb[i] = ldg(c.tqf[i].p + c.tqf[i].ob2);
      qb[i] = ldg(c.qf[i].p + c.qf[i].ob2);
    }
    rew = ldg(S.rew + b); term = ldg(S.term + b);
    if (use_entropy) { lp = ldg(S.logpi + b); alpha = ldg(&c.dyn->alpha); }
  }
  cta_sync();
  if (b < S.B) {
    const float tq0 = vdot_s(ht[0], tw[0], Hd, lane, nl) + tb[0];
    const float tq1 = vdot_s(ht[1], tw[1], Hd, lane, nl) + tb[1];
    const float tmin = fminf(tq0, tq1);
    const float rs = c.hp.reward_scale * rew;
    const float inner = use_entropy ? tmin - alpha * lp : tmin;
    const float y = rs + (1.0f - term) * c.hp.discount * inner;
    const float invB = 1.0f / (float)S.B;
    for (int i = 0; i < 2; ++i) {
      const float q = vdot_s(hq[i], qw[i], Hd, lane, nl) + qb[i];
      const float diff = q - y;
      const float dq = loss_grad_factor * diff * invB;
      if (lane == 0) { S.qp[i][b] = q; S.dq[i][b] = dq; S.lossterm[i][b] = diff * diff; }
      float* d1 = S.d1q[i] + (size_t)b * Hd;
#pragma unroll
      for (int x = 0; x < ILSW_VL; ++x) {
        const int k = lane + x * nl;
        if (k < Hd) d1[k] = hq[i].v[x] > 0.f ? dq * lds(qw[i] + k) : 0.f;
      }
    }
    if (lane == 0) { S.tq[0][b] = tq0; S.tq[1][b] = tq1; S.y[b] = y; }
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// SAC policy-loss terms + output-layer backward through min(Q1,Q2)(obs, a~)
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_ploss(const Ctx& c, int job, const RowEnv& e) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, A = S.A, B = S.B, lane = e.lane, nl = e.nl;
  SPtr qw[2];
  for (int i = 0; i < 2; ++i) qw[i] = cta_stage(e, c.qf[i].p + c.qf[i].oW2, Hd, i * Hd);
  const int b = job * kRowsPerJob + e.warp;
  Vec h[2];
  float qb[2] = {0.f, 0.f}, lp = 0.f, alpha = 0.f, smu = 0.f, sls = 0.f;
  if (b < B) {
    for (int i = 0; i < 2; ++i) { vload(h[i], S.h1n[i] + (size_t)b * Hd, Hd, lane, nl); qb[i] = ldg(c.qf[i].p + c.qf[i].ob2); }
    lp = ldg(S.logpi + B + b); alpha = ldg(&c.dyn->alpha);
    for (int j = lane; j < A; j += nl) {
      const float mu = ldg(S.mean + (size_t)(B + b) * A + j), ls = ldg(S.lstd + (size_t)(B + b) * A + j);
      smu += mu * mu; sls += ls * ls;
    }
  }
  cta_sync();
  if (b < B) {
    float q[2];
    for (int i = 0; i < 2; ++i) q[i] = vdot_s(h[i], qw[i], Hd, lane, nl) + qb[i];
    smu = wsum(smu); sls = wsum(sls);
    const float invB = 1.0f / (float)B;
    const float w0 = q[0] < q[1] ? 1.f : (q[0] == q[1] ? 0.5f : 0.f);
    const float wq[2] = {w0, 1.f - w0};
    if (lane == 0) {
      S.qn[0][b] = q[0]; S.qn[1][b] = q[1];
      S.plterm[b] = alpha * lp - fminf(q[0], q[1]);
      S.regmu[b] = smu; S.regls[b] = sls;
    }
    for (int i = 0; i < 2; ++i) {
      const float dq = -invB * wq[i];
      float* e1 = S.e1[i] + (size_t)b * Hd;
#pragma unroll
      for (int x = 0; x < ILSW_VL; ++x) {
        const int k = lane + x * nl;
        if (k < Hd) e1[k] = h[i].v[x] > 0.f ? dq * lds(qw[i] + k) : 0.f;
      }
    }
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// backward through the tanh-Gaussian head
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_sac_pibwd(const Ctx& c, int job, const RowEnv& e, bool with_da) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int Hd = S.Hd, A = S.A, B = S.B, lane = e.lane, nl = e.nl;
  const SPtr Wm = cta_stage(e, P.p + P.oW2, A * Hd, 0);
  const SPtr Ws = cta_stage(e, P.p + P.oW3, A * Hd, A * Hd);
  SPtr Wa[2] = {sptr_null(), sptr_null()};
  if (with_da) for (int i = 0; i < 2; ++i) Wa[i] = cta_stage_w0a(e, c.qf[i], S.O, A, Hd, (2 + i) * A * Hd);
  const int b = job * kRowsPerJob + e.warp;
  const int r = B + b;
  Vec h;
  const SPtr sc = warp_scratch(e);
  float lpi = 0.f;
  if (with_da) {
    // dA[b,j] = sum_n e0_i[b,n] * W0_i[n, O+j], both critics (formerly a GEMM phase of its own)
    Vec e0[2];
    if (b < B) for (int i = 0; i < 2; ++i) vload(e0[i], S.e0[i] + (size_t)b * Hd, Hd, lane, nl);
    cta_sync();
    if (b < B) {
#pragma unroll 1
      for (int j = 0; j < A; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int x = 0; x < ILSW_VL; ++x) {
          const int n = lane + x * nl;
          if (n < Hd) acc += e0[0].v[x] * w0a(Wa[0], c.qf[0], S.O, A, Hd, j, n) + e0[1].v[x] * w0a(Wa[1], c.qf[1], S.O, A, Hd, j, n);
        }
        acc = wsum(acc);
        if (lane == 0) sts(sc + 2 * A + j, acc);
      }
    }
    wsync();
  }
  if (b < B) {
    vload(h, S.h1p + (size_t)r * Hd, Hd, lane, nl);
    const float alpha = ldg(&c.dyn->alpha);
    lpi = ldg(S.logpi + r);
    const float invB = 1.0f / (float)B, invBA = 1.0f / (float)(B * A);
    for (int j = lane; j < A; j += nl) {
      const float gA = with_da ? lds(sc + 2 * A + j) : ldg(S.dA[0] + (size_t)b * A + j) + ldg(S.dA[1] + (size_t)b * A + j);
      const float t = ldg(S.act + (size_t)r * A + j);
      const float mu = ldg(S.mean + (size_t)r * A + j), ls = ldg(S.lstd + (size_t)r * A + j), lr = ldg(S.lraw + (size_t)r * A + j);
      const float ep = ldg(S.eps + (size_t)r * A + j);
      const float om = 1.0f - t * t;
      const float J = 2.0f * t * om / (om + 1e-6f);
      const float dz = gA * om + alpha * invB * J;
      const float dmu = dz + 2.0f * c.hp.mean_reg * mu * invBA;
      const float dl = dz * ep * expf(ls) - alpha * invB + 2.0f * c.hp.std_reg * ls * invBA;
      const float dlr = (lr >= -20.0f && lr <= 2.0f) ? dl : 0.f;
      sts(sc + j, dmu); sts(sc + A + j, dlr);
    }
  }
  cta_sync();   // staged heads visible; also orders the scratch writes inside each warp
  if (b < B) {
    for (int j = lane; j < A; j += nl) { S.dmean[(size_t)b * A + j] = lds(sc + j); S.dlraw[(size_t)b * A + j] = lds(sc + A + j); }
    float* d1 = S.d1p + (size_t)b * Hd;
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const int k = lane + x * nl;
      if (k < Hd) {
        float acc = 0.f;
#pragma unroll 1
        for (int j = 0; j < A; ++j) acc += lds(sc + j) * lds(Wm + (size_t)j * Hd + k) + lds(sc + A + j) * lds(Ws + (size_t)j * Hd + k);
        d1[k] = h.v[x] > 0.f ? acc : 0.f;
      }
    }
    if (lane == 0) S.aterm[b] = lpi + c.hp.target_entropy;
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// TD3 heads / losses
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_td3_head(const Ctx& c, int job, const RowEnv& e, bool target) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = target ? c.tpolicy : c.policy;
  const int A = S.A, Hd = S.Hd, O = S.O, lane = e.lane, nl = e.nl;
  const SPtr W = cta_stage(e, P.p + P.oW2, A * Hd, 0);
  const SPtr bb = cta_stage(e, P.p + P.ob2, A, A * Hd);
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  if (b < S.B) vload(h, (target ? S.h1tp : S.h1p) + (size_t)b * Hd, Hd, lane, nl);
  cta_sync();
  if (b < S.B) {
    const SPtr sc = warp_scratch(e);
    for (int j = 0; j < A; ++j) {
      const float pre = vdot_s(h, W + (size_t)j * Hd, Hd, lane, nl) + lds(bb + j);
      if (lane == 0) sts(sc + j, pre);
    }
    wsync();
    for (int j = lane; j < A; j += nl) {
      const float t = tanhf(lds(sc + j));
      if (target) {
        float nz = c.hp.policy_noise * ldg(S.noise + (size_t)b * A + j);
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
  const SPtr qw = cta_stage(e, Q.p + Q.oW2, Hd, 0);
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  float qb = 0.f;
  if (b < S.B) { vload(h, S.h1n[0] + (size_t)b * Hd, Hd, lane, nl); qb = ldg(Q.p + Q.ob2); }
  cta_sync();
  if (b < S.B) {
    const float q = vdot_s(h, qw, Hd, lane, nl) + qb;
    const float dq = -1.0f / (float)S.B;
    if (lane == 0) { S.qn[0][b] = q; S.plterm[b] = -q; }
    float* e1 = S.e1[0] + (size_t)b * Hd;
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const int k = lane + x * nl;
      if (k < Hd) e1[k] = h.v[x] > 0.f ? dq * lds(qw + k) : 0.f;
    }
  }
  cta_sync();
}

ILSW_HDN void job_td3_pibwd(const Ctx& c, int job, const RowEnv& e, bool with_da) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int Hd = S.Hd, A = S.A, lane = e.lane, nl = e.nl;
  const SPtr W = cta_stage(e, P.p + P.oW2, A * Hd, 0);
  SPtr Wa = sptr_null();
  if (with_da) Wa = cta_stage_w0a(e, c.qf[0], S.O, A, Hd, A * Hd);
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  const SPtr sc = warp_scratch(e);
  if (with_da) {
    Vec e0;
    if (b < S.B) vload(e0, S.e0[0] + (size_t)b * Hd, Hd, lane, nl);
    cta_sync();
    if (b < S.B) {
#pragma unroll 1
      for (int j = 0; j < A; ++j) {
        float acc = 0.f;
#pragma unroll
        for (int x = 0; x < ILSW_VL; ++x) {
          const int n = lane + x * nl;
          if (n < Hd) acc += e0.v[x] * w0a(Wa, c.qf[0], S.O, A, Hd, j, n);
        }
        acc = wsum(acc);
        if (lane == 0) sts(sc + A + j, acc);
      }
    }
    wsync();
  }
  if (b < S.B) {
    vload(h, S.h1p + (size_t)b * Hd, Hd, lane, nl);
    for (int j = lane; j < A; j += nl) {
      const float t = ldg(S.act + (size_t)b * A + j);
      const float gA = with_da ? lds(sc + A + j) : ldg(S.dA[0] + (size_t)b * A + j);
      sts(sc + j, gA * c.hp.max_act * (1.0f - t * t));
    }
  }
  cta_sync();
  if (b < S.B) {
    for (int j = lane; j < A; j += nl) S.dmean[(size_t)b * A + j] = lds(sc + j);
    float* d1 = S.d1p + (size_t)b * Hd;
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const int k = lane + x * nl;
      if (k < Hd) {
        float acc = 0.f;
#pragma unroll 1
        for (int j = 0; j < A; ++j) acc += lds(sc + j) * lds(W + (size_t)j * Hd + k);
        d1[k] = h.v[x] > 0.f ? acc : 0.f;
      }
    }
  }
  cta_sync();
}

// ---------------------------------------------------------------------------------------------
// discriminator rows
// ---------------------------------------------------------------------------------------------
ILSW_HDN void job_disc_head(const Ctx& c, int job, const RowEnv& e, int rows) {
  const DiscBufs& Dd = c.d;
  const MlpPtrs& N = c.disc;
  const int B = Dd.B, Hd = Dd.Hd, lane = e.lane, nl = e.nl;
  const SPtr w3 = cta_stage(e, N.p + N.oW2, Hd, 0);
  const int r = job * kRowsPerJob + e.warp;
  Vec h2;
  float b3 = 0.f;
  if (r < rows) { vload(h2, Dd.h2 + (size_t)r * Hd, Hd, lane, nl); b3 = ldg(N.p + N.ob2); }
  cta_sync();
  if (r < rows) {
    const float y = vdot_s(h2, w3, Hd, lane, nl) + b3;
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
      float* d2 = Dd.d2 + (size_t)r * Hd;
#pragma unroll
      for (int xx = 0; xx < ILSW_VL; ++xx) {
        const int k = lane + xx * nl;
        if (k < Hd) d2[k] = dl * lds(w3 + k) * (1.0f - h2.v[xx] * h2.v[xx]);
      }
    } else {
      const int b = r - 2 * B;
      if (lane == 0) Dd.cmask[b] = pass;
      float* dl2 = Dd.dl2 + (size_t)b * Hd;
#pragma unroll
      for (int xx = 0; xx < ILSW_VL; ++xx) {
        const int k = lane + xx * nl;
        if (k < Hd) dl2[k] = pass * lds(w3 + k) * (1.0f - h2.v[xx] * h2.v[xx]);
      }
    }
  }
  cta_sync();
}

ILSW_HDN void job_disc_ew(const Ctx& c, int kind, int job, const RowEnv& e) {
  const DiscBufs& Dd = c.d;
  const int Hd = Dd.Hd, B = Dd.B, lane = e.lane, nl = e.nl;
  const int b = job * kRowsPerJob + e.warp;
  if (b >= B) return;
  const size_t ro = (size_t)b * Hd, ri = (size_t)(2 * B + b) * Hd;
  if (kind == ROW_DISC_EW1) {
    Vec h1, db, u1;
    vload(h1, Dd.h1 + ri, Hd, lane, nl); vload(db, Dd.db1 + ro, Hd, lane, nl); vload(u1, Dd.u1 + ro, Hd, lane, nl);
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const int k = lane + x * nl;
      if (k < Hd) { Dd.ub1[ro + k] = db.v[x] * (1.0f - h1.v[x] * h1.v[x]); Dd.sb1[ro + k] = db.v[x] * u1.v[x]; }
    }
  } else if (kind == ROW_DISC_EW2) {
    Vec h2, db, w3;
    vload(h2, Dd.h2 + ri, Hd, lane, nl); vload(db, Dd.db2 + ro, Hd, lane, nl); vload(w3, c.disc.p + c.disc.oW2, Hd, lane, nl);
    const float cmk = ldg(Dd.cmask + b);
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const int k = lane + x * nl;
      if (k < Hd) {
        const float s2 = 1.0f - h2.v[x] * h2.v[x];
        Dd.t3[ro + k] = db.v[x] * s2;
        const float sb2 = db.v[x] * (cmk * w3.v[x]);
        Dd.zb2[ro + k] = (-2.0f * h2.v[x] * sb2) * s2;
      }
    }
  } else {
    Vec h1, hb, sb;
    vload(h1, Dd.h1 + ri, Hd, lane, nl); vload(hb, Dd.hb1 + ro, Hd, lane, nl); vload(sb, Dd.sb1 + ro, Hd, lane, nl);
#pragma unroll
    for (int x = 0; x < ILSW_VL; ++x) {
      const int k = lane + x * nl;
      if (k < Hd) Dd.zb1[ro + k] = (hb.v[x] - 2.0f * h1.v[x] * sb.v[x]) * (1.0f - h1.v[x] * h1.v[x]);
    }
  }
}

ILSW_HDN void job_disc_reward(const Ctx& c, int job, const RowEnv& e) {
  const DiscBufs& Dd = c.d;
  const MlpPtrs& N = c.disc;
  const int Hd = Dd.Hd, lane = e.lane, nl = e.nl;
  const SPtr w3 = cta_stage(e, N.p + N.oW2, Hd, 0);
  const int b = job * kRowsPerJob + e.warp;
  Vec h;
  float b3 = 0.f;
  if (b < Dd.B) { vload(h, Dd.rh2 + (size_t)b * Hd, Hd, lane, nl); b3 = ldg(N.p + N.ob2); }
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
