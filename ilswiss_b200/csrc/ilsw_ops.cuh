// ilswiss_b200 -- op semantics of the fused step engine.
//
// Every function here is __host__ __device__: the sm_100a engine kernel (ilsw_engine.cu)
// runs them with one warp per batch row / one CTA per GEMM tile, and the TEST-ONLY host
// simulator (tests/hostsim) runs the very same code with a 1-lane "warp" so that program
// wiring and the hand-derived backward formulas can be checked against the oracle without a
// GPU.  The host simulator is never linked into the product library.
//
// Reference semantics restated (file:line relative to the reference checkout):
//   tanh-Gaussian policy + log-prob ... rlkit/torch/common/policies.py:248-307,
//                                       rlkit/torch/common/distributions.py:43-50,74-95
//   SAC-alpha step ..................... rlkit/torch/algorithms/sac/sac_alpha.py:78-181
//   TD3 step ........................... rlkit/torch/algorithms/td3/td3.py:72-124
//   discriminator step / reward ........ rlkit/torch/algorithms/adv_irl/adv_irl.py:133-314
//   Adam ............................... torch.optim.Adam as configured at sac_alpha.py:65-76
//   Polyak ............................. rlkit/torch/utils/pytorch_util.py:10-12
#pragma once
#include <math.h>
#include "ilsw_types.h"

namespace ilsw {

// ------------------------------------------------------------------------------------------
// lane helpers (device: 32 lanes; host simulator: 1 lane)
// ------------------------------------------------------------------------------------------
ILSW_HD float wsum(float v) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
#endif
  return v;
}
ILSW_HD float wmaxf(float v) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
  return v;
}
ILSW_HD float wminf(float v) {
#ifdef __CUDA_ARCH__
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
#endif
  return v;
}
ILSW_HD void wsync() {
#ifdef __CUDA_ARCH__
  __syncwarp();
#endif
}
// Loads of data produced by OTHER CTAs in an earlier phase go through L2 (ld.global.cg) so a
// stale L1 line can never be observed, independent of what the grid barrier's fence does to L1.
ILSW_HD float ldg(const float* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);
#else
  return *p;
#endif
}
ILSW_HD int ldgi(const int* p) {
#ifdef __CUDA_ARCH__
  return __ldcg(p);
#else
  return *p;
#endif
}

ILSW_HDN float wdot(const float* h, const float* w, int n, int lane, int nl) {
  float s = 0.f;
  for (int k = lane; k < n; k += nl) s += ldg(h + k) * ldg(w + k);
  return wsum(s);
}

// ------------------------------------------------------------------------------------------
// Philox4x32-10 counter RNG (speed mode; parity mode injects indices / eps instead)
// ------------------------------------------------------------------------------------------
struct U4 { uint32_t x, y, z, w; };
ILSW_HD uint32_t mulhi32(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
ILSW_HDN U4 philox(uint64_t seed, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  U4 c = {c0, c1, c2, c3};
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    uint32_t hi1 = mulhi32(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    U4 n = {hi1 ^ c.y ^ k0, lo1, hi0 ^ c.w ^ k1, lo0};
    c = n;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return c;
}
ILSW_HD float u01(uint32_t u) { return ((float)(u >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // (0,1)
ILSW_HDN float philox_normal(uint64_t seed, uint32_t gstep, uint32_t row, uint32_t col, uint32_t stream) {
  U4 r = philox(seed, gstep, row, col, stream);
  float u1 = u01(r.x), u2 = u01(r.y);
  return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}
ILSW_HD int philox_index(uint64_t seed, uint32_t gstep, uint32_t row, uint32_t stream, int size) {
  U4 r = philox(seed, gstep, row, 0u, stream);
  return (int)mulhi32(r.x, (uint32_t)size);  // uniform with replacement in [0,size)
}
ILSW_HD float philox_uniform(uint64_t seed, uint32_t gstep, uint32_t row, uint32_t stream) {
  U4 r = philox(seed, gstep, row, 1u, stream);
  return (float)(r.x >> 8) * (1.0f / 16777216.0f);  // [0,1) like torch.rand
}

// ------------------------------------------------------------------------------------------
// operand rounding of the tensor-core modes (the host simulator applies the same rounding so
// the precision of a mode can be evaluated against the oracle without a GPU)
ILSW_HD float round_tf32(float x) {   // cvt.rna.tf32.f32: nearest, ties away, 10-bit mantissa
  union { float f; uint32_t u; } v; v.f = x;
  v.u = (v.u + 0x1000u) & 0xFFFFE000u;
  return v.f;
}
ILSW_HD float round_bf16(float x) {   // cvt.rn.bf16.f32
  union { float f; uint32_t u; } v; v.f = x;
  v.u = (v.u + 0x7FFFu + ((v.u >> 16) & 1u)) & 0xFFFF0000u;
  return v.f;
}

// GEMM operand accessors + epilogue (shared by the device tile kernel and the host simulator)
// ------------------------------------------------------------------------------------------
ILSW_HD float act_apply(float v, int act) {
  if (act == ACT_RELU) return v > 0.f ? v : 0.f;
  if (act == ACT_TANH) return tanhf(v);
  return v;
}
// element (m,k) of a fused first layer (GemmOp::a0 -> L0FuseOp): fp32 FMA chain in input order
ILSW_HD float gemm_A_fused(const L0FuseOp& f, int m, int k) {
  float acc = ldg(f.b + k);
  const float* x = f.X + (size_t)m * f.ldx;
  const float* w = f.W + (size_t)k * f.K0;
  for (int j = 0; j < f.K0; ++j) acc = fmaf(ldg(x + j), ldg(w + j), acc);
  return act_apply(acc, f.act);
}
ILSW_HD float gemm_A(const GemmOp& o, int m, int k) {
  return ldg(o.A + (o.a_mc ? (size_t)k * o.lda + m : (size_t)m * o.lda + k));
}
ILSW_HD float gemm_B(const GemmOp& o, int k, int n) {
  if (o.aug_ones && n == o.N) return 1.0f;
  return ldg(o.B + (o.b_nc ? (size_t)k * o.ldb + n : (size_t)n * o.ldb + k));
}
// The epilogue is split into a LOAD half and a STORE half so that a tile can issue the loads of
// all its outputs (bias, mask source, previous value) before any store: a store between two
// loads would serialise them into separate L2 round trips.
struct EpiIn { float bias, h, prev; };
ILSW_HD EpiIn epi_load(const GemmOp& o, int m, int n) {
  EpiIn e; e.bias = 0.f; e.h = 0.f; e.prev = 0.f;
  if (o.aug_ones && n == o.N) {
    if (o.accumulate) e.prev = ldg(o.bias_out + m);
    return e;
  }
  if (o.bias) e.bias = ldg(o.bias + n);
  if (o.mask != ACT_NONE) e.h = ldg(o.H + (size_t)m * o.ldh + n);
  if (o.accumulate) e.prev = ldg(o.C + (size_t)m * o.ldc + n);
  return e;
}
// device only: the produced gradient element also goes to every replica's receive slot (plain stores on peer-mapped
// memory; the CTA fences and signals once, after its jobs of the phase)
ILSW_HD void push_elem(const PushCtx* push, const float* where, float v) {
#ifdef __CUDA_ARCH__
  if (push) {
    const size_t gi = (size_t)(where - push->grad_base);
    if (push->mc) asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(push->mc + gi), "f"(v) : "memory");
    else
      for (int r = 0; r < push->world; ++r) push->peer[r][gi] = v;
  }
#else
  (void)push; (void)where; (void)v;
#endif
}
ILSW_HD void epi_store(const GemmOp& o, int m, int n, float v, const EpiIn& e, const PushCtx* push = nullptr) {
  if (o.aug_ones && n == o.N) {
    const float bv = o.accumulate ? v + e.prev : v;
    o.bias_out[m] = bv;
    push_elem(push, o.bias_out + m, bv);
    return;
  }
  if (o.bias) v += e.bias;
  if (o.act == ACT_RELU) v = v > 0.f ? v : 0.f;
  else if (o.act == ACT_TANH) v = tanhf(v);
  const size_t ci = (size_t)m * o.ldc + n;
  if (o.C2) o.C2[ci] = v;
  if (o.mask == ACT_RELU) v = e.h > 0.f ? v : 0.f;
  else if (o.mask == ACT_TANH) v *= (1.0f - e.h * e.h);
  if (o.accumulate) v += e.prev;
  o.C[ci] = v;
  push_elem(push, o.C + ci, v);
}
ILSW_HD void gemm_epilogue(const GemmOp& o, int m, int n, float v) { epi_store(o, m, n, v, epi_load(o, m, n)); }

// ------------------------------------------------------------------------------------------
// Adam / Polyak element kernels
// ------------------------------------------------------------------------------------------
struct AdamCoef { float w1, one_m_w1, beta2, one_m_beta2, neg_step, bc2_sqrt, eps, gscale, tau, one_m_tau; };

// Adam step count of `slot` for step index `s` of this launch (1-based count AFTER the update).
ILSW_HD int adam_t(const RunArgs& a, const Hyper& hp, int slot, int s) {
  if (hp.algo == 2 && slot == SLOT_POLICY) {  // TD3 delayed policy: updates when (step0+s) % period == 0
    int per = hp.period > 0 ? hp.period : 1;
    int first = ((per - (a.step0 % per)) % per);  // first s >= 0 with update
    if (s < first) return a.t0[slot];
    return a.t0[slot] + (s - first) / per + 1;
  }
  return a.t0[slot] + s + 1;
}

// p1 = beta1^t, p2 = beta2^t
ILSW_HD AdamCoef adam_coef_pw(const AdamOp& o, double p1, double p2, int world) {
  AdamCoef c;
  double b1 = o.beta1, b2 = o.beta2;
  double bc1 = 1.0 - p1;
  double bc2 = 1.0 - p2;
  c.w1 = (float)(1.0 - b1);
  c.one_m_w1 = 1.0f - c.w1;
  c.beta2 = (float)b2;
  c.one_m_beta2 = (float)(1.0 - b2);
  c.neg_step = (float)(-(o.lr / bc1));
  c.bc2_sqrt = (float)sqrt(bc2);
  c.eps = (float)o.eps;
  c.gscale = (o.grad_scale_world && world > 1) ? 1.0f / (float)world : 1.0f;
  c.tau = o.tau;
  c.one_m_tau = (float)(1.0 - (double)o.tau);
  return c;
}
ILSW_HD AdamCoef adam_coef(const AdamOp& o, int t, int world) {
  return adam_coef_pw(o, pow(o.beta1, (double)t), pow(o.beta2, (double)t), world);
}

// one element of torch.optim.Adam (+ the Polyak update of a target network from the NEW parameter) with the
// state already in registers: shared by the flat Adam jobs and the fused GEMM epilogues (identical arithmetic)
// The element update is written with EXPLICIT roundings on the device: it is inlined into several call sites (flat Adam
// jobs, three tile epilogues) and nvcc's FMA contraction picks a different fusion per site (a*b + c*d has two), which made
// a single-replica run (Adam in the tile epilogues) differ from R identical replicas (flat Adam job after the exchange)
// by an ulp.  With intrinsics every site executes the same sequence (tools/replica_check.py test A: bitwise equal).
#ifdef __CUDA_ARCH__
#define ILSW_MUL(a, b) __fmul_rn(a, b)
#define ILSW_ADD(a, b) __fadd_rn(a, b)
#define ILSW_SUB(a, b) __fsub_rn(a, b)
#define ILSW_FMA(a, b, c) __fmaf_rn(a, b, c)
// Division and square root: ONE approximate instruction each (div.approx / sqrt.approx, <= 2 ulp) instead of the IEEE
// sequences (__fdiv_rn / __fsqrt_rn: ~15 dependent instructions plus a slow-path call each).  An Adam update moves a
// parameter by <= lr, so 2 ulp of the update is ~1e-10 of a parameter (the parity bar is 1e-5); the IEEE forms made
// the optimiser the most expensive part of every weight-gradient epilogue (1 us per element and thread: 14 us per
// 2048-element flat job on the B200).  Still one fixed instruction sequence at every call site -> bit-reproducible.
__device__ __forceinline__ float ilsw_div_approx(float a, float b) {
  float r;
  asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ float ilsw_sqrt_approx(float a) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(a));
  return r;
}
#define ILSW_DIV(a, b) ilsw_div_approx(a, b)
#define ILSW_SQRT(a) ilsw_sqrt_approx(a)
#else
#define ILSW_MUL(a, b) ((a) * (b))
#define ILSW_ADD(a, b) ((a) + (b))
#define ILSW_SUB(a, b) ((a) - (b))
#define ILSW_FMA(a, b, c) fmaf(a, b, c)
#define ILSW_DIV(a, b) ((a) / (b))
#define ILSW_SQRT(a) sqrtf(a)
#endif
// sh: false when the caller has checked once (per tile / job) that the optimiser has no aligned W0 copies to maintain
ILSW_HD void adam_math_store(const AdamOp& o, const AdamCoef& c, int i, float g, float m, float v, float p, float tg, bool sh = true) {
  // exp_avg.lerp_(grad, 1-beta1)
  const float d = ILSW_SUB(g, m);
  m = (c.w1 < 0.5f) ? ILSW_FMA(c.w1, d, m) : ILSW_SUB(g, ILSW_MUL(d, c.one_m_w1));
  // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1-beta2)
  v = ILSW_FMA(ILSW_MUL(c.one_m_beta2, g), g, ILSW_MUL(v, c.beta2));
  const float denom = ILSW_ADD(ILSW_DIV(ILSW_SQRT(v), c.bc2_sqrt), c.eps);
  p = ILSW_ADD(p, ILSW_DIV(ILSW_MUL(c.neg_step, m), denom));  // addcdiv_(exp_avg, denom, value=-step_size)
  o.m[i] = m;
  o.v[i] = v;
  o.p[i] = p;
  if (sh) shadow_store(o.sh_p, i, p);
  if (o.target) {
    const float tn = ILSW_FMA(tg, c.one_m_tau, ILSW_MUL(p, c.tau));
    o.target[i] = tn;
    if (sh) shadow_store(o.sh_t, i, tn);
  }
}
ILSW_HD bool adam_has_shadow(const AdamOp& o) { return o.sh_p.ptr != nullptr || o.sh_t.ptr != nullptr; }
ILSW_HD void adam_elem_g(const AdamOp& o, const AdamCoef& c, int i, float g, bool sh = true) {
  adam_math_store(o, c, i, g, o.m[i], o.v[i], o.p[i], o.target ? o.target[i] : 0.f, sh);
}
// gradient element i: the sum of the split-K partial arenas in a fixed order (one arena without splits)
// All partial loads are issued before the first add (a runtime-trip-count loop of load + add serialises one L2 round
// trip per split: measured 17 us per 2048-element Adam job on the B200).
ILSW_HD float adam_grad(const AdamOp& o, int i) {
  float part[kMaxGradSplits];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int s = 0; s < kMaxGradSplits; ++s) part[s] = (s == 0 || s < o.g_splits) ? ldg(o.g + (size_t)s * o.g_split_stride + i) : 0.f;
  float g = part[0];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int s = 1; s < kMaxGradSplits; ++s) g += part[s];
  return g;
}
ILSW_HD void adam_elem(const AdamOp& o, const AdamCoef& c, int i) { adam_elem_g(o, c, i, adam_grad(o, i) * c.gscale); }

// fused optimiser epilogue: element (m,n) of GEMM `o` is gradient element `gi` of the arena described by `ad`
ILSW_HD int gemm_grad_index(const GemmOp& o, const AdamOp& ad, int m, int n) {
  if (o.aug_ones && n == o.N) return (int)(o.bias_out - ad.g) + m;
  return (int)(o.C - ad.g) + m * o.ldc + n;
}

ILSW_HD void polyak_elem(const PolyakOp& o, int i) {
  float om = (float)(1.0 - (double)o.tau);
  const float tn = o.target[i] * om + ldg(o.src + i) * o.tau;
  o.target[i] = tn;
  shadow_store(o.sh_t, i, tn);
}
ILSW_HD void shadow_refresh_elem(const ShadowOp& o, int i) { shadow_store(o.dst, i, ldg(o.src + i)); }

// ------------------------------------------------------------------------------------------
// Row kernels.  (c: context, a: launch args, s: step index within launch, r: row)
// ------------------------------------------------------------------------------------------
ILSW_HD const float* ring_row(const RingView& rv, int idx) { return rv.rows + (size_t)idx * rv.stride; }

// ---- SAC-alpha -------------------------------------------------------------------------
// R3/R4: sample index -> packed batch tiles.  Ring row layout: [obs(O) | act(A) | rew | term | next_obs(O)].
ILSW_HDN void row_sac_gather(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const int O = S.O, A = S.A, B = S.B;
  const bool td3 = (c.hp.algo == 2);
  const float *obs, *act, *nobs;
  float rew, term;
  int idx = b, idx_her = -1;
  if (a.has_direct) {
    obs = a.direct.obs + (size_t)b * O;
    act = a.direct.act + (size_t)b * A;
    nobs = a.direct.next_obs + (size_t)b * O;
    rew = a.direct.rew[b];
    term = a.direct.term[b];
    if (lane == 0) S.idx[b] = b;
  } else {
    // adv_irl.py:239-255: the last n_from_expert rows of the policy batch are expert transitions
    const bool from_expert = c.hp.has_disc && b >= B - c.hp.n_from_expert;
    const RingView& rv = from_expert ? a.ring_expert : a.ring_policy;
    if (a.her.enabled) {
      // relabel_replay_buffer.py:70-95: uniform finished trajectory -> uniform step in it -> uniform FUTURE step [step, end)
      if (a.has_inject) {
        idx = a.inj.idx[(size_t)s * B + b];
        idx_her = a.her.inj_idx_her[(size_t)s * B + b];
      } else {
        const int t = philox_index(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 6u, a.her.n_traj);
        const int st = a.her.traj_start[t], len = a.her.traj_len[t];
        const int off = philox_index(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 0u, len);
        const int fut = off + philox_index(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 7u, len - off);
        idx = (st + off) % rv.size;
        idx_her = (st + fut) % rv.size;
      }
    } else {
      idx = a.has_inject ? a.inj.idx[(size_t)s * B + b]
                         : philox_index(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 0u, rv.size);
    }
    if (lane == 0) S.idx[b] = idx;
    const float* src = ring_row(rv, idx);
    obs = src;
    act = src + O;
    rew = src[O + A];
    term = src[O + A + 1];
    nobs = src + O + A + 2;
  }
  // hindsight relabel (relabel_replay_buffer.py:99-131): goal part of the first relabel_num rows <- next achieved goal of
  // the future step; sparse goal reward recomputed for EVERY row from its next achieved goal and (new) desired goal
  const bool her = a.her.enabled && !a.has_direct;
  const int G = her ? a.her.G : 0, g0 = O - G;
  const bool relabel = her && b < a.her.relabel_num;
  const float* goal_src = relabel ? a.her.ag_next + (size_t)idx_her * G : nullptr;
  if (her && a.her.relabel_num > 0) {
    const float* ag = a.her.ag_next + (size_t)idx * G;
    float ss = 0.f;
    for (int k = lane; k < G; k += nl) {
      const float d = ag[k] - (relabel ? goal_src[k] : obs[g0 + k]);
      ss += d * d;
    }
    ss = wsum(ss);
    rew = (sqrtf(ss) > a.her.threshold) ? -1.0f : -0.0f;
  }
  if (lane == 0) { S.rew[b] = rew; S.term[b] = term; }
  // ring rows come from HBM: all loads of a batch of 8 elements per lane are issued before the first store (a plain
  // load/store loop serialises one DRAM round trip per 32 floats: 25 us for 1024 Humanoid rows)
  constexpr int kGB = 8;
  for (int k0 = 0; k0 < O; k0 += kGB * nl) {
    float ov[kGB], nv[kGB];
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int u = 0; u < kGB; ++u) {
      const int k = k0 + lane + u * nl;
      if (k < O) { ov[u] = obs[k]; nv[u] = nobs[k]; }
    }
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
    for (int u = 0; u < kGB; ++u) {
      const int k = k0 + lane + u * nl;
      if (k >= O) continue;
      float o = ov[u], n = nv[u];
      if (relabel && k >= g0) o = n = goal_src[k - g0];
      S.Xoa[(size_t)b * S.ld_oa + k] = o;
      S.Xon[(size_t)b * S.ld_oa + k] = o;
      S.Xna[(size_t)b * S.ld_oa + k] = n;
      if (c.hp.has_disc && c.hp.state_only) {          // reward relabel input cat(obs, next_obs) (adv_irl.py:265-269)
        c.d.Xsn[(size_t)b * c.d.ld_sn + k] = o;
        c.d.Xsn[(size_t)b * c.d.ld_sn + O + k] = n;
      }
      if (!td3) {
        S.Xpi[(size_t)b * S.ld_o + k] = n;            // rows [0,B): next_obs
        S.Xpi[(size_t)(B + b) * S.ld_o + k] = o;      // rows [B,2B): obs
      }
    }
  }
  for (int j = lane; j < A; j += nl) {
    S.Xoa[(size_t)b * S.ld_oa + O + j] = act[j];
    float e0, e1 = 0.f;
    if (a.has_inject) {
      e0 = a.inj.eps_next[((size_t)s * B + b) * A + j];
      if (!td3) e1 = a.inj.eps_cur[((size_t)s * B + b) * A + j];
    } else {
      e0 = philox_normal(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, (uint32_t)j, 1u);
      if (!td3) e1 = philox_normal(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, (uint32_t)j, 2u);
    }
    if (td3) {
      S.noise[(size_t)b * A + j] = e0;
    } else {
      S.eps[(size_t)b * A + j] = e0;
      S.eps[(size_t)(B + b) * A + j] = e1;
    }
  }
}

// N2: policy heads on 2B rows (rows [0,B) = next_obs, [B,2B) = obs).
ILSW_HDN void row_sac_heads(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int A = S.A, Hd = S.Hd, B = S.B, O = S.O;
  const float* h = S.h1p + (size_t)r * Hd;
  for (int j = 0; j < A; ++j) {
    float mu = wdot(h, P.p + P.oW2 + (size_t)j * Hd, Hd, lane, nl) + ldg(P.p + P.ob2 + j);
    float lr = wdot(h, P.p + P.oW3 + (size_t)j * Hd, Hd, lane, nl) + ldg(P.p + P.ob3 + j);
    if (lane == 0) { S.mean[(size_t)r * A + j] = mu; S.lraw[(size_t)r * A + j] = lr; }
  }
  wsync();
  float s1 = 0.f, s2 = 0.f, s3 = 0.f;
  for (int j = lane; j < A; j += nl) {
    float mu = S.mean[(size_t)r * A + j];
    float ls = S.lraw[(size_t)r * A + j];
    ls = fminf(fmaxf(ls, -20.0f), 2.0f);  // clamp(LOG_SIG_MIN, LOG_SIG_MAX), policies.py:265
    float sig = expf(ls);
    float cov = expf(2.0f * ls);
    float z = ldg(S.eps + (size_t)r * A + j) * sig + mu;
    float t = tanhf(z);
    float d = mu - z;
    s1 += d * d / cov;
    s2 += ls;
    s3 += logf(1.0f - t * t + 1e-6f);
    S.lstd[(size_t)r * A + j] = ls;
    S.act[(size_t)r * A + j] = t;
    if (r < B) S.Xna[(size_t)r * S.ld_oa + O + j] = t;
    else S.Xon[(size_t)(r - B) * S.ld_oa + O + j] = t;
  }
  s1 = wsum(s1); s2 = wsum(s2); s3 = wsum(s3);
  if (lane == 0) {
    float lp = -0.5f * s1;
    lp -= (s2 + 0.5f * 1.8378770664093453f);  // + 0.5*log(2*pi) ONCE (distributions.py:47-49)
    lp -= s3;
    S.logpi[r] = lp;
  }
}

// shared by SAC and TD3: target value, critic loss terms, output-layer backward of both critics
ILSW_HDN void row_critic_target(const Ctx& c, int b, int lane, int nl, bool use_entropy, float loss_grad_factor) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd;
  float tq0 = wdot(S.h1t[0] + (size_t)b * Hd, c.tqf[0].p + c.tqf[0].oW2, Hd, lane, nl) + ldg(c.tqf[0].p + c.tqf[0].ob2);
  float tq1 = wdot(S.h1t[1] + (size_t)b * Hd, c.tqf[1].p + c.tqf[1].oW2, Hd, lane, nl) + ldg(c.tqf[1].p + c.tqf[1].ob2);
  float tmin = fminf(tq0, tq1);
  if (c.hp.her) tmin = fminf(fmaxf(tmin, c.hp.clip_l), c.hp.clip_r);      // her/td3.py:116-120 (torch.clip)
  float rs = c.hp.reward_scale * ldg(S.rew + b);
  float inner = tmin;
  if (use_entropy) inner = tmin - ldg(&c.dyn->alpha) * ldg(S.logpi + b);
  float y = rs + (1.0f - ldg(S.term + b)) * c.hp.discount * inner;
  float invB = 1.0f / (float)S.B;
  for (int i = 0; i < 2; ++i) {
    const MlpPtrs& Q = c.qf[i];
    const float* h = S.h1q[i] + (size_t)b * Hd;
    float q = wdot(h, Q.p + Q.oW2, Hd, lane, nl) + ldg(Q.p + Q.ob2);
    float diff = q - y;
    float dq = loss_grad_factor * diff * invB;
    if (lane == 0) {
      S.qp[i][b] = q;
      S.dq[i][b] = dq;
      S.lossterm[i][b] = diff * diff;
    }
    float* d1 = S.d1q[i] + (size_t)b * Hd;
    for (int k = lane; k < Hd; k += nl) d1[k] = (ldg(h + k) > 0.f) ? dq * ldg(Q.p + Q.oW2 + k) : 0.f;
  }
  if (lane == 0) { S.tq[0][b] = tq0; S.tq[1][b] = tq1; S.y[b] = y; }
}

ILSW_HDN void row_sac_target(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  row_critic_target(c, b, lane, nl, true, 1.0f);  // d/dq of 0.5*mean((q-y)^2)
}

// policy loss terms + output-layer backward through min(Q1,Q2)(obs, a~) with UPDATED critics
ILSW_HDN void row_sac_ploss(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, A = S.A, B = S.B;
  float q[2];
  for (int i = 0; i < 2; ++i) {
    const MlpPtrs& Q = c.qf[i];
    q[i] = wdot(S.h1n[i] + (size_t)b * Hd, Q.p + Q.oW2, Hd, lane, nl) + ldg(Q.p + Q.ob2);
  }
  float invB = 1.0f / (float)B;
  // torch.min backward: all to the smaller, split evenly on ties
  float w0 = q[0] < q[1] ? 1.f : (q[0] == q[1] ? 0.5f : 0.f);
  float wq[2] = {w0, 1.f - w0};
  float qmin = fminf(q[0], q[1]);
  float smu = 0.f, sls = 0.f;
  for (int j = lane; j < A; j += nl) {
    float mu = ldg(S.mean + (size_t)(B + b) * A + j), ls = ldg(S.lstd + (size_t)(B + b) * A + j);
    smu += mu * mu;
    sls += ls * ls;
  }
  smu = wsum(smu); sls = wsum(sls);
  if (lane == 0) {
    S.qn[0][b] = q[0]; S.qn[1][b] = q[1];
    S.plterm[b] = ldg(&c.dyn->alpha) * ldg(S.logpi + B + b) - qmin;
    S.regmu[b] = smu; S.regls[b] = sls;
  }
  for (int i = 0; i < 2; ++i) {
    const MlpPtrs& Q = c.qf[i];
    float dq = -invB * wq[i];
    const float* h = S.h1n[i] + (size_t)b * Hd;
    float* e1 = S.e1[i] + (size_t)b * Hd;
    for (int k = lane; k < Hd; k += nl) e1[k] = (ldg(h + k) > 0.f) ? dq * ldg(Q.p + Q.oW2 + k) : 0.f;
  }
}

// backward through the tanh-Gaussian head; produces dmean, dlraw and delta of the last hidden layer
ILSW_HDN void row_sac_pibwd(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int Hd = S.Hd, A = S.A, B = S.B;
  const int r = B + b;
  const float alpha = ldg(&c.dyn->alpha);
  const float invB = 1.0f / (float)B, invBA = 1.0f / (float)(B * A);
  for (int j = lane; j < A; j += nl) {
    float gA = ldg(S.dA[0] + (size_t)b * A + j) + ldg(S.dA[1] + (size_t)b * A + j);
    float t = ldg(S.act + (size_t)r * A + j);
    float mu = ldg(S.mean + (size_t)r * A + j), ls = ldg(S.lstd + (size_t)r * A + j), lr = ldg(S.lraw + (size_t)r * A + j);
    float ep = ldg(S.eps + (size_t)r * A + j);
    float om = 1.0f - t * t;
    float J = 2.0f * t * om / (om + 1e-6f);           // d/dz of -log(1 - tanh(z)^2 + 1e-6)
    float dz = gA * om + alpha * invB * J;
    float dmu = dz + 2.0f * c.hp.mean_reg * mu * invBA;
    float dl = dz * ep * expf(ls) - alpha * invB + 2.0f * c.hp.std_reg * ls * invBA;
    float dlr = (lr >= -20.0f && lr <= 2.0f) ? dl : 0.f;  // clamp backward
    S.dmean[(size_t)b * A + j] = dmu;
    S.dlraw[(size_t)b * A + j] = dlr;
  }
  wsync();
  const float* h = S.h1p + (size_t)r * Hd;
  float* d1 = S.d1p + (size_t)b * Hd;
  for (int k = lane; k < Hd; k += nl) {
    float acc = 0.f;
    for (int j = 0; j < A; ++j)
      acc += S.dmean[(size_t)b * A + j] * ldg(P.p + P.oW2 + (size_t)j * Hd + k) +
             S.dlraw[(size_t)b * A + j] * ldg(P.p + P.oW3 + (size_t)j * Hd + k);
    d1[k] = (ldg(h + k) > 0.f) ? acc : 0.f;
  }
  if (lane == 0) S.aterm[b] = ldg(S.logpi + r) + c.hp.target_entropy;
}

ILSW_HDN float wmean(const float* x, int n, int lane, int nl) {
  float s = 0.f;
  for (int i = lane; i < n; i += nl) s += ldg(x + i);
  return wsum(s) / (float)n;
}

ILSW_HDN void stats_copy(float* dst, const float* src, int n, int lane, int nl) {
  for (int i = lane; i < n; i += nl) dst[i] = ldg(src + i);
}

// stats snapshot layout (floats): see ilsw_stats_offsets() in ilsw_program.h
ILSW_HDN void snapshot_sac(const Ctx& c, int lane, int nl) {
  const SacBufs& S = c.s;
  const int B = S.B, A = S.A;
  const bool td3 = c.hp.algo == 2;
  float* st = c.stats;
  stats_copy(st, S.qp[0], B, lane, nl); st += B;
  stats_copy(st, S.qp[1], B, lane, nl); st += B;
  stats_copy(st, S.y, B, lane, nl); st += B;
  stats_copy(st, S.lossterm[0], B, lane, nl); st += B;
  stats_copy(st, S.lossterm[1], B, lane, nl); st += B;
  if (c.hp.algo == 3) stats_copy(st, S.vp, B, lane, nl);   // SAC-V: V predictions (slot unused otherwise)
  st += B;
  if (td3) {
    stats_copy(st, S.act, B * A, lane, nl); st += B * A;
  } else {
    stats_copy(st, S.logpi + B, B, lane, nl); st += B;
    stats_copy(st, S.mean + (size_t)B * A, B * A, lane, nl); st += B * A;
    stats_copy(st, S.lstd + (size_t)B * A, B * A, lane, nl); st += B * A;
  }
}

// losses, alpha update (float64 scalar Adam), loss log, optional stats snapshot.  Single warp.
ILSW_HDN void row_sac_final(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  if (r != 0) return;
  const SacBufs& S = c.s;
  const int B = S.B, A = S.A;
  float l1 = 0.5f * wmean(S.lossterm[0], B, lane, nl);
  float l2 = 0.5f * wmean(S.lossterm[1], B, lane, nl);
  float pl = wmean(S.plterm, B, lane, nl);
  float rm = wmean(S.regmu, B, lane, nl) / (float)A;
  float rl = wmean(S.regls, B, lane, nl) / (float)A;
  pl = pl + (c.hp.mean_reg * rm + c.hp.std_reg * rl);
  float am = wmean(S.aterm, B, lane, nl);
  float q1m = wmean(S.qp[0], B, lane, nl);
  float lpm = wmean(S.logpi + B, B, lane, nl);
  float ytm = wmean(S.y, B, lane, nl);
  DynState* d = c.dyn;
  float alpha_loss = 0.f;
  if (lane == 0) {
    float* L = c.loss_log + (size_t)(a.loss_log_offset + s) * kLossSlots;
    L[L_QF1] = l1; L[L_QF2] = l2; L[L_POLICY] = pl;
    L[L_Q1_MEAN] = q1m; L[L_LOGPI_MEAN] = lpm; L[L_QT_MEAN] = ytm;
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

// ---- TD3 -------------------------------------------------------------------------------
// target policy head with the policy module's clipped noise (policies.py:176-186, td3.py:82-83)
ILSW_HDN void row_td3_thead(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.tpolicy;
  const int A = S.A, Hd = S.Hd, O = S.O;
  if (c.hp.her) {     // her/td3.py:103-112: next action = clamp(sigma * N(0,1), min_act, max_act)
    for (int j = lane; j < A; j += nl) {
      float nz = c.hp.her_sigma * ldg(S.noise + (size_t)b * A + j);
      S.Xna[(size_t)b * S.ld_oa + O + j] = fminf(fmaxf(nz, c.hp.min_act), c.hp.max_act);
    }
    return;
  }
  const float* h = S.h1tp + (size_t)b * Hd;
  for (int j = 0; j < A; ++j) {
    float pre = wdot(h, P.p + P.oW2 + (size_t)j * Hd, Hd, lane, nl) + ldg(P.p + P.ob2 + j);
    if (lane == 0) {
      float act = c.hp.max_act * tanhf(pre);
      float nz = c.hp.policy_noise * ldg(S.noise + (size_t)b * A + j);
      nz = fminf(fmaxf(nz, -c.hp.noise_clip), c.hp.noise_clip);
      S.Xna[(size_t)b * S.ld_oa + O + j] = act + nz;  // no re-clip
    }
  }
}
ILSW_HDN void row_td3_target(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  row_critic_target(c, b, lane, nl, false, 2.0f);  // d/dq of mean((q-y)^2), no 1/2 (td3.py:93-98)
}
ILSW_HDN void row_td3_phead(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int A = S.A, Hd = S.Hd, O = S.O;
  const float* h = S.h1p + (size_t)b * Hd;
  for (int j = 0; j < A; ++j) {
    float pre = wdot(h, P.p + P.oW2 + (size_t)j * Hd, Hd, lane, nl) + ldg(P.p + P.ob2 + j);
    if (lane == 0) {
      float t = tanhf(pre);
      S.act[(size_t)b * A + j] = t;
      S.Xon[(size_t)b * S.ld_oa + O + j] = c.hp.max_act * t;
    }
  }
}
ILSW_HDN void row_td3_ploss(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const MlpPtrs& Q = c.qf[0];
  const int Hd = S.Hd;
  const float* h = S.h1n[0] + (size_t)b * Hd;
  float q = wdot(h, Q.p + Q.oW2, Hd, lane, nl) + ldg(Q.p + Q.ob2);
  float dq = -1.0f / (float)S.B;
  float l2 = 0.f;
  if (c.hp.her) {     // + mean(action^2) over B x A elements (her/td3.py:150-152)
    for (int j = lane; j < S.A; j += nl) { float av = c.hp.max_act * ldg(S.act + (size_t)b * S.A + j); l2 += av * av; }
    l2 = wsum(l2) / (float)S.A;
  }
  if (lane == 0) { S.qn[0][b] = q; S.plterm[b] = -q + l2; }
  float* e1 = S.e1[0] + (size_t)b * Hd;
  for (int k = lane; k < Hd; k += nl) e1[k] = (ldg(h + k) > 0.f) ? dq * ldg(Q.p + Q.oW2 + k) : 0.f;
}
ILSW_HDN void row_td3_pibwd(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const MlpPtrs& P = c.policy;
  const int Hd = S.Hd, A = S.A;
  for (int j = lane; j < A; j += nl) {
    float t = ldg(S.act + (size_t)b * A + j);
    float da = ldg(S.dA[0] + (size_t)b * A + j);
    if (c.hp.her) da += 2.0f * c.hp.max_act * t / (float)(S.B * A);     // d mean(a^2) / da
    S.dmean[(size_t)b * A + j] = da * c.hp.max_act * (1.0f - t * t);
  }
  wsync();
  const float* h = S.h1p + (size_t)b * Hd;
  float* d1 = S.d1p + (size_t)b * Hd;
  for (int k = lane; k < Hd; k += nl) {
    float acc = 0.f;
    for (int j = 0; j < A; ++j) acc += S.dmean[(size_t)b * A + j] * ldg(P.p + P.oW2 + (size_t)j * Hd + k);
    d1[k] = (ldg(h + k) > 0.f) ? acc : 0.f;
  }
}
ILSW_HDN void row_td3_final(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  if (r != 0) return;
  const SacBufs& S = c.s;
  const int B = S.B;
  float l1 = wmean(S.lossterm[0], B, lane, nl);
  float l2 = wmean(S.lossterm[1], B, lane, nl);
  float q1m = wmean(S.qp[0], B, lane, nl);
  float ytm = wmean(S.y, B, lane, nl);
  if (lane == 0) {
    float* L = c.loss_log + (size_t)(a.loss_log_offset + s) * kLossSlots;
    L[L_QF1] = l1; L[L_QF2] = l2; L[L_Q1_MEAN] = q1m; L[L_QT_MEAN] = ytm;
    L[L_POLICY] = nanf("");  // overwritten by row_td3_final_policy on policy steps
  }
  if (s == a.stats_step) snapshot_sac(c, lane, nl);
}
ILSW_HDN void row_td3_final_policy(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  if (r != 0) return;
  float pl = wmean(c.s.plterm, c.s.B, lane, nl);
  if (lane == 0) c.loss_log[(size_t)(a.loss_log_offset + s) * kLossSlots + L_POLICY] = pl;
  // "Policy Action" of the logged step = the deterministic policy actions of THIS step (td3.py:113-136), computed after
  // row_td3_final took its snapshot
  if (s == a.stats_step) stats_copy(c.stats + 6 * c.s.B, c.s.act, c.s.B * c.s.A, lane, nl);
}

// ---- AdvIRL discriminator --------------------------------------------------------------
// hidden activation of the discriminator blocks, from its OUTPUT h: act'(z) and act''(z)/act'(z) (tanh: 1-h^2, -2h; relu: [h>0], 0)
ILSW_HD float disc_dact(int act, float h) { return act == ACT_RELU ? (h > 0.f ? 1.f : 0.f) : 1.0f - h * h; }
ILSW_HD float disc_curv(int act, float h) { return act == ACT_RELU ? 0.f : -2.0f * h; }
ILSW_HDN void row_disc_gather(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const int B = Dd.B, D = Dd.D;
  int ie = a.has_inject ? a.inj.idx_expert[(size_t)s * B + b]
                        : philox_index(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 3u, a.ring_expert.size);
  int ip = a.has_inject ? a.inj.idx_policy_d[(size_t)s * B + b]
                        : philox_index(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 4u, a.ring_policy.size);
  if (lane == 0) { Dd.idx_e[b] = ie; Dd.idx_p[b] = ip; }
  const float* xe = ring_row(a.ring_expert, ie);   // first D = O+A floats are cat(obs, act)
  const float* xp = ring_row(a.ring_policy, ip);
  float ep = 0.f;
  if (c.hp.use_gp) {
    ep = a.has_inject ? a.inj.gp_eps[(size_t)s * B + b]
                      : philox_uniform(a.seed, (uint32_t)(a.step0 + s), (uint32_t)b, 5u);
    if (lane == 0) Dd.gp_eps[b] = ep;
  }
  const int O = c.s.O, skip = c.hp.state_only ? c.s.A + 2 : 0;     // state_only: [obs | next_obs] of the ring row
  for (int k = lane; k < D; k += nl) {
    const int ks = k < O ? k : k + skip;
    float e = xe[ks], p = xp[ks];
    Dd.X3[(size_t)b * Dd.ld_d + k] = e;
    Dd.X3[(size_t)(B + b) * Dd.ld_d + k] = p;
    if (c.hp.use_gp) Dd.X3[(size_t)(2 * B + b) * Dd.ld_d + k] = ep * e + (1.0f - ep) * p;  // adv_irl.py:191
  }
}

// output layer on all rows; BCE-with-logits terms + CE output backward (rows < 2B);
// clamp mask and GP delta2 = c * w3 * (1-h2^2) for the interpolated rows (>= 2B)
ILSW_HDN void row_disc_head(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const MlpPtrs& N = c.disc;
  const int B = Dd.B, Hd = Dd.Hd;
  const float* h2 = Dd.h2 + (size_t)r * Hd;
  const float* w3 = N.p + N.oW2;
  float y = wdot(h2, w3, Hd, lane, nl) + ldg(N.p + N.ob2);
  const float cm = c.hp.disc_clamp;
  float pass = (y >= -cm && y <= cm) ? 1.f : 0.f;
  if (r < 2 * B) {
    float x = fminf(fmaxf(y, -cm), cm);
    float t = r < B ? 1.f : 0.f;
    // BCEWithLogits: (1-t)*x + max(-x,0) + log(exp(-max(-x,0)) + exp(-x-max(-x,0)))
    float mv = fmaxf(-x, 0.f);
    float ce = (1.0f - t) * x + mv + logf(expf(-mv) + expf(-x - mv));
    float sg = 1.0f / (1.0f + expf(-x));
    float dl = (sg - t) / (float)(2 * B) * pass;
    if (lane == 0) {
      Dd.y[r] = x;
      Dd.dlogit[r] = dl;
      Dd.ceterm[r] = ce;
      Dd.accterm[r] = ((x > 0.f ? 1.f : 0.f) == t) ? 1.f : 0.f;
    }
    float* d2 = Dd.d2 + (size_t)r * Hd;
    for (int k = lane; k < Hd; k += nl) {
      float h = ldg(h2 + k);
      d2[k] = dl * ldg(w3 + k) * disc_dact(c.hp.disc_act, h);
    }
  } else {
    int b = r - 2 * B;
    if (lane == 0) Dd.cmask[b] = pass;
    float* dl2 = Dd.dl2 + (size_t)b * Hd;
    for (int k = lane; k < Hd; k += nl) {
      float h = ldg(h2 + k);
      dl2[k] = pass * ldg(w3 + k) * disc_dact(c.hp.disc_act, h);
    }
  }
}

// Gulrajani penalty: n = ||g||, term (n-1)^2, gbar = dL/dg = (2*lambda/B)(n-1) g/n  (0 at n=0)
ILSW_HDN void row_disc_gnorm(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const int D = Dd.D;
  const float* g = Dd.g + (size_t)b * Dd.ld_d;
  float ss = 0.f;
  for (int k = lane; k < D; k += nl) { float v = ldg(g + k); ss += v * v; }
  ss = wsum(ss);
  float n = sqrtf(ss);
  float coef = n > 0.f ? (2.0f * c.hp.gp_weight / (float)Dd.B) * (n - 1.0f) / n : 0.f;
  float* gb = Dd.gbar + (size_t)b * Dd.ld_d;
  for (int k = lane; k < D; k += nl) gb[k] = coef * ldg(g + k);
  if (lane == 0) { Dd.nrm[b] = n; Dd.gpterm[b] = (n - 1.0f) * (n - 1.0f); }
}
// ubar1 = dbar1 * s1 ; sbar1 = dbar1 * u1
ILSW_HDN void row_disc_ew1(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const int Hd = Dd.Hd, B = Dd.B;
  for (int k = lane; k < Hd; k += nl) {
    size_t i = (size_t)b * Hd + k;
    float h1 = ldg(Dd.h1 + (size_t)(2 * B + b) * Hd + k);
    float db = ldg(Dd.db1 + i);
    Dd.ub1[i] = db * disc_dact(c.hp.disc_act, h1);
    Dd.sb1[i] = db * ldg(Dd.u1 + i);
  }
}
// t3 = c*dbar2*s2 ; sbar2 = dbar2*(c*w3) ; zbar2 = (-2 h2 sbar2) * s2
ILSW_HDN void row_disc_ew2(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const int Hd = Dd.Hd, B = Dd.B;
  const float* w3 = c.disc.p + c.disc.oW2;
  float cmk = ldg(Dd.cmask + b);
  for (int k = lane; k < Hd; k += nl) {
    size_t i = (size_t)b * Hd + k;
    float h2 = ldg(Dd.h2 + (size_t)(2 * B + b) * Hd + k);
    float s2 = disc_dact(c.hp.disc_act, h2);
    float db = ldg(Dd.db2 + i);
    Dd.t3[i] = db * s2;                       // multiplied by c through the A operand (cmask)
    float sb2 = db * (cmk * ldg(w3 + k));
    Dd.zb2[i] = (disc_curv(c.hp.disc_act, h2) * sb2) * s2;
  }
}
// zbar1 = (hbar1_raw - 2 h1 sbar1) * s1
ILSW_HDN void row_disc_ew3(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const int Hd = Dd.Hd, B = Dd.B;
  for (int k = lane; k < Hd; k += nl) {
    size_t i = (size_t)b * Hd + k;
    float h1 = ldg(Dd.h1 + (size_t)(2 * B + b) * Hd + k);
    Dd.zb1[i] = (ldg(Dd.hb1 + i) + disc_curv(c.hp.disc_act, h1) * ldg(Dd.sb1 + i)) * disc_dact(c.hp.disc_act, h1);
  }
}
ILSW_HDN void row_disc_final(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  if (r != 0) return;
  const DiscBufs& Dd = c.d;
  float ce = wmean(Dd.ceterm, 2 * Dd.B, lane, nl);
  float acc = wmean(Dd.accterm, 2 * Dd.B, lane, nl);
  float gp = c.hp.use_gp ? wmean(Dd.gpterm, Dd.B, lane, nl) : 0.f;
  if (lane == 0) {
    float* L = c.loss_log + (size_t)(a.loss_log_offset + s) * kLossSlots;
    L[L_DISC_CE] = ce; L[L_DISC_ACC] = acc; L[L_GRAD_PEN] = gp;
  }
}
// D2: reward relabel of the policy batch (adv_irl.py:266-298); overwrites the SAC batch reward
ILSW_HDN void row_disc_reward(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const DiscBufs& Dd = c.d;
  const MlpPtrs& N = c.disc;
  const int Hd = Dd.Hd;
  float y = wdot(Dd.rh2 + (size_t)b * Hd, N.p + N.oW2, Hd, lane, nl) + ldg(N.p + N.ob2);
  float x = fminf(fmaxf(y, -c.hp.disc_clamp), c.hp.disc_clamp);
  float r;
  switch (c.hp.disc_mode) {
    case 0: r = x; break;                                              // airl
    case 1: r = (x > 20.f) ? x : log1pf(expf(x)); break;                // gail: softplus(x, beta=1)
    case 2: r = (-x > 20.f) ? x : -log1pf(expf(-x)); break;             // gail2: softplus(x, beta=-1)
    default: r = expf(x) * (-1.0f * x); break;                          // fairl
  }
  if (c.hp.clip_max_on) r = fminf(r, c.hp.rew_clip_max);
  if (c.hp.clip_min_on) r = fmaxf(r, c.hp.rew_clip_min);
  if (lane == 0) { c.s.rew[b] = r; Dd.rewraw[b] = r; }
}
ILSW_HDN void row_disc_reward_final(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  if (r != 0) return;
  const DiscBufs& Dd = c.d;
  const int B = Dd.B;
  float mean = wmean(Dd.rewraw, B, lane, nl);
  float ss = 0.f, mx = -INFINITY, mn = INFINITY;
  for (int i = lane; i < B; i += nl) {
    float v = ldg(Dd.rewraw + i), dv = v - mean;
    ss += dv * dv; mx = fmaxf(mx, v); mn = fminf(mn, v);
  }
  ss = wsum(ss); mx = wmaxf(mx); mn = wminf(mn);
  if (lane == 0) {
    float* L = c.loss_log + (size_t)(a.loss_log_offset + s) * kLossSlots;
    L[L_REW_MEAN] = mean; L[L_REW_STD] = sqrtf(ss / (float)B); L[L_REW_MAX] = mx; L[L_REW_MIN] = mn;
  }
}

// ---- SAC, V-function variant (sac.py:70-179) ------------------------------------------------
// q_target from the TARGET V net; V regression target = min Q(obs, a~) - alpha*log pi with the
// PRE-update critics; output-layer backward of Q1, Q2 and V.
ILSW_HDN void row_sacv_target(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, B = S.B;
  const float invB = 1.0f / (float)B;
  const float tv = wdot(S.h1tv + (size_t)b * Hd, c.tvf.p + c.tvf.oW2, Hd, lane, nl) + ldg(c.tvf.p + c.tvf.ob2);
  const float y = c.hp.reward_scale * ldg(S.rew + b) + (1.0f - ldg(S.term + b)) * c.hp.discount * tv;
  float qold[2];
  for (int i = 0; i < 2; ++i) {
    const MlpPtrs& Q = c.qf[i];
    qold[i] = wdot(S.h1n[i] + (size_t)b * Hd, Q.p + Q.oW2, Hd, lane, nl) + ldg(Q.p + Q.ob2);
    const float* h = S.h1q[i] + (size_t)b * Hd;
    const float q = wdot(h, Q.p + Q.oW2, Hd, lane, nl) + ldg(Q.p + Q.ob2);
    const float diff = q - y, dq = diff * invB;
    if (lane == 0) { S.qp[i][b] = q; S.dq[i][b] = dq; S.lossterm[i][b] = diff * diff; S.qn_old[i][b] = qold[i]; }
    float* d1 = S.d1q[i] + (size_t)b * Hd;
    for (int k = lane; k < Hd; k += nl) d1[k] = (ldg(h + k) > 0.f) ? dq * ldg(Q.p + Q.oW2 + k) : 0.f;
  }
  const float vtarget = fminf(qold[0], qold[1]) - ldg(&c.dyn->alpha) * ldg(S.logpi + B + b);
  const float* hv = S.h1v + (size_t)b * Hd;
  const float v = wdot(hv, c.vf.p + c.vf.oW2, Hd, lane, nl) + ldg(c.vf.p + c.vf.ob2);
  const float dvd = v - vtarget, dv = dvd * invB;
  if (lane == 0) { S.vp[b] = v; S.dv[b] = dv; S.lossterm_v[b] = dvd * dvd; S.tv[b] = tv; S.y[b] = y; }
  float* d1 = S.d1v + (size_t)b * Hd;
  for (int k = lane; k < Hd; k += nl) d1[k] = (ldg(hv + k) > 0.f) ? dv * ldg(c.vf.p + c.vf.oW2 + k) : 0.f;
}
ILSW_HDN void row_sacv_final(const Ctx& c, const RunArgs& a, int s, int r, int lane, int nl) {
  if (r != 0) return;
  const SacBufs& S = c.s;
  const int B = S.B, A = S.A;
  float l1 = 0.5f * wmean(S.lossterm[0], B, lane, nl);
  float l2 = 0.5f * wmean(S.lossterm[1], B, lane, nl);
  float lv = 0.5f * wmean(S.lossterm_v, B, lane, nl);
  float pl = wmean(S.plterm, B, lane, nl);
  float rm = wmean(S.regmu, B, lane, nl) / (float)A;
  float rl = wmean(S.regls, B, lane, nl) / (float)A;
  pl = pl + (c.hp.mean_reg * rm + c.hp.std_reg * rl);
  float q1m = wmean(S.qp[0], B, lane, nl), lpm = wmean(S.logpi + B, B, lane, nl), ytm = wmean(S.y, B, lane, nl);
  if (lane == 0) {
    float* L = c.loss_log + (size_t)(a.loss_log_offset + s) * kLossSlots;
    L[L_QF1] = l1; L[L_QF2] = l2; L[L_VF] = lv; L[L_POLICY] = pl; L[L_ALPHA] = c.dyn->alpha; L[L_ALPHA_LOSS] = 0.f;
    L[L_Q1_MEAN] = q1m; L[L_LOGPI_MEAN] = lpm; L[L_QT_MEAN] = ytm;
  }
  if (s == a.stats_step) snapshot_sac(c, lane, nl);
}

// generic form of the fused "dA = e0 . W0[:, O:O+A]" + head backward rows (the fast jobs are in
// ilsw_rows_fast.cuh); used for shapes outside the fast path and as cross-check in the host simulator
ILSW_HDN void row_da(const Ctx& c, int b, int nets, int lane, int nl) {
  const SacBufs& S = c.s;
  const int Hd = S.Hd, A = S.A, O = S.O, K0 = O + A;
  for (int i = 0; i < nets; ++i) {
    const MlpPtrs& Q = c.qf[i];
    for (int j = 0; j < A; ++j) {
      float acc = 0.f;
      for (int n = lane; n < Hd; n += nl) acc += ldg(S.e0[i] + (size_t)b * Hd + n) * ldg(Q.p + Q.oW0 + (size_t)n * K0 + O + j);
      acc = wsum(acc);
      if (lane == 0) S.dA[i][(size_t)b * A + j] = acc;
    }
  }
  wsync();
}
ILSW_HDN void row_sac_pibwd_da(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  row_da(c, b, 2, lane, nl);
  row_sac_pibwd(c, a, s, b, lane, nl);
}
ILSW_HDN void row_td3_pibwd_da(const Ctx& c, const RunArgs& a, int s, int b, int lane, int nl) {
  row_da(c, b, 1, lane, nl);
  row_td3_pibwd(c, a, s, b, lane, nl);
}

// ---- dispatcher --------------------------------------------------------------------------
ILSW_HDN void run_row(const Ctx& c, const RunArgs& a, int kind, int s, int r, int lane, int nl) {
  switch (kind) {
    case ROW_SAC_GATHER: case ROW_TD3_GATHER: row_sac_gather(c, a, s, r, lane, nl); break;
    case ROW_SAC_HEADS: row_sac_heads(c, a, s, r, lane, nl); break;
    case ROW_SAC_TARGET: row_sac_target(c, a, s, r, lane, nl); break;
    case ROW_SAC_PLOSS: row_sac_ploss(c, a, s, r, lane, nl); break;
    case ROW_SAC_PIBWD: row_sac_pibwd(c, a, s, r, lane, nl); break;
    case ROW_SAC_PIBWD_DA: row_sac_pibwd_da(c, a, s, r, lane, nl); break;
    case ROW_SACV_TARGET: row_sacv_target(c, a, s, r, lane, nl); break;
    case ROW_SACV_FINAL: row_sacv_final(c, a, s, r, lane, nl); break;
    case ROW_TD3_PIBWD_DA: row_td3_pibwd_da(c, a, s, r, lane, nl); break;
    case ROW_SAC_FINAL: row_sac_final(c, a, s, r, lane, nl); break;
    case ROW_TD3_THEAD: row_td3_thead(c, a, s, r, lane, nl); break;
    case ROW_TD3_TARGET: row_td3_target(c, a, s, r, lane, nl); break;
    case ROW_TD3_PHEAD: row_td3_phead(c, a, s, r, lane, nl); break;
    case ROW_TD3_PLOSS: row_td3_ploss(c, a, s, r, lane, nl); break;
    case ROW_TD3_PIBWD: row_td3_pibwd(c, a, s, r, lane, nl); break;
    case ROW_TD3_FINAL: row_td3_final(c, a, s, r, lane, nl); break;
    case ROW_TD3_FINAL_POLICY: row_td3_final_policy(c, a, s, r, lane, nl); break;
    case ROW_DISC_GATHER: row_disc_gather(c, a, s, r, lane, nl); break;
    case ROW_DISC_HEAD: row_disc_head(c, a, s, r, lane, nl); break;
    case ROW_DISC_GNORM: row_disc_gnorm(c, a, s, r, lane, nl); break;
    case ROW_DISC_EW1: row_disc_ew1(c, a, s, r, lane, nl); break;
    case ROW_DISC_EW2: row_disc_ew2(c, a, s, r, lane, nl); break;
    case ROW_DISC_EW3: row_disc_ew3(c, a, s, r, lane, nl); break;
    case ROW_DISC_FINAL: row_disc_final(c, a, s, r, lane, nl); break;
    case ROW_DISC_REWARD: row_disc_reward(c, a, s, r, lane, nl); break;
    case ROW_DISC_REWARD_FINAL: row_disc_reward_final(c, a, s, r, lane, nl); break;
    default: break;
  }
}

ILSW_HD bool phase_active(const Phase& ph, const Hyper& hp, const RunArgs& a, int s) {
  if ((ph.cond & COND_DISC_PART) && a.update_mode == UPDATE_POLICY_ONLY) return false;
  if ((ph.cond & COND_POLICY_PART) && a.update_mode == UPDATE_DISC_ONLY) return false;
  if ((ph.cond & COND_FIRST_STEP) && s != 0) return false;
  if ((ph.cond & COND_WORLD_1) && a.world > 1) return false;
  if ((ph.cond & COND_WORLD_N) && a.world <= 1) return false;
  if (ph.cond & COND_TD3_POLICY) {
    int per = hp.period > 0 ? hp.period : 1;
    return ((a.step0 + s) % per) == 0;
  }
  if (ph.cond & COND_TD3_POLICY_OR_STATS) {
    int per = hp.period > 0 ? hp.period : 1;
    return ((a.step0 + s) % per) == 0 || s == a.stats_step;
  }
  return true;
}

// Host mirror of the counters a launch of n_steps advanced (shared by the C ABI and the test-only host simulator):
// a disc-only launch steps only the discriminator's Adam, a policy-only launch everything else.
inline void commit_counters(int (&t)[kMaxNets], int& n_total, const RunArgs& a, const Hyper& hp, int n_steps) {
  for (int slot = 0; slot < kMaxNets; ++slot) {
    const bool disc_slot = slot == SLOT_DISC;
    if (a.update_mode == UPDATE_DISC_ONLY && !disc_slot) continue;
    if (a.update_mode == UPDATE_POLICY_ONLY && disc_slot) continue;
    t[slot] = adam_t(a, hp, slot, n_steps - 1);
  }
  n_total += n_steps;
}

}  // namespace ilsw
