// ilswiss_b200 -- host-side compiler from (algorithm, shapes, hyper-parameters) to the phase
// program executed by the persistent engine kernel.  Header-only so that the product library
// (device pointers) and the test-only host simulator (host pointers) build the identical
// program.  See ilsw_types.h for the program model.
#pragma once
#include <cstdio>
#include <cstring>
#include <string>

#include "../../include/ilswiss_b200.h"
#include "ilsw_types.h"

namespace ilsw {

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
constexpr int kDiscGradSplits = kMaxGradSplits;     // partial gradient arenas of the discriminator (all GEMM modes but the exact one)
constexpr int kMaxKSplits = kMaxGradSplits;   // partial gradient arenas of a tcgen05 program (adam_grad sums exactly this many)

// two-pass bump allocator: pass 1 (base==nullptr) measures, pass 2 hands out pointers
struct Bump {
  char* base = nullptr;
  size_t off = 0;
  template <class T>
  T* take(size_t n) {
    size_t bytes = (n * sizeof(T) + 255) / 256 * 256;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += bytes;
    return p;
  }
  float* f(size_t n) { return take<float>(n); }
};

inline int mlp_num_params(int in_dim, int hid, int out_dim, int heads2) {
  int n = hid * in_dim + hid + hid * hid + hid + out_dim * hid + out_dim;
  if (heads2) n += out_dim * hid + out_dim;
  return n;
}

inline MlpPtrs make_mlp(const ilsw_mlp& n, float* grad) {
  MlpPtrs m;
  memset(&m, 0, sizeof(m));
  m.p = n.p; m.m = n.m; m.v = n.v; m.g = grad;
  m.in_dim = n.in_dim; m.hid = n.hidden; m.out_dim = n.out_dim; m.heads = n.log_std_head ? 2 : 1;
  m.oW0 = 0;
  m.ob0 = m.oW0 + n.hidden * n.in_dim;
  m.oW1 = m.ob0 + n.hidden;
  m.ob1 = m.oW1 + n.hidden * n.hidden;
  m.oW2 = m.ob1 + n.hidden;
  m.ob2 = m.oW2 + n.out_dim * n.hidden;
  m.oW3 = m.ob2 + n.out_dim;
  m.ob3 = m.oW3 + (n.log_std_head ? n.out_dim * n.hidden : 0);
  m.n_params = mlp_num_params(n.in_dim, n.hidden, n.out_dim, n.log_std_head);
  return m;
}

struct Builder {
  Program& P;
  explicit Builder(Program& p) : P(p) { P.n_phases = 0; P.n_ops = 0; }
  bool overflow = false;
  int base_cond = COND_ALWAYS;   // OR-ed into every phase (COND_DISC_PART / COND_POLICY_PART of AdvIRL programs)

  void phase(int cond = COND_ALWAYS, int collective = 0, int push = 0) {
    if (P.n_phases >= kMaxPhases) { overflow = true; return; }
    Phase& ph = P.phases[P.n_phases++];
    ph.op_begin = P.n_ops; ph.op_count = 0; ph.total_jobs = 0; ph.cond = cond | base_cond; ph.collective = collective; ph.push = push;
  }
  Op* add(int kind, int n_jobs) {
    if (P.n_ops >= kMaxOps || P.n_phases == 0) { overflow = true; return nullptr; }
    Op& o = P.ops[P.n_ops++];
    memset(&o, 0, sizeof(o));
    o.kind = kind; o.n_jobs = n_jobs;
    Phase& ph = P.phases[P.n_phases - 1];
    ph.op_count++; ph.total_jobs += n_jobs;
    return &o;
  }
  // tcgen05/TMA tile (ilsw_tc5.cuh): both operands must be TMA-addressable (16-byte aligned base and row stride); the
  // skinny shapes (M <= 8 weight gradients) and accumulating GEMMs keep their own tiles
  static bool tc5_operand_ok(const float* base, int ld) { return (reinterpret_cast<uintptr_t>(base) & 15) == 0 && (ld & 3) == 0; }
  bool tc5_eligible(const GemmOp& g) const {
    if (!P.ctx.hp.use_tc5 || g.accumulate || g.C2) return false;
    if (g.M <= 8 && g.a_mc && g.b_nc) return false;
    if (g.aug_ones && !(g.a_mc && g.b_nc)) return false;
    return tc5_operand_ok(g.A, g.lda) && tc5_operand_ok(g.B, g.ldb);
  }
  void gemm(GemmOp g) {
    g.tc5 = tc5_eligible(g) ? 1 : 0;
    g.tmapA = g.tmapB = nullptr;
    if (g.tc5) { g.tiles_m = (g.M + 127) / 128; g.tiles_n = (g.N + kTc5BN - 1) / kTc5BN; }
    else { g.tiles_m = (g.M + 31) / 32; g.tiles_n = (g.N + 31) / 32; }   // the aug (bias-gradient) column is produced by the tn==0 tiles
    const bool skinny = g.M <= 8 && g.a_mc && g.b_nc && g.tiles_m == 1;        // gemm_is_skinny (ilsw_engine.cuh)
    // K splits: tcgen05 tiles, skinny tiles of tcgen05 programs, and plain weight-gradient tiles (k-major operands, optimiser not fused)
    const bool split_ok = g.tc5 || skinny || (g.a_mc && g.b_nc && !g.adam && P.ctx.hp.gemm_precision != 0);
    if (!split_ok || g.ksplit < 1) g.ksplit = 1;
    if (skinny && !P.ctx.hp.use_tc5) g.ksplit = 1;          // gemm_tile_skinny_split exists in the tcgen05 engine variant only
    g.ksplit = g.ksplit < (g.K + 63) / 64 ? g.ksplit : (g.K + 63) / 64;
    g.kind = g.tc5 ? GK_TC5 : skinny ? (g.ksplit > 1 ? GK_SKINNY_SPLIT : GK_SKINNY) : (g.ksplit > 1 ? GK_TILE_SPLIT : GK_TILE);
    Op* o = add(OP_GEMM, g.ksplit * g.tiles_m * g.tiles_n);
    if (o) o->gemm = g;
  }
  // the aligned copy of a first-layer weight matrix, if `W` is one that has a copy (ShadowRef in ilsw_types.h)
  const MlpPtrs* shadowed_net(const float* W) const {
    const Ctx& c = P.ctx;
    const MlpPtrs* nets[] = {&c.policy, &c.qf[0], &c.qf[1], &c.tqf[0], &c.tqf[1], &c.tpolicy, &c.vf, &c.tvf, &c.disc};
    for (const MlpPtrs* n : nets)
      if (n->p && n->w0p && W == n->p + n->oW0) return n;
    return nullptr;
  }
  static ShadowRef shadow_of(const MlpPtrs* n) {
    ShadowRef r; r.ptr = nullptr; r.in = r.ld = r.n = 0; r.inv_in = 0.f;
    if (n && n->w0p) { r.ptr = n->w0p; r.in = n->in_dim; r.ld = n->ld_w0p; r.n = n->hid * n->in_dim; r.inv_in = 1.0f / (float)n->in_dim; }
    return r;
  }
  void shadow_refresh(const MlpPtrs& n) {
    if (!n.p || !n.w0p) return;
    Op* o = add(OP_SHADOW, (n.hid * n.in_dim + kAdamChunk - 1) / kAdamChunk);
    if (o) { o->shadow.src = n.p + n.oW0; o->shadow.dst = shadow_of(&n); }
  }
  // Y[M,N] = act(X[M,K] W[N,K]^T + b)          (N1: networks.py:85-101)
  void fwd(const float* X, int ldx, int M, int K, const float* W, const float* b, int N, float* Y, int ldy, int act) {
    GemmOp g; memset(&g, 0, sizeof(g));
    g.A = X; g.lda = ldx; g.a_mc = 0; g.B = W; g.ldb = K; g.b_nc = 0; g.M = M; g.N = N; g.K = K;
    if (const MlpPtrs* sn = shadowed_net(W)) { g.B = sn->w0p; g.ldb = sn->ld_w0p; }   // packed W0 rows are not TMA-addressable
    g.C = Y; g.ldc = ldy; g.bias = b; g.act = act;
    gemm(g);
  }
  // Two layers in one phase: H0 = act0(X W0^T + b0) [M, hid] is produced inside the tiles of Y = act(H0 W1^T + b1)
  // (GemmOp::a0) when the first layer's input is narrow; otherwise two phases.  H0 is materialised for the backward pass.
  bool fuse_l0_ok(int K0, int hid) const {
    return P.ctx.hp.fuse_l0 && !P.ctx.hp.use_tc5 && K0 <= kFuseL0MaxK && hid <= 256 && (hid & 1) == 0 && P.ctx.s.B < 512;
  }
  void fwd2_fused(const float* X, int ldx, int M, int K0, const float* W0, const float* b0, int act0, int hid, float* H0,
                  const float* W1, const float* b1, int N, float* Y, int ldy, int act) {
    const int fidx = P.n_ops;
    Op* fo = add(OP_L0FUSE, 0);
    if (!fo) return;
    fo->l0.X = X; fo->l0.ldx = ldx; fo->l0.K0 = K0; fo->l0.W = W0; fo->l0.b = b0; fo->l0.act = act0; fo->l0.out = H0; fo->l0.ldo = hid;
    GemmOp g; memset(&g, 0, sizeof(g));
    g.A = H0; g.lda = hid; g.a_mc = 0; g.B = W1; g.ldb = hid; g.b_nc = 0; g.M = M; g.N = N; g.K = hid;
    g.C = Y; g.ldc = ldy; g.bias = b1; g.act = act;
    g.a0 = fidx + 1;
    gemm(g);
  }
  // the first two layers of several networks, side by side: one fused phase when every first layer is narrow, else a
  // phase of first layers followed by a phase of second layers.  Opens its own phase(s); the caller may add further
  // ops to the last one.
  struct TwoLayers { const float* X; int ldx, M, K0; const MlpPtrs* net; float* H0; float* H1; int act; };
  void two_layers(const TwoLayers* L, int n, int cond = COND_ALWAYS) {
    bool fz = true;
    for (int i = 0; i < n; ++i) fz = fz && fuse_l0_ok(L[i].K0, L[i].net->hid);
    phase(cond);
    if (fz) {
      for (int i = 0; i < n; ++i) {
        const MlpPtrs& N = *L[i].net;
        fwd2_fused(L[i].X, L[i].ldx, L[i].M, L[i].K0, N.p + N.oW0, N.p + N.ob0, L[i].act, N.hid, L[i].H0, N.p + N.oW1, N.p + N.ob1,
                   N.hid, L[i].H1, N.hid, L[i].act);
      }
      return;
    }
    for (int i = 0; i < n; ++i) {
      const MlpPtrs& N = *L[i].net;
      fwd(L[i].X, L[i].ldx, L[i].M, L[i].K0, N.p + N.oW0, N.p + N.ob0, N.hid, L[i].H0, N.hid, L[i].act);
    }
    phase(cond);
    for (int i = 0; i < n; ++i) {
      const MlpPtrs& N = *L[i].net;
      fwd(L[i].H0, N.hid, L[i].M, N.hid, N.p + N.oW1, N.p + N.ob1, N.hid, L[i].H1, N.hid, L[i].act);
    }
  }
  // dX[M,Nin] = (D[M,Kred] W[Kred, ldw(:Nin)]) (.) act'(Hm)      (backward through a Linear)
  void dx(const float* D, int ldd, int M, int Kred, const float* W, int ldw, int Nin, const float* Hm, int ldh,
          int mask, float* out, int ldo, float* raw = nullptr) {
    GemmOp g; memset(&g, 0, sizeof(g));
    g.A = D; g.lda = ldd; g.a_mc = 0; g.B = W; g.ldb = ldw; g.b_nc = 1; g.M = M; g.N = Nin; g.K = Kred;
    g.C = out; g.ldc = ldo; g.C2 = raw; g.H = Hm; g.ldh = ldh; g.mask = mask;
    gemm(g);
  }
  // G[Mout,Nin] (+)= D[Kb,Mout]^T X[Kb,Nin] ; gbias[Mout] (+)= colsum(D)   (weight gradient)
  // adam_op >= 0: index of an adam_desc() op -- the epilogue applies the optimiser step to what it produces
  // net != nullptr (tcgen05 programs): the GEMM may be split along K over net->g_splits partial gradient arenas
  int dw_splits = 1;    // K splits requested for the following dw() calls (set per phase by the program builders)
  void dw(const float* D, int ldd, int Mout, const float* X, int ldx, int Nin, int Kb, float* G, float* gbias,
          int accumulate = 0, int adam_op = -1, const MlpPtrs* net = nullptr) {
    GemmOp g; memset(&g, 0, sizeof(g));
    g.A = D; g.lda = ldd; g.a_mc = 1; g.B = X; g.ldb = ldx; g.b_nc = 1; g.M = Mout; g.N = Nin; g.K = Kb;
    g.C = G; g.ldc = Nin; g.aug_ones = gbias ? 1 : 0; g.bias_out = gbias; g.accumulate = accumulate;
    g.adam = adam_op >= 0 ? adam_op + 1 : 0;
    if (net && net->g_splits > 1 && adam_op < 0 && (!accumulate || !P.ctx.hp.use_tc5)) {     // mma.sync tiles accumulate per arena
      g.ksplit = dw_splits < net->g_splits ? dw_splits : net->g_splits;
      g.split_stride = net->g_stride;
    }
    gemm(g);
  }
  // K splits that bring a weight-gradient phase of `tiles` tcgen05 tiles close to one wave of the 148-SM grid
  static int pick_splits(int tiles, int Kb, int max_splits) {
    int s = tiles > 0 ? 148 / tiles : 1;
    const int nkb = (Kb + 31) / 32;
    if (s > nkb / 4) s = nkb / 4;          // at least 4 K blocks (one pipeline depth) per split
    if (s > max_splits) s = max_splits;
    return s < 1 ? 1 : s;
  }
  static int tc5_tiles(int M, int N) { return ((M + 127) / 128) * ((N + kTc5BN - 1) / kTc5BN); }
  // next-step row op that runs with the last active phase of every step (Ctx::tail_op1).  Must be the LAST op added: the ops
  // of a phase are contiguous, and this one belongs to none.
  void tail_row(int kind, int rows) {
    if (P.n_ops >= kMaxOps) { overflow = true; return; }
    Op& o = P.ops[P.n_ops++];
    memset(&o, 0, sizeof(o));
    o.kind = OP_ROW; o.n_jobs = (rows + kRowsPerJob - 1) / kRowsPerJob;
    o.row.kind = kind; o.row.rows = rows; o.row.arg0 = 1; o.row.arg1 = 0;
    P.ctx.tail_op1 = P.n_ops;
  }
  void row(int kind, int rows, int next_step = 0, int row_offset = 0) {
    Op* o = add(OP_ROW, (rows + kRowsPerJob - 1) / kRowsPerJob);
    if (o) { o->row.kind = kind; o->row.rows = rows + row_offset; o->row.arg0 = next_step; o->row.arg1 = row_offset; }
  }
  // descriptor for fused GEMM epilogues (no jobs of its own); returns its op index
  int adam_desc(const MlpPtrs& n, const MlpPtrs* target, double lr, double b1, double b2, double eps, float tau, int slot) {
    const int idx = P.n_ops;
    Op* o = add(OP_ADAM, 0);
    if (!o) return -1;
    o->adam.p = n.p; o->adam.g = n.g; o->adam.m = n.m; o->adam.v = n.v;
    o->adam.target = target ? target->p : nullptr;
    o->adam.n = n.n_params; o->adam.lr = lr; o->adam.beta1 = b1; o->adam.beta2 = b2; o->adam.eps = eps;
    o->adam.tau = tau; o->adam.slot = slot; o->adam.grad_scale_world = 0; o->adam.begin = 0; o->adam.fused_only = 1;
    o->adam.sh_p = shadow_of(&n); o->adam.sh_t = shadow_of(target);
    o->adam.g_splits = n.g_splits; o->adam.g_split_stride = n.g_stride;
    return idx;
  }
  void adam(const MlpPtrs& n, const MlpPtrs* target, double lr, double b1, double b2, double eps, float tau, int slot,
            int world_scale = 0) {
    Op* o = add(OP_ADAM, (n.n_params + kAdamChunk - 1) / kAdamChunk);
    if (!o) return;
    o->adam.p = n.p; o->adam.g = n.g; o->adam.m = n.m; o->adam.v = n.v;
    o->adam.target = target ? target->p : nullptr;
    o->adam.n = n.n_params; o->adam.lr = lr; o->adam.beta1 = b1; o->adam.beta2 = b2; o->adam.eps = eps;
    o->adam.tau = tau; o->adam.slot = slot; o->adam.grad_scale_world = world_scale;
    o->adam.sh_p = shadow_of(&n); o->adam.sh_t = shadow_of(target);
    o->adam.g_splits = n.g_splits; o->adam.g_split_stride = n.g_stride;
  }
  void polyak(const MlpPtrs& src, const MlpPtrs& tgt, float tau) {
    Op* o = add(OP_POLYAK, (src.n_params + kAdamChunk - 1) / kAdamChunk);
    if (!o) return;
    o->polyak.target = tgt.p; o->polyak.src = src.p; o->polyak.n = src.n_params; o->polyak.tau = tau;
    o->polyak.sh_t = shadow_of(&tgt);
  }
};

// ------------------------------------------------------------------------------------------
// scratch layout
// ------------------------------------------------------------------------------------------
inline void alloc_sac_bufs(Bump& mem, SacBufs& S, int algo, int B, int O, int A, int Hd) {
  S.B = B; S.O = O; S.A = A; S.Hd = Hd;
  S.ld_oa = round_up(O + A, 4);
  S.ld_o = round_up(O, 4);
  const bool td3 = algo == ILSW_ALGO_TD3;
  const int PB = td3 ? B : 2 * B;  // policy rows per step
  S.idx = mem.take<int>(B);
  S.Xoa = mem.f((size_t)B * S.ld_oa); S.rew = mem.f(B); S.term = mem.f(B);
  S.Xpi = mem.f((size_t)2 * B * S.ld_o);
  S.Xna = mem.f((size_t)B * S.ld_oa); S.Xon = mem.f((size_t)B * S.ld_oa);
  S.eps = mem.f((size_t)2 * B * A); S.noise = mem.f((size_t)B * A);
  for (int i = 0; i < 2; ++i) {
    S.h0q[i] = mem.f((size_t)B * Hd); S.h1q[i] = mem.f((size_t)B * Hd);
    S.h0t[i] = mem.f((size_t)B * Hd); S.h1t[i] = mem.f((size_t)B * Hd);
    S.h0n[i] = mem.f((size_t)B * Hd); S.h1n[i] = mem.f((size_t)B * Hd);
    S.qp[i] = mem.f(B); S.tq[i] = mem.f(B); S.qn[i] = mem.f(B); S.dq[i] = mem.f(B);
    S.d1q[i] = mem.f((size_t)B * Hd); S.d0q[i] = mem.f((size_t)B * Hd);
    S.e1[i] = mem.f((size_t)B * Hd); S.e0[i] = mem.f((size_t)B * Hd);
    S.dA[i] = mem.f((size_t)B * A);
    S.lossterm[i] = mem.f(B);
    S.qn_old[i] = mem.f(B);
  }
  S.h0p = mem.f((size_t)PB * Hd); S.h1p = mem.f((size_t)PB * Hd);
  S.mean = mem.f((size_t)PB * A); S.lraw = mem.f((size_t)PB * A); S.lstd = mem.f((size_t)PB * A);
  S.act = mem.f((size_t)PB * A); S.logpi = mem.f(PB);
  S.y = mem.f(B);
  S.dmean = mem.f((size_t)B * A); S.dlraw = mem.f((size_t)B * A);
  S.d1p = mem.f((size_t)B * Hd); S.d0p = mem.f((size_t)B * Hd);
  S.plterm = mem.f(B); S.regmu = mem.f(B); S.regls = mem.f(B); S.aterm = mem.f(B);
  S.h0tp = mem.f((size_t)B * Hd); S.h1tp = mem.f((size_t)B * Hd);
  S.h0v = S.h1v = S.h0tv = S.h1tv = S.vp = S.tv = S.dv = S.d1v = S.d0v = S.lossterm_v = nullptr;
  if (algo == ILSW_ALGO_SAC_V) {
    S.h0v = mem.f((size_t)B * Hd); S.h1v = mem.f((size_t)B * Hd); S.h0tv = mem.f((size_t)B * Hd); S.h1tv = mem.f((size_t)B * Hd);
    S.vp = mem.f(B); S.tv = mem.f(B); S.dv = mem.f(B); S.lossterm_v = mem.f(B);
    S.d1v = mem.f((size_t)B * Hd); S.d0v = mem.f((size_t)B * Hd);
  }
}

inline void alloc_disc_bufs(Bump& mem, DiscBufs& D, int B, int Din, int Hd) {
  D.B = B; D.D = Din; D.Hd = Hd; D.ld_d = round_up(Din, 4);
  D.idx_e = mem.take<int>(B); D.idx_p = mem.take<int>(B);
  D.X3 = mem.f((size_t)3 * B * D.ld_d); D.gp_eps = mem.f(B);
  D.h1 = mem.f((size_t)3 * B * Hd); D.h2 = mem.f((size_t)3 * B * Hd);
  D.y = mem.f(3 * B); D.dlogit = mem.f(2 * B); D.cmask = mem.f(B);
  D.d2 = mem.f((size_t)2 * B * Hd); D.d1 = mem.f((size_t)2 * B * Hd);
  D.dl2 = mem.f((size_t)B * Hd); D.u1 = mem.f((size_t)B * Hd); D.dl1 = mem.f((size_t)B * Hd);
  D.g = mem.f((size_t)B * D.ld_d); D.gbar = mem.f((size_t)B * D.ld_d); D.nrm = mem.f(B);
  D.db1 = mem.f((size_t)B * Hd); D.ub1 = mem.f((size_t)B * Hd); D.sb1 = mem.f((size_t)B * Hd);
  D.db2 = mem.f((size_t)B * Hd); D.t3 = mem.f((size_t)B * Hd); D.zb2 = mem.f((size_t)B * Hd);
  D.hb1 = mem.f((size_t)B * Hd); D.zb1 = mem.f((size_t)B * Hd);
  D.ceterm = mem.f(2 * B); D.accterm = mem.f(2 * B); D.gpterm = mem.f(B);
  D.rh1 = mem.f((size_t)B * Hd); D.rh2 = mem.f((size_t)B * Hd); D.rewraw = mem.f(B);
  D.ld_sn = D.ld_d; D.Xsn = mem.f((size_t)B * D.ld_sn);
}

inline int stats_floats_for(int algo, int B, int A) {
  return algo == ILSW_ALGO_TD3 ? 6 * B + B * A : 7 * B + 2 * B * A;   // SAC-alpha and SAC-V share a layout
}

// ------------------------------------------------------------------------------------------
// program builders
// ------------------------------------------------------------------------------------------
// D1: discriminator update (adv_irl.py:133-216) incl. the closed-form double backward of the
// gradient penalty (formula block: SURVEY.md section 8a row D1).
inline void build_disc_step(Builder& b, const Ctx& c) {
  const DiscBufs& D = c.d;
  const MlpPtrs& N = c.disc;
  const int B = D.B, Hd = D.Hd, Din = D.D, ld = D.ld_d;
  const bool gp = c.hp.use_gp != 0;
  const int act = c.hp.disc_act;
  // ReLU blocks have no second derivative: every term of the penalty's double backward that carries act'' vanishes
  // (zbar2 = 0, hence hbar1 = 0 and zbar1 = 0), so the three GEMMs and the row phase that only move those zeros are left out
  const bool curved = act == ACT_TANH;
  const int R = gp ? 3 * B : 2 * B;
  const float* W1 = N.p + N.oW0; const float* b1 = N.p + N.ob0;
  const float* W2 = N.p + N.oW1; const float* b2 = N.p + N.ob1;
  float* G1 = N.g + N.oW0; float* gb1 = N.g + N.ob0;
  float* G2 = N.g + N.oW1; float* gb2 = N.g + N.ob1;
  float* G3 = N.g + N.oW2; float* gb3 = N.g + N.ob2;
  const float* xhat = D.X3 + (size_t)2 * B * ld;
  const float* h1i = D.h1 + (size_t)2 * B * Hd;

  b.phase(); b.row(ROW_DISC_GATHER, B);
  { Builder::TwoLayers L[1] = {{D.X3, ld, R, Din, &N, D.h1, D.h2, act}}; b.two_layers(L, 1); }
  b.phase(); b.row(ROW_DISC_HEAD, R);
  b.phase();
  b.dx(D.d2, Hd, 2 * B, Hd, W2, Hd, Hd, D.h1, Hd, act, D.d1, Hd);               // CE: delta1
  b.dw(D.d2, Hd, Hd, D.h1, Hd, Hd, 2 * B, G2, gb2);                             // CE: dW2,db2
  b.dw(D.dlogit, 1, 1, D.h2, Hd, Hd, 2 * B, G3, gb3);                            // CE: dw3,db3
  if (gp) b.dx(D.dl2, Hd, B, Hd, W2, Hd, Hd, h1i, Hd, act, D.dl1, Hd, D.u1);       // u1, delta1_gp
  b.phase();
  // W1's gradient is a 128 x Din output over K = 2B / B rows: 4 tiles.  Split along K over the partial arenas (the flat Adam
  // job sums them); the first, non-accumulating GEMM uses every arena, so the accumulating ones below find fresh partials.
  b.dw_splits = kDiscGradSplits;
  b.dw(D.d1, Hd, Hd, D.X3, ld, Din, 2 * B, G1, gb1, 0, -1, &N);                  // CE: dW1,db1
  if (gp) b.dx(D.dl1, Hd, B, Hd, W1, Din, Din, nullptr, 0, ACT_NONE, D.g, ld);   // g = delta1 W1
  if (gp) {
    b.phase(); b.row(ROW_DISC_GNORM, B);
    b.phase();
    b.fwd(D.gbar, ld, B, Din, W1, nullptr, Hd, D.db1, Hd, ACT_NONE);             // dbar1 = gbar W1^T
    b.dw(D.dl1, Hd, Hd, D.gbar, ld, Din, B, G1, nullptr, 1, -1, &N);             // dW1 += delta1^T gbar
    b.phase(); b.row(ROW_DISC_EW1, B);
    b.phase();
    b.fwd(D.ub1, Hd, B, Hd, W2, nullptr, Hd, D.db2, Hd, ACT_NONE);               // dbar2 = ubar1 W2^T
    b.dw(D.dl2, Hd, Hd, D.ub1, Hd, Hd, B, G2, nullptr, 1);                       // dW2 += delta2^T ubar1
    b.phase(); b.row(ROW_DISC_EW2, B);
    b.phase();
    if (curved) b.dw(D.zb2, Hd, Hd, h1i, Hd, Hd, B, G2, gb2, 1);                 // dW2 += zbar2^T h1 ; db2
    b.dw(D.cmask, 1, 1, D.t3, Hd, Hd, B, G3, nullptr, 1);                        // dw3 += c^T (dbar2*s2)
    if (curved) {
      b.dx(D.zb2, Hd, B, Hd, W2, Hd, Hd, nullptr, 0, ACT_NONE, D.hb1, Hd);       // hbar1_raw = zbar2 W2
      b.phase(); b.row(ROW_DISC_EW3, B);
      b.phase(); b.dw(D.zb1, Hd, Hd, xhat, ld, Din, B, G1, gb1, 1, -1, &N);      // dW1 += zbar1^T xhat ; db1
    }
  }
  b.phase();
  b.adam(N, nullptr, c.hp.disc_lr, c.hp.disc_beta1, 0.999, c.hp.adam_eps, 0.f, SLOT_DISC);
  b.row(ROW_DISC_FINAL, 1);
}

// S1 (+ D2 reward relabel when a discriminator is attached)
// Two critics' backward-data tiles plus their skinny output-layer gradient tiles fit one wave of a 148-SM grid?
inline bool early_out_layer_grad(int B, int Hd) {
  const int dx_tiles = 2 * ((B + 31) / 32) * ((Hd + 31) / 32), skinny_tiles = 2 * ((Hd + 31) / 32);
  return dx_tiles + skinny_tiles <= 148;
}

inline void build_sac_alpha(Builder& b, const Ctx& c) {
  const SacBufs& S = c.s;
  const int B = S.B, O = S.O, A = S.A, Hd = S.Hd, K0 = O + A;
  const MlpPtrs& P = c.policy;
  const bool disc = c.hp.has_disc != 0;
  const double b1 = c.hp.beta1, b2 = c.hp.beta2, eps = c.hp.adam_eps;
  const float* obs_rows = S.Xpi + (size_t)B * S.ld_o;
  const float* h0p_obs = S.h0p + (size_t)B * Hd;
  const float* h1p_obs = S.h1p + (size_t)B * Hd;

  // the batch of step s+1 is gathered in the LAST phase of step s (independent of everything that
  // phase does); only the first step of a launch needs a gather phase of its own
  b.phase(COND_FIRST_STEP); b.row(ROW_SAC_GATHER, B);
  // narrow first layers (Hopper / Walker: 14 .. 23 inputs) are folded into the tiles of the second layer (fwd2_fused)
  const bool fz = b.fuse_l0_ok(K0, Hd) && (!disc || b.fuse_l0_ok(c.hp.state_only ? 2 * O : K0, c.d.Hd));
  const float* dX = (disc && c.hp.state_only) ? c.d.Xsn : S.Xoa;
  const int dld = (disc && c.hp.state_only) ? c.d.ld_sn : S.ld_oa, dK = (disc && c.hp.state_only) ? 2 * O : K0;
  b.phase();
  if (fz) {
    for (int i = 0; i < 2; ++i)
      b.fwd2_fused(S.Xoa, S.ld_oa, B, K0, c.qf[i].p + c.qf[i].oW0, c.qf[i].p + c.qf[i].ob0, ACT_RELU, Hd, S.h0q[i],
                   c.qf[i].p + c.qf[i].oW1, c.qf[i].p + c.qf[i].ob1, Hd, S.h1q[i], Hd, ACT_RELU);
    b.fwd2_fused(S.Xpi, S.ld_o, 2 * B, O, P.p + P.oW0, P.p + P.ob0, ACT_RELU, Hd, S.h0p, P.p + P.oW1, P.p + P.ob1, Hd, S.h1p, Hd, ACT_RELU);
    if (disc)
      b.fwd2_fused(dX, dld, B, dK, c.disc.p + c.disc.oW0, c.disc.p + c.disc.ob0, c.hp.disc_act, c.d.Hd, c.d.rh1,
                   c.disc.p + c.disc.oW1, c.disc.p + c.disc.ob1, c.d.Hd, c.d.rh2, c.d.Hd, c.hp.disc_act);
  } else {
    for (int i = 0; i < 2; ++i) b.fwd(S.Xoa, S.ld_oa, B, K0, c.qf[i].p + c.qf[i].oW0, c.qf[i].p + c.qf[i].ob0, Hd, S.h0q[i], Hd, ACT_RELU);
    b.fwd(S.Xpi, S.ld_o, 2 * B, O, P.p + P.oW0, P.p + P.ob0, Hd, S.h0p, Hd, ACT_RELU);
    if (disc) b.fwd(dX, dld, B, dK, c.disc.p + c.disc.oW0, c.disc.p + c.disc.ob0, c.d.Hd, c.d.rh1, c.d.Hd, c.hp.disc_act);
    b.phase();
    for (int i = 0; i < 2; ++i) b.fwd(S.h0q[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, c.qf[i].p + c.qf[i].ob1, Hd, S.h1q[i], Hd, ACT_RELU);
    b.fwd(S.h0p, Hd, 2 * B, Hd, P.p + P.oW1, P.p + P.ob1, Hd, S.h1p, Hd, ACT_RELU);
    if (disc) b.fwd(c.d.rh1, c.d.Hd, B, c.d.Hd, c.disc.p + c.disc.oW1, c.disc.p + c.disc.ob1, c.d.Hd, c.d.rh2, c.d.Hd, c.hp.disc_act);
  }
  b.phase();
  b.row(ROW_SAC_HEADS, 2 * B);
  if (disc) b.row(ROW_DISC_REWARD, B);
  b.phase();
  if (fz) {
    for (int i = 0; i < 2; ++i)
      b.fwd2_fused(S.Xna, S.ld_oa, B, K0, c.tqf[i].p + c.tqf[i].oW0, c.tqf[i].p + c.tqf[i].ob0, ACT_RELU, Hd, S.h0t[i],
                   c.tqf[i].p + c.tqf[i].oW1, c.tqf[i].p + c.tqf[i].ob1, Hd, S.h1t[i], Hd, ACT_RELU);
    if (disc) b.row(ROW_DISC_REWARD_FINAL, 1);
  } else {
    for (int i = 0; i < 2; ++i) b.fwd(S.Xna, S.ld_oa, B, K0, c.tqf[i].p + c.tqf[i].oW0, c.tqf[i].p + c.tqf[i].ob0, Hd, S.h0t[i], Hd, ACT_RELU);
    if (disc) b.row(ROW_DISC_REWARD_FINAL, 1);
    b.phase();
    for (int i = 0; i < 2; ++i) b.fwd(S.h0t[i], Hd, B, Hd, c.tqf[i].p + c.tqf[i].oW1, c.tqf[i].p + c.tqf[i].ob1, Hd, S.h1t[i], Hd, ACT_RELU);
  }
  b.phase(); b.row(ROW_SAC_TARGET, B);
  const bool tc5 = c.hp.use_tc5 != 0;
  if (tc5) {
    // tcgen05 programs: weight gradients are split along K (= batch) over partial arenas and the optimiser runs as flat
    // jobs over the whole grid (see ilsw_tc5.cuh); the Polyak update of the targets rides in the same Adam jobs
    b.phase();
    for (int i = 0; i < 2; ++i)
      b.dx(S.d1q[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0q[i], Hd, ACT_RELU, S.d0q[i], Hd);
    b.dw_splits = kMaxKSplits;       // skinny output-layer gradients: K = batch split over the partial arenas too
    for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, -1, &c.qf[i]);
    b.phase();
    b.dw_splits = Builder::pick_splits(2 * (Builder::tc5_tiles(Hd, Hd) + Builder::tc5_tiles(Hd, K0)), B, kMaxKSplits);
    for (int i = 0; i < 2; ++i) b.dw(S.d1q[i], Hd, Hd, S.h0q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW1, c.qf[i].g + c.qf[i].ob1, 0, -1, &c.qf[i]);
    for (int i = 0; i < 2; ++i) b.dw(S.d0q[i], Hd, Hd, S.Xoa, S.ld_oa, K0, B, c.qf[i].g + c.qf[i].oW0, c.qf[i].g + c.qf[i].ob0, 0, -1, &c.qf[i]);
    b.phase();
    for (int i = 0; i < 2; ++i) b.adam(c.qf[i], &c.tqf[i], c.hp.qf_lr, b1, b2, eps, c.hp.tau, i == 0 ? SLOT_QF1 : SLOT_QF2);
  } else {
    b.phase();   // backward-data of both critics: one full tile per CTA
    {
      int ad[2];
      for (int i = 0; i < 2; ++i) ad[i] = b.adam_desc(c.qf[i], &c.tqf[i], c.hp.qf_lr, b1, b2, eps, c.hp.tau, i == 0 ? SLOT_QF1 : SLOT_QF2);
      for (int i = 0; i < 2; ++i)
        b.dx(S.d1q[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0q[i], Hd, ACT_RELU, S.d0q[i], Hd);
      // the output-layer gradient (+ Adam/Polyak of W2, b2) needs only dq and h1 and nothing reads W2 any more this
      // step: when the backward-data phase leaves CTAs idle its skinny tiles run there, which keeps the weight-gradient
      // phase within one wave of the grid (160 -> 144 jobs at B = 256)
      const bool early_w2 = early_out_layer_grad(B, Hd);
      if (early_w2)
        for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, ad[i]);
      b.phase();   // the other weight gradients of both critics; Adam + Polyak of the targets fused into the tile epilogues
      for (int i = 0; i < 2; ++i) b.dw(S.d1q[i], Hd, Hd, S.h0q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW1, c.qf[i].g + c.qf[i].ob1, 0, ad[i]);
      for (int i = 0; i < 2; ++i) b.dw(S.d0q[i], Hd, Hd, S.Xoa, S.ld_oa, K0, B, c.qf[i].g + c.qf[i].oW0, c.qf[i].g + c.qf[i].ob0, 0, ad[i]);
      if (!early_w2)
        for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, ad[i]);
    }
  }
  b.phase();
  if (fz) {
    for (int i = 0; i < 2; ++i)
      b.fwd2_fused(S.Xon, S.ld_oa, B, K0, c.qf[i].p + c.qf[i].oW0, c.qf[i].p + c.qf[i].ob0, ACT_RELU, Hd, S.h0n[i],
                   c.qf[i].p + c.qf[i].oW1, c.qf[i].p + c.qf[i].ob1, Hd, S.h1n[i], Hd, ACT_RELU);
  } else {
    for (int i = 0; i < 2; ++i) b.fwd(S.Xon, S.ld_oa, B, K0, c.qf[i].p + c.qf[i].oW0, c.qf[i].p + c.qf[i].ob0, Hd, S.h0n[i], Hd, ACT_RELU);
    b.phase();
    for (int i = 0; i < 2; ++i) b.fwd(S.h0n[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, c.qf[i].p + c.qf[i].ob1, Hd, S.h1n[i], Hd, ACT_RELU);
  }
  b.phase(); b.row(ROW_SAC_PLOSS, B);
  b.phase();
  for (int i = 0; i < 2; ++i) b.dx(S.e1[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0n[i], Hd, ACT_RELU, S.e0[i], Hd);
  b.phase(); b.row(ROW_SAC_PIBWD_DA, B);      // dA = e0 . W0[:, O:O+A] fused into the head backward rows
  b.phase();
  b.dx(S.d1p, Hd, B, Hd, P.p + P.oW1, Hd, Hd, h0p_obs, Hd, ACT_RELU, S.d0p, Hd);
  if (tc5) {
    // gradients (split along K), then the exchange (replicas only) + flat Adam of the (averaged) gradient
    b.phase();
    b.dw_splits = Builder::pick_splits(Builder::tc5_tiles(Hd, Hd) + Builder::tc5_tiles(Hd, O) + (A > 8 ? 2 * Builder::tc5_tiles(A, Hd) : 0), B, kMaxKSplits);
    b.dw(S.d1p, Hd, Hd, h0p_obs, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1, 0, -1, &P);
    b.dw(S.d0p, Hd, Hd, obs_rows, S.ld_o, O, B, P.g + P.oW0, P.g + P.ob0, 0, -1, &P);
    b.dw(S.dmean, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2, 0, -1, &P);
    b.dw(S.dlraw, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW3, P.g + P.ob3, 0, -1, &P);
    b.phase(COND_ALWAYS, 1);
    b.adam(P, nullptr, c.hp.policy_lr, b1, b2, eps, 0.f, SLOT_POLICY, 1);
    b.row(ROW_SAC_FINAL, 1);
    b.row(ROW_SAC_GATHER, B, /*next_step=*/1);
    return;
  }
  // one replica: the policy's Adam step is fused into its weight-gradient tiles
  b.phase(COND_WORLD_1);
  {
    const int ad = b.adam_desc(P, nullptr, c.hp.policy_lr, b1, b2, eps, 0.f, SLOT_POLICY);
    b.dw(S.d1p, Hd, Hd, h0p_obs, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1, 0, ad);
    b.dw(S.d0p, Hd, Hd, obs_rows, S.ld_o, O, B, P.g + P.oW0, P.g + P.ob0, 0, ad);
    b.dw(S.dmean, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2, 0, ad);
    b.dw(S.dlraw, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW3, P.g + P.ob3, 0, ad);
  }
  b.phase(COND_WORLD_1);
  b.row(ROW_SAC_FINAL, 1);
  b.row(ROW_SAC_GATHER, B, /*next_step=*/1);
  // replicas: the weight-gradient tiles push what they produce into every replica's receive slot from their epilogues (the
  // NVLink transfer overlaps the rest of the phase), then the Adam phase waits for the peers' pushes and applies the
  // rank-ordered sum
  b.phase(COND_WORLD_N, 0, /*push=*/1);
  b.dw(S.d1p, Hd, Hd, h0p_obs, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1);
  b.dw(S.d0p, Hd, Hd, obs_rows, S.ld_o, O, B, P.g + P.oW0, P.g + P.ob0);
  b.dw(S.dmean, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2);
  b.dw(S.dlraw, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW3, P.g + P.ob3);
  b.phase(COND_WORLD_N, 2);
  b.adam(P, nullptr, c.hp.policy_lr, b1, b2, eps, 0.f, SLOT_POLICY, 1);
  b.row(ROW_SAC_FINAL, 1);
  b.row(ROW_SAC_GATHER, B, /*next_step=*/1);
}

// S3
inline void build_td3(Builder& b, const Ctx& c) {
  const SacBufs& S = c.s;
  const int B = S.B, O = S.O, A = S.A, Hd = S.Hd, K0 = O + A;
  const MlpPtrs& P = c.policy; const MlpPtrs& TP = c.tpolicy;
  const double b1 = 0.9, b2 = 0.999, eps = c.hp.adam_eps;   // optimizer defaults (td3.py:56-67)

  // tcgen05 programs (and only those: their last phase on every kind of step is a flat Adam / statistics phase that reads
  // none of the gathered buffers) gather the batch of step s + 1 alongside the last active phase of step s (Builder::tail_row):
  // 1024 random Humanoid rows are a 19 us phase of DRAM latency when gathered on the critical path
  const bool hoist_gather = c.hp.use_tc5 != 0;
  b.phase(hoist_gather ? COND_FIRST_STEP : COND_ALWAYS); b.row(ROW_TD3_GATHER, B);
  // HER-TD3 overwrites the target policy's action with clipped noise (her/td3.py:103-112): its forward pass is dead code
  const bool her = c.hp.her != 0;
  {
    Builder::TwoLayers L[3] = {{S.Xoa, S.ld_oa, B, K0, &c.qf[0], S.h0q[0], S.h1q[0], ACT_RELU},
                               {S.Xoa, S.ld_oa, B, K0, &c.qf[1], S.h0q[1], S.h1q[1], ACT_RELU},
                               {S.Xna, S.ld_oa, B, O, &TP, S.h0tp, S.h1tp, ACT_RELU}};
    b.two_layers(L, her ? 2 : 3);
  }
  b.phase(); b.row(ROW_TD3_THEAD, B);
  {
    Builder::TwoLayers L[2] = {{S.Xna, S.ld_oa, B, K0, &c.tqf[0], S.h0t[0], S.h1t[0], ACT_RELU},
                               {S.Xna, S.ld_oa, B, K0, &c.tqf[1], S.h0t[1], S.h1t[1], ACT_RELU}};
    b.two_layers(L, 2);
  }
  b.phase(); b.row(ROW_TD3_TARGET, B);
  const bool tc5 = c.hp.use_tc5 != 0;
  if (tc5) {        // see build_sac_alpha: split-K weight gradients, flat Adam jobs
    b.phase();
    for (int i = 0; i < 2; ++i)
      b.dx(S.d1q[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0q[i], Hd, ACT_RELU, S.d0q[i], Hd);
    b.dw_splits = kMaxKSplits;       // skinny output-layer gradients: K = batch split over the partial arenas too
    for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, -1, &c.qf[i]);
    b.phase();
    b.dw_splits = Builder::pick_splits(2 * (Builder::tc5_tiles(Hd, Hd) + Builder::tc5_tiles(Hd, K0)), B, kMaxKSplits);
    for (int i = 0; i < 2; ++i) b.dw(S.d1q[i], Hd, Hd, S.h0q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW1, c.qf[i].g + c.qf[i].ob1, 0, -1, &c.qf[i]);
    for (int i = 0; i < 2; ++i) b.dw(S.d0q[i], Hd, Hd, S.Xoa, S.ld_oa, K0, B, c.qf[i].g + c.qf[i].oW0, c.qf[i].g + c.qf[i].ob0, 0, -1, &c.qf[i]);
    b.phase();
    for (int i = 0; i < 2; ++i) b.adam(c.qf[i], nullptr, c.hp.qf_lr, b1, b2, eps, 0.f, i == 0 ? SLOT_QF1 : SLOT_QF2);
  } else {
  b.phase();
    {
      int ad[2];
      for (int i = 0; i < 2; ++i) ad[i] = b.adam_desc(c.qf[i], nullptr, c.hp.qf_lr, b1, b2, eps, 0.f, i == 0 ? SLOT_QF1 : SLOT_QF2);
      for (int i = 0; i < 2; ++i)
        b.dx(S.d1q[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0q[i], Hd, ACT_RELU, S.d0q[i], Hd);
      const bool early_w2 = early_out_layer_grad(B, Hd);      // see build_sac_alpha
      if (early_w2)
        for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, ad[i]);
      b.phase();   // weight gradients of both critics with the Adam step fused into the tile epilogues
      for (int i = 0; i < 2; ++i) b.dw(S.d1q[i], Hd, Hd, S.h0q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW1, c.qf[i].g + c.qf[i].ob1, 0, ad[i]);
      for (int i = 0; i < 2; ++i) b.dw(S.d0q[i], Hd, Hd, S.Xoa, S.ld_oa, K0, B, c.qf[i].g + c.qf[i].oW0, c.qf[i].g + c.qf[i].ob0, 0, ad[i]);
      if (!early_w2)
        for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, ad[i]);
    }
}
  b.row(ROW_TD3_FINAL, 1);
  // delayed policy + target update (td3.py:113-124), steps with (n_train_steps_total % period)==0
  const int PC = COND_TD3_POLICY;
  // the forward half also runs on the statistics step of a launch when that is not a policy step: the reference logs a
  // stats-only policy loss -mean(Q1(obs, pi(obs))) and its actions there (td3.py:131-136)
  const int PS = COND_TD3_POLICY_OR_STATS;
  { Builder::TwoLayers L[1] = {{S.Xoa, S.ld_oa, B, O, &P, S.h0p, S.h1p, ACT_RELU}}; b.two_layers(L, 1, PS); }
  b.phase(PS); b.row(ROW_TD3_PHEAD, B);
  { Builder::TwoLayers L[1] = {{S.Xon, S.ld_oa, B, K0, &c.qf[0], S.h0n[0], S.h1n[0], ACT_RELU}}; b.two_layers(L, 1, PS); }
  b.phase(PS); b.row(ROW_TD3_PLOSS, B);
  // the policy loss of this step is complete: log it here (policy steps and the statistics step alike; on a stats-only
  // step the backward-data tile of this phase is dead work, once per epoch)
  b.phase(PS); b.dx(S.e1[0], Hd, B, Hd, c.qf[0].p + c.qf[0].oW1, Hd, Hd, S.h0n[0], Hd, ACT_RELU, S.e0[0], Hd);
  b.row(ROW_TD3_FINAL_POLICY, 1);
  b.phase(PC); b.row(ROW_TD3_PIBWD_DA, B);
  b.phase(PC);
  b.dx(S.d1p, Hd, B, Hd, P.p + P.oW1, Hd, Hd, S.h0p, Hd, ACT_RELU, S.d0p, Hd);
  if (tc5) {
    b.phase(PC);
    b.dw_splits = Builder::pick_splits(Builder::tc5_tiles(Hd, Hd) + Builder::tc5_tiles(Hd, O) + (A > 8 ? Builder::tc5_tiles(A, Hd) : 0), B, kMaxKSplits);
    b.dw(S.d1p, Hd, Hd, S.h0p, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1, 0, -1, &P);
    b.dw(S.d0p, Hd, Hd, S.Xoa, S.ld_oa, O, B, P.g + P.oW0, P.g + P.ob0, 0, -1, &P);
    b.dw(S.dmean, A, A, S.h1p, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2, 0, -1, &P);
    b.phase(PC, 1);
    b.adam(P, &c.tpolicy, c.hp.policy_lr, b1, b2, eps, c.hp.tau, SLOT_POLICY, 1);
    b.polyak(c.qf[0], c.tqf[0], c.hp.tau);
    b.polyak(c.qf[1], c.tqf[1], c.hp.tau);
    if (hoist_gather) b.tail_row(ROW_TD3_GATHER, B);
    return;
  }
  b.phase(PC | COND_WORLD_1);
  {
    const int ad = b.adam_desc(P, &c.tpolicy, c.hp.policy_lr, b1, b2, eps, c.hp.tau, SLOT_POLICY);
    b.dw(S.d1p, Hd, Hd, S.h0p, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1, 0, ad);
    b.dw(S.d0p, Hd, Hd, S.Xoa, S.ld_oa, O, B, P.g + P.oW0, P.g + P.ob0, 0, ad);
    b.dw(S.dmean, A, A, S.h1p, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2, 0, ad);
  }
  b.polyak(c.qf[0], c.tqf[0], c.hp.tau);
  b.polyak(c.qf[1], c.tqf[1], c.hp.tau);
  b.phase(PC | COND_WORLD_N, 0, /*push=*/1);
  b.dw(S.d1p, Hd, Hd, S.h0p, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1);
  b.dw(S.d0p, Hd, Hd, S.Xoa, S.ld_oa, O, B, P.g + P.oW0, P.g + P.ob0);
  b.dw(S.dmean, A, A, S.h1p, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2);
  b.phase(PC | COND_WORLD_N, 2);
  b.adam(P, &c.tpolicy, c.hp.policy_lr, b1, b2, eps, c.hp.tau, SLOT_POLICY, 1);
  b.polyak(c.qf[0], c.tqf[0], c.hp.tau);
  b.polyak(c.qf[1], c.tqf[1], c.hp.tau);
}

// S2: SAC with a V function and a fixed entropy coefficient (sac.py:70-179, 242-243)
inline void build_sac_v(Builder& b, const Ctx& c) {
  const SacBufs& S = c.s;
  const int B = S.B, O = S.O, A = S.A, Hd = S.Hd, K0 = O + A;
  const MlpPtrs& P = c.policy; const MlpPtrs& V = c.vf; const MlpPtrs& TV = c.tvf;
  const double b1 = c.hp.beta1, b2 = c.hp.beta2, eps = c.hp.adam_eps;
  float* h0p_obs = S.h0p + (size_t)B * Hd;      // the policy runs on the obs rows [B,2B) of the 2B layout
  float* h1p_obs = S.h1p + (size_t)B * Hd;
  const float* obs_rows = S.Xpi + (size_t)B * S.ld_o;
  b.phase(); b.row(ROW_SAC_GATHER, B);
  {
    Builder::TwoLayers L[5] = {{S.Xoa, S.ld_oa, B, K0, &c.qf[0], S.h0q[0], S.h1q[0], ACT_RELU},
                               {S.Xoa, S.ld_oa, B, K0, &c.qf[1], S.h0q[1], S.h1q[1], ACT_RELU},
                               {S.Xna, S.ld_oa, B, O, &TV, S.h0tv, S.h1tv, ACT_RELU},
                               {S.Xoa, S.ld_oa, B, O, &V, S.h0v, S.h1v, ACT_RELU},
                               {obs_rows, S.ld_o, B, O, &P, h0p_obs, h1p_obs, ACT_RELU}};
    b.two_layers(L, 5);
  }
  b.phase(); b.row(ROW_SAC_HEADS, B, 0, /*row_offset=*/B);          // one eps draw: a~, log pi on the obs rows
  {   // min Q(obs, a~) with the PRE-update critics (V regression target, sac.py:121-129)
    Builder::TwoLayers L[2] = {{S.Xon, S.ld_oa, B, K0, &c.qf[0], S.h0n[0], S.h1n[0], ACT_RELU},
                               {S.Xon, S.ld_oa, B, K0, &c.qf[1], S.h0n[1], S.h1n[1], ACT_RELU}};
    b.two_layers(L, 2);
  }
  b.phase(); b.row(ROW_SACV_TARGET, B);
  b.phase();
  for (int i = 0; i < 2; ++i)
    b.dx(S.d1q[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0q[i], Hd, ACT_RELU, S.d0q[i], Hd);
  b.dx(S.d1v, Hd, B, Hd, V.p + V.oW1, Hd, Hd, S.h0v, Hd, ACT_RELU, S.d0v, Hd);
  // all three backward passes first, then the three Adam steps (sac.py:132-139): the weight-gradient tiles read
  // only deltas and activations computed above, so fusing each Adam (+ Polyak of the target V, :242-243) into
  // their epilogues keeps that order
  const bool tc5 = c.hp.use_tc5 != 0;
  if (tc5) {        // see build_sac_alpha: split-K weight gradients, flat Adam jobs (all three after all three backwards)
    b.phase();
    b.dw_splits = Builder::pick_splits(3 * Builder::tc5_tiles(Hd, Hd) + 2 * Builder::tc5_tiles(Hd, K0) + Builder::tc5_tiles(Hd, O), B, kMaxKSplits);
    for (int i = 0; i < 2; ++i) b.dw(S.d1q[i], Hd, Hd, S.h0q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW1, c.qf[i].g + c.qf[i].ob1, 0, -1, &c.qf[i]);
    b.dw(S.d1v, Hd, Hd, S.h0v, Hd, Hd, B, V.g + V.oW1, V.g + V.ob1, 0, -1, &V);
    for (int i = 0; i < 2; ++i) b.dw(S.d0q[i], Hd, Hd, S.Xoa, S.ld_oa, K0, B, c.qf[i].g + c.qf[i].oW0, c.qf[i].g + c.qf[i].ob0, 0, -1, &c.qf[i]);
    b.dw(S.d0v, Hd, Hd, S.Xoa, S.ld_oa, O, B, V.g + V.oW0, V.g + V.ob0, 0, -1, &V);
    b.dw_splits = kMaxKSplits;       // skinny output-layer gradients: K = batch split over the partial arenas too
    for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, -1, &c.qf[i]);
    b.dw(S.dv, 1, 1, S.h1v, Hd, Hd, B, V.g + V.oW2, V.g + V.ob2, 0, -1, &V);
    b.phase();
    b.adam(c.qf[0], nullptr, c.hp.qf_lr, b1, b2, eps, 0.f, SLOT_QF1);
    b.adam(c.qf[1], nullptr, c.hp.qf_lr, b1, b2, eps, 0.f, SLOT_QF2);
    b.adam(V, &c.tvf, c.hp.vf_lr, b1, b2, eps, c.hp.tau, SLOT_VF);
  } else {
  b.phase();
    {
      const int aq0 = b.adam_desc(c.qf[0], nullptr, c.hp.qf_lr, b1, b2, eps, 0.f, SLOT_QF1);
      const int aq1 = b.adam_desc(c.qf[1], nullptr, c.hp.qf_lr, b1, b2, eps, 0.f, SLOT_QF2);
      const int av = b.adam_desc(V, &c.tvf, c.hp.vf_lr, b1, b2, eps, c.hp.tau, SLOT_VF);
      const int ad[2] = {aq0, aq1};
      for (int i = 0; i < 2; ++i) b.dw(S.d1q[i], Hd, Hd, S.h0q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW1, c.qf[i].g + c.qf[i].ob1, 0, ad[i]);
      b.dw(S.d1v, Hd, Hd, S.h0v, Hd, Hd, B, V.g + V.oW1, V.g + V.ob1, 0, av);
      for (int i = 0; i < 2; ++i) b.dw(S.d0q[i], Hd, Hd, S.Xoa, S.ld_oa, K0, B, c.qf[i].g + c.qf[i].oW0, c.qf[i].g + c.qf[i].ob0, 0, ad[i]);
      b.dw(S.d0v, Hd, Hd, S.Xoa, S.ld_oa, O, B, V.g + V.oW0, V.g + V.ob0, 0, av);
      for (int i = 0; i < 2; ++i) b.dw(S.dq[i], 1, 1, S.h1q[i], Hd, Hd, B, c.qf[i].g + c.qf[i].oW2, c.qf[i].g + c.qf[i].ob2, 0, ad[i]);
      b.dw(S.dv, 1, 1, S.h1v, Hd, Hd, B, V.g + V.oW2, V.g + V.ob2, 0, av);
    }
}
  {   // policy loss re-evaluates the UPDATED critics on the SAME action sample (:150-153)
    Builder::TwoLayers L[2] = {{S.Xon, S.ld_oa, B, K0, &c.qf[0], S.h0n[0], S.h1n[0], ACT_RELU},
                               {S.Xon, S.ld_oa, B, K0, &c.qf[1], S.h0n[1], S.h1n[1], ACT_RELU}};
    b.two_layers(L, 2);
  }
  b.phase(); b.row(ROW_SAC_PLOSS, B);
  b.phase();
  for (int i = 0; i < 2; ++i) b.dx(S.e1[i], Hd, B, Hd, c.qf[i].p + c.qf[i].oW1, Hd, Hd, S.h0n[i], Hd, ACT_RELU, S.e0[i], Hd);
  b.phase(); b.row(ROW_SAC_PIBWD_DA, B);
  b.phase();
  b.dx(S.d1p, Hd, B, Hd, P.p + P.oW1, Hd, Hd, h0p_obs, Hd, ACT_RELU, S.d0p, Hd);
  if (tc5) {
    b.phase();
    b.dw_splits = Builder::pick_splits(Builder::tc5_tiles(Hd, Hd) + Builder::tc5_tiles(Hd, O) + (A > 8 ? 2 * Builder::tc5_tiles(A, Hd) : 0), B, kMaxKSplits);
    b.dw(S.d1p, Hd, Hd, h0p_obs, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1, 0, -1, &P);
    b.dw(S.d0p, Hd, Hd, obs_rows, S.ld_o, O, B, P.g + P.oW0, P.g + P.ob0, 0, -1, &P);
    b.dw(S.dmean, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2, 0, -1, &P);
    b.dw(S.dlraw, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW3, P.g + P.ob3, 0, -1, &P);
    b.phase(COND_ALWAYS, 1);
    b.adam(P, nullptr, c.hp.policy_lr, b1, b2, eps, 0.f, SLOT_POLICY, 1);
    b.row(ROW_SACV_FINAL, 1);
    return;
  }
  b.phase(COND_WORLD_1);
  {
    const int ad = b.adam_desc(P, nullptr, c.hp.policy_lr, b1, b2, eps, 0.f, SLOT_POLICY);
    b.dw(S.d1p, Hd, Hd, h0p_obs, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1, 0, ad);
    b.dw(S.d0p, Hd, Hd, obs_rows, S.ld_o, O, B, P.g + P.oW0, P.g + P.ob0, 0, ad);
    b.dw(S.dmean, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2, 0, ad);
    b.dw(S.dlraw, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW3, P.g + P.ob3, 0, ad);
  }
  b.row(ROW_SACV_FINAL, 1);
  b.phase(COND_WORLD_N, 0, /*push=*/1);
  b.dw(S.d1p, Hd, Hd, h0p_obs, Hd, Hd, B, P.g + P.oW1, P.g + P.ob1);
  b.dw(S.d0p, Hd, Hd, obs_rows, S.ld_o, O, B, P.g + P.oW0, P.g + P.ob0);
  b.dw(S.dmean, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW2, P.g + P.ob2);
  b.dw(S.dlraw, A, A, h1p_obs, Hd, Hd, B, P.g + P.oW3, P.g + P.ob3);
  b.phase(COND_WORLD_N, 2);
  b.adam(P, nullptr, c.hp.policy_lr, b1, b2, eps, 0.f, SLOT_POLICY, 1);
  b.row(ROW_SACV_FINAL, 1);
}

inline Hyper make_hyper(const ilsw_trainer_config& cfg) {
  Hyper h; memset(&h, 0, sizeof(h));
  h.algo = cfg.algo;
  h.gemm_precision = cfg.gemm_precision;
  h.reward_scale = (float)cfg.reward_scale; h.discount = (float)cfg.discount; h.tau = (float)cfg.soft_target_tau;
  h.policy_lr = cfg.policy_lr; h.qf_lr = cfg.qf_lr; h.vf_lr = cfg.vf_lr; h.alpha_lr = cfg.alpha_lr;
  h.beta1 = cfg.beta_1; h.beta2 = cfg.beta_2 > 0 ? cfg.beta_2 : 0.999; h.adam_eps = cfg.adam_eps > 0 ? cfg.adam_eps : 1e-8;
  h.mean_reg = (float)cfg.policy_mean_reg_weight; h.std_reg = (float)cfg.policy_std_reg_weight;
  h.target_entropy = (float)cfg.target_entropy; h.train_alpha = cfg.train_alpha; h.fixed_alpha = (float)cfg.alpha;
  h.period = cfg.policy_and_target_update_period > 0 ? cfg.policy_and_target_update_period : 1;
  h.policy_noise = (float)cfg.policy_noise; h.noise_clip = (float)cfg.policy_noise_clip;
  h.max_act = cfg.max_act != 0 ? (float)cfg.max_act : 1.0f;
  h.her = cfg.algo == ILSW_ALGO_TD3 ? cfg.her : 0;
  h.her_sigma = (float)cfg.her_sigma; h.min_act = (float)cfg.min_act;
  h.clip_l = (float)cfg.clip_return_l; h.clip_r = (float)cfg.clip_return_r;
  return h;
}

inline void apply_disc_hyper(Hyper& h, const ilsw_disc_config& d) {
  h.has_disc = 1; h.disc_mode = d.mode; h.disc_lr = d.disc_lr; h.disc_beta1 = d.disc_momentum;
  h.gp_weight = (float)d.grad_pen_weight; h.disc_clamp = (float)d.clamp_magnitude; h.use_gp = d.use_grad_pen;
  h.clip_min_on = d.rew_clip_min_on; h.clip_max_on = d.rew_clip_max_on;
  h.rew_clip_min = (float)d.rew_clip_min; h.rew_clip_max = (float)d.rew_clip_max;
  h.state_only = d.state_only; h.n_from_expert = d.policy_batch_from_expert;
  h.disc_act = d.hid_act == ILSW_DISC_ACT_RELU ? ACT_RELU : ACT_TANH;
}

inline int build_program(Program& P);

struct TrainerSpec {
  ilsw_trainer_config cfg;
  ilsw_mlp nets[6];
  int n_nets;
  int has_disc;
  ilsw_disc_config dcfg;
  ilsw_mlp disc;
};

inline int validate_spec(const TrainerSpec& sp, std::string* why) {
  const ilsw_trainer_config& c = sp.cfg;
  auto fail = [&](const char* m) { if (why) *why = m; return (int)ILSW_ERR_ARG; };
  if (c.obs_dim <= 0 || c.act_dim <= 0 || c.batch <= 0) return fail("obs_dim/act_dim/batch must be positive");
  if (c.act_dim > 64) return fail("act_dim > 64 unsupported");
  if (c.max_steps_per_call <= 0) return fail("max_steps_per_call must be positive");
  int need = c.algo == ILSW_ALGO_SAC_ALPHA ? 5 : (c.algo == ILSW_ALGO_TD3 ? 6 : (c.algo == ILSW_ALGO_SAC_V ? 5 : -1));
  if (need < 0) return fail("unsupported algo");
  if (sp.n_nets != need) return fail("wrong number of networks for this algorithm");
  const int Hd = sp.nets[0].hidden;
  for (int i = 0; i < sp.n_nets; ++i) {
    const ilsw_mlp& n = sp.nets[i];
    if (!n.p) return fail("null parameter arena");
    if (n.hidden != Hd || Hd <= 0) return fail("all networks must share one hidden width");
    bool is_policy = (i == 0) || (c.algo == ILSW_ALGO_TD3 && i == 5);
    bool is_vnet = c.algo == ILSW_ALGO_SAC_V && i >= 3;          // vf, target_vf: obs -> 1
    int in_dim = (is_policy || is_vnet) ? c.obs_dim : c.obs_dim + c.act_dim;
    int out_dim = is_policy ? c.act_dim : 1;
    if (n.in_dim != in_dim || n.out_dim != out_dim) return fail("network in/out dims do not match obs/act dims");
    if ((n.log_std_head != 0) != (is_policy && c.algo != ILSW_ALGO_TD3)) return fail("log_std_head mismatch");
    bool trainable = i < 3 || (c.algo == ILSW_ALGO_SAC_V && i == 3);
    if (trainable && (!n.m || !n.v)) return fail("trainable network needs Adam moment arenas");
  }
  if (sp.has_disc) {
    if (c.algo != ILSW_ALGO_SAC_ALPHA) return fail("discriminator requires the SAC-alpha trainer");
    if (sp.dcfg.batch != c.batch) return fail("disc batch must equal policy batch");
    const int disc_in = sp.dcfg.state_only ? 2 * c.obs_dim : c.obs_dim + c.act_dim;
    if (sp.disc.in_dim != disc_in || sp.disc.out_dim != 1 || sp.disc.log_std_head) return fail("disc dims");
    if (sp.dcfg.policy_batch_from_expert < 0 || sp.dcfg.policy_batch_from_expert >= c.batch) return fail("policy_batch_from_expert must be in [0, batch)");
    if (!sp.disc.p || !sp.disc.m || !sp.disc.v) return fail("disc arenas");
    if (sp.dcfg.mode < 0 || sp.dcfg.mode > 3) return fail("disc mode");
    if (sp.dcfg.hid_act != ILSW_DISC_ACT_TANH && sp.dcfg.hid_act != ILSW_DISC_ACT_RELU) return fail("disc hid_act");
  }
  return ILSW_OK;
}

// Lays out scratch in `mem` (two-pass capable), fills P.ctx and compiles the program.
// tcgen05/TMA tiles pay where the layer is a dense GEMM: batch >= 512 (TD3-Humanoid B = 1024, HER B = 4096), tensor-core
// precision modes, no discriminator program (B = 256 everywhere)
inline bool tc5_wanted(const TrainerSpec& sp) {
  return sp.cfg.batch >= 512 && sp.cfg.gemm_precision != 0 && !sp.has_disc;
}

inline int assemble(Program& P, const TrainerSpec& sp, Bump& mem, bool use_tc5 = false, bool fuse_l0 = true) {
  memset(&P, 0, sizeof(P));
  Ctx& c = P.ctx;
  const ilsw_trainer_config& cfg = sp.cfg;
  c.hp = make_hyper(cfg);
  c.hp.use_tc5 = use_tc5 ? 1 : 0;
  c.hp.fuse_l0 = fuse_l0 ? 1 : 0;
  if (sp.has_disc) apply_disc_hyper(c.hp, sp.dcfg);
  const int B = cfg.batch, O = cfg.obs_dim, A = cfg.act_dim, Hd = sp.nets[0].hidden;
  c.dyn = mem.take<DynState>(1);
  c.loss_log = mem.f((size_t)cfg.max_steps_per_call * kLossSlots);
  c.stats_floats = stats_floats_for(cfg.algo, B, A);
  c.stats = mem.f(c.stats_floats);
  c.phase_ns = mem.take<unsigned long long>(2 * (kMaxPhases + 1));
  c.cta_ns = mem.take<unsigned long long>((size_t)kMaxPhases * kMaxGrid);
  alloc_sac_bufs(mem, c.s, cfg.algo, B, O, A, Hd);
  // tcgen05 programs: kMaxKSplits partial gradient arenas per trainable net (split-K weight gradients; the scratch is
  // zero-filled once, and elements no split GEMM produces stay zero in the arenas s >= 1)
  const int nsplit = use_tc5 ? kMaxKSplits : 1;
  auto grad = [&](const ilsw_mlp& n) { return mem.f((size_t)nsplit * round_up(mlp_num_params(n.in_dim, n.hidden, n.out_dim, n.log_std_head), 4)); };
  auto mk = [&](const ilsw_mlp& n) {
    MlpPtrs m = make_mlp(n, grad(n));
    m.g_splits = nsplit; m.g_stride = round_up(m.n_params, 4);
    return m;
  };
  c.policy = mk(sp.nets[0]);
  c.qf[0] = mk(sp.nets[1]);
  c.qf[1] = mk(sp.nets[2]);
  if (cfg.algo == ILSW_ALGO_SAC_V) {
    c.vf = mk(sp.nets[3]);
    c.tvf = make_mlp(sp.nets[4], nullptr);
  } else {
    c.tqf[0] = make_mlp(sp.nets[3], nullptr);
    c.tqf[1] = make_mlp(sp.nets[4], nullptr);
  }
  if (cfg.algo == ILSW_ALGO_TD3) c.tpolicy = make_mlp(sp.nets[5], nullptr);
  if (sp.has_disc) {
    alloc_disc_bufs(mem, c.d, B, sp.dcfg.state_only ? 2 * O : O + A, sp.disc.hidden);
    {   // the discriminator's gradient lives in kDiscGradSplits partial arenas (split-K first-layer weight gradient, build_disc_step)
      const int ns = cfg.gemm_precision != 0 ? kDiscGradSplits : 1;
      c.disc = make_mlp(sp.disc, mem.f((size_t)ns * round_up(mlp_num_params(sp.disc.in_dim, sp.disc.hidden, sp.disc.out_dim, sp.disc.log_std_head), 4)));
      c.disc.g_splits = ns; c.disc.g_stride = round_up(c.disc.n_params, 4);
    }
  }
  if (use_tc5) {
    MlpPtrs* nets[] = {&c.policy, &c.qf[0], &c.qf[1], &c.tqf[0], &c.tqf[1], &c.tpolicy, &c.vf, &c.tvf};
    for (MlpPtrs* n : nets)
      if (n->p && (n->in_dim & 3)) { n->ld_w0p = round_up(n->in_dim, 4); n->w0p = mem.f((size_t)n->hid * n->ld_w0p); }
  }
  return build_program(P);
}

// Assembles the whole step program.  Returns 0 or a negative ilsw_status.
inline int build_program(Program& P) {
  Builder b(P);
  const Ctx& c = P.ctx;
  if (c.hp.use_tc5) {      // parameters may have been written by the host since the last launch: refresh the aligned W0 copies
    const MlpPtrs* nets[] = {&c.policy, &c.qf[0], &c.qf[1], &c.tqf[0], &c.tqf[1], &c.tpolicy, &c.vf, &c.tvf};
    bool any = false;
    for (const MlpPtrs* n : nets) any = any || (n->p && n->w0p);
    if (any) {
      b.phase(COND_FIRST_STEP);
      for (const MlpPtrs* n : nets) b.shadow_refresh(*n);
    }
  }
  if (c.hp.has_disc) { b.base_cond = COND_DISC_PART; build_disc_step(b, c); b.base_cond = COND_POLICY_PART; }
  if (c.hp.algo == ILSW_ALGO_SAC_ALPHA) build_sac_alpha(b, c);
  else if (c.hp.algo == ILSW_ALGO_TD3) build_td3(b, c);
  else if (c.hp.algo == ILSW_ALGO_SAC_V) build_sac_v(b, c);
  else return ILSW_ERR_UNSUPPORTED;
  return b.overflow ? ILSW_ERR_STATE : ILSW_OK;
}

inline std::string describe_program(const Program& P) {
  std::string out;
  char line[256];
  static const char* kinds[] = {"?", "GEMM", "ADAM", "ROW", "POLYAK"};
  for (int i = 0; i < P.n_phases; ++i) {
    const Phase& ph = P.phases[i];
    snprintf(line, sizeof(line), "phase %2d jobs=%4d%s%s%s%s%s:", i, ph.total_jobs, (ph.cond & COND_TD3_POLICY) ? " [td3-policy-step]" : ((ph.cond & COND_TD3_POLICY_OR_STATS) ? " [td3-policy-or-stats-step]" : ""),
             (ph.cond & COND_FIRST_STEP) ? " [first-step]" : "", (ph.cond & COND_WORLD_1) ? " [1-replica]" : "",
             (ph.cond & COND_WORLD_N) ? " [n-replicas]" : "", ph.collective ? " [replica-exchange]" : (ph.push ? " [pushes to replicas]" : ""));
    out += line;
    for (int j = 0; j < ph.op_count; ++j) {
      const Op& o = P.ops[ph.op_begin + j];
      if (o.kind == OP_GEMM)
        snprintf(line, sizeof(line), " GEMM(%dx%dx%d%s%s%s%s%s)", o.gemm.M, o.gemm.N, o.gemm.K, o.gemm.aug_ones ? "+1" : "",
                 o.gemm.accumulate ? ",acc" : "", o.gemm.adam ? ",adam" : "", o.gemm.tc5 ? ",tcgen05" : "", o.gemm.a0 ? ",fused-L0" : "");
      else if (o.kind == OP_ROW)
        snprintf(line, sizeof(line), " ROW(k%d,%d)", o.row.kind, o.row.rows);
      else if (o.kind == OP_ADAM && o.adam.fused_only)
        line[0] = 0;
      else if (o.kind == OP_ADAM)
        snprintf(line, sizeof(line), " ADAM(%d%s)", o.adam.n, o.adam.target ? ",polyak" : "");
      else if (o.kind == OP_SHADOW)
        snprintf(line, sizeof(line), " W0COPY(%d)", o.shadow.dst.n);
      else if (o.kind == OP_L0FUSE)
        line[0] = 0;
      else
        snprintf(line, sizeof(line), " %s(%d)", kinds[o.kind], o.polyak.n);
      out += line;
    }
    out += "\n";
  }
  if (P.ctx.tail_op1) {
    const Op& o = P.ops[P.ctx.tail_op1 - 1];
    snprintf(line, sizeof(line), "with the last active phase of every step: ROW(k%d,%d) for step s+1, jobs=%d\n", o.row.kind, o.row.rows, o.n_jobs);
    out += line;
  }
  return out;
}

}  // namespace ilsw
