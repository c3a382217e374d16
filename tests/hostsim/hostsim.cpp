// TEST INFRASTRUCTURE ONLY -- host simulator of the ilswiss_b200 step engine.
//
// Compiles the product's op semantics (ilswiss_b200/csrc/ilsw_ops.cuh) and program builder
// (ilsw_program.h) for the CPU with a 1-lane "warp", and executes the phase program
// sequentially on host memory.  tests/test_hostsim.py compares it with the oracle so that the
// program wiring and the hand-derived backward passes are verified without a GPU.  This file
// is never linked into libilswiss_b200.so and the product never falls back to it.
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../ilswiss_b200/csrc/ilsw_ops.cuh"
#include "../../ilswiss_b200/csrc/ilsw_rows_fast.cuh"
#include "../../ilswiss_b200/csrc/ilsw_program.h"

using namespace ilsw;

struct HostSim {
  Program prog;
  std::vector<char> scratch;
  TrainerSpec spec;
  int t[kMaxNets];
  int n_total;
  int generic_rows;     // 1: force the generic per-row kernels (cross-check of the fast row jobs)
  int update_mode = 0;  // UpdateMode of the next hs_train calls
  HerSampling her = {};  // relabel-at-sample of the next hs_train calls (hs_set_her)
  int world = 1;         // > 1: run the N-replica program (COND_WORLD_N phases) with R IDENTICAL replicas emulated
  std::vector<float> rowbuf;
};

static void run_op(HostSim* hs, const Program& P, const Op& o, const RunArgs& a, int s) {
  const Ctx& c = P.ctx;
  if (o.kind == OP_GEMM) {
    GemmOp g = o.gemm;
    if (g.a0) {          // fused first layer: materialise it (the device's tn == 0 tiles do), then read it as a plain operand
      const L0FuseOp& f = P.ops[g.a0 - 1].l0;
      for (int m = 0; m < g.M; ++m)
        for (int k = 0; k < g.K; ++k) f.out[(size_t)m * f.ldo + k] = gemm_A_fused(f, m, k);
      g.A = f.out; g.lda = f.ldo; g.a_mc = 0; g.a0 = 0;
    }
    int Nt = g.N + g.aug_ones;
    const AdamOp* ad = g.adam ? &P.ops[g.adam - 1].adam : nullptr;     // fused optimiser epilogue
    AdamCoef acf;
    if (ad) acf = adam_coef(*ad, adam_t(a, c.hp, ad->slot, s), a.world);
    for (int m = 0; m < g.M; ++m)
      for (int n = 0; n < Nt; ++n) {
        float acc = 0.f;
        const int prec = c.hp.gemm_precision;
        for (int k = 0; k < g.K; ++k) {
          float x = gemm_A(g, m, k), y = gemm_B(g, k, n);
          if (prec == 1) { x = round_tf32(x); y = round_tf32(y); }
          else if (prec == 2) { x = round_bf16(x); y = round_bf16(y); }
          acc += x * y;
        }
        gemm_epilogue(g, m, n, acc);
        if (ad) {
          const int gi = gemm_grad_index(g, *ad, m, n);
          adam_elem_g(*ad, acf, gi, ad->g[gi] * acf.gscale);
        }
      }
  } else if (o.kind == OP_ROW) {
    const bool fast = !hs->generic_rows && fast_rows_ok(c);
    s += o.row.arg0;                               // arg0 = 1: prefetch job for the next step
    if (s >= a.n_steps) return;
    for (int job = 0; job < o.n_jobs; ++job)
      for (int w = 0; w < kRowsPerJob; ++w) {
        RowEnv env; env.lane = 0; env.nl = 1; env.warp = w; env.sm = hs->rowbuf.data(); env.prof = -1;
        if (fast && run_row_job_fast(c, a, o.row.kind, o.row.rows, s, job + o.row.arg1 / kRowsPerJob, env)) continue;
        const int r = o.row.arg1 + job * kRowsPerJob + w;
        if (r < o.row.rows) run_row(c, a, o.row.kind, s, r, 0, 1);
      }
  } else if (o.kind == OP_ADAM) {
    if (o.adam.fused_only) return;                 // applied by the GEMM epilogues that reference it
    AdamCoef cf = adam_coef(o.adam, adam_t(a, c.hp, o.adam.slot, s), a.world);
    if (a.world > 1 && o.adam.grad_scale_world) {
      // the exchange of R identical replicas: every receive slot holds this replica's gradient; summed in rank order and
      // scaled by 1/R exactly as replica_reduced_grad + adam_job do on the device
      for (int i = o.adam.begin; i < o.adam.n; ++i) {
        float gs = 0.f;
        for (int r = 0; r < a.world; ++r) gs += o.adam.g[i];
        adam_elem_g(o.adam, cf, i, gs * cf.gscale);
      }
    } else {
      for (int i = o.adam.begin; i < o.adam.n; ++i) adam_elem(o.adam, cf, i);
    }
  } else if (o.kind == OP_POLYAK) {
    for (int i = 0; i < o.polyak.n; ++i) polyak_elem(o.polyak, i);
  } else if (o.kind == OP_SHADOW) {
    for (int i = 0; i < o.shadow.dst.n; ++i) shadow_refresh_elem(o.shadow, i);
  }
}

extern "C" {

void* hs_create(const ilsw_trainer_config* cfg, const ilsw_mlp* nets, int n_nets, const ilsw_disc_config* dcfg,
                const ilsw_mlp* disc, char* err, int err_len) {
  HostSim* h = new HostSim();
  memset(&h->spec, 0, sizeof(h->spec));
  h->spec.cfg = *cfg;
  for (int i = 0; i < n_nets && i < 6; ++i) h->spec.nets[i] = nets[i];
  h->spec.n_nets = n_nets;
  if (dcfg) { h->spec.has_disc = 1; h->spec.dcfg = *dcfg; h->spec.disc = *disc; }
  std::string why;
  if (validate_spec(h->spec, &why) != 0) {
    if (err) snprintf(err, err_len, "%s", why.c_str());
    delete h;
    return nullptr;
  }
  // ILSW_HOSTSIM_TC5=1: compile the program variant of the tcgen05/TMA engine (aligned W0 copies, 128-row tile counts) --
  // the host executes GEMMs element by element, so this checks the program wiring of that variant, not the tile
  const char* tv = getenv("ILSW_HOSTSIM_TC5");
  const bool tc5 = tv && atoi(tv) != 0 && tc5_wanted(h->spec);
  Bump measure;
  Program tmp;
  assemble(tmp, h->spec, measure, tc5);
  h->scratch.assign(measure.off + 256, 0);
  Bump mem;
  mem.base = h->scratch.data();
  if (assemble(h->prog, h->spec, mem, tc5) != 0) { delete h; return nullptr; }
  DynState* d = h->prog.ctx.dyn;
  memset(d, 0, sizeof(*d));
  d->log_alpha = log(cfg->alpha);
  d->alpha = (float)exp(d->log_alpha);
  d->alpha_p1 = d->alpha_p2 = 1.0;
  memset(h->t, 0, sizeof(h->t));
  h->n_total = 0;
  h->generic_rows = 0;
  h->rowbuf.assign(kRowStageFloats + 8 * kRowScratchPerWarp + 64, 0.f);
  return h;
}

void hs_destroy(void* p) { delete (HostSim*)p; }

int hs_train(void* p, const float* ring, int stride, int size, const float* ering, int estride, int esize, int n_steps,
             const ilsw_inject* inj, const ilsw_batch* batch, uint64_t seed, int stats_step) {
  HostSim* h = (HostSim*)p;
  RunArgs a;
  memset(&a, 0, sizeof(a));
  a.n_steps = n_steps; a.step0 = h->n_total; a.stats_step = stats_step; a.seed = seed;
  for (int i = 0; i < kMaxNets; ++i) a.t0[i] = h->t[i];
  a.ring_policy = {ring, stride, size};
  a.ring_expert = {ering, estride, esize};
  if (inj) {
    a.has_inject = 1;
    a.inj.idx = inj->idx; a.inj.eps_next = inj->eps_next; a.inj.eps_cur = inj->eps_cur;
    a.inj.idx_expert = inj->idx_expert; a.inj.idx_policy_d = inj->idx_policy_d; a.inj.gp_eps = inj->gp_eps;
  }
  if (batch) {
    a.has_direct = 1;
    a.direct = {batch->obs, batch->act, batch->rew, batch->term, batch->next_obs};
  }
  a.world = h->world; a.rank = 0; a.loss_log_offset = 0;
  a.update_mode = h->update_mode;
  a.her = h->her;
  const Program& P = h->prog;
  for (int s = 0; s < n_steps; ++s) {
    for (int ph = 0; ph < P.n_phases; ++ph) {
      const Phase& phs = P.phases[ph];
      if (!phase_active(phs, P.ctx.hp, a, s)) continue;
      for (int j = 0; j < phs.op_count; ++j) run_op(h, P, P.ops[phs.op_begin + j], a, s);
    }
    // the next-step row op (Ctx::tail_op1) runs with the last active phase of the step: sequentially, after it
    if (P.ctx.tail_op1) run_op(h, P, P.ops[P.ctx.tail_op1 - 1], a, s);
  }
  commit_counters(h->t, h->n_total, a, P.ctx.hp, n_steps);
  return 0;
}

void hs_set_generic_rows(void* p, int flag) { ((HostSim*)p)->generic_rows = flag; }
void hs_set_update_mode(void* p, int mode) { ((HostSim*)p)->update_mode = mode; }
void hs_set_world(void* p, int world) { ((HostSim*)p)->world = world > 1 ? world : 1; }
void hs_set_her(void* p, const ilsw_her_sampling* her) {
  HostSim* h = (HostSim*)p;
  memset(&h->her, 0, sizeof(h->her));
  if (!her || !her->enabled) return;
  h->her.enabled = 1; h->her.n_traj = her->n_traj; h->her.traj_start = her->traj_start; h->her.traj_len = her->traj_len;
  h->her.ag_next = her->next_achieved_goal; h->her.G = her->goal_dim; h->her.relabel_num = her->relabel_num;
  h->her.threshold = her->distance_threshold; h->her.inj_idx_her = her->inj_idx_her;
}
void hs_set_precision(void* p, int prec) { ((HostSim*)p)->prog.ctx.hp.gemm_precision = prec; }
const float* hs_losses(void* p) { return ((HostSim*)p)->prog.ctx.loss_log; }
const float* hs_stats(void* p) { return ((HostSim*)p)->prog.ctx.stats; }
int hs_stats_floats(void* p) { return ((HostSim*)p)->prog.ctx.stats_floats; }
double hs_log_alpha(void* p) { return ((HostSim*)p)->prog.ctx.dyn->log_alpha; }
int hs_num_phases(void* p) { return ((HostSim*)p)->prog.n_phases; }
int hs_describe(void* p, char* buf, int n) {
  std::string s = describe_program(((HostSim*)p)->prog);
  snprintf(buf, n, "%s", s.c_str());
  return (int)s.size();
}
// access to a gradient arena for formula tests: which = 0 policy, 1 qf1, 2 qf2, 3 disc
const float* hs_grad(void* p, int which, int* n) {
  const Ctx& c = ((HostSim*)p)->prog.ctx;
  const MlpPtrs* m = which == 0 ? &c.policy : which == 1 ? &c.qf[0] : which == 2 ? &c.qf[1] : &c.disc;
  *n = m->n_params;
  return m->g;
}
}
