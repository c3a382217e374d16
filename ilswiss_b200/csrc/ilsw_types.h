// ilswiss_b200 -- plain-old-data structures shared by the program builder (host C++),
// the persistent engine kernel (sm_100a) and the test-only host simulator.
//
// One "program" = the whole gradient step of one algorithm (SAC-alpha: sac_alpha.py:78-181,
// TD3: td3.py:72-124, AdvIRL disc step + reward relabel: adv_irl.py:133-314) expressed as an
// ordered list of PHASES; a phase is a set of independent tile jobs (GEMM tiles, row jobs,
// flat Adam/Polyak chunks) separated from the next phase by a grid-wide barrier.  The engine
// kernel is launched ONCE per train call and loops `n_steps x phases` on-chip.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ILSW_HD __host__ __device__ __forceinline__
#define ILSW_HDN __host__ __device__ __noinline__
#else
#define ILSW_HD inline
#define ILSW_HDN inline
#endif

namespace ilsw {

constexpr int kMaxPhases = 96;
constexpr int kMaxOps = 128;
constexpr int kMaxGrid = 304;     // CTAs of the persistent grid (148 SMs x 2)
constexpr int kMaxNets = 8;       // Adam step-counter slots
constexpr int kLossSlots = 16;    // floats per step in the loss log
constexpr int kThreads = 256;     // CTA size of the engine kernel
constexpr int kRowsPerJob = 8;    // one warp per batch row
constexpr int kAdamChunk = 2048;  // elements per flat Adam/Polyak job
constexpr int kFuseL0MaxK = 32;   // widest first-layer input folded into the second layer's GEMM tiles (GemmOp::a0)
constexpr int kMaxGradSplits = 4;  // split-K partial gradient arenas summed by the flat Adam jobs (adam_grad)
constexpr int kTc5BN = 64;        // column width of the tcgen05 tile (ilsw_tc5.cuh is instantiated with it; rows: 128)

enum OpKind : int { OP_GEMM = 1, OP_ADAM = 2, OP_ROW = 3, OP_POLYAK = 4, OP_SHADOW = 5, OP_L0FUSE = 6 };

enum Act : int { ACT_NONE = 0, ACT_RELU = 1, ACT_TANH = 2 };

// phase conditions (bit mask: every set condition must hold)
enum Cond : int { COND_ALWAYS = 0, COND_TD3_POLICY = 1, COND_FIRST_STEP = 2, COND_WORLD_1 = 4, COND_WORLD_N = 8,
                  COND_DISC_PART = 16,      // phase of the discriminator update (D1); skipped by policy-only launches
                  COND_POLICY_PART = 32,    // phase of the policy update (D2 + S1) of an AdvIRL program; skipped by disc-only launches
                  COND_TD3_POLICY_OR_STATS = 64 };   // TD3 policy steps AND the statistics step of a launch: td3.py:131-136 evaluates a
                                                     // stats-only policy loss when the logged step is not a policy step
// AdvIRL launches (adv_irl.py:126-131): one engine step = one disc update + one policy update (UPDATE_BOTH, the
// num_*_updates_per_loop_iter = 1 case of every shipped yaml but one), or n disc-only / n policy-only steps
enum UpdateMode : int { UPDATE_BOTH = 0, UPDATE_DISC_ONLY = 1, UPDATE_POLICY_ONLY = 2 };

// loss-log slots (per step)
enum LossSlot : int {
  L_QF1 = 0, L_QF2 = 1, L_POLICY = 2, L_ALPHA_LOSS = 3, L_ALPHA = 4, L_VF = 5,
  L_DISC_CE = 6, L_DISC_ACC = 7, L_GRAD_PEN = 8, L_REW_MEAN = 9, L_REW_STD = 10,
  L_REW_MAX = 11, L_REW_MIN = 12, L_Q1_MEAN = 13, L_LOGPI_MEAN = 14, L_QT_MEAN = 15
};

struct GemmOp {
  // C[m,n] = epilogue( sum_k A(m,k) * B(k,n) )
  const float* A; int lda; int a_mc;  // a_mc=0: A(m,k)=A[m*lda+k] (k contiguous); 1: A[k*lda+m]
  const float* B; int ldb; int b_nc;  // b_nc=0: B(k,n)=B[n*ldb+k] (k contiguous); 1: B[k*ldb+n]
  int M, N, K;
  int aug_ones;         // 1: logical extra column n==N with B(k,N)=1, routed to bias_out[m] (= sum_k A(m,k): the bias
                        // gradient of a weight-gradient GEMM; requires a_mc).  The device tiles produce it in their
                        // tn==0 tile from the A panel -- it costs no extra column tile.
  float* C; int ldc;
  float* C2;            // optional second output: value BEFORE the mask (ldc shared)
  float* bias_out;      // aug_ones destination (bias gradient)
  const float* bias;    // forward: + bias[n]
  const float* H; int ldh;  // mask source (post-activation values of the layer below)
  int act;              // forward activation
  int mask;             // backward mask: ACT_RELU -> (H>0), ACT_TANH -> (1-H^2)
  int accumulate;       // C += value (and bias_out +=)
  int tiles_m, tiles_n;
  int adam;             // 0: none; k+1: ops[k] is an OP_ADAM descriptor (fused_only) covering this GEMM's outputs --
                        // the epilogue applies the Adam (+Polyak) update to every gradient element it produces, so the
                        // weight-gradient phase needs no separate optimiser phase (valid only without accumulate
                        // partners and without a cross-replica exchange)
  int a0;               // 0: none; k+1: ops[k] is an OP_L0FUSE descriptor -- the A operand is PRODUCED inside the tile (fused
                        // first layer, see L0FuseOp) instead of read from memory
  int tc5;              // 1: 128 x 64 tcgen05/TMA tile (ilsw_tc5.cuh); tiles_m/tiles_n then count those tiles
  int ksplit;           // tcgen05 weight-gradient GEMMs: number of K splits (jobs = ksplit * tiles_m * tiles_n); split s writes
  int split_stride;     //   its partial sums to C + s * split_stride / bias_out + s * split_stride (floats); 0 / 1: no split
  const void* tmapA;    // device CUtensorMap of the A / B operand (SWIZZLE_128B boxes, see ilsw_tc5.cuh)
  const void* tmapB;
  int kbase;            // first K index of the job's K range in the tensor maps (0 except in the shifted copies of split-K jobs)
  int kind;             // tile routine, resolved by the program builder (GemmKind): the engine's dispatch is one load + compare
  int tma;              // mma.sync tile (ilsw_engine.cuh): bit 0 / bit 1 = the A / B panels of a stage are brought in by TMA boxes
                        // {32 floats, 32 rows} of tmapA / tmapB instead of per-thread cp.async (16-byte aligned operands only)
};

// 16-byte-aligned copy of a first-layer weight matrix W0[hid x in] whose rows are not (in % 4 != 0: critics on Hopper /
// Humanoid ...): row stride `ld` = in rounded up to 4.  TMA cannot address the packed nn.Linear layout, so the tcgen05
// programs read the copy; whoever writes parameter element i < n (= hid * in) also writes its copy.
struct ShadowRef { float* ptr; int in, ld, n; float inv_in; };
ILSW_HD void shadow_store(const ShadowRef& sh, int i, float v) {
  if (sh.ptr && i < sh.n) {
    // row = i / in without the integer-division sequence: float estimate (exact for i < 2^23) + one correction step
    int r = (int)(((float)i + 0.5f) * sh.inv_in);
    int c = i - r * sh.in;
    if (c < 0) { --r; c += sh.in; } else if (c >= sh.in) { ++r; c -= sh.in; }
    sh.ptr[(size_t)r * sh.ld + c] = v;
  }
}

enum GemmKind : int { GK_TILE = 0, GK_TC5 = 1, GK_SKINNY = 2, GK_SKINNY_SPLIT = 3, GK_TILE_SPLIT = 4 };

struct AdamOp {
  float* p; const float* g; float* m; float* v;
  float* target;        // optional Polyak target updated from the NEW p (nullptr: none)
  ShadowRef sh_p, sh_t; // aligned copies of W0 of the network / of its target (ptr == nullptr: none)
  int n;
  double lr, beta1, beta2, eps; float tau;
  int slot;             // Adam step-counter slot
  int grad_scale_world; // 1: divide g by world size (replica-averaged policy gradient)
  int begin;            // first element of the flat job range [begin, n)
  int fused_only;       // 1: descriptor for GEMM epilogues (GemmOp::adam); owns no jobs
  int g_splits;         // > 1: the gradient is the sum of g_splits partial arenas g + s * g_split_stride (split-K GEMMs)
  int g_split_stride;
};

// replica exchange from the tile epilogues (Phase::push): where element `p` of the local policy gradient arena goes on
// every replica -- peer[r] is this rank's receive slot on replica r (NVLink peer mapping), offset by the arena position
// `mc` != nullptr: the same slot through an NVLS multicast mapping (one store reaches every replica: the NVSwitch replicates
// it), the per-peer pointers are then unused
struct PushCtx { float* peer[8]; const float* grad_base; int world; float* mc; };

// Fused first layer (A-operand producer of a second-layer GEMM, small-input nets: K0 = O + A <= kFuseL0MaxK):
//   A(m,k) = act( sum_j X[m*ldx + j] * W[k*K0 + j] + b[k] ),  j < K0,  k < K (= hidden width),
// i.e. Linear(K0 -> hidden) + activation of networks.py:85-101 folded into the GEMM of the second layer: one phase (and
// one grid barrier) less per forward pass.  The tn == 0 tile of every row block also stores A(m, :) to `out` (the
// backward pass needs the first-layer activations).  Descriptor op without jobs, referenced by GemmOp::a0.
struct L0FuseOp { const float* X; int ldx; int K0; const float* W; const float* b; int act; float* out; int ldo; };

struct PolyakOp { float* target; const float* src; int n; float tau; ShadowRef sh_t; };
struct ShadowOp { const float* src; ShadowRef dst; };   // refresh of one aligned copy (first step of a launch)

struct RowOp {
  int kind;             // algorithm-specific row kernel id
  int rows;             // number of rows (jobs = ceil(rows / kRowsPerJob))
  int arg0, arg1;       // arg0 = 1: the job works for step s+1 (batch prefetch), skipped on the last step
};

struct Op {
  int kind;
  int n_jobs;
  union {
    GemmOp gemm;
    AdamOp adam;
    PolyakOp polyak;
    RowOp row;
    ShadowOp shadow;
    L0FuseOp l0;
  };
};

struct Phase {
  int op_begin, op_count;
  int total_jobs;
  int cond;
  int collective;       // cross-replica gradient exchange at the START of this phase: 1 = push loop + flags (gradient arenas
                        // pushed whole), 2 = only wait for the pushes the weight-gradient tiles made from their epilogues
  int push;             // 1: the GEMM tiles of this phase also store what they produce into every replica's receive slot
                        // (NVLink peer stores) and every CTA signals the peers when its jobs are done
};

// ---- per-algorithm buffer tables (device or host pointers) --------------------------------

struct MlpPtrs {        // canonical 2-hidden-layer layout inside one flat arena
  float* p; float* m; float* v; float* g;   // params, Adam moments, gradient arena
  int in_dim, hid, out_dim, heads;          // heads=2: extra log-std head (policy)
  int n_params;
  float* w0p; int ld_w0p;                   // aligned copy of W0 (tcgen05 programs, in_dim % 4 != 0), else nullptr
  int g_splits, g_stride;                   // split-K partial gradient arenas g + s * g_stride, s < g_splits (0 / 1: one arena)
  // element offsets
  int oW0, ob0, oW1, ob1, oW2, ob2, oW3, ob3;
};

struct SacBufs {        // SAC-alpha / TD3 / SAC-V scratch (all fp32, row-major)
  int B, O, A, Hd, ld_oa, ld_o;
  int* idx;
  float *Xoa, *rew, *term, *Xpi, *Xna, *Xon, *eps;
  float *h0q[2], *h1q[2], *h0t[2], *h1t[2], *h0n[2], *h1n[2];
  float *h0p, *h1p, *mean, *lraw, *lstd, *act, *logpi;
  float *qp[2], *tq[2], *y, *qn[2], *dq[2];
  float *d1q[2], *d0q[2], *e1[2], *e0[2], *dA[2];
  float *dmean, *dlraw, *d1p, *d0p;
  float *lossterm[2], *plterm, *regmu, *regls, *aterm;
  // TD3 extras
  float *h0tp, *h1tp;   // target policy activations
  float *noise;
  // SAC-V extras
  float *h0v, *h1v, *h0tv, *h1tv, *vp, *tv, *dv, *d1v, *d0v, *lossterm_v, *qn_old[2];
};

struct DiscBufs {
  int B, D, Hd, ld_d;
  int *idx_e, *idx_p;
  float *X3;            // [3B x ld_d]: expert rows, policy rows, interpolated rows
  float *gp_eps;
  float *h1, *h2;       // [3B x Hd]
  float *y, *dlogit, *cmask;      // [3B], [2B], [B]
  float *d2, *d1;       // CE deltas [2B x Hd]
  float *dl2, *u1, *dl1, *g, *gbar, *nrm;   // GP: delta2, u1, delta1, g [B x ld_d], gbar, norms
  float *db1, *ub1, *sb1, *db2, *t3, *zb2, *hb1, *zb1;
  float *ceterm, *accterm, *gpterm;
  // reward relabel (D2)
  float *rh1, *rh2, *rewraw;
  float *Xsn; int ld_sn;     // state_only: cat(obs, next_obs) rows of the policy batch (reward relabel input)
};

struct DynState {       // device-resident mutable scalars (persist across launches)
  double log_alpha, alpha_m, alpha_v;
  double alpha_p1, alpha_p2;   // running beta1^t, beta2^t of the alpha optimiser (avoids pow() on the device)
  int alpha_t;
  float alpha;          // exp(log_alpha) as used by the fp32 graph
  int abort_flag;
  int error_code;
};

struct Hyper {
  int algo;             // 1 sac_alpha, 2 td3, 3 sac_v
  int gemm_precision;   // 0 fp32 (SIMT, parity gate), 1 tf32 tensor cores, 2 bf16 tensor cores, 3 3xtf32
  float reward_scale, discount, tau;
  double policy_lr, qf_lr, vf_lr, alpha_lr, beta1, beta2, adam_eps;
  float mean_reg, std_reg, target_entropy;
  int train_alpha;
  float fixed_alpha;
  // td3
  int period; float policy_noise, noise_clip, max_act;
  // HER-TD3 (her/td3.py): see ilsw_trainer_config
  int her; float her_sigma, min_act, clip_l, clip_r;
  // disc
  int has_disc; int disc_mode;    // 0 airl 1 gail 2 gail2 3 fairl
  double disc_lr, disc_beta1; float gp_weight, disc_clamp; int use_gp;
  int clip_min_on, clip_max_on; float rew_clip_min, rew_clip_max;
  int state_only;       // disc input = cat(obs, next_obs) (adv_irl.py:139-179)
  int disc_act;         // MLPDisc hid_act (simple_disc_models.py:19-24): ACT_TANH (every shipped yaml) or ACT_RELU
  int n_from_expert;    // last n rows of the policy batch come from the expert ring (adv_irl.py:239-255)
  int use_tc5;          // dense GEMM phases run on the tcgen05/TMA tile (ilsw_tc5.cuh): batch >= 512, tensor-core modes
  int fuse_l0;          // narrow first layers are produced inside the second layer's tiles (GemmOp::a0)
};

struct Ctx {            // everything a row kernel needs
  Hyper hp;
  SacBufs s;
  DiscBufs d;
  MlpPtrs policy, qf[2], tqf[2], tpolicy, vf, tvf, disc;
  DynState* dyn;
  float* loss_log;      // [max_steps x kLossSlots]
  float* stats;         // snapshot area (see ilsw_stats layout in include/ilswiss_b200.h)
  int stats_floats;
  int tail_op1;         // 0: none; k + 1: ops[k] (a next-step row op, RowOp::arg0 = 1, owned by no phase) runs alongside the LAST
                        // ACTIVE phase of every step -- which phase that is depends on the step (TD3 policy / statistics steps)
  unsigned long long* phase_ns;  // [2*(kMaxPhases+1)] globaltimer stamps of the LAST step of a launch (profiling):
                                 // [i] = after the barrier of phase i-1; [kMaxPhases+1+i] = CTA 0 finished its jobs of phase i
  unsigned long long* cta_ns;    // [kMaxPhases x kMaxGrid] (profiling on): every CTA's jobs-done time per phase, last step
};

struct RingView {       // replay ring as seen by the gather row kernel
  const float* rows; int stride; int size;   // hot rows [capacity x stride]
};

struct Inject {         // optional injected randomness (parity mode); all device pointers
  const int* idx;       // [T x B]
  const float* eps_next;  // [T x B x A]   (TD3: target-policy noise)
  const float* eps_cur;   // [T x B x A]
  const int* idx_expert;  // [T x B]   (disc step)
  const int* idx_policy_d;  // [T x B]
  const float* gp_eps;    // [T x B]
};

struct HerSampling {    // relabel-at-sample (see ilsw_her_sampling in include/ilswiss_b200.h)
  int enabled, n_traj;
  const int* traj_start; const int* traj_len;
  const float* ag_next; int G;
  int relabel_num; float threshold;
  const int* inj_idx_her;
};

struct DirectBatch {    // optional: train on a caller-provided dense batch instead of the ring
  const float *obs, *act, *rew, *term, *next_obs;   // [B x O], [B x A], [B], [B], [B x O]
};

struct RunArgs {
  int n_steps;
  int step0;            // global train-step counter at launch (TD3 delay phase, Philox offset)
  int stats_step;       // step index (within launch) whose vectors are snapshotted; -1 none
  int t0[kMaxNets];     // Adam step counts at launch, per slot
  uint64_t seed;
  RingView ring_policy, ring_expert;
  Inject inj; int has_inject;
  DirectBatch direct; int has_direct;
  // replicas
  int world, rank;
  int loss_log_offset;  // first row of loss_log to write
  int profile;          // 1: CTA 0 stamps the stages of its GEMM tiles (tools/phase_profile.py); 0 in production
  int update_mode;      // UpdateMode (AdvIRL programs only)
  HerSampling her;
  // host mailbox (pinned, mapped): after the last step CTA 0 copies the launch's loss rows there and publishes `mail_seq`
  // with a system-scope release store -- the host reads a launch's losses by polling memory, without a stream
  // synchronisation or a copy of its own (ilsw_read_losses)
  float* mail_losses;                  // [max_steps x kLossSlots] device view of the mapped buffer; nullptr: off
  unsigned long long* mail_done;       // device view of the sequence word
  unsigned long long mail_seq;         // value to publish for this launch
  unsigned* bar_other;                 // counter of the barrier state the NEXT launch uses: zeroed by this launch
};

struct Program {
  int n_phases, n_ops;
  Phase phases[kMaxPhases];
  Op ops[kMaxOps];
  Ctx ctx;
};

// Adam slots
enum Slot : int { SLOT_QF1 = 0, SLOT_QF2 = 1, SLOT_POLICY = 2, SLOT_VF = 3, SLOT_DISC = 4 };

// row kernel ids
enum RowKind : int {
  ROW_SAC_GATHER = 1, ROW_SAC_HEADS, ROW_SAC_TARGET, ROW_SAC_PLOSS, ROW_SAC_PIBWD, ROW_SAC_FINAL,
  ROW_TD3_GATHER, ROW_TD3_THEAD, ROW_TD3_TARGET, ROW_TD3_PHEAD, ROW_TD3_PLOSS, ROW_TD3_PIBWD,
  ROW_TD3_FINAL, ROW_TD3_FINAL_POLICY,
  ROW_DISC_GATHER, ROW_DISC_HEAD, ROW_DISC_GNORM, ROW_DISC_EW1, ROW_DISC_EW2, ROW_DISC_EW3,
  ROW_DISC_FINAL, ROW_DISC_REWARD, ROW_DISC_REWARD_FINAL,
  ROW_SACV_HEADS, ROW_SACV_TARGET, ROW_SACV_VTARGET, ROW_SACV_PLOSS, ROW_SACV_FINAL,
  ROW_SAC_PIBWD_DA, ROW_TD3_PIBWD_DA      // policy-head backward fused with dA = e0 . W0[:, O:O+A]
};

}  // namespace ilsw
