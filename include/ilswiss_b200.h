/* ilswiss_b200 -- C ABI of the B200-native replacement for ILSwiss's off-policy hot path.
 *
 * The reference (Ericonaldo/ILSwiss) is 100% Python and has no FFI of its own; its boundary
 * for this path is three duck-typed Python interfaces:
 *   ReplayBuffer ... rlkit/data_management/replay_buffer.py:4-83,
 *                    rlkit/data_management/simple_replay_buffer.py:17-323
 *   Trainer ........ rlkit/core/trainer.py:4-28 (sac_alpha.py:21-76, td3.py:20-70)
 *   AdvIRL ......... rlkit/torch/algorithms/adv_irl/adv_irl.py:34-131
 * The Python classes in ilswiss_b200/ mirror those interfaces and bind to the entry points
 * below through ctypes (INTEGRATION.md shows the stub).  Plain pointers and sizes only: no
 * torch types cross this boundary.  Every function returns 0 on success or a negative
 * ilsw_status; ilsw_last_error() gives a message.  All device pointers must belong to the
 * current CUDA device; `stream` is a cudaStream_t passed as void*.
 */
#ifndef ILSWISS_B200_H
#define ILSWISS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ILSW_ABI_VERSION 1

typedef enum {
  ILSW_OK = 0,
  ILSW_ERR_ARG = -1,
  ILSW_ERR_CUDA = -2,
  ILSW_ERR_STATE = -3,
  ILSW_ERR_UNSUPPORTED = -4,
  ILSW_ERR_ABORTED = -5
} ilsw_status;

int ilsw_abi_version(void);
const char* ilsw_last_error(void);
/* number of SMs / name of the current device; fails loudly (ILSW_ERR_CUDA) without a GPU */
int ilsw_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len);

/* ---------------------------------------------------------------------------------------
 * Replay ring in HBM.  Replaces SimpleReplayBuffer's numpy arrays
 * (simple_replay_buffer.py:17-68): one packed fp32 "hot" row per transition
 *     [ obs(O) | act(A) | reward | terminal | next_obs(O) | pad to 4 floats ]
 * plus a "cold" row [absorbing0, absorbing1, timeout, 0] that only random_batch() returns.
 * Host rows passed to append are packed WITHOUT padding:
 *     [ obs(O) | act(A) | reward | terminal | next_obs(O) | absorbing0 | absorbing1 | timeout ]
 * ------------------------------------------------------------------------------------- */
typedef struct ilsw_rb ilsw_rb;

int ilsw_rb_create(ilsw_rb** out, int64_t capacity, int obs_dim, int act_dim);
int ilsw_rb_destroy(ilsw_rb* rb);
int ilsw_rb_host_row_floats(const ilsw_rb* rb);   /* 2O + A + 5 */
int ilsw_rb_row_stride(const ilsw_rb* rb);        /* padded hot-row stride in floats */
int64_t ilsw_rb_capacity(const ilsw_rb* rb);
int64_t ilsw_rb_size(const ilsw_rb* rb);          /* num_steps_can_sample(), replay_buffer.py:39 */
int64_t ilsw_rb_top(const ilsw_rb* rb);
float* ilsw_rb_rows_ptr(ilsw_rb* rb);             /* device pointer to the hot rows */
/* add_sample/add_path (simple_replay_buffer.py:78-108,134-216): stage n host rows (pinned or
 * pageable) with cudaMemcpyAsync on `copy_stream`; the rows enter the ring at `top` (with
 * wrap-around, _advance :228-237) when ilsw_rb_commit() runs on the compute stream. */
int ilsw_rb_append(ilsw_rb* rb, const float* host_rows, int64_t n, void* copy_stream);
int ilsw_rb_commit(ilsw_rb* rb, void* stream);
/* bulk device-to-device load of already packed hot rows (synthetic fills, snapshot restore) */
int ilsw_rb_load_device(ilsw_rb* rb, const float* dev_hot_rows, int64_t n, void* stream);
/* _get_batch_using_indices (:255-293): gather `B` rows by index into dense tiles.
 * out_hot: [B x stride], out_cold: [B x 4] (nullable).  idx on the device (int32). */
int ilsw_rb_gather(ilsw_rb* rb, const int32_t* idx_dev, int B, float* out_hot, float* out_cold,
                   void* stream);
/* random_batch (:239-253) with the in-kernel Philox stream instead of numpy's MT19937:
 * writes the drawn indices to idx_out_dev (nullable) and the gathered tiles to out_hot. */
int ilsw_rb_sample(ilsw_rb* rb, int B, uint64_t seed, uint64_t counter, int32_t* idx_out_dev,
                   float* out_hot, void* stream);
int ilsw_rb_clear(ilsw_rb* rb);
/* snapshot restore: set the write cursor / fill level after a physical-layout load */
int ilsw_rb_set_cursor(ilsw_rb* rb, int64_t top, int64_t size);

/* ---------------------------------------------------------------------------------------
 * Networks: one flat fp32 arena per network in nn.Module.parameters() order of the reference
 * classes (FlattenMlp / policies: fc0.w, fc0.b, fc1.w, fc1.b, last_fc.w, last_fc.b
 * [, last_fc_log_std.w, last_fc_log_std.b]; MLPDisc: mod_list.{0,2,4}.{weight,bias}),
 * nn.Linear (out,in) row-major.  Exactly two hidden layers of equal width (all shipped
 * exp_specs).  p/m/v are caller-owned device arenas of n_params floats (m,v = Adam moments).
 * ------------------------------------------------------------------------------------- */
typedef struct {
  float* p;
  float* m;
  float* v;
  int in_dim, hidden, out_dim;
  int log_std_head;   /* 1: ReparamTanhMultivariateGaussianPolicy (policies.py:191-243) */
} ilsw_mlp;

int ilsw_mlp_num_params(int in_dim, int hidden, int out_dim, int log_std_head);

typedef enum { ILSW_ALGO_SAC_ALPHA = 1, ILSW_ALGO_TD3 = 2, ILSW_ALGO_SAC_V = 3 } ilsw_algo;
typedef enum { ILSW_DISC_AIRL = 0, ILSW_DISC_GAIL = 1, ILSW_DISC_GAIL2 = 2, ILSW_DISC_FAIRL = 3 } ilsw_disc_mode;
typedef enum { ILSW_DISC_ACT_TANH = 0, ILSW_DISC_ACT_RELU = 1 } ilsw_disc_act;

typedef struct {
  int algo;                 /* ilsw_algo */
  int obs_dim, act_dim, batch;
  int max_steps_per_call;   /* capacity of the per-step loss log */
  int gemm_precision;       /* 0: fp32 SIMT (exact parity gate); 1: TF32 tensor cores; 3: 3xTF32 (fp32-level) */
  /* SoftActorCritic.__init__ (sac_alpha.py:21-40) / TD3.__init__ (td3.py:20-36) */
  double reward_scale, discount, soft_target_tau;
  double policy_lr, qf_lr, vf_lr, alpha_lr;
  double beta_1, beta_2, adam_eps;
  double alpha;             /* initial / fixed entropy coefficient */
  int train_alpha;
  double target_entropy;
  double policy_mean_reg_weight, policy_std_reg_weight;
  /* TD3: policy MODULE noise parameters (policies.py:150-152), td3.py:113 period */
  int policy_and_target_update_period;
  double policy_noise, policy_noise_clip, max_act;
  /* HER-TD3 (rlkit/torch/algorithms/her/td3.py:88-160; observations are cat(obs, desired_goal)): target action =
   * clamp(her_sigma * N(0,1), min_act, max_act) (:103-112 -- the target policy's output is overwritten), min target Q
   * clipped to [clip_return_l, clip_return_r] (:116-120), policy loss + mean(action^2) (:150-152). */
  int her;
  double her_sigma, min_act, clip_return_l, clip_return_r;
} ilsw_trainer_config;

typedef struct {
  int mode;                 /* ilsw_disc_mode (adv_irl.py:277-289) */
  int batch;                /* disc_optim_batch_size == policy_optim_batch_size */
  double disc_lr, disc_momentum;      /* Adam betas=(disc_momentum, 0.999), adv_irl.py:75-77 */
  int use_grad_pen;
  double grad_pen_weight;
  double clamp_magnitude;   /* MLPDisc clamp (simple_disc_models.py:43-48) */
  int rew_clip_min_on, rew_clip_max_on;
  double rew_clip_min, rew_clip_max;
  /* adv_irl.py:139-179,265-269: discriminator input cat(obs, next_obs) instead of cat(obs, act); disc.in_dim = 2*obs_dim */
  int state_only;
  /* adv_irl.py:239-255 policy_optim_batch_size_from_expert: the LAST n rows of every policy batch are sampled from the expert
   * ring (injected idx rows >= batch - n index the expert ring) */
  int policy_batch_from_expert;
  /* MLPDisc hid_act (simple_disc_models.py:19-24), use_bn=False, num_layer_blocks=2: ilsw_disc_act.  0 = tanh, what every
   * shipped exp_specs/gail yaml sets; relu blocks run the same step program minus the act'' terms of the penalty */
  int hid_act;
} ilsw_disc_config;

typedef struct ilsw_trainer ilsw_trainer;

/* SAC-alpha: nets = {policy, qf1, qf2, target_qf1, target_qf2};
 * TD3:       nets = {policy, qf1, qf2, target_qf1, target_qf2, target_policy};
 * SAC-V:     nets = {policy, qf1, qf2, vf, target_vf}.  (m,v of targets may be NULL) */
int ilsw_trainer_create(ilsw_trainer** out, const ilsw_trainer_config* cfg, const ilsw_mlp* nets,
                        int n_nets);
/* Turns a SAC-alpha trainer into an AdvIRL engine: one engine step = one loop iteration of
 * adv_irl.py:126-131 with 1 discriminator update + 1 policy update. */
int ilsw_trainer_attach_disc(ilsw_trainer* tr, const ilsw_disc_config* cfg, const ilsw_mlp* disc);
/* num_disc_updates_per_loop_iter / num_policy_updates_per_loop_iter != 1 (adv_irl.py:126-131; exp_specs/gail/
 * gail_humanoid.yaml uses 100/100): subsequent ilsw_train calls run n_steps discriminator updates only (mode 1:
 * _do_reward_training, :133-236) or n_steps policy updates only (mode 2: _do_policy_training, :238-314, rewards from the
 * current discriminator); mode 0 restores the fused 1+1 iteration.  Adam step counts advance only for the part run. */
/* Hindsight relabel-at-sample (rlkit/data_management/relabel_replay_buffer.py:63-131, relabel_type "future"): ring rows hold
 * obs = cat(observation, desired_goal); the gather of every step (i) samples a finished trajectory uniformly, a step in it,
 * and a future step of the same trajectory, (ii) overwrites the goal part of obs / next_obs of the FIRST relabel_num rows of
 * the batch with the next achieved goal of the future step (:103-118) and (iii) recomputes every row's reward as the
 * sparse goal reward -(||next_achieved_goal - desired_goal|| > distance_threshold) (:127-131, Fetch envs' compute_reward).
 * All pointers are DEVICE pointers that must stay valid while set; NULL / enabled = 0 switches it off.  With injected
 * randomness (ilsw_inject.idx = step rows) inj_idx_her [T,B] supplies the future rows. */
typedef struct {
  int enabled;
  int n_traj;                   /* finished trajectories */
  const int32_t* traj_start;    /* [n_traj] ring index of the first transition */
  const int32_t* traj_len;      /* [n_traj] transitions in the trajectory (may wrap around the ring) */
  const float* next_achieved_goal;   /* [capacity, goal_dim] side array, same slot numbering as the ring */
  int goal_dim;
  int relabel_num;              /* int(her_ratio * batch) */
  float distance_threshold;
  const int32_t* inj_idx_her;   /* parity mode only */
} ilsw_her_sampling;
int ilsw_trainer_set_her(ilsw_trainer* tr, const ilsw_her_sampling* her);

enum { ILSW_UPDATE_BOTH = 0, ILSW_UPDATE_DISC_ONLY = 1, ILSW_UPDATE_POLICY_ONLY = 2 };
int ilsw_trainer_set_update_mode(ilsw_trainer* tr, int mode);
int ilsw_trainer_destroy(ilsw_trainer* tr);

/* Injected randomness for parity runs (all DEVICE pointers, T = n_steps of the call):
 *   idx [T,B] int32 ........ RandomState.randint stream of the policy buffer (policy batch)
 *   eps_next/eps_cur [T,B,A]  N(0,1) draws of distributions.py:24 (TD3: eps_next = noise)
 *   idx_expert/idx_policy_d [T,B], gp_eps [T,B] .... discriminator step (adv_irl.py:147-188) */
typedef struct {
  const int32_t* idx;
  const float* eps_next;
  const float* eps_cur;
  const int32_t* idx_expert;
  const int32_t* idx_policy_d;
  const float* gp_eps;
} ilsw_inject;

/* Caller-provided dense batch (Trainer.train_step(batch) with arbitrary tensors); DEVICE ptrs */
typedef struct {
  const float* obs;       /* [B,O] */
  const float* act;       /* [B,A] */
  const float* rew;       /* [B]   */
  const float* term;      /* [B]   */
  const float* next_obs;  /* [B,O] */
} ilsw_batch;

/* Runs n_steps gradient steps (TorchRLAlgorithm._do_training, torch_rl_algorithm.py:28-34;
 * AdvIRL._do_training, adv_irl.py:126-131) in ONE persistent kernel launch.
 *   policy_rb : replay ring sampled for policy/critic batches (may be NULL iff batch != NULL)
 *   expert_rb : expert ring (AdvIRL only)
 *   inject    : NULL -> in-kernel Philox sampling keyed by (seed, global step)
 *   batch     : NULL -> sample from the ring; else train on this batch (n_steps must be 1)
 *   stats_step: step index whose batch vectors are snapshotted for eval statistics, -1 none */
int ilsw_train(ilsw_trainer* tr, ilsw_rb* policy_rb, ilsw_rb* expert_rb, int n_steps,
               const ilsw_inject* inject, const ilsw_batch* batch, uint64_t seed, int stats_step,
               void* stream);

/* per-step scalars of the LAST ilsw_train call: out[n_steps][ILSW_LOSS_SLOTS]; synchronises */
#define ILSW_LOSS_SLOTS 16
enum {
  ILSW_L_QF1 = 0, ILSW_L_QF2 = 1, ILSW_L_POLICY = 2, ILSW_L_ALPHA_LOSS = 3, ILSW_L_ALPHA = 4,
  ILSW_L_VF = 5, ILSW_L_DISC_CE = 6, ILSW_L_DISC_ACC = 7, ILSW_L_GRAD_PEN = 8,
  ILSW_L_REW_MEAN = 9, ILSW_L_REW_STD = 10, ILSW_L_REW_MAX = 11, ILSW_L_REW_MIN = 12,
  ILSW_L_Q1_MEAN = 13, ILSW_L_LOGPI_MEAN = 14, ILSW_L_QT_MEAN = 15
};
int ilsw_read_losses(ilsw_trainer* tr, float* host_out, int n_steps, void* stream);
/* asynchronous variant: enqueue D2H of the last `n_steps` rows into pinned host memory */
int ilsw_read_losses_async(ilsw_trainer* tr, float* pinned_out, int n_steps, void* stream);
/* synchronises `stream` and reports ILSW_ERR_ABORTED if any launch since the last check hit the in-kernel watchdog
 * (the synchronous readers check on every call; users of the async variant call this when they collect) */
int ilsw_check_abort(ilsw_trainer* tr, void* stream);

/* eval-statistics snapshot (sac_alpha.py:186-233, td3.py:126-177): vectors of the batch of
 * `stats_step`.  Layout (floats): q1_pred[B] q2_pred[B] q_target[B] err1[B] err2[B] reward[B]
 * then SAC: log_pi[B] policy_mean[B*A] policy_log_std[B*A];  TD3: policy_action(tanh)[B*A]. */
int ilsw_stats_floats(const ilsw_trainer* tr);
int ilsw_read_stats(ilsw_trainer* tr, float* host_out, int n_floats, void* stream);

/* scalar optimiser state (for get_snapshot/load_snapshot round trips) */
typedef struct {
  double log_alpha, alpha_exp_avg, alpha_exp_avg_sq;
  int alpha_step;
  int adam_step[8];       /* slots: 0 qf1, 1 qf2, 2 policy, 3 vf, 4 disc */
  int n_train_steps_total;
} ilsw_state;
int ilsw_get_state(ilsw_trainer* tr, ilsw_state* out, void* stream);
int ilsw_set_state(ilsw_trainer* tr, const ilsw_state* in, void* stream);

/* human-readable phase table of the compiled step program (host only; no GPU work) */
int ilsw_describe_program(const ilsw_trainer* tr, char* buf, int buf_len);
int ilsw_num_phases(const ilsw_trainer* tr);
/* profiling: %globaltimer (ns) at the start and after every phase barrier of the LAST step of the
 * most recent launch: out[0..n_phases] */
int ilsw_read_phase_ns(ilsw_trainer* tr, unsigned long long* host_out, int n, void* stream);
/* profiling: stage stamps of CTA 0's last tensor-core GEMM tile in every phase: out[96][8] (ns);
 * only recorded after ilsw_trainer_set_profiling(tr, 1) (off by default: the stamps cost CTA 0 an L2
 * round trip per tile) */
int ilsw_read_tile_ns(unsigned long long* host_out);
int ilsw_trainer_set_profiling(ilsw_trainer* tr, int on);
/* profiling on: jobs-done time (ns) of every CTA in every phase of the last step: out[96][304] */
int ilsw_read_cta_ns(ilsw_trainer* tr, unsigned long long* host_out, void* stream);
int64_t ilsw_kernel_launches(const ilsw_trainer* tr);   /* engine launches so far */
int ilsw_trainer_uses_tc5(const ilsw_trainer* tr);      /* 1: the program runs its dense GEMM phases on the tcgen05/TMA tile (batch >= 512) */

/* A1: sampler-side policy inference for <= 4096 env rows (policies.py:245-246, core.py:74-89).
 * obs_dev [n,O] -> act_dev [n,A]; deterministic: tanh(mean) (SAC) / no noise (TD3). */
int ilsw_policy_act(ilsw_trainer* tr, const float* obs_dev, int n, int deterministic,
                    uint64_t seed, float* act_dev, void* stream);
/* the same with HOST buffers -- replaces the per-env-step round trip of exploration_policy.get_actions
 * (rlkit/torch/core.py:74-89 eval_np: torch_ify -> forward -> np_ify; caller base_algorithm.py:369-380):
 * pinned H2D of obs_host [n,O], one kernel, pinned D2H into act_host [n,A], one stream synchronisation. */
int ilsw_policy_act_host(ilsw_trainer* tr, const float* obs_host, int n, int deterministic,
                         uint64_t seed, float* act_host, void* stream);

/* ---------------------------------------------------------------------------------------
 * Replicas: one process per GPU; policy gradients are averaged across ranks
 * INSIDE the step kernel over NVLink peer memory (CUDA IPC), fused with the policy Adam.
 * ------------------------------------------------------------------------------------- */
#define ILSW_IPC_HANDLE_BYTES 64
int ilsw_replica_export(ilsw_trainer* tr, void* handle_out /* ILSW_IPC_HANDLE_BYTES */);
int ilsw_replica_connect(ilsw_trainer* tr, int rank, int world,
                         const void* all_handles /* world * ILSW_IPC_HANDLE_BYTES */);
/* Same exchange over a caller-provided symmetric buffer (e.g. torch.distributed._symmetric_memory): peer_ptrs[r] = the
 * buffer of rank r mapped into THIS process (peer_ptrs[rank] = the local one), multicast_ptr = its NVLS multicast mapping
 * or 0.  With a multicast mapping the gradient push is one multimem store per element (the NVSwitch replicates it) instead
 * of `world` peer stores.  Replaces the torch.distributed all-reduce a data-parallel port of
 * torch_rl_algorithm.py:28-34 would issue after policy_loss.backward() (sac_alpha.py:150-152). */
int64_t ilsw_replica_buffer_bytes(ilsw_trainer* tr);
int ilsw_replica_connect_symm(ilsw_trainer* tr, int rank, int world, const uint64_t* peer_ptrs /* world */,
                              uint64_t multicast_ptr, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* ILSWISS_B200_H */
