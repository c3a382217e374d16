"""TEST INFRASTRUCTURE: drives the reference's OWN training loop (rlkit BaseAlgorithm.start_training through
TorchRLAlgorithm.train() / AdvIRL.train(), base_algorithm.py:150-291) on a synthetic vec-env, either with the reference's
classes (CPU) or with ilswiss_b200.dropin installed (GPU).  Real MuJoCo / envpool environments do not exist in this image
(SURVEY.md 8c); the fake vec-env implements the interface the loop uses: len(), reset(ids), step(actions, ids),
observation_space / action_space."""
import csv
import os

import numpy as np


class _Box:
    def __init__(self, n, seed):
        self.shape = (n,)
        self.low, self.high = -np.ones(n), np.ones(n)
        self._rs = np.random.RandomState(seed)

    def sample(self):
        return self._rs.uniform(-1, 1, self.shape)


class FakeVecEnv:
    """n linear-Gaussian environments; episodes end (terminal) after `episode_len` steps."""

    def __init__(self, n, obs_dim, act_dim, episode_len=25, seed=0):
        self.n, self.O, self.A, self.T = n, obs_dim, act_dim, episode_len
        self.rs = np.random.RandomState(seed)
        try:                                 # the reference's get_dim() wants gym.spaces.Box instances (stubbed by oracle/ref_shim)
            import gym.spaces as gs
            self.observation_space, self.action_space = gs.Box(obs_dim), gs.Box(act_dim)
        except Exception:
            self.observation_space, self.action_space = _Box(obs_dim, seed + 1), _Box(act_dim, seed + 2)
        self.M = 0.9 * np.eye(obs_dim) + 0.02 * self.rs.randn(obs_dim, obs_dim)
        self.Bm = 0.3 * self.rs.randn(obs_dim, act_dim)
        self.state = self.rs.randn(n, obs_dim)
        self.t = np.zeros(n, dtype=int)

    def __len__(self):
        return self.n

    def seed(self, s):
        self.rs = np.random.RandomState(s)

    def reset(self, ids=None):
        ids = np.arange(self.n) if ids is None else np.asarray(ids, dtype=int)
        self.state[ids] = self.rs.randn(len(ids), self.O)
        self.t[ids] = 0
        return self.state[ids].copy()

    def step(self, actions, ids):
        ids = np.asarray(ids, dtype=int)
        a = np.asarray(actions, dtype=np.float64).reshape(len(ids), self.A)
        nxt = np.clip(self.state[ids] @ self.M.T + a @ self.Bm.T + 0.05 * self.rs.randn(len(ids), self.O), -10, 10)
        rew = -0.1 * (nxt ** 2).sum(1) - 0.01 * (a ** 2).sum(1)
        self.state[ids] = nxt
        self.t[ids] += 1
        term = self.t[ids] >= self.T
        return nxt.copy(), rew, term, [dict() for _ in ids]


def _gpu_mode(ptu, on):
    """ptu.set_gpu_mode(False) raises in the reference (pytorch_util.py:61-62 assigns None to os.environ); the scripts only
    ever call it with True."""
    import torch
    if on:
        ptu.set_gpu_mode(True, 0)
    else:
        ptu._use_gpu, ptu.device = False, torch.device("cpu")


def _setup_logger(log_dir):
    from rlkit.core import logger

    os.makedirs(log_dir, exist_ok=True)
    logger.reset() if hasattr(logger, "reset") else None
    csv_path = os.path.join(log_dir, "progress.csv")
    for f in list(getattr(logger, "_tabular_outputs", [])):
        logger.remove_tabular_output(f)
    logger.add_tabular_output(csv_path)
    logger.set_snapshot_dir(log_dir, dict(exp_name="ref_loop"), log_tboard=False, log_wandb=False)
    logger.set_snapshot_mode("none")
    return logger, csv_path


def read_progress(csv_path):
    with open(csv_path) as f:
        rows = list(csv.reader(f))
    return rows[0], rows[1:]


def run_sac_loop(log_dir, device, epochs=2, env_num=2, O=11, A=3, B=64, hidden=256, steps_per_epoch=200, seed=0):
    """TorchRLAlgorithm.train() with a SAC-alpha trainer, written the way run_scripts/sac_alpha_exp_script.py:60-105 builds
    it.  device=True: ilswiss_b200.dropin is installed first, so the SAME constructor lines bind the B200 classes."""
    import torch
    from oracle import ref_shim

    ref_shim.install()
    if device:
        import ilswiss_b200.dropin as dropin
        dropin.install()
    try:
        from rlkit.torch.common.networks import FlattenMlp
        from rlkit.torch.common.policies import ReparamTanhMultivariateGaussianPolicy
        import rlkit.torch.algorithms.sac.sac_alpha as sac_mod
        import rlkit.torch.algorithms.torch_rl_algorithm as alg_mod
        import rlkit.torch.utils.pytorch_util as ptu

        _gpu_mode(ptu, bool(device))
        np.random.seed(seed); torch.manual_seed(seed)
        logger, csv_path = _setup_logger(log_dir)
        env, train_env, eval_env = FakeVecEnv(1, O, A, seed=seed), FakeVecEnv(env_num, O, A, seed=seed + 10), FakeVecEnv(env_num, O, A, seed=seed + 20)
        qf1 = FlattenMlp(hidden_sizes=[hidden, hidden], input_size=O + A, output_size=1)
        qf2 = FlattenMlp(hidden_sizes=[hidden, hidden], input_size=O + A, output_size=1)
        policy = ReparamTanhMultivariateGaussianPolicy(hidden_sizes=[hidden, hidden], obs_dim=O, action_dim=A)
        trainer = sac_mod.SoftActorCritic(policy=policy, qf1=qf1, qf2=qf2, env=env, reward_scale=1.0, discount=0.99, policy_lr=3e-4,
                                          qf_lr=3e-4, soft_target_tau=0.005)
        algorithm = alg_mod.TorchRLAlgorithm(trainer=trainer, env=env, training_env=train_env, eval_env=eval_env,
                                             exploration_policy=policy, batch_size=B, num_train_steps_per_train_call=20,
                                             num_epochs=epochs - 1, num_steps_per_epoch=steps_per_epoch, num_steps_between_train_calls=20,
                                             num_steps_per_eval=60, max_path_length=1000, min_steps_before_training=60,
                                             replay_buffer_size=5000, freq_saving=1, save_replay_buffer=False)
        if device:
            algorithm.to(ptu.device)
        algorithm.train()
        header, rows = read_progress(csv_path)
        return dict(header=header, rows=rows, algorithm=algorithm, trainer=trainer)
    finally:
        if device:
            dropin.uninstall()


def run_advirl_loop(log_dir, device, epochs=2, env_num=2, O=11, A=3, B=64, hidden=256, steps_per_epoch=200, seed=0,
                    disc_kind="mlp_tanh", use_grad_pen=True):
    """AdvIRL.train() (GAIL) the way run_scripts/adv_irl_exp_script.py builds it, with a synthetic expert buffer."""
    import torch
    from oracle import ref_shim

    ref_shim.install()
    if device:
        import ilswiss_b200.dropin as dropin
        dropin.install()
    try:
        from rlkit.torch.common.networks import FlattenMlp
        from rlkit.torch.common.policies import ReparamTanhMultivariateGaussianPolicy
        from rlkit.torch.algorithms.adv_irl.disc_models.simple_disc_models import MLPDisc, ResNetAIRLDisc
        import rlkit.torch.algorithms.sac.sac_alpha as sac_mod
        import rlkit.torch.algorithms.adv_irl.adv_irl as irl_mod
        import rlkit.data_management.env_replay_buffer as erb
        import rlkit.torch.utils.pytorch_util as ptu

        _gpu_mode(ptu, bool(device))
        np.random.seed(seed); torch.manual_seed(seed)
        logger, csv_path = _setup_logger(log_dir)
        env, train_env, eval_env = FakeVecEnv(1, O, A, seed=seed), FakeVecEnv(env_num, O, A, seed=seed + 10), FakeVecEnv(env_num, O, A, seed=seed + 20)
        expert = erb.EnvReplayBuffer(5000, env, random_seed=3)
        ers = np.random.RandomState(8)
        for _ in range(4):
            T = 50
            path = dict(observations=ers.randn(T, O), actions=ers.uniform(-1, 1, (T, A)), rewards=ers.randn(T, 1),
                        next_observations=ers.randn(T, O), terminals=np.zeros((T, 1)), absorbings=np.zeros((T, 2)),
                        env_infos=[dict() for _ in range(T)], agent_infos=[dict() for _ in range(T)])
            expert.add_path(path, env=env)
        qf1 = FlattenMlp(hidden_sizes=[hidden, hidden], input_size=O + A, output_size=1)
        qf2 = FlattenMlp(hidden_sizes=[hidden, hidden], input_size=O + A, output_size=1)
        policy = ReparamTanhMultivariateGaussianPolicy(hidden_sizes=[hidden, hidden], obs_dim=O, action_dim=A)
        if disc_kind == "mlp_tanh":          # every shipped yaml (exp_specs/gail/*.yaml:24-28): the fused program
            disc = MLPDisc(O + A, num_layer_blocks=2, hid_dim=128, hid_act="tanh", use_bn=False, clamp_magnitude=10.0)
        elif disc_kind == "mlp_relu":        # relu blocks without BatchNorm: the same step program minus the act'' terms
            disc = MLPDisc(O + A, num_layer_blocks=2, hid_dim=128, hid_act="relu", use_bn=False, clamp_magnitude=10.0)
        elif disc_kind == "mlp_bn_relu":     # the class defaults (simple_disc_models.py:9-16)
            disc = MLPDisc(O + A, num_layer_blocks=2, hid_dim=128, hid_act="relu", use_bn=True, clamp_magnitude=10.0)
        elif disc_kind == "resnet":
            disc = ResNetAIRLDisc(O + A, num_layer_blocks=3, hid_dim=128, hid_act="relu", use_bn=True, clamp_magnitude=10.0)
        else:
            raise ValueError(disc_kind)
        trainer = sac_mod.SoftActorCritic(policy=policy, qf1=qf1, qf2=qf2, env=env, reward_scale=2.0, discount=0.99, policy_lr=3e-4,
                                          qf_lr=3e-4, soft_target_tau=0.005, beta_1=0.25)
        algorithm = irl_mod.AdvIRL(mode="gail2", discriminator=disc, policy_trainer=trainer, expert_replay_buffer=expert,
                                   state_only=False, disc_optim_batch_size=B, policy_optim_batch_size=B,
                                   policy_optim_batch_size_from_expert=0, num_update_loops_per_train_call=10,
                                   num_disc_updates_per_loop_iter=1, num_policy_updates_per_loop_iter=1, disc_lr=3e-4,
                                   disc_momentum=0.9, use_grad_pen=use_grad_pen, grad_pen_weight=8.0, env=env, training_env=train_env,
                                   eval_env=eval_env, exploration_policy=policy, num_epochs=epochs - 1,
                                   num_steps_per_epoch=steps_per_epoch, num_steps_between_train_calls=20, num_steps_per_eval=60,
                                   max_path_length=1000, min_steps_before_training=60, replay_buffer_size=5000, freq_saving=1,
                                   no_terminal=False)
        if device:
            algorithm.to(ptu.device)
        algorithm.train()
        header, rows = read_progress(csv_path)
        return dict(header=header, rows=rows, algorithm=algorithm, trainer=trainer)
    finally:
        if device:
            dropin.uninstall()
