"""Drop-in Trainer classes (rlkit/core/trainer.py:4-28 interface) backed by the fused engine.

    SoftActorCritic  <->  rlkit/torch/algorithms/sac/sac_alpha.py:13-284
    SoftActorCriticV <->  rlkit/torch/algorithms/sac/sac.py:13-243
    TD3              <->  rlkit/torch/algorithms/td3/td3.py:13-223
    HerTD3           <->  rlkit/torch/algorithms/her/td3.py:14-245
    HerSAC           <->  rlkit/torch/algorithms/her/sac.py:12-251

Same constructor signatures, same attributes read from outside (.policy, .networks,
.eval_statistics, get_eval_statistics(), end_epoch(), get_snapshot(), load_snapshot(), to()),
same eval-statistics keys.  The nn.Modules passed in stay valid: their Parameters are re-pointed
to views of the device parameter arenas the kernels update in place, so exploration /
evaluation policies and pickled snapshots keep seeing the trained weights.
"""
import copy
import os
from collections import OrderedDict

import numpy as np
import torch
import torch.optim as optim

from . import _abi, layout
from .engine import NetArena, StepEngine, require_cuda


def default_gemm_precision():
    """GEMM tile mode.  3 (default): 3xTF32 split on the tensor cores -- fp32-level accuracy, every parity
    test passes at the 1e-4 loss bar with room to spare; 0: fp32 FFMA tiles (the exact gate);
    1: single-pass TF32 (fastest; losses within 5e-4 of the reference: OUTSIDE the 1e-4 bar on the
    small-batch AdvIRL cases, opt-in only).  Override with ILSW_GEMM_PRECISION or gemm_precision=."""
    return int(os.environ.get("ILSW_GEMM_PRECISION", "3"))


def _stats(name, data):
    """rlkit/core/eval_util.py:91-142 create_stats_ordered_dict for ndarray data."""
    data = np.asarray(data)
    return OrderedDict([(name + " Mean", np.mean(data)), (name + " Std", np.std(data)),
                        (name + " Max", np.max(data)), (name + " Min", np.min(data))])


def _fn_name(fn):
    return (getattr(fn, "__name__", None) or type(fn).__name__).lower()


def check_activations(module, output=None):
    """The step programs hard-wire what the run scripts build (networks.py:23-101): ReLU hidden layers, identity output
    on the critics, tanh output on the TD3 policy (td3_exp_script.py:71-78).  A module built with anything else has the same
    parameter list, so it is refused here rather than trained as something it is not."""
    hid = getattr(module, "hidden_activation", None)
    if hid is not None and _fn_name(hid) != "relu":
        raise NotImplementedError("hidden_activation %s: the fused engine implements ReLU hidden layers" % _fn_name(hid))
    out = getattr(module, "output_activation", None)
    if output == "identity" and out is not None and _fn_name(out) != "identity":
        raise NotImplementedError("output_activation %s on a critic: the fused engine implements identity" % _fn_name(out))
    if output == "tanh" and (out is None or _fn_name(out) != "tanh"):
        raise NotImplementedError("TD3 policy output_activation %s: the fused engine implements max_act * tanh "
                                  "(policies.py:178, td3_exp_script.py:75)" % (None if out is None else _fn_name(out)))


def module_dims(module):
    """(in_dim, hidden, out_dim, log_std_head) of a 2-hidden-layer reference MLP from its
    parameter list (networks.py:23-83 / policies.py:191-243 / simple_disc_models.py:8-41)."""
    check_activations(module)
    ps = list(module.parameters())
    if len(ps) not in (6, 8):
        raise NotImplementedError("fused engine supports MLPs with exactly two hidden layers (got %d parameter tensors)" % len(ps))
    H, in_dim = ps[0].shape
    if tuple(ps[2].shape) != (H, H) or ps[4].shape[1] != H:
        raise NotImplementedError("fused engine needs equal hidden widths")
    out_dim = ps[4].shape[0]
    if len(ps) == 8 and tuple(ps[6].shape) != (out_dim, H):
        raise NotImplementedError("unexpected log-std head shape")
    for flag in ("layer_norm", "batch_norm"):
        if getattr(module, flag, False):
            raise NotImplementedError("%s is not supported by the fused engine" % flag)
    return int(in_dim), int(H), int(out_dim), len(ps) == 8


def adopt_module(module, trainable=True):
    """Moves a module's parameters into a flat device arena and re-points them at views of it."""
    in_dim, H, out_dim, ls = module_dims(module)
    arena = NetArena(in_dim, H, out_dim, ls, trainable=trainable)
    with torch.no_grad():
        for p, view in zip(module.parameters(), arena.views("p")):
            view.copy_(p.detach().to(device="cuda", dtype=torch.float32))
            p.data = view
    module._ilsw_arena = arena
    return arena


def clone_module(module):
    """PyTorchModule.copy() (rlkit/torch/core.py:32-35) when available, else deepcopy."""
    if hasattr(module, "copy") and callable(module.copy) and not isinstance(module, dict):
        try:
            return module.copy()
        except Exception:
            pass
    return copy.deepcopy(module)


class _FusedTrainer:
    eval_statistics = None
    _n_built = 0

    def _finish_init(self, cfg, nets):
        require_cuda()
        self.engine = StepEngine(cfg, nets)
        self._cfg, self._nets = cfg, list(nets)
        _FusedTrainer._n_built += 1
        self._seed = layout.derive_seed(salt=_FusedTrainer._n_built)        # reproducible under np.random.seed(...)
        self._launch = 0

    def ensure_batch(self, batch_size, max_steps_per_call=None):
        """The step program is compiled for ONE batch size, but the reference hands `batch_size` to the ALGORITHM
        (rl_alg_params, torch_rl_algorithm.py:16-18), not to the trainer: adopt the algorithm's value the first time it
        is seen.  Rebuilds the engine for the new shapes and carries the optimiser state over."""
        B = int(batch_size)
        M = max(int(self._cfg.max_steps_per_call), int(max_steps_per_call or 0))
        if B == self._cfg.batch and M == self._cfg.max_steps_per_call:
            return
        if getattr(self.engine, "disc", None) is not None:
            raise ValueError("batch size %d != trainer batch %d, and a discriminator is attached: construct the trainer "
                             "with batch_size=%d" % (B, self._cfg.batch, B))
        # the optimiser state (Adam step counters, log_alpha and its moments, the train-step counter) lives in the engine:
        # carry it over, so that load_snapshot() followed by the first _do_training with the algorithm's batch size resumes
        # exactly (the parameter / moment arenas are shared between the two engines)
        st = self.engine.get_state()
        self._cfg.batch, self._cfg.max_steps_per_call = B, M
        self.engine = StepEngine(self._cfg, self._nets)
        self.engine.set_state(st)

    # -- the two ways to run gradient steps ------------------------------------------------
    def train_step(self, batch):
        """Trainer.train_step(batch): batch = dict of (B,.) float tensors with the reference's keys
        (np_to_pytorch_batch output).  Runs ONE fused step on that batch."""
        B = self._cfg.batch
        b = {}
        for key, name in (("observations", "obs"), ("actions", "act"), ("rewards", "rew"), ("terminals", "term"),
                          ("next_observations", "next_obs")):
            t = batch[key]
            if not torch.is_tensor(t):
                t = torch.as_tensor(np.asarray(t))
            t = t.to(device="cuda", dtype=torch.float32).contiguous()
            if t.shape[0] != B:
                self.ensure_batch(t.shape[0])          # first use only; raises once the engine has trained
                B = self._cfg.batch
            b[name] = t
        want = self.eval_statistics is None
        self._launch += 1
        self.engine.train(None, 1, batch=b, seed=self._seed + self._launch, stats_step=0 if want else -1)
        self._after_launch(1, want)

    def train_from_buffer(self, replay_buffer, n_steps, inject=None):
        """n_steps gradient steps, each on a fresh uniform sample of `replay_buffer`
        (a DeviceReplayBuffer), in ONE kernel launch -- TorchRLAlgorithm._do_training
        (torch_rl_algorithm.py:28-34) without the per-step host round trips."""
        replay_buffer.flush()
        want = self.eval_statistics is None
        done = 0
        while done < n_steps:
            k = min(n_steps - done, self._cfg.max_steps_per_call)
            self._launch += 1
            sub = None if inject is None else {kk: v[done:done + k].contiguous() for kk, v in inject.items()}
            self.engine.train(replay_buffer.ring, k, inject=sub, seed=self._seed + self._launch,
                              stats_step=0 if (want and done == 0) else -1)
            if want and done == 0:
                self._after_launch(k, True)
            done += k

    def _after_launch(self, n_steps, want_stats):
        if want_stats:
            self.eval_statistics = self._build_stats(self.engine.losses(n_steps)[0], self.engine.stats())

    # -- Trainer interface -----------------------------------------------------------------
    def get_eval_statistics(self):
        return self.eval_statistics

    def end_epoch(self):
        self.eval_statistics = None

    def to(self, device):
        # parameters already live on the GPU inside the arenas; nothing to move
        return self

    def _sync_optimizer(self, opt, arena, step):
        opt.state.clear()
        for p, m, v in zip(opt.param_groups[0]["params"], arena.views("m"), arena.views("v")):
            opt.state[p] = dict(step=torch.tensor(float(step)), exp_avg=m, exp_avg_sq=v)

    def _load_optimizer(self, opt, arena):
        step = 0
        for p, m, v in zip(opt.param_groups[0]["params"], arena.views("m"), arena.views("v")):
            st = opt.state.get(p)
            if st:
                m.copy_(st["exp_avg"].to(m.device))
                v.copy_(st["exp_avg_sq"].to(v.device))
                step = int(float(st["step"]))
        return step


class SoftActorCritic(_FusedTrainer):
    """sac_alpha.py:13-284 (reparameterised SAC, twin Q, auto-tuned alpha)."""

    def __init__(self, policy, qf1, qf2, reward_scale=1.0, discount=0.99, policy_lr=1e-3, qf_lr=1e-3,
                 alpha_lr=3e-4, soft_target_tau=1e-2, alpha=0.2, train_alpha=True,
                 policy_mean_reg_weight=1e-3, policy_std_reg_weight=1e-3, optimizer_class=optim.Adam,
                 beta_1=0.9, target_entropy=None, batch_size=256, max_steps_per_call=1000, gemm_precision=None,
                 **kwargs):
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("the fused step implements torch.optim.Adam only")
        self.policy, self.qf1, self.qf2 = policy, qf1, qf2
        self.reward_scale, self.discount, self.soft_target_tau = reward_scale, discount, soft_target_tau
        self.policy_mean_reg_weight, self.policy_std_reg_weight = policy_mean_reg_weight, policy_std_reg_weight
        self.train_alpha = train_alpha
        in_dim, H, act_dim, ls = module_dims(policy)
        if not ls:
            raise NotImplementedError("SAC needs a ReparamTanhMultivariateGaussianPolicy (conditioned_std=True)")
        for q in (qf1, qf2):
            check_activations(q, "identity")
        self.target_entropy = target_entropy
        if target_entropy is None:   # sac_alpha.py:55-58
            if "env" in kwargs:
                self.target_entropy = -np.prod(kwargs["env"].action_space.shape) / 2.0
            else:
                self.target_entropy = -act_dim / 2.0
        self.target_qf1, self.target_qf2 = clone_module(qf1), clone_module(qf2)   # :60-61
        self._arenas = OrderedDict(
            policy=adopt_module(policy), qf1=adopt_module(qf1), qf2=adopt_module(qf2),
            target_qf1=adopt_module(self.target_qf1, False), target_qf2=adopt_module(self.target_qf2, False))
        self.policy_optimizer = optim.Adam(self.policy.parameters(), lr=policy_lr, betas=(beta_1, 0.999))
        self.qf1_optimizer = optim.Adam(self.qf1.parameters(), lr=qf_lr, betas=(beta_1, 0.999))
        self.qf2_optimizer = optim.Adam(self.qf2.parameters(), lr=qf_lr, betas=(beta_1, 0.999))
        self._log_alpha0 = float(np.log(alpha))
        self._log_alpha_param = torch.tensor(np.log(alpha), requires_grad=train_alpha, device="cuda")
        self.alpha_optimizer = optim.Adam([self._log_alpha_param], lr=alpha_lr, betas=(beta_1, 0.999))
        cfg = _abi.TrainerConfig()
        cfg.algo = _abi.ALGO_SAC_ALPHA
        cfg.gemm_precision = default_gemm_precision() if gemm_precision is None else int(gemm_precision)
        cfg.obs_dim, cfg.act_dim, cfg.batch = in_dim, act_dim, int(batch_size)
        cfg.max_steps_per_call = int(max_steps_per_call)
        cfg.reward_scale, cfg.discount, cfg.soft_target_tau = reward_scale, discount, soft_target_tau
        cfg.policy_lr, cfg.qf_lr, cfg.alpha_lr = policy_lr, qf_lr, alpha_lr
        cfg.beta_1, cfg.beta_2, cfg.adam_eps = beta_1, 0.999, 1e-8
        cfg.alpha, cfg.train_alpha = alpha, int(bool(train_alpha))
        cfg.target_entropy = float(self.target_entropy)
        cfg.policy_mean_reg_weight, cfg.policy_std_reg_weight = policy_mean_reg_weight, policy_std_reg_weight
        cfg.max_act = 1.0
        self._finish_init(cfg, list(self._arenas.values()))
        self.eval_statistics = None

    @property
    def log_alpha(self):
        """float64 0-dim tensor like sac_alpha.py:51-53, refreshed from the device state."""
        with torch.no_grad():
            self._log_alpha_param.fill_(self.engine.get_state().log_alpha)
        return self._log_alpha_param

    @property
    def alpha(self):
        return self.log_alpha.detach().exp()

    @property
    def networks(self):
        return [self.policy, self.qf1, self.qf2, self.target_qf1, self.target_qf2]

    def _build_stats(self, L, vec):
        B, A = self._cfg.batch, self._cfg.act_dim
        st = OrderedDict()
        st["Reward Scale"] = self.reward_scale
        st["QF1 Loss"] = float(L[_abi.L_QF1])
        st["QF2 Loss"] = float(L[_abi.L_QF2])
        if self.train_alpha:
            st["Alpha Loss"] = float(L[_abi.L_ALPHA_LOSS])
        st["Policy Loss"] = float(L[_abi.L_POLICY])
        st.update(_stats("Q1 Predictions", vec[0:B]))
        st.update(_stats("Q2 Predictions", vec[B:2 * B]))
        st.update(_stats("Alpha", [float(L[_abi.L_ALPHA])]))
        o = 6 * B
        st.update(_stats("Log Pis", vec[o:o + B]))
        st.update(_stats("Policy mu", vec[o + B:o + B + B * A]))
        st.update(_stats("Policy log std", vec[o + B + B * A:o + B + 2 * B * A]))
        return st

    def get_snapshot(self):
        st = self.engine.get_state()
        self._sync_optimizer(self.policy_optimizer, self._arenas["policy"], st.adam_step[2])
        self._sync_optimizer(self.qf1_optimizer, self._arenas["qf1"], st.adam_step[0])
        self._sync_optimizer(self.qf2_optimizer, self._arenas["qf2"], st.adam_step[1])
        la = self.log_alpha
        self.alpha_optimizer.state.clear()
        self.alpha_optimizer.state[la] = dict(step=torch.tensor(float(st.alpha_step)),
                                              exp_avg=torch.tensor(st.alpha_exp_avg, dtype=torch.float64),
                                              exp_avg_sq=torch.tensor(st.alpha_exp_avg_sq, dtype=torch.float64))
        return dict(qf1=self.qf1, qf2=self.qf2, policy=self.policy, target_qf1=self.target_qf1,
                    target_qf2=self.target_qf2, log_alpha=la, policy_optimizer=self.policy_optimizer,
                    qf1_optimizer=self.qf1_optimizer, qf2_optimizer=self.qf2_optimizer,
                    alpha_optimizer=self.alpha_optimizer)

    def load_snapshot(self, snapshot):
        """sac_alpha.py:262-272 -- values are copied INTO the arenas (the modules keep their identity)."""
        with torch.no_grad():
            for name in ("policy", "qf1", "qf2", "target_qf1", "target_qf2"):
                for dst, src in zip(getattr(self, name).parameters(), snapshot[name].parameters()):
                    dst.copy_(src.to(dst.device))
        st = self.engine.get_state()
        for slot, key, arena in ((2, "policy_optimizer", "policy"), (0, "qf1_optimizer", "qf1"), (1, "qf2_optimizer", "qf2")):
            src = snapshot[key]
            mine = getattr(self, key)
            for p_dst, p_src in zip(mine.param_groups[0]["params"], src.param_groups[0]["params"]):
                if p_src in src.state:
                    mine.state[p_dst] = src.state[p_src]
            st.adam_step[slot] = self._load_optimizer(mine, self._arenas[arena])
        st.log_alpha = float(snapshot["log_alpha"].detach().cpu())
        a_src = snapshot["alpha_optimizer"]
        for p_src in a_src.param_groups[0]["params"]:
            if p_src in a_src.state:
                s = a_src.state[p_src]
                st.alpha_exp_avg, st.alpha_exp_avg_sq = float(s["exp_avg"]), float(s["exp_avg_sq"])
                st.alpha_step = int(float(s["step"]))
        self.engine.set_state(st)


class SoftActorCriticV(_FusedTrainer):
    """sac.py:12-273 -- the older SAC with a V function, a target V and a FIXED entropy coefficient
    (what run_scripts/sac_exp_script.py instantiates)."""

    def __init__(self, policy, qf1, qf2, vf, reward_scale=1.0, discount=0.99, alpha=1.0, policy_lr=1e-3, qf_lr=1e-3,
                 vf_lr=1e-3, soft_target_tau=1e-2, policy_mean_reg_weight=1e-3, policy_std_reg_weight=1e-3,
                 optimizer_class=optim.Adam, beta_1=0.9, batch_size=256, max_steps_per_call=1000, gemm_precision=None,
                 **kwargs):
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("the fused step implements torch.optim.Adam only")
        self.policy, self.qf1, self.qf2, self.vf = policy, qf1, qf2, vf
        self.reward_scale, self.discount, self.soft_target_tau = reward_scale, discount, soft_target_tau
        self.policy_mean_reg_weight, self.policy_std_reg_weight = policy_mean_reg_weight, policy_std_reg_weight
        self.alpha = alpha
        in_dim, H, act_dim, ls = module_dims(policy)
        if not ls:
            raise NotImplementedError("SAC needs a ReparamTanhMultivariateGaussianPolicy (conditioned_std=True)")
        for q in (qf1, qf2, vf):
            check_activations(q, "identity")
        self.target_vf = clone_module(vf)
        self._arenas = OrderedDict(policy=adopt_module(policy), qf1=adopt_module(qf1), qf2=adopt_module(qf2),
                                   vf=adopt_module(vf), target_vf=adopt_module(self.target_vf, False))
        self.policy_optimizer = optim.Adam(self.policy.parameters(), lr=policy_lr, betas=(beta_1, 0.999))
        self.qf1_optimizer = optim.Adam(self.qf1.parameters(), lr=qf_lr, betas=(beta_1, 0.999))
        self.qf2_optimizer = optim.Adam(self.qf2.parameters(), lr=qf_lr, betas=(beta_1, 0.999))
        self.vf_optimizer = optim.Adam(self.vf.parameters(), lr=vf_lr, betas=(beta_1, 0.999))
        cfg = _abi.TrainerConfig()
        cfg.algo = _abi.ALGO_SAC_V
        cfg.gemm_precision = default_gemm_precision() if gemm_precision is None else int(gemm_precision)
        cfg.obs_dim, cfg.act_dim, cfg.batch = in_dim, act_dim, int(batch_size)
        cfg.max_steps_per_call = int(max_steps_per_call)
        cfg.reward_scale, cfg.discount, cfg.soft_target_tau = reward_scale, discount, soft_target_tau
        cfg.policy_lr, cfg.qf_lr, cfg.vf_lr = policy_lr, qf_lr, vf_lr
        cfg.beta_1, cfg.beta_2, cfg.adam_eps = beta_1, 0.999, 1e-8
        cfg.alpha, cfg.train_alpha = alpha, 0
        cfg.policy_mean_reg_weight, cfg.policy_std_reg_weight = policy_mean_reg_weight, policy_std_reg_weight
        cfg.max_act = 1.0
        self._finish_init(cfg, list(self._arenas.values()))
        self.eval_statistics = None

    @property
    def networks(self):
        return [self.policy, self.qf1, self.qf2, self.vf, self.target_vf]

    def _build_stats(self, L, vec):
        B, A = self._cfg.batch, self._cfg.act_dim
        st = OrderedDict()
        st["Reward Scale"] = self.reward_scale
        st["QF1 Loss"] = float(L[_abi.L_QF1])
        st["QF2 Loss"] = float(L[_abi.L_QF2])
        st["VF Loss"] = float(L[_abi.L_VF])
        st["Policy Loss"] = float(L[_abi.L_POLICY])
        st.update(_stats("Q1 Predictions", vec[0:B]))
        st.update(_stats("Q2 Predictions", vec[B:2 * B]))
        st.update(_stats("V Predictions", vec[5 * B:6 * B]))
        o = 6 * B
        st.update(_stats("Log Pis", vec[o:o + B]))
        st.update(_stats("Policy mu", vec[o + B:o + B + B * A]))
        st.update(_stats("Policy log std", vec[o + B + B * A:o + B + 2 * B * A]))
        return st

    def get_snapshot(self):
        st = self.engine.get_state()
        for slot, name in ((2, "policy"), (0, "qf1"), (1, "qf2"), (3, "vf")):
            self._sync_optimizer(getattr(self, name + "_optimizer"), self._arenas[name], st.adam_step[slot])
        return dict(qf1=self.qf1, qf2=self.qf2, policy=self.policy, vf=self.vf, target_vf=self.target_vf,
                    policy_optimizer=self.policy_optimizer, qf1_optimizer=self.qf1_optimizer,
                    qf2_optimizer=self.qf2_optimizer, vf_optimizer=self.vf_optimizer)

    def load_snapshot(self, snapshot):
        """sac.py:258-268 (whose 'self.vf_optimizer' key typo is NOT reproduced)."""
        with torch.no_grad():
            for name in ("policy", "qf1", "qf2", "vf", "target_vf"):
                for dst, src in zip(getattr(self, name).parameters(), snapshot[name].parameters()):
                    dst.copy_(src.to(dst.device))
        st = self.engine.get_state()
        for slot, name in ((2, "policy"), (0, "qf1"), (1, "qf2"), (3, "vf")):
            key = name + "_optimizer"
            if key not in snapshot:
                continue
            src, mine = snapshot[key], getattr(self, key)
            for p_dst, p_src in zip(mine.param_groups[0]["params"], src.param_groups[0]["params"]):
                if p_src in src.state:
                    mine.state[p_dst] = src.state[p_src]
            st.adam_step[slot] = self._load_optimizer(mine, self._arenas[name])
        self.engine.set_state(st)


class TD3(_FusedTrainer):
    """td3.py:13-223.  The target-action noise is the policy MODULE's (policy.noise /
    policy.noise_clip, policies.py:150-152); the trainer arguments target_policy_noise* are stored
    but unused, exactly as in the reference (td3.py:46-47,82-83)."""

    def __init__(self, policy, qf1, qf2, reward_scale=1.0, discount=0.99, target_policy_noise=0.2,
                 target_policy_noise_clip=0.5, policy_lr=1e-3, qf_lr=1e-3, policy_and_target_update_period=2,
                 soft_target_tau=0.005, qf_criterion=None, optimizer_class=optim.Adam, batch_size=256,
                 max_steps_per_call=1000, gemm_precision=None, **kwargs):
        if optimizer_class is not optim.Adam:
            raise NotImplementedError("the fused step implements torch.optim.Adam only")
        if qf_criterion is not None and not isinstance(qf_criterion, torch.nn.MSELoss):
            raise NotImplementedError("the fused step implements the default MSELoss criterion")
        self.qf1, self.qf2, self.policy = qf1, qf2, policy
        self.reward_scale, self.discount = reward_scale, discount
        self.target_policy_noise, self.target_policy_noise_clip = target_policy_noise, target_policy_noise_clip
        self.policy_and_target_update_period, self.soft_target_tau = policy_and_target_update_period, soft_target_tau
        in_dim, H, act_dim, ls = module_dims(policy)
        if ls:
            raise NotImplementedError("TD3 needs a deterministic MlpGaussianNoisePolicy")
        check_activations(policy, "tanh")
        for q in (qf1, qf2):
            check_activations(q, "identity")
        self.target_policy = clone_module(policy)
        self.target_qf1, self.target_qf2 = clone_module(qf1), clone_module(qf2)
        self._arenas = OrderedDict(
            policy=adopt_module(policy), qf1=adopt_module(qf1), qf2=adopt_module(qf2),
            target_qf1=adopt_module(self.target_qf1, False), target_qf2=adopt_module(self.target_qf2, False),
            target_policy=adopt_module(self.target_policy, False))
        self.qf1_optimizer = optim.Adam(self.qf1.parameters(), lr=qf_lr)
        self.qf2_optimizer = optim.Adam(self.qf2.parameters(), lr=qf_lr)
        self.policy_optimizer = optim.Adam(self.policy.parameters(), lr=policy_lr)
        cfg = _abi.TrainerConfig()
        cfg.algo = _abi.ALGO_TD3
        cfg.gemm_precision = default_gemm_precision() if gemm_precision is None else int(gemm_precision)
        cfg.obs_dim, cfg.act_dim, cfg.batch = in_dim, act_dim, int(batch_size)
        cfg.max_steps_per_call = int(max_steps_per_call)
        cfg.reward_scale, cfg.discount, cfg.soft_target_tau = reward_scale, discount, soft_target_tau
        cfg.policy_lr, cfg.qf_lr = policy_lr, qf_lr
        cfg.beta_1, cfg.beta_2, cfg.adam_eps, cfg.alpha = 0.9, 0.999, 1e-8, 1.0
        cfg.policy_and_target_update_period = int(policy_and_target_update_period)
        cfg.policy_noise = float(getattr(policy, "noise", 0.1))
        cfg.policy_noise_clip = float(getattr(policy, "noise_clip", 0.5))
        cfg.max_act = float(getattr(policy, "max_act", 1.0))
        self._configure(cfg, policy, kwargs)
        self._finish_init(cfg, list(self._arenas.values()))
        self.eval_statistics = None

    def _configure(self, cfg, policy, kwargs):
        pass

    @property
    def _n_train_steps_total(self):
        return self.engine.get_state().n_train_steps_total

    @property
    def networks(self):
        return [self.policy, self.qf1, self.qf2, self.target_policy, self.target_qf1, self.target_qf2]

    def _build_stats(self, L, vec):
        B, A = self._cfg.batch, self._cfg.act_dim
        st = OrderedDict()
        st["QF1 Loss"] = float(L[_abi.L_QF1])
        st["QF2 Loss"] = float(L[_abi.L_QF2])
        # on non-policy steps the reference evaluates a stats-only policy loss (td3.py:131-136): the step program runs the
        # policy / Q1 forward passes on the statistics step of a launch too (COND_TD3_POLICY_OR_STATS)
        st["Policy Loss"] = float(L[_abi.L_POLICY])
        st.update(_stats("Q1 Predictions", vec[0:B]))
        st.update(_stats("Q2 Predictions", vec[B:2 * B]))
        st.update(_stats("Q Targets", vec[2 * B:3 * B]))
        st.update(_stats("Bellman Errors 1", vec[3 * B:4 * B]))
        st.update(_stats("Bellman Errors 2", vec[4 * B:5 * B]))
        st.update(_stats("Policy Action", self._cfg.max_act * vec[6 * B:6 * B + B * A]))
        return st

    def get_snapshot(self):
        st = self.engine.get_state()
        self._sync_optimizer(self.policy_optimizer, self._arenas["policy"], st.adam_step[2])
        self._sync_optimizer(self.qf1_optimizer, self._arenas["qf1"], st.adam_step[0])
        self._sync_optimizer(self.qf2_optimizer, self._arenas["qf2"], st.adam_step[1])
        return dict(qf1=self.qf1, qf2=self.qf2, policy=self.policy, target_policy=self.target_policy,
                    target_qf1=self.target_qf1, target_qf2=self.target_qf2, policy_optimizer=self.policy_optimizer,
                    qf1_optimizer=self.qf1_optimizer, qf2_optimizer=self.qf2_optimizer)

    def load_snapshot(self, snapshot):
        """td3.py:198-206 (whose 'self.qf2_optimizer' key typo is NOT reproduced)."""
        with torch.no_grad():
            for name in ("policy", "qf1", "qf2", "target_qf1", "target_qf2", "target_policy"):
                if name in snapshot:
                    for dst, src in zip(getattr(self, name).parameters(), snapshot[name].parameters()):
                        dst.copy_(src.to(dst.device))
        st = self.engine.get_state()
        for slot, key, arena in ((2, "policy_optimizer", "policy"), (0, "qf1_optimizer", "qf1"), (1, "qf2_optimizer", "qf2")):
            if key not in snapshot:
                continue
            src, mine = snapshot[key], getattr(self, key)
            for p_dst, p_src in zip(mine.param_groups[0]["params"], src.param_groups[0]["params"]):
                if p_src in src.state:
                    mine.state[p_dst] = src.state[p_src]
            st.adam_step[slot] = self._load_optimizer(mine, self._arenas[arena])
        self.engine.set_state(st)


class _HindsightMixin:
    """Goal-conditioned batches (desired_goals keys) for train_step; hindsight sampling inside the gather for
    train_from_buffer on a DeviceHindsightReplayBuffer."""

    def train_step(self, batch):
        super().train_step(_concat_goals(batch))

    def train_from_buffer(self, replay_buffer, n_steps, inject=None):
        """With a DeviceHindsightReplayBuffer the hindsight sampling + relabel (relabel_replay_buffer.py:63-131) runs
        inside the step program's gather phase; inject (parity mode) then also carries `idx_her` [T, B]."""
        if hasattr(replay_buffer, "her_desc"):
            her_idx = None
            if inject is not None:
                inject = dict(inject)
                her_idx = inject.pop("idx_her").contiguous()
            if her_idx is not None and n_steps > self._cfg.max_steps_per_call:
                raise ValueError("injected hindsight runs must fit one launch")
            self.engine.set_her(replay_buffer.her_desc(self._cfg.batch, her_idx))
            try:
                super().train_from_buffer(replay_buffer, n_steps, inject=inject)
            finally:
                self.engine.set_her(None)
        else:
            super().train_from_buffer(replay_buffer, n_steps, inject=inject)


class HerTD3(_HindsightMixin, TD3):
    """rlkit/torch/algorithms/her/td3.py:14-245 -- goal-conditioned TD3 (exp_specs/her/her_*_td3.yaml through
    run_scripts/her_td3_exp_script.py:69-88).  Networks take cat(observation, desired_goal); the three differences from
    TD3 are those of the reference (her/td3.py:103-112, :116-120, :150-152): the next action is the clipped noise alone
    (`clamp(sigma * N(0,1), min_act, max_act)` -- the target policy's output is overwritten there), the min target Q is
    clipped to [clip_return_l, clip_return_r], and the policy loss carries + mean(action^2).

    train_step(batch) takes the reference's goal-conditioned batch (keys desired_goals / next_desired_goals next to the
    usual five) and concatenates on the device; train_from_buffer() works on a DeviceReplayBuffer whose rows already
    hold cat(obs, goal) (observation_dim = obs_dim + goal_dim)."""

    def __init__(self, policy, qf1, qf2, clip_return_l=None, clip_return_r=None, **kwargs):
        self._her_clip = (clip_return_l, clip_return_r)
        super().__init__(policy, qf1, qf2, **kwargs)
        self.clip_return_l, self.clip_return_r = self._cfg.clip_return_l, self._cfg.clip_return_r

    def _configure(self, cfg, policy, kwargs):
        cl, cr = self._her_clip
        gamma_sum = 1.0 / (1.0 - cfg.discount)                       # her/td3.py:81-85
        cfg.her = 1
        cfg.her_sigma = float(getattr(policy, "sigma", 0.2))         # target_policy.sigma (policies.py:506)
        cfg.min_act = float(getattr(policy, "min_act", -1.0))
        cfg.clip_return_l = -gamma_sum if cl is None else float(cl)
        cfg.clip_return_r = 0.0 if cr is None else float(cr)


def _concat_goals(batch):
    """her/td3.py:94-98, her/sac.py:80-84: networks see cat(obs, desired_goal) / cat(next_obs, next_desired_goal)."""
    if "desired_goals" not in batch:
        return batch

    def dev(x):
        return torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device="cuda", dtype=torch.float32)
    batch = dict(batch)
    batch["observations"] = torch.cat([dev(batch["observations"]), dev(batch["desired_goals"])], dim=-1)
    batch["next_observations"] = torch.cat([dev(batch["next_observations"]), dev(batch["next_desired_goals"])], dim=-1)
    return batch


class HerSAC(_HindsightMixin, SoftActorCritic):
    """rlkit/torch/algorithms/her/sac.py:12-251 (run_scripts/her_sac_exp_script.py) -- goal-conditioned SAC with the
    auto-tuned alpha.  The step is sac_alpha's on cat(observation, desired_goal) (her/sac.py:80-143); the one numeric
    difference is the default target entropy: -prod(action_space.shape) (her/sac.py:52), not half of it."""

    def __init__(self, policy, qf1, qf2, **kwargs):
        if kwargs.get("target_entropy") is None:
            if "env" in kwargs:
                kwargs["target_entropy"] = -float(np.prod(kwargs["env"].action_space.shape))
            else:
                kwargs["target_entropy"] = -float(list(policy.parameters())[4].shape[0])
        super().__init__(policy, qf1, qf2, **kwargs)
