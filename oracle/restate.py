"""TEST INFRASTRUCTURE ONLY -- CPU restatement (torch fp32 on CPU + numpy) of the
reference's off-policy hot path.  It is the parity oracle for the CUDA kernels.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference`
legs may import this module.  The product path (ilswiss_b200/) never does.

Pinning status: the reference ships NO golden vectors or known-answer tests for this
path.  The oracle is therefore pinned by executing the unmodified
reference itself in the build container (oracle/ref_shim.py + oracle/make_golden.py) and
committing the transcripts under tests/golden/; tests/test_oracle_golden.py replays
them against this file.  The arithmetic itself lives in PyTorch (reference pins
torch==1.9.0, requirements.txt:26; this image has 2.11.0) and numpy's legacy MT19937
`RandomState.randint` -- both are used here exactly as the reference's call sites do.

Every function cites the reference file:line it restates (paths relative to the
reference checkout root).
"""
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

LOG_SIG_MAX = 2.0  # rlkit/torch/common/policies.py:15
LOG_SIG_MIN = -20.0  # rlkit/torch/common/policies.py:16
LOG_2PI = float(np.log(2 * np.pi))  # rlkit/torch/common/distributions.py:8


# --------------------------------------------------------------------------------------
# Synthetic data -- draw order fixed HERE and used by every consumer
# --------------------------------------------------------------------------------------
def synth_transitions(n, obs_dim, act_dim, seed, terminal_p=0.01):
    """obs,next_obs ~ N(0,1); actions ~ U(-1,1); rewards ~ N(0,1); terminals ~ Bern(p).
    float64 / uint8 like the reference buffer fields (simple_replay_buffer.py:48-60)."""
    rs = np.random.RandomState(seed)
    obs = rs.randn(n, obs_dim)
    act = rs.uniform(-1.0, 1.0, (n, act_dim))
    rew = rs.randn(n, 1)
    nobs = rs.randn(n, obs_dim)
    term = (rs.rand(n, 1) < terminal_p).astype(np.uint8)
    return OrderedDict(
        observations=obs, actions=act, rewards=rew, terminals=term, next_observations=nobs
    )


# --------------------------------------------------------------------------------------
# R1-R4: replay buffer + numpy->torch batch conversion
# --------------------------------------------------------------------------------------
class ReplayOracle:
    """Restates SimpleReplayBuffer (rlkit/data_management/simple_replay_buffer.py:17-323),
    flat-observation case only (observation_dim is an int)."""

    ALL_KEYS = (
        "observations",
        "actions",
        "rewards",
        "terminals",
        "next_observations",
        "absorbing",
    )

    def __init__(self, max_replay_buffer_size, observation_dim, action_dim, random_seed=1995):
        # simple_replay_buffer.py:17-68
        self._np_rand_state = np.random.RandomState(random_seed)
        self._observation_dim = observation_dim
        self._action_dim = action_dim
        self._max_replay_buffer_size = max_replay_buffer_size
        n = max_replay_buffer_size
        self._observations = np.zeros((n, observation_dim))
        self._next_obs = np.zeros((n, observation_dim))
        self._actions = np.zeros((n, action_dim))
        self._rewards = np.zeros((n, 1))
        self._terminals = np.zeros((n, 1), dtype="uint8")
        self._timeouts = np.zeros((n, 1), dtype="uint8")
        self._absorbing = np.zeros((n, 2))
        self._top = 0
        self._size = 0
        self._trajs = 0
        self._cur_start = 0
        self._traj_endpoints = {}

    def add_sample(self, observation, action, reward, terminal, next_observation,
                   timeout=False, **kwargs):
        # simple_replay_buffer.py:78-108
        self._actions[self._top] = action
        self._rewards[self._top] = reward
        self._terminals[self._top] = terminal
        self._timeouts[self._top] = timeout
        if "absorbing" in kwargs:
            self._absorbing[self._top] = kwargs["absorbing"]
        if terminal:
            next_start = (self._top + 1) % self._max_replay_buffer_size
            self._traj_endpoints[self._cur_start] = next_start
            self._cur_start = next_start
        self._observations[self._top] = observation
        self._next_obs[self._top] = next_observation
        self._advance()

    def terminate_episode(self):
        # simple_replay_buffer.py:125-132
        if self._cur_start != self._top:
            self._traj_endpoints[self._cur_start] = self._top
            self._cur_start = self._top

    def add_path(self, path):
        # simple_replay_buffer.py:134-162 (absorbing=False branch)
        for ob, action, reward, next_ob, terminal in zip(
            path["observations"], path["actions"], path["rewards"],
            path["next_observations"], path["terminals"],
        ):
            self.add_sample(observation=ob, action=action, reward=reward,
                            terminal=terminal, next_observation=next_ob)
        self.terminate_episode()
        self._trajs += 1

    def _advance(self):
        # simple_replay_buffer.py:228-237
        if self._top in self._traj_endpoints:
            del self._traj_endpoints[self._top]
        self._top = (self._top + 1) % self._max_replay_buffer_size
        if self._size < self._max_replay_buffer_size:
            self._size += 1

    def num_steps_can_sample(self):
        return self._size

    def load_bulk(self, data):
        """Test convenience (not in the reference): equivalent to add_sample() over the rows,
        without the per-row python loop.  Leaves trajectory bookkeeping empty."""
        n = len(data["observations"])
        assert n <= self._max_replay_buffer_size
        self._observations[:n] = data["observations"]
        self._actions[:n] = data["actions"]
        self._rewards[:n] = data["rewards"]
        self._terminals[:n] = data["terminals"]
        self._next_obs[:n] = data["next_observations"]
        self._top = n % self._max_replay_buffer_size
        self._size = n

    def sample_indices(self, batch_size):
        # simple_replay_buffer.py:242 (uniform WITH replacement, legacy MT19937)
        return self._np_rand_state.randint(0, self._size, batch_size)

    def random_batch(self, batch_size, keys=None):
        # simple_replay_buffer.py:239-253
        return self.get_batch_using_indices(self.sample_indices(batch_size), keys=keys)

    def get_batch_using_indices(self, indices, keys=None):
        # simple_replay_buffer.py:255-293
        if keys is None:
            keys = set(self.ALL_KEYS)
        ret = {}
        if "observations" in keys:
            ret["observations"] = self._observations[indices]
        if "actions" in keys:
            ret["actions"] = self._actions[indices]
        if "rewards" in keys:
            ret["rewards"] = self._rewards[indices]
        if "terminals" in keys:
            ret["terminals"] = self._terminals[indices]
        if "next_observations" in keys:
            ret["next_observations"] = self._next_obs[indices]
        if "absorbing" in keys:
            ret["absorbing"] = self._absorbing[indices]
        return ret


# --------------------------------------------------------------------------------------
# HER: HindsightReplayBuffer (rlkit/data_management/relabel_replay_buffer.py:12-131) on top of the dict-observation mode
# of SimpleReplayBuffer (simple_replay_buffer.py:31-43,99-104,270-279).  Observations are dicts with the keys
# observation / achieved_goal / desired_goal; the reward function is the goal environments' sparse compute_reward
# (-(||achieved - desired|| > distance_threshold), e.g. gym robotics FetchEnv) that the reference takes from
# env.compute_reward (:36-37).
# --------------------------------------------------------------------------------------
class HindsightOracle:
    OBS_KEYS = ("observation", "achieved_goal", "desired_goal")

    def __init__(self, max_replay_buffer_size, obs_dim, goal_dim, action_dim, random_seed=1995, relabel_type="future",
                 her_ratio=0.8, distance_threshold=0.05):
        self._np_rand_state = np.random.RandomState(random_seed)
        N = self._max_replay_buffer_size = max_replay_buffer_size
        dims = dict(observation=obs_dim, achieved_goal=goal_dim, desired_goal=goal_dim)
        self._observations = {k: np.zeros((N, d)) for k, d in dims.items()}
        self._next_obs = {k: np.zeros((N, d)) for k, d in dims.items()}
        self._actions = np.zeros((N, action_dim))
        self._rewards = np.zeros((N, 1))
        self._terminals = np.zeros((N, 1), dtype="uint8")
        self._top = self._size = self._cur_start = 0
        self._traj_endpoints = {}
        self.relabel_type, self.her_ratio, self.distance_threshold = relabel_type, her_ratio, distance_threshold

    def compute_reward(self, achieved, goal, info=None):
        d = np.linalg.norm(achieved - goal, axis=-1)
        return -(d > self.distance_threshold).astype(np.float32)

    def add_sample(self, observation, action, reward, terminal, next_observation, **kwargs):
        # simple_replay_buffer.py:78-108
        self._actions[self._top] = action
        self._rewards[self._top] = reward
        self._terminals[self._top] = terminal
        if terminal:
            next_start = (self._top + 1) % self._max_replay_buffer_size
            self._traj_endpoints[self._cur_start] = next_start
            self._cur_start = next_start
        for k, v in observation.items():
            self._observations[k][self._top] = v
        for k, v in next_observation.items():
            self._next_obs[k][self._top] = v
        if self._top in self._traj_endpoints:              # _advance, :228-237
            del self._traj_endpoints[self._top]
        self._top = (self._top + 1) % self._max_replay_buffer_size
        if self._size < self._max_replay_buffer_size:
            self._size += 1

    def terminate_episode(self):
        if self._cur_start != self._top:                   # :125-132
            self._traj_endpoints[self._cur_start] = self._top
            self._cur_start = self._top

    def sample_indices(self, batch_size):
        """relabel_replay_buffer.py:70-95 -- the index draws of random_batch, in the reference's RNG order (the buffer's
        RandomState for trajectory shuffle / trajectory / step, the GLOBAL numpy RNG for the future step)."""
        relabel = (self.relabel_type is not None) and (self.her_ratio > 0)
        keys_list = list(self._traj_endpoints.keys())
        starts = self._np_rand_state.choice(keys_list, size=len(keys_list), replace=False)
        ends = [self._traj_endpoints[k] for k in starts]
        traj_indice = self._np_rand_state.randint(0, len(starts), batch_size)
        indices, indices_relabel = [], []
        for i in traj_indice:
            traj_len = (ends[i] - starts[i]) % self._size
            step = (self._np_rand_state.randint(0, traj_len, 1)[0] + starts[i]) % self._size
            indices.append(step)
            if relabel:
                step_her = {"final": ends[i] - 1,
                            "future": np.random.randint(step, (traj_len + starts[i])) % self._size}[self.relabel_type]
                indices_relabel.append(step_her)
        return np.asarray(indices), np.asarray(indices_relabel)

    def batch_from_indices(self, indices, indices_relabel):
        """:97-131 -- gather, relabel the first int(her_ratio * B) rows, recompute every reward."""
        relabel = (self.relabel_type is not None) and (self.her_ratio > 0)
        B = len(indices)
        obs = {k: v[indices].copy() for k, v in self._observations.items()}
        nobs = {k: v[indices].copy() for k, v in self._next_obs.items()}
        out = dict(actions=self._actions[indices], rewards=self._rewards[indices], terminals=self._terminals[indices])
        if relabel:
            n = int(self.her_ratio * B)
            src = self._next_obs["achieved_goal"][indices_relabel]
            obs["desired_goal"][:n] = src[:n]
            nobs["desired_goal"][:n] = src[:n]
        out["achieved_goals"], out["desired_goals"] = obs["achieved_goal"], obs["desired_goal"]
        out["next_achieved_goals"], out["next_desired_goals"] = nobs["achieved_goal"], nobs["desired_goal"]
        out["observations"], out["next_observations"] = obs["observation"], nobs["observation"]
        if relabel:
            out["rewards"] = self.compute_reward(out["next_achieved_goals"], out["desired_goals"]).reshape(-1, 1)
        return out

    def random_batch(self, batch_size):
        return self.batch_from_indices(*self.sample_indices(batch_size))


def synth_goal_episodes(rs, n_episodes, T, O0, G, A):
    """Synthetic goal-env episodes: random walks so that future achieved goals are sometimes within the threshold."""
    eps = []
    for _ in range(n_episodes):
        goal = rs.randn(G) * 0.05
        ag = np.cumsum(rs.randn(T + 1, G) * 0.02, axis=0)
        ob = rs.randn(T + 1, O0)
        eps.append([(dict(observation=ob[t], achieved_goal=ag[t], desired_goal=goal), rs.uniform(-1, 1, A), -1.0,
                     t == T - 1 and rs.rand() < 0.5, dict(observation=ob[t + 1], achieved_goal=ag[t + 1], desired_goal=goal))
                    for t in range(T)])
    return eps


def np_to_torch_batch(np_batch):
    """rlkit/torch/core.py:124-143 + pytorch_util.py:84-88: per key astype(float32)."""
    return {k: torch.from_numpy(v.astype(np.float32)) for k, v in np_batch.items()}


# --------------------------------------------------------------------------------------
# Parameter containers + deterministic init (nets "arrive constructed" in the reference;
# this init only mirrors the *shape* of the reference init so goldens are reproducible
# anywhere from a seed.  networks.py:58-83, pytorch_util.py:20-29)
# --------------------------------------------------------------------------------------
def mlp_param_names(n_hidden, extra_heads=()):
    names = []
    for i in range(n_hidden):
        names += ["fc%d.weight" % i, "fc%d.bias" % i]
    names += ["last_fc.weight", "last_fc.bias"]
    for h in extra_heads:
        names += [h + ".weight", h + ".bias"]
    return names


def init_mlp(rs, in_dim, hidden, out_dim, init_w=3e-3, b_init=0.1, log_std_head=False):
    """Returns OrderedDict name -> float32 ndarray in the reference's parameter order
    (policies.py param order: fc0.w, fc0.b, fc1.w, fc1.b, last_fc.w, last_fc.b,
    last_fc_log_std.w, last_fc_log_std.b)."""
    p = OrderedDict()
    d = in_dim
    for i, h in enumerate(hidden):
        bound = 1.0 / np.sqrt(h)  # fanin_init uses size[0] == out_features (pytorch_util.py:23)
        p["fc%d.weight" % i] = rs.uniform(-bound, bound, (h, d)).astype(np.float32)
        p["fc%d.bias" % i] = np.full((h,), b_init, np.float32)
        d = h
    p["last_fc.weight"] = rs.uniform(-init_w, init_w, (out_dim, d)).astype(np.float32)
    p["last_fc.bias"] = rs.uniform(-init_w, init_w, (out_dim,)).astype(np.float32)
    if log_std_head:
        p["last_fc_log_std.weight"] = rs.uniform(-init_w, init_w, (out_dim, d)).astype(np.float32)
        p["last_fc_log_std.bias"] = rs.uniform(-init_w, init_w, (out_dim,)).astype(np.float32)
    return p


def init_disc(rs, in_dim, hid=128):
    """MLPDisc with num_layer_blocks=2, no BN (simple_disc_models.py:29-41): parameters()
    de-dups to mod_list.{0,2,4}.{weight,bias}; default nn.Linear init U(+-1/sqrt(fan_in))."""
    p = OrderedDict()
    dims = [(hid, in_dim), (hid, hid), (1, hid)]
    for idx, (o, i) in zip((0, 2, 4), dims):
        b = 1.0 / np.sqrt(i)
        p["mod_list.%d.weight" % idx] = rs.uniform(-b, b, (o, i)).astype(np.float32)
        p["mod_list.%d.bias" % idx] = rs.uniform(-b, b, (o,)).astype(np.float32)
    return p


class Net:
    """Parameters + Adam moments of one network, as torch fp32 CPU tensors."""

    def __init__(self, params):
        self.p = OrderedDict((k, torch.tensor(np.asarray(v), dtype=torch.float32)) for k, v in params.items())
        self.m = OrderedDict((k, torch.zeros_like(v)) for k, v in self.p.items())
        self.v = OrderedDict((k, torch.zeros_like(v)) for k, v in self.p.items())
        self.t = 0  # torch.optim.Adam per-optimizer step count

    def clone(self):
        n = Net({k: v.numpy().copy() for k, v in self.p.items()})
        n.m = OrderedDict((k, v.clone()) for k, v in self.m.items())
        n.v = OrderedDict((k, v.clone()) for k, v in self.v.items())
        n.t = self.t
        return n

    def leaves(self):
        return [t.detach().clone().requires_grad_(True) for t in self.p.values()]

    def flat(self):
        return np.concatenate([v.numpy().ravel() for v in self.p.values()])

    def flat_m(self):
        return np.concatenate([v.numpy().ravel() for v in self.m.values()])

    def flat_v(self):
        return np.concatenate([v.numpy().ravel() for v in self.v.values()])


# --------------------------------------------------------------------------------------
# N1-N3, D0: network forwards
# --------------------------------------------------------------------------------------
def mlp_forward(w, x, n_hidden=2):
    """Mlp.forward, relu hidden, identity output (networks.py:85-101).  w: list of leaves
    [W0,b0,W1,b1,W2,b2,...]."""
    h = x
    for i in range(n_hidden):
        h = F.relu(F.linear(h, w[2 * i], w[2 * i + 1]))
    return F.linear(h, w[2 * n_hidden], w[2 * n_hidden + 1]), h


def q_forward(w, obs, act):
    """FlattenMlp.forward (networks.py:108-115): cat(inputs, dim=1) then Mlp.forward."""
    return mlp_forward(w, torch.cat([obs, act], dim=1))[0]


def tanh_gaussian_forward(w, obs, eps):
    """ReparamTanhMultivariateGaussianPolicy.forward(return_log_prob=True)
    (policies.py:248-307) with the N(0,1) draw of distributions.py:24 injected as `eps`;
    log-prob per distributions.py:43-50,74-95 (0.5*log(2pi) added ONCE, not x A)."""
    _, h = mlp_forward(w[:4] + w[4:6], obs)  # h = last hidden
    mean = F.linear(h, w[4], w[5])
    log_std = torch.clamp(F.linear(h, w[6], w[7]), LOG_SIG_MIN, LOG_SIG_MAX)
    sig = torch.exp(log_std)
    cov = torch.exp(2.0 * log_std)
    z = eps * sig + mean
    a = torch.tanh(z)
    log_prob = -0.5 * torch.sum((mean - z) ** 2 / cov, 1, keepdim=True)
    log_prob = log_prob - (torch.sum(log_std, 1, keepdim=True) + 0.5 * LOG_2PI)
    log_prob = log_prob - torch.sum(torch.log(1 - a ** 2 + 1e-6), 1, keepdim=True)
    return a, mean, log_std, log_prob


def td3_policy_forward(w, obs, noise=None, policy_noise=0.2, noise_clip=0.5, max_act=1.0):
    """MlpGaussianNoisePolicy.forward (policies.py:166-188) built with
    output_activation=tanh (td3_exp_script.py:71-78).  `noise` is the injected N(0,1)
    draw of policies.py:182-184; None == deterministic.  No re-clip of the action."""
    pre, _ = mlp_forward(w, obs)
    a = max_act * torch.tanh(pre)
    if noise is not None:
        a = a + torch.clamp(policy_noise * noise, -noise_clip, noise_clip)
    return a


def disc_forward(w, x, clamp=10.0, hid_act="tanh"):
    """MLPDisc.forward, 2 Linear + hid_act blocks, no BN (simple_disc_models.py:19-24,29-48)."""
    act = {"tanh": torch.tanh, "relu": torch.relu}[hid_act]
    h = act(F.linear(x, w[0], w[1]))
    h = act(F.linear(h, w[2], w[3]))
    return torch.clamp(F.linear(h, w[4], w[5]), -clamp, clamp)


# --------------------------------------------------------------------------------------
# torch.optim.Adam (single-tensor path, eps=1e-8, no weight decay) and Polyak (G1)
# --------------------------------------------------------------------------------------
def adam_update(net, grads, lr, beta1, beta2=0.999, eps=1e-8):
    """Restates torch/optim/adam.py::_single_tensor_adam as called from
    sac_alpha.py:65-76 / td3.py:56-67 / adv_irl.py:75-77."""
    net.t += 1
    t = net.t
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    step_size = lr / bc1
    bc2_sqrt = bc2 ** 0.5
    with torch.no_grad():
        for (k, p), g in zip(net.p.items(), grads):
            m, v = net.m[k], net.v[k]
            m.lerp_(g, 1 - beta1)
            v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
            denom = (v.sqrt() / bc2_sqrt).add_(eps)
            p.addcdiv_(m, denom, value=-step_size)


def polyak(source, target, tau):
    """ptu.soft_update_from_to (pytorch_util.py:10-12)."""
    with torch.no_grad():
        for k in target.p:
            target.p[k].copy_(target.p[k] * (1.0 - tau) + source.p[k] * tau)


class ScalarAdam:
    """Adam state for the float64 0-dim log_alpha (sac_alpha.py:51-53,74-76)."""

    def __init__(self):
        self.m = 0.0
        self.v = 0.0
        self.t = 0

    def update(self, p, g, lr, beta1, beta2=0.999, eps=1e-8):
        self.t += 1
        w = 1 - beta1
        # torch lerp_: weight<0.5 ? a + w(b-a) : b - (b-a)(1-w)
        self.m = self.m + w * (g - self.m) if w < 0.5 else g - (g - self.m) * (1 - w)
        self.v = self.v * beta2 + (1 - beta2) * g * g
        bc1 = 1 - beta1 ** self.t
        bc2 = 1 - beta2 ** self.t
        denom = np.sqrt(self.v) / np.sqrt(bc2) + eps
        return p + (-(lr / bc1)) * self.m / denom


# --------------------------------------------------------------------------------------
# S1: SAC with auto-tuned alpha (sac_alpha.py:78-181)
# --------------------------------------------------------------------------------------
class SacAlphaOracle:
    def __init__(self, policy, qf1, qf2, action_dim, reward_scale=1.0, discount=0.99,
                 policy_lr=1e-3, qf_lr=1e-3, alpha_lr=3e-4, soft_target_tau=1e-2, alpha=0.2,
                 train_alpha=True, policy_mean_reg_weight=1e-3, policy_std_reg_weight=1e-3,
                 beta_1=0.9, target_entropy=None):
        # sac_alpha.py:21-76
        self.policy, self.qf1, self.qf2 = policy, qf1, qf2
        self.target_qf1, self.target_qf2 = qf1.clone(), qf2.clone()
        self.reward_scale, self.discount = reward_scale, discount
        self.policy_lr, self.qf_lr, self.alpha_lr = policy_lr, qf_lr, alpha_lr
        self.tau, self.beta_1 = soft_target_tau, beta_1
        self.mean_reg, self.std_reg = policy_mean_reg_weight, policy_std_reg_weight
        self.train_alpha = train_alpha
        self.log_alpha = float(np.log(alpha))  # float64 0-dim tensor in the reference
        self.alpha_opt = ScalarAdam()
        self.target_entropy = (-action_dim / 2.0) if target_entropy is None else target_entropy
        self.last = {}

    @property
    def alpha(self):
        return float(np.exp(self.log_alpha))

    def train_step(self, batch, eps_next, eps_cur):
        rewards = self.reward_scale * batch["rewards"]
        terminals, obs = batch["terminals"], batch["observations"]
        actions, next_obs = batch["actions"], batch["next_observations"]
        alpha = torch.tensor(self.alpha, dtype=torch.float64)  # 0-dim f64, promotes to f32

        w1, w2 = self.qf1.leaves(), self.qf2.leaves()
        wp = self.policy.leaves()
        tw1 = [t.detach() for t in self.target_qf1.p.values()]
        tw2 = [t.detach() for t in self.target_qf2.p.values()]
        q1_pred = q_forward(w1, obs, actions)
        q2_pred = q_forward(w2, obs, actions)
        with torch.no_grad():
            na, _, _, nlogpi = tanh_gaussian_forward(wp, next_obs, eps_next)  # :105-112
            tmin = torch.min(q_forward(tw1, next_obs, na), q_forward(tw2, next_obs, na))
            q_target = rewards + (1.0 - terminals) * self.discount * (tmin - alpha * nlogpi)
        qf1_loss = 0.5 * torch.mean((q1_pred - q_target) ** 2)  # :126-127
        qf2_loss = 0.5 * torch.mean((q2_pred - q_target) ** 2)
        g1 = torch.autograd.grad(qf1_loss, w1)
        g2 = torch.autograd.grad(qf2_loss, w2)
        adam_update(self.qf1, g1, self.qf_lr, self.beta_1)  # :136-137
        adam_update(self.qf2, g2, self.qf_lr, self.beta_1)

        # policy loss with the UPDATED critics (:146-160)
        w1n = [t.detach() for t in self.qf1.p.values()]
        w2n = [t.detach() for t in self.qf2.p.values()]
        new_a, mean, log_std, log_pi = tanh_gaussian_forward(wp, obs, eps_cur)
        q_new = torch.min(q_forward(w1n, obs, new_a), q_forward(w2n, obs, new_a))
        policy_loss = torch.mean(alpha * log_pi - q_new)
        policy_loss = policy_loss + self.mean_reg * (mean ** 2).mean() + self.std_reg * (log_std ** 2).mean()
        gp = torch.autograd.grad(policy_loss, wp)
        adam_update(self.policy, gp, self.policy_lr, self.beta_1)

        alpha_loss = None
        if self.train_alpha:  # :165-171
            log_prob = log_pi.detach() + self.target_entropy
            la = torch.tensor(self.log_alpha, dtype=torch.float64, requires_grad=True)
            alpha_loss = -(la * log_prob).mean()
            (ga,) = torch.autograd.grad(alpha_loss, la)
            self.log_alpha = float(self.alpha_opt.update(self.log_alpha, float(ga), self.alpha_lr, self.beta_1))

        polyak(self.qf1, self.target_qf1, self.tau)  # :245-247
        polyak(self.qf2, self.target_qf2, self.tau)
        self.last = dict(
            qf1_loss=float(qf1_loss), qf2_loss=float(qf2_loss), policy_loss=float(policy_loss),
            alpha_loss=None if alpha_loss is None else float(alpha_loss), alpha=self.alpha,
            q1_pred=q1_pred.detach().numpy().ravel(), q2_pred=q2_pred.detach().numpy().ravel(),
            q_target=q_target.numpy().ravel(), log_pi=log_pi.detach().numpy().ravel(),
            policy_mean=mean.detach().numpy(), policy_log_std=log_std.detach().numpy(),
            grads=dict(qf1=[g.numpy() for g in g1], qf2=[g.numpy() for g in g2],
                       policy=[g.numpy() for g in gp]),
        )
        return self.last


# --------------------------------------------------------------------------------------
# S2: SAC, V-function variant, fixed alpha (sac.py:70-179,242-243)
# --------------------------------------------------------------------------------------
class SacVOracle:
    def __init__(self, policy, qf1, qf2, vf, reward_scale=1.0, discount=0.99, alpha=1.0,
                 policy_lr=1e-3, qf_lr=1e-3, vf_lr=1e-3, soft_target_tau=1e-2,
                 policy_mean_reg_weight=1e-3, policy_std_reg_weight=1e-3, beta_1=0.9):
        self.policy, self.qf1, self.qf2, self.vf = policy, qf1, qf2, vf
        self.target_vf = vf.clone()
        self.reward_scale, self.discount, self.alpha = reward_scale, discount, alpha
        self.policy_lr, self.qf_lr, self.vf_lr = policy_lr, qf_lr, vf_lr
        self.tau, self.beta_1 = soft_target_tau, beta_1
        self.mean_reg, self.std_reg = policy_mean_reg_weight, policy_std_reg_weight

    def train_step(self, batch, eps_cur):
        rewards = self.reward_scale * batch["rewards"]
        terminals, obs = batch["terminals"], batch["observations"]
        actions, next_obs = batch["actions"], batch["next_observations"]
        w1, w2, wv, wp = self.qf1.leaves(), self.qf2.leaves(), self.vf.leaves(), self.policy.leaves()
        twv = [t.detach() for t in self.target_vf.p.values()]
        q1_pred, q2_pred = q_forward(w1, obs, actions), q_forward(w2, obs, actions)
        with torch.no_grad():
            q_target = rewards + (1.0 - terminals) * self.discount * mlp_forward(twv, next_obs)[0]
        qf1_loss = 0.5 * torch.mean((q1_pred - q_target) ** 2)
        qf2_loss = 0.5 * torch.mean((q2_pred - q_target) ** 2)
        v_pred = mlp_forward(wv, obs)[0]
        new_a, mean, log_std, log_pi = tanh_gaussian_forward(wp, obs, eps_cur)
        with torch.no_grad():  # v_target detached (sac.py:128-129), pre-update critics
            w1o = [t.detach() for t in self.qf1.p.values()]
            w2o = [t.detach() for t in self.qf2.p.values()]
            q_new_old = torch.min(q_forward(w1o, obs, new_a), q_forward(w2o, obs, new_a))
            v_target = q_new_old - self.alpha * log_pi
        vf_loss = 0.5 * torch.mean((v_pred - v_target) ** 2)
        g1 = torch.autograd.grad(qf1_loss, w1)
        g2 = torch.autograd.grad(qf2_loss, w2)
        gv = torch.autograd.grad(vf_loss, wv)
        adam_update(self.qf1, g1, self.qf_lr, self.beta_1)  # sac.py:136-139
        adam_update(self.qf2, g2, self.qf_lr, self.beta_1)
        adam_update(self.vf, gv, self.vf_lr, self.beta_1)
        # policy loss: updated critics, SAME action sample / log_pi graph (sac.py:150-163)
        w1n = [t.detach() for t in self.qf1.p.values()]
        w2n = [t.detach() for t in self.qf2.p.values()]
        q_new = torch.min(q_forward(w1n, obs, new_a), q_forward(w2n, obs, new_a))
        policy_loss = torch.mean(self.alpha * log_pi - q_new)
        policy_loss = policy_loss + self.mean_reg * (mean ** 2).mean() + self.std_reg * (log_std ** 2).mean()
        gp = torch.autograd.grad(policy_loss, wp)
        adam_update(self.policy, gp, self.policy_lr, self.beta_1)
        polyak(self.vf, self.target_vf, self.tau)
        return dict(qf1_loss=float(qf1_loss), qf2_loss=float(qf2_loss), vf_loss=float(vf_loss),
                    policy_loss=float(policy_loss))


# --------------------------------------------------------------------------------------
# S3: TD3 (td3.py:72-178)
# --------------------------------------------------------------------------------------
class TD3Oracle:
    def __init__(self, policy, qf1, qf2, reward_scale=1.0, discount=0.99, policy_lr=1e-3,
                 qf_lr=1e-3, policy_and_target_update_period=2, soft_target_tau=0.005,
                 policy_noise=0.2, policy_noise_clip=0.5):
        # td3.py:20-70.  NB the trainer's own target_policy_noise* args are unused; the
        # noise parameters are those of the policy MODULE (policies.py:150-152).
        self.policy, self.qf1, self.qf2 = policy, qf1, qf2
        self.target_policy, self.target_qf1, self.target_qf2 = policy.clone(), qf1.clone(), qf2.clone()
        self.reward_scale, self.discount = reward_scale, discount
        self.policy_lr, self.qf_lr = policy_lr, qf_lr
        self.period, self.tau = policy_and_target_update_period, soft_target_tau
        self.noise, self.noise_clip = policy_noise, policy_noise_clip
        self.n_total = 0

    def train_step(self, batch, noise):
        rewards = self.reward_scale * batch["rewards"]
        terminals, obs = batch["terminals"], batch["observations"]
        actions, next_obs = batch["actions"], batch["next_observations"]
        w1, w2, wp = self.qf1.leaves(), self.qf2.leaves(), self.policy.leaves()
        with torch.no_grad():
            twp = list(self.target_policy.p.values())
            tw1 = list(self.target_qf1.p.values())
            tw2 = list(self.target_qf2.p.values())
            na = td3_policy_forward(twp, next_obs, noise, self.noise, self.noise_clip)
            tq = torch.min(q_forward(tw1, next_obs, na), q_forward(tw2, next_obs, na))
            q_target = rewards + (1.0 - terminals) * self.discount * tq
        q1_pred = q_forward(w1, obs, actions)
        qf1_loss = ((q1_pred - q_target) ** 2).mean()  # no 1/2 (td3.py:93-98)
        q2_pred = q_forward(w2, obs, actions)
        qf2_loss = ((q2_pred - q_target) ** 2).mean()
        g1 = torch.autograd.grad(qf1_loss, w1)
        adam_update(self.qf1, g1, self.qf_lr, 0.9)  # default betas (td3.py:56-67)
        g2 = torch.autograd.grad(qf2_loss, w2)
        adam_update(self.qf2, g2, self.qf_lr, 0.9)
        policy_loss = None
        if self.n_total % self.period == 0:  # td3.py:113 (steps 0,2,4,...)
            w1n = [t.detach() for t in self.qf1.p.values()]
            pa = td3_policy_forward(wp, obs, None)
            policy_loss = -q_forward(w1n, obs, pa).mean()
            gp = torch.autograd.grad(policy_loss, wp)
            adam_update(self.policy, gp, self.policy_lr, 0.9)
            polyak(self.policy, self.target_policy, self.tau)  # td3.py:180-183
            polyak(self.qf1, self.target_qf1, self.tau)
            polyak(self.qf2, self.target_qf2, self.tau)
        # td3.py:126-136: when statistics are collected on a step without a policy update, the reference evaluates the
        # policy loss once more, for logging only (updated critics, unchanged policy)
        stats_pl = policy_loss
        if stats_pl is None:
            with torch.no_grad():
                w1n = [t.detach() for t in self.qf1.p.values()]
                stats_pl = -q_forward(w1n, obs, td3_policy_forward([t.detach() for t in wp], obs, None)).mean()
        self.n_total += 1
        return dict(qf1_loss=float(qf1_loss), qf2_loss=float(qf2_loss),
                    policy_loss=None if policy_loss is None else float(policy_loss), stats_policy_loss=float(stats_pl),
                    q1_pred=q1_pred.detach().numpy().ravel(), q_target=q_target.numpy().ravel())


# --------------------------------------------------------------------------------------
# HER-TD3 (rlkit/torch/algorithms/her/td3.py:88-160).  Observations are cat(obs, desired_goal) /
# cat(next_obs, next_desired_goal) (:97-98); everything else as TD3 except the three marked lines.
# --------------------------------------------------------------------------------------
class HerTD3Oracle(TD3Oracle):
    def __init__(self, policy, qf1, qf2, sigma=0.2, min_act=-1.0, max_act=1.0, clip_return_l=None,
                 clip_return_r=None, **kw):
        kw.pop("policy_noise", None), kw.pop("policy_noise_clip", None)
        super().__init__(policy, qf1, qf2, **kw)
        self.sigma, self.min_act, self.max_act = sigma, min_act, max_act
        gamma_sum = 1.0 / (1.0 - self.discount)                      # :81-85
        self.clip_l = -gamma_sum if clip_return_l is None else clip_return_l
        self.clip_r = 0.0 if clip_return_r is None else clip_return_r

    def train_step(self, batch, noise):
        rewards = self.reward_scale * batch["rewards"]
        terminals, obs = batch["terminals"], batch["observations"]
        actions, next_obs = batch["actions"], batch["next_observations"]
        w1, w2, wp = self.qf1.leaves(), self.qf2.leaves(), self.policy.leaves()
        with torch.no_grad():
            tw1 = list(self.target_qf1.p.values())
            tw2 = list(self.target_qf2.p.values())
            # :103-112 -- `noisy_next_actions = clamp(noise, min_act, max_act)`: the target policy's action is
            # overwritten, the next action is the clipped noise alone
            na = torch.clamp(self.sigma * noise, self.min_act, self.max_act)
            tq = torch.clip(torch.min(q_forward(tw1, next_obs, na), q_forward(tw2, next_obs, na)),
                            self.clip_l, self.clip_r)                # :116-120
            q_target = rewards + (1.0 - terminals) * self.discount * tq
        q1_pred = q_forward(w1, obs, actions)
        qf1_loss = ((q1_pred - q_target) ** 2).mean()
        q2_pred = q_forward(w2, obs, actions)
        qf2_loss = ((q2_pred - q_target) ** 2).mean()
        g1 = torch.autograd.grad(qf1_loss, w1)
        adam_update(self.qf1, g1, self.qf_lr, 0.9)
        g2 = torch.autograd.grad(qf2_loss, w2)
        adam_update(self.qf2, g2, self.qf_lr, 0.9)
        policy_loss = None
        if self.n_total % self.period == 0:
            w1n = [t.detach() for t in self.qf1.p.values()]
            pa = td3_policy_forward(wp, obs, None, max_act=self.max_act)
            policy_loss = -q_forward(w1n, obs, pa).mean() + torch.square(pa).mean()      # :150-152
            gp = torch.autograd.grad(policy_loss, wp)
            adam_update(self.policy, gp, self.policy_lr, 0.9)
            polyak(self.policy, self.target_policy, self.tau)
            polyak(self.qf1, self.target_qf1, self.tau)
            polyak(self.qf2, self.target_qf2, self.tau)
        self.n_total += 1
        return dict(qf1_loss=float(qf1_loss), qf2_loss=float(qf2_loss),
                    policy_loss=None if policy_loss is None else float(policy_loss),
                    q1_pred=q1_pred.detach().numpy().ravel(), q_target=q_target.numpy().ravel())


# --------------------------------------------------------------------------------------
# D1 + D2: AdvIRL discriminator step and reward relabel (adv_irl.py:133-314)
# --------------------------------------------------------------------------------------
def disc_reward(logits, mode, rew_clip_min=None, rew_clip_max=None):
    """adv_irl.py:277-298."""
    if mode == "airl":
        r = logits
    elif mode == "gail":
        r = F.softplus(logits, beta=1)
    elif mode == "gail2":
        r = F.softplus(logits, beta=-1)
    elif mode == "fairl":
        r = torch.exp(logits) * (-1.0 * logits)
    else:
        raise ValueError(mode)
    if rew_clip_max is not None:
        r = torch.clamp(r, max=rew_clip_max)
    if rew_clip_min is not None:
        r = torch.clamp(r, min=rew_clip_min)
    return r


class DiscOracle:
    def __init__(self, disc, disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True,
                 grad_pen_weight=10.0, clamp=10.0, hid_act="tanh"):
        self.disc, self.hid_act = disc, hid_act
        self.lr, self.beta1 = disc_lr, disc_momentum  # adv_irl.py:75-77
        self.use_grad_pen, self.gp_w, self.clamp = use_grad_pen, grad_pen_weight, clamp

    def reward_step(self, expert_x, policy_x, gp_eps):
        """adv_irl.py:133-216.  expert_x/policy_x = cat(obs, act) rows (B,D)."""
        w = self.disc.leaves()
        B = expert_x.shape[0]
        x = torch.cat([expert_x, policy_x], dim=0)
        targets = torch.cat([torch.ones(B, 1), torch.zeros(B, 1)], dim=0)
        logits = disc_forward(w, x, self.clamp, self.hid_act)
        ce = F.binary_cross_entropy_with_logits(logits, targets)
        acc = ((logits > 0).float() == targets).float().mean()
        gp = torch.zeros(())
        total = ce
        if self.use_grad_pen:
            interp = (gp_eps * expert_x + (1 - gp_eps) * policy_x).detach().requires_grad_(True)
            (g,) = torch.autograd.grad(disc_forward(w, interp, self.clamp, self.hid_act).sum(), [interp],
                                       create_graph=True, retain_graph=True)
            gp = ((g.norm(2, dim=1) - 1) ** 2).mean()
            total = ce + gp * self.gp_w
        grads = torch.autograd.grad(total, w)
        adam_update(self.disc, grads, self.lr, self.beta1)
        return dict(disc_ce_loss=float(ce), disc_acc=float(acc), grad_pen=float(gp),
                    grads=[g.numpy() for g in grads])

    def rewards(self, obs, act, mode, rew_clip_min=None, rew_clip_max=None):
        """adv_irl.py:266-298 (disc in eval mode, logits detached).  `act` is next_obs in state_only mode (:265-269)."""
        with torch.no_grad():
            w = list(self.disc.p.values())
            return disc_reward(disc_forward(w, torch.cat([obs, act], dim=1), self.clamp, self.hid_act), mode,
                               rew_clip_min, rew_clip_max)


def param_digest(net):
    """Small per-tensor fingerprint stored in golden files: (sum, abs-sum, first 4 values)."""
    out = []
    for v in net.p.values():
        a = v.numpy().astype(np.float64).ravel()
        out.append(np.concatenate([[a.sum(), np.abs(a).sum()], a[:4] if a.size >= 4 else np.resize(a, 4)]))
    return np.stack(out)
