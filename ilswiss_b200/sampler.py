"""Sampler <-> learner coupling.

    DevicePolicy       <->  the exploration policy's numpy interface: get_action / get_actions
                            (rlkit/torch/common/policies.py:241-246, 154-164; eval_np rlkit/torch/core.py:74-89;
                            called every env step by BaseAlgorithm, base_algorithm.py:369-380)
    MakeDeterministic  <->  rlkit/torch/common/policies.py:19-36 (the eval policy, base_algorithm.py:82-110)

The reference moves the observations of all vec-envs to the device, runs ~12 eager kernels and copies the
actions back, with two blocking copies, on every env step.  Here the same round trip is ONE C-ABI call
(`ilsw_policy_act_host`): pinned H2D, one kernel that reads the policy arena the step engine trains in place,
pinned D2H, one stream synchronisation -- stream ordered behind the gradient steps already queued, so the
sampler always acts with the latest policy, and transitions go back through DeviceReplayBuffer.add_sample
(pinned cudaMemcpyAsync on a side stream).

Exploration noise comes from an in-kernel Philox stream keyed by (seed, call counter, row): the same
distribution as the reference's CPU torch.normal draw (policies.py:276-283 / distributions.py:23-28), not the
same bit stream (documented deviation, as for the in-kernel index sampling).
"""
import numpy as np

from . import layout


class DevicePolicy:
    """Numpy-interface exploration policy backed by a fused trainer's policy arena.

    `DevicePolicy(trainer)` replaces `exploration_policy=policy` in the experiment scripts
    (sac_alpha_exp_script.py:97-104); `trainer.policy` (the nn.Module) stays valid and shares the weights."""

    def __init__(self, trainer, seed=None):
        self.trainer = trainer
        self.stochastic_policy = trainer.policy
        # the numpy / python global RNG streams belong to the reference's code (exploration noise, hindsight sampling) and
        # must not be advanced here: the default seed is derived from numpy's state without drawing from it
        self._seed = layout.derive_seed(salt=0x5a17) if seed is None else int(seed)
        self._calls = 0

    def get_actions(self, obs_np, deterministic=False):
        obs = np.asarray(obs_np, dtype=np.float32)
        self._calls += 1
        return self.trainer.engine.policy_act_host(obs, deterministic=deterministic,
                                                   seed=(self._seed << 20) + self._calls)

    def get_action(self, obs_np, deterministic=False):
        actions = self.get_actions(np.asarray(obs_np)[None], deterministic=deterministic)
        return actions[0, :], {}

    # ExplorationPolicy / Policy surface used by BaseAlgorithm and the samplers (rlkit/policies/base.py)
    def reset(self):
        pass

    def set_num_steps_total(self, t):
        pass

    def train(self, mode=True):
        self.stochastic_policy.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        return self

    def parameters(self):
        return self.stochastic_policy.parameters()

    def state_dict(self):
        return self.stochastic_policy.state_dict()


class MakeDeterministic:
    """policies.py:19-36 on top of a DevicePolicy (or any policy with the same numpy interface)."""

    def __init__(self, stochastic_policy):
        self.stochastic_policy = stochastic_policy

    def get_action(self, observation):
        return self.stochastic_policy.get_action(observation, deterministic=True)

    def get_actions(self, observations):
        return self.stochastic_policy.get_actions(observations, deterministic=True)

    def reset(self):
        pass

    def train(self, mode):
        pass

    def set_num_steps_total(self, num):
        pass

    def to(self, device):
        self.stochastic_policy.to(device)


class HerDevicePolicy(DevicePolicy):
    """Goal-conditioned exploration policy of the HER scripts (MlpGaussianAndEpsilonConditionPolicy,
    rlkit/torch/common/policies.py:481-560 + ConditionPolicy.get_actions :618-642) on the device round trip.

    Observations may be the goal environments' dicts (or lists of dicts): cat(observation, desired_goal) goes to the kernel
    (deterministic forward, max_act * tanh), and the exploration rule of the reference runs on the host with the SAME
    random sources and call order (:546-560): with probability epsilon (python `random`) the whole batch is replaced by
    action_space samples, otherwise Gaussian noise (numpy global RNG) with the decayed sigma is added and the result
    clipped to [min_act, max_act]."""

    def __init__(self, trainer, action_space, epsilon=0.3, max_sigma=0.2, min_sigma=0.2, decay_period=1000000,
                 max_act=1.0, min_act=-1.0, observation_key="observation", desired_goal_key="desired_goal",
                 achieved_goal_key="achieved_goal", seed=None):
        super().__init__(trainer, seed=seed)
        self._action_space, self._epsilon = action_space, epsilon
        self._max_sigma, self._min_sigma = max_sigma, (max_sigma if min_sigma is None else min_sigma)
        self._decay_period, self.max_act, self.min_act = decay_period, max_act, min_act
        self.sigma, self.t = max_sigma, 0
        self.observation_key, self.desired_goal_key, self.achieved_goal_key = observation_key, desired_goal_key, achieved_goal_key

    def set_num_steps_total(self, t):
        self.t = t

    def _flatten(self, obs):
        if isinstance(obs, dict):
            return np.concatenate([obs[self.observation_key], obs[self.desired_goal_key]], axis=-1)
        if len(obs) and isinstance(obs[0], dict):
            return np.array([np.concatenate([x[self.observation_key], x[self.desired_goal_key]], axis=-1) for x in obs])
        return np.asarray(obs)

    def _deterministic(self, obs2d):
        self._calls += 1
        return self.trainer.engine.policy_act_host(obs2d, deterministic=True, seed=(self._seed << 20) + self._calls)

    def get_actions(self, obs_np, deterministic=False):
        import random

        obs = np.asarray(self._flatten(obs_np), dtype=np.float32)
        single = obs.ndim == 1
        action = self._deterministic(obs[None] if single else obs)
        action = action[0] if single else action
        if deterministic:
            return action
        if random.random() < self._epsilon:                                           # :546-550
            action = self._action_space.sample()             # drawn (and discarded for a batch) exactly as the reference does
            if not single:
                action = [self._action_space.sample() for _ in range(obs.shape[0])]
            return action
        self.sigma = self._max_sigma - (self._max_sigma - self._min_sigma) * min(1.0, self.t * 1.0 / self._decay_period)
        return np.clip(action + np.random.normal(size=np.shape(action)) * self.sigma, self.min_act, self.max_act)

    def get_action(self, obs_np, deterministic=False):
        obs = self._flatten(obs_np)
        actions = self.get_actions(np.asarray(obs)[None], deterministic=deterministic)
        return np.asarray(actions)[0, :], {}
