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


class DevicePolicy:
    """Numpy-interface exploration policy backed by a fused trainer's policy arena.

    `DevicePolicy(trainer)` replaces `exploration_policy=policy` in the experiment scripts
    (sac_alpha_exp_script.py:97-104); `trainer.policy` (the nn.Module) stays valid and shares the weights."""

    def __init__(self, trainer, seed=None):
        self.trainer = trainer
        self.stochastic_policy = trainer.policy
        self._seed = int(np.random.randint(1, 2 ** 31 - 1)) if seed is None else int(seed)
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
