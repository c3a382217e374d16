"""HerDevicePolicy's host-side exploration rule against the executed reference policy class (build container only): the
device kernel is replaced by the reference module's own deterministic forward, so only the rule is under test."""
import random

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout absent")


def test_her_exploration_rule_matches_reference_policy():
    ref = ref_shim.import_reference()
    from ilswiss_b200.sampler import HerDevicePolicy

    O0, G, A = 10, 3, 4
    space = ref_shim.FakeEnv(O0 + G, A).action_space
    torch.manual_seed(0)
    pol = ref.MlpGaussianAndEpsilonConditionPolicy(hidden_sizes=[64, 64], action_space=space, obs_dim=O0, condition_dim=G,
                                                   action_dim=A, output_activation=torch.tanh, epsilon=0.3, max_sigma=0.3,
                                                   min_sigma=0.1, decay_period=1000)

    class Engine:                                   # stand-in for StepEngine.policy_act_host: the module's own forward
        def policy_act_host(self, obs, deterministic=False, seed=0):
            with torch.no_grad():
                return pol(torch.as_tensor(obs, dtype=torch.float32), deterministic=True)[0].numpy()

    class Trainer:
        policy, engine = pol, Engine()

    dp = HerDevicePolicy(Trainer(), space, epsilon=0.3, max_sigma=0.3, min_sigma=0.1, decay_period=1000, seed=1)
    rs = np.random.RandomState(0)
    for trial in range(12):
        obs = [dict(observation=rs.randn(O0), achieved_goal=rs.randn(G), desired_goal=rs.randn(G)) for _ in range(5)]
        t = 100 * trial
        pol.set_num_steps_total(t)
        dp.set_num_steps_total(t)
        random.seed(trial); np.random.seed(trial)
        want = pol.get_actions(obs)
        random.seed(trial); np.random.seed(trial)
        got = dp.get_actions(obs)
        np.testing.assert_allclose(np.asarray(got), np.asarray(want), rtol=0, atol=1e-6)
        assert dp.sigma == pol.sigma
    random.seed(2); np.random.seed(2)      # a seed whose first random.random() >= epsilon: the reference get_action cannot index the eps-branch list
    want, _ = pol.get_action(obs[0])
    random.seed(2); np.random.seed(2)      # a seed whose first random.random() >= epsilon: the reference get_action cannot index the eps-branch list
    got, info = dp.get_action(obs[0])
    np.testing.assert_allclose(got, want, rtol=0, atol=1e-6)
    assert info == {}
    np.testing.assert_allclose(dp.get_actions(obs, deterministic=True), pol.get_actions(obs, deterministic=True), atol=1e-6)


def test_her_mixin_builds_the_device_policy_from_the_reference_module():
    """DeviceHERMixin._get_action_and_info: exploration policy == the trainer's MlpGaussianAndEpsilonConditionPolicy ->
    HerDevicePolicy with the module's own exploration parameters; anything else -> the reference path."""
    ref = ref_shim.import_reference()
    from ilswiss_b200.adv_irl import DeviceHERMixin

    O0, G, A = 6, 2, 3
    space = ref_shim.FakeEnv(O0 + G, A).action_space
    pol = ref.MlpGaussianAndEpsilonConditionPolicy(hidden_sizes=[32, 32], action_space=space, obs_dim=O0, condition_dim=G,
                                                   action_dim=A, output_activation=torch.tanh, epsilon=0.0, max_sigma=0.25,
                                                   min_sigma=0.05, decay_period=500)

    class Engine:
        def policy_act_host(self, obs, deterministic=False, seed=0):
            with torch.no_grad():
                return pol(torch.as_tensor(obs, dtype=torch.float32), deterministic=True)[0].numpy()

    class Trainer:
        policy, engine = pol, Engine()

    class RefHER:
        def _get_action_and_info(self, observation):
            return "reference path"

    class HER(DeviceHERMixin, RefHER):
        pass

    alg = HER()
    alg.trainer, alg.exploration_policy, alg._n_env_steps_total = Trainer(), pol, 250
    obs = [dict(observation=np.ones(O0) * i, achieved_goal=np.zeros(G), desired_goal=np.ones(G)) for i in range(4)]
    np.random.seed(0); random.seed(0)
    got = alg._get_action_and_info(obs)
    np.random.seed(0); random.seed(0)
    pol.set_num_steps_total(250)
    want = pol.get_actions(obs)
    np.testing.assert_allclose(got, want, atol=1e-6)
    assert alg._ilsw_dp.sigma == pytest.approx(0.25 - 0.2 * 0.5) and alg._ilsw_dp._epsilon == 0.0
    other = HER()
    other.trainer, other.exploration_policy, other._n_env_steps_total = Trainer(), object(), 0
    assert other._get_action_and_info(obs) == "reference path"
