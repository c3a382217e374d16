"""Sampler <-> learner coupling: DevicePolicy.get_actions through ilsw_policy_act_host."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sac(O=11, A=3, B=64):
    from ilswiss_b200 import modules
    from ilswiss_b200.trainers import SoftActorCritic

    torch.manual_seed(0)
    pol = modules.TanhGaussianPolicy([256, 256], O, A)
    tr = SoftActorCritic(pol, modules.FlattenMlp([256, 256], 1, O + A), modules.FlattenMlp([256, 256], 1, O + A), batch_size=B)
    return tr, pol


def test_device_policy_deterministic_matches_module_and_follows_training():
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.sampler import DevicePolicy, MakeDeterministic

    O, A = 11, 3
    tr, pol = _sac(O, A)
    dp = DevicePolicy(tr, seed=5)
    ev = MakeDeterministic(dp)
    rs = np.random.RandomState(0)
    obs = rs.randn(4, O)                      # float64 host observations, as the vec-env hands them over
    with torch.no_grad():
        ref = pol(torch.as_tensor(obs, dtype=torch.float32, device="cuda"), deterministic=True)[0].cpu().numpy()
    got = ev.get_actions(obs)
    assert got.shape == (4, A) and got.dtype == np.float32
    assert np.allclose(got, ref, atol=2e-6, rtol=1e-5)
    a1, info = ev.get_action(obs[0])
    assert info == {} and np.allclose(a1, ref[0], atol=2e-6, rtol=1e-5)
    # the sampler sees the parameters the engine trains in place (stream ordered behind the gradient steps)
    buf = DeviceReplayBuffer(1000, O, A, random_seed=1)
    buf.add_samples(rs.randn(500, O), rs.uniform(-1, 1, (500, A)), rs.randn(500), np.zeros(500), rs.randn(500, O))
    tr.train_from_buffer(buf, 20)
    after = ev.get_actions(obs)
    with torch.no_grad():
        ref2 = pol(torch.as_tensor(obs, dtype=torch.float32, device="cuda"), deterministic=True)[0].cpu().numpy()
    assert not np.allclose(after, got, atol=1e-5)
    assert np.allclose(after, ref2, atol=2e-6, rtol=1e-5)


def test_device_policy_exploration_noise_has_the_policy_distribution():
    """Stochastic actions: atanh(a) ~ N(mean, std) per dimension (policies.py:276-283), fresh noise per call."""
    from ilswiss_b200.sampler import DevicePolicy

    O, A = 11, 3
    tr, pol = _sac(O, A)
    dp = DevicePolicy(tr, seed=7)
    obs1 = np.random.RandomState(1).randn(1, O).astype(np.float32)
    obs = np.repeat(obs1, 4096, axis=0)
    a = dp.get_actions(obs)
    b = dp.get_actions(obs)
    assert np.abs(a).max() < 1.0 and not np.array_equal(a, b)            # new Philox counter per call
    assert len(np.unique(a[:, 0])) > 4000                                 # independent noise per row
    with torch.no_grad():
        out = pol(torch.as_tensor(obs1, device="cuda"))
    mean, log_std = out[1][0].cpu().numpy(), out[2][0].cpu().numpy()
    z = np.arctanh(np.clip(a.astype(np.float64), -1 + 1e-7, 1 - 1e-7))
    std = np.exp(log_std)
    assert np.all(np.abs(z.mean(0) - mean) < 5 * std / np.sqrt(4096))
    assert np.all(np.abs(z.std(0) / std - 1) < 0.05)
    with pytest.raises(ValueError):
        dp.get_actions(np.zeros((2, O + 1)))


def test_td3_device_policy():
    from ilswiss_b200 import modules
    from ilswiss_b200.sampler import DevicePolicy
    from ilswiss_b200.trainers import TD3

    torch.manual_seed(1)
    O, A = 17, 6
    pol = modules.DeterministicNoisePolicy([256, 256], O, A, policy_noise=0.2, policy_noise_clip=0.5)
    tr = TD3(pol, modules.FlattenMlp([256, 256], 1, O + A), modules.FlattenMlp([256, 256], 1, O + A), batch_size=64)
    dp = DevicePolicy(tr, seed=2)
    obs = np.random.RandomState(2).randn(8, O)
    with torch.no_grad():
        ref = pol(torch.as_tensor(obs, dtype=torch.float32, device="cuda"), deterministic=True)[0].cpu().numpy()
    assert np.allclose(dp.get_actions(obs, deterministic=True), ref, atol=2e-6, rtol=1e-5)
    noisy = dp.get_actions(obs)
    assert np.abs(noisy - ref).max() <= 0.5 + 1e-6 and np.abs(noisy - ref).max() > 0     # clipped Gaussian noise
