"""TEST INFRASTRUCTURE ONLY -- import shim that makes the *unmodified* reference
(Ericonaldo/ILSwiss, mounted read-only at /root/reference) importable in this
container so it can be executed as the parity oracle and to generate the golden
vectors under tests/golden/.  Nothing in the product path may import this module.

Why a shim is needed:
  * rlkit/core/eval_util.py:12 -> vistools.py imports matplotlib/seaborn (absent);
  * rlkit/core/base_algorithm.py:5 imports gtimer (absent);
  * rlkit/data_management/env_replay_buffer.py:4 imports gym.spaces (absent);
  * rlkit/torch/algorithms/torch_base_algorithm.py:24 uses `torch` without
    importing it (NameError on import) -> we inject builtins.torch.

/root/reference does NOT exist on the GPU box: only oracle/make_golden.py and
the `-m "not gpu"` validation tests (skipped when the directory is absent) use it.
"""
import builtins
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("ILSWISS_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "rlkit"))


class _Space:
    """Just enough of gym.spaces.Box for EnvReplayBuffer/get_dim and sac_alpha.py:55-58."""

    def __init__(self, n):
        self.shape = (n,)
        self.low = -np.ones(n)
        self.high = np.ones(n)

    def sample(self):
        return np.random.uniform(-1, 1, self.shape)


class FakeEnv:
    def __init__(self, obs_dim, act_dim):
        import gym.spaces as gs  # the stub below

        self.observation_space = gs.Box(obs_dim)
        self.action_space = gs.Box(act_dim)


class FakeGoalEnv:
    """Goal-environment stand-in for HindsightReplayBuffer (relabel_replay_buffer.py:27-48): Dict observation space and the
    gym-robotics sparse compute_reward."""

    def __init__(self, obs_dim, goal_dim, act_dim, distance_threshold=0.05):
        import gym.spaces as gs

        d = gs.Dict(1)
        d.spaces = dict(observation=gs.Box(obs_dim), achieved_goal=gs.Box(goal_dim), desired_goal=gs.Box(goal_dim))
        self.observation_space, self.action_space = d, gs.Box(act_dim)
        self.distance_threshold = distance_threshold

    def compute_reward(self, achieved_goal, goal, info):
        d = np.linalg.norm(achieved_goal - goal, axis=-1)
        return -(d > self.distance_threshold).astype(np.float32)


_installed = False


def install():
    """Idempotent.  Stubs absent third-party modules and puts the reference on sys.path."""
    global _installed
    if _installed:
        return
    if not reference_available():
        raise RuntimeError(
            "reference checkout not found at %s (expected only in the build container)"
            % REFERENCE_ROOT
        )
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in [
        "matplotlib",
        "matplotlib.pyplot",
        "matplotlib.animation",
        "seaborn",
        "gtimer",
        "gym",
        "gym.spaces",
        "envpool",
    ]:
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].animation = sys.modules["matplotlib.animation"]
    sys.modules["matplotlib.animation"].FuncAnimation = object
    sys.modules["seaborn"].set = lambda *a, **k: None
    gs = sys.modules["gym.spaces"]
    sys.modules["gym"].spaces = gs
    for c in ["Box", "Discrete", "Tuple", "Dict"]:
        if not hasattr(gs, c):
            setattr(gs, c, type(c, (_Space,), {}))
    sys.modules["gym"].Env = object
    sys.modules["gym"].Wrapper = object
    sys.modules["gym"].Space = _Space
    builtins.torch = torch  # torch_base_algorithm.py:24 lacks `import torch`
    _installed = True


def import_reference():
    """Returns a namespace with the reference classes on the hot path."""
    install()
    import rlkit.torch.utils.pytorch_util as ptu
    from rlkit.data_management.simple_replay_buffer import SimpleReplayBuffer
    from rlkit.torch.core import np_to_pytorch_batch
    from rlkit.torch.common.networks import FlattenMlp
    from rlkit.torch.common.policies import (
        ReparamTanhMultivariateGaussianPolicy,
        MlpGaussianNoisePolicy,
        MlpGaussianAndEpsilonConditionPolicy,
    )
    from rlkit.torch.common.policies import ReparamTanhMultivariateGaussianConditionPolicy
    from rlkit.torch.algorithms.her.td3 import TD3 as HerTD3
    from rlkit.torch.algorithms.her.sac import SAC as HerSAC
    from rlkit.torch.algorithms.sac.sac_alpha import SoftActorCritic as SacAlpha
    from rlkit.torch.algorithms.sac.sac import SoftActorCritic as SacV
    from rlkit.torch.algorithms.td3.td3 import TD3
    from rlkit.torch.algorithms.adv_irl.adv_irl import AdvIRL
    from rlkit.torch.algorithms.adv_irl.disc_models.simple_disc_models import MLPDisc

    ptu.set_gpu_mode.__doc__  # touch
    ptu._use_gpu = False
    ptu.device = torch.device("cpu")
    ns = types.SimpleNamespace(**{k: v for k, v in locals().items() if k != "ns"})
    return ns
