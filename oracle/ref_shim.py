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

/root/reference does NOT exist on the GPU box: oracle/make_golden.py and the `-m "not gpu"` validation
tests use it here; on the GPU box the same unmodified sources are importable from baseline/_ref (pip
--target install, see _find_root) for bench.py's reference arm and the drop-in loop tests.
"""
import builtins
import os
import sys
import types

import numpy as np
import torch

# Where the unmodified reference lives: the read-only checkout of the build container, or -- on the GPU boxes, where that
# does not exist -- the pip --target install of the same sources under baseline/_ref (git-ignored, ships with the lease;
# installed from a /tmp copy with a 3-line setup.py because the reference has no packaging metadata).
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_root():
    env = os.environ.get("ILSWISS_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", os.path.join(_REPO, "baseline", "_ref")):
        if os.path.isdir(os.path.join(cand, "rlkit")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()


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


def _install_gtimer(mod):
    """Just enough of `gtimer` for BaseAlgorithm.start_training (base_algorithm.py:152-157,166-172,281-291,330-340):
    timed_for(iterable, save_itrs), stamp(name), get_times().stamps.itrs[name][-1], get_times().total, reset(),
    set_def_unique().  Per iteration of timed_for every stamp name gets one entry (the time since the previous stamp,
    summed when a name is stamped more than once in the iteration)."""
    import time as _time

    state = types.SimpleNamespace(itrs={}, cur={}, last=_time.time(), start=_time.time())

    def reset():
        state.itrs, state.cur = {}, {}
        state.last = state.start = _time.time()

    def stamp(name, *a, **k):
        now = _time.time()
        state.cur[name] = state.cur.get(name, 0.0) + (now - state.last)
        state.last = now

    def _close_iteration():
        for k, v in state.cur.items():
            state.itrs.setdefault(k, []).append(v)
        state.cur = {}

    def timed_for(iterable, *a, **k):
        for x in iterable:
            yield x
            _close_iteration()

    def get_times():
        # _try_to_eval runs INSIDE an iteration: the entries of the running iteration are visible as the last element
        itrs = {k: list(v) for k, v in state.itrs.items()}
        for k, v in state.cur.items():
            itrs.setdefault(k, []).append(v)
        return types.SimpleNamespace(stamps=types.SimpleNamespace(itrs=itrs), total=_time.time() - state.start)

    mod.reset, mod.stamp, mod.timed_for, mod.get_times = reset, stamp, timed_for, get_times
    mod.set_def_unique = lambda *a, **k: None


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
    _install_gtimer(sys.modules["gtimer"])
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


def time_reference(w, steps, warmup, threads):
    """bench.py's reference arm: the reference's OWN classes (SimpleReplayBuffer.random_batch -> np_to_pytorch_batch ->
    trainer.train_step; AdvIRL._do_reward_training + _do_policy_training for GAIL) on the host cores, on the synthetic
    workload `w` of bench.WORKLOADS.  Returns (gradient steps / s, seconds)."""
    import time

    from oracle import restate as R

    ref = import_reference()
    torch.set_num_threads(threads)
    O, A, B = w["O"], w["A"], w["B"]
    H = w.get("H", 256)
    n = min(w["N"], 1_000_000)
    env = FakeEnv(O, A)
    torch.manual_seed(0)
    qf1 = ref.FlattenMlp(hidden_sizes=[H, H], input_size=O + A, output_size=1)
    qf2 = ref.FlattenMlp(hidden_sizes=[H, H], input_size=O + A, output_size=1)

    def fill(buf, m, seed, term_p):
        d = R.synth_transitions(m, O, A, seed, term_p)
        buf._observations[:m] = d["observations"]; buf._actions[:m] = d["actions"]; buf._rewards[:m] = d["rewards"]
        buf._terminals[:m] = d["terminals"]; buf._next_obs[:m] = d["next_observations"]
        buf._top, buf._size = m % buf._max_replay_buffer_size, m

    buf = ref.SimpleReplayBuffer(n, O, A, random_seed=1)
    fill(buf, n, 7, 0.0 if w["algo"] == "gail" else 0.01)
    if w["algo"] == "td3" and w.get("her"):
        raise RuntimeError("the HER workload is timed with the oracle port")
    if w["algo"] == "td3":
        policy = ref.MlpGaussianNoisePolicy(hidden_sizes=[H, H], obs_dim=O, action_dim=A, policy_noise=0.2, policy_noise_clip=0.5,
                                            output_activation=torch.tanh)
        trainer = ref.TD3(policy=policy, qf1=qf1, qf2=qf2, reward_scale=1.0, discount=0.99, policy_lr=3e-4, qf_lr=3e-4,
                          policy_and_target_update_period=2, tau=0.005)
    else:
        policy = ref.ReparamTanhMultivariateGaussianPolicy(hidden_sizes=[H, H], obs_dim=O, action_dim=A)
        kw = dict(policy=policy, qf1=qf1, qf2=qf2, env=env, reward_scale=w["reward_scale"], discount=0.99, policy_lr=3e-4, qf_lr=3e-4,
                  alpha_lr=3e-4, soft_target_tau=0.005, alpha=0.2, train_alpha=True, policy_mean_reg_weight=1e-3,
                  policy_std_reg_weight=1e-3, beta_1=w["beta_1"])
        if w["target_entropy"] is not None:
            kw["target_entropy"] = w["target_entropy"]
        trainer = ref.SacAlpha(**kw)
    alg = None
    if w["algo"] == "gail":
        ebuf = ref.SimpleReplayBuffer(n, O, A, random_seed=3)
        fill(ebuf, w["NE"], 8, 0.0)
        disc = ref.MLPDisc(O + A, num_layer_blocks=2, hid_dim=128, hid_act="tanh", use_bn=False, clamp_magnitude=10.0)
        alg = ref.AdvIRL(mode="gail2", discriminator=disc, policy_trainer=trainer, expert_replay_buffer=ebuf, state_only=False,
                         disc_optim_batch_size=B, policy_optim_batch_size=B, policy_optim_batch_size_from_expert=0,
                         num_update_loops_per_train_call=1, num_disc_updates_per_loop_iter=1, num_policy_updates_per_loop_iter=1,
                         disc_lr=3e-4, disc_momentum=0.9, use_grad_pen=True, grad_pen_weight=8.0, env=env, training_env=None,
                         exploration_policy=policy, replay_buffer=buf, max_path_length=1000, no_terminal=True)

    def one():
        if alg is not None:
            alg._do_training(0)       # adv_irl.py:126-131: one reward update + one policy update
        else:
            trainer.train_step(ref.np_to_pytorch_batch(buf.random_batch(B)))     # torch_rl_algorithm.py:28-34

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return steps / dt, dt
