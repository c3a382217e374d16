"""Zero-edit drop-in: patches an importable ILSwiss checkout (package `rlkit`) so that its own, unmodified
`run_experiment.py` / `run_scripts/*_exp_script.py` use the B200 path.

    import ilswiss_b200.dropin as dropin
    dropin.install("/path/to/ILSwiss")          # before the experiment script is imported / run
    runpy.run_path("run_scripts/sac_alpha_exp_script.py", run_name="__main__")

The scripts import their classes by module path at run time (`from rlkit.torch.algorithms.sac.sac_alpha import
SoftActorCritic`, sac_alpha_exp_script.py:21), so replacing the module attributes is enough:

    rlkit.torch.algorithms.sac.sac_alpha.SoftActorCritic        -> trainers.SoftActorCritic
    rlkit.torch.algorithms.sac.sac.SoftActorCritic              -> trainers.SoftActorCriticV
    rlkit.torch.algorithms.td3.td3.TD3                          -> trainers.TD3
    rlkit.torch.algorithms.her.td3.TD3 / her.sac.SAC            -> trainers.HerTD3 / trainers.HerSAC
    rlkit.torch.algorithms.torch_rl_algorithm.TorchRLAlgorithm  -> DeviceTorchRLAlgorithmMixin in front of the original
    rlkit.torch.algorithms.adv_irl.adv_irl.AdvIRL               -> DeviceAdvIRLMixin in front of the original
    rlkit.data_management.env_replay_buffer.EnvReplayBuffer     -> replay_buffer.DeviceEnvReplayBuffer
      (also the name BaseAlgorithm bound at import time, base_algorithm.py:9,116-123)
    rlkit.torch.algorithms.her.her.HER                          -> DeviceHERMixin in front of the original
    rlkit.torch.algorithms.her.her.HindsightReplayBuffer        -> replay_buffer.DeviceEnvHindsightReplayBuffer
      (the name HER.__init__ uses when no buffer is passed, her.py:16-25)

The HER class subclasses the ORIGINAL TorchRLAlgorithm (bound when her.py was imported); with any other replay buffer or
trainer the mixin falls through to the reference's host relabel buffer and per-step `train_step(batch)` calls.
The trainers adopt the algorithm's `batch_size` on first use (`ensure_batch`).  `uninstall()` restores everything.
"""
import importlib
import sys

_saved = []


def _patch(module_name, attr, value):
    mod = importlib.import_module(module_name)
    _saved.append((mod, attr, getattr(mod, attr)))
    setattr(mod, attr, value)


def install(reference_root=None):
    """Idempotent.  `reference_root`: directory that contains the `rlkit` package (put on sys.path if given)."""
    if _saved:
        return
    if reference_root and reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    from . import adv_irl, replay_buffer, trainers

    _patch("rlkit.torch.algorithms.sac.sac_alpha", "SoftActorCritic", trainers.SoftActorCritic)
    _patch("rlkit.torch.algorithms.sac.sac", "SoftActorCritic", trainers.SoftActorCriticV)
    _patch("rlkit.torch.algorithms.td3.td3", "TD3", trainers.TD3)
    _patch("rlkit.torch.algorithms.her.td3", "TD3", trainers.HerTD3)
    _patch("rlkit.torch.algorithms.her.sac", "SAC", trainers.HerSAC)
    # her.py must bind the ORIGINAL TorchRLAlgorithm: import it before the class is replaced
    ref_her = importlib.import_module("rlkit.torch.algorithms.her.her").HER

    class HER(adv_irl.DeviceHERMixin, ref_her):
        __doc__ = ref_her.__doc__

    _patch("rlkit.torch.algorithms.her.her", "HindsightReplayBuffer", replay_buffer.DeviceEnvHindsightReplayBuffer)
    _patch("rlkit.torch.algorithms.her.her", "HER", HER)
    ref_alg = importlib.import_module("rlkit.torch.algorithms.torch_rl_algorithm").TorchRLAlgorithm
    ref_irl = importlib.import_module("rlkit.torch.algorithms.adv_irl.adv_irl").AdvIRL

    class TorchRLAlgorithm(adv_irl.DeviceTorchRLAlgorithmMixin, ref_alg):
        __doc__ = ref_alg.__doc__

    class AdvIRL(adv_irl.DeviceAdvIRLMixin, ref_irl):
        __doc__ = ref_irl.__doc__

    _patch("rlkit.torch.algorithms.torch_rl_algorithm", "TorchRLAlgorithm", TorchRLAlgorithm)
    _patch("rlkit.torch.algorithms.adv_irl.adv_irl", "AdvIRL", AdvIRL)
    _patch("rlkit.data_management.env_replay_buffer", "EnvReplayBuffer", replay_buffer.DeviceEnvReplayBuffer)
    _patch("rlkit.core.base_algorithm", "EnvReplayBuffer", replay_buffer.DeviceEnvReplayBuffer)


def uninstall():
    while _saved:
        mod, attr, value = _saved.pop()
        setattr(mod, attr, value)
