"""ilswiss_b200.dropin against the real reference package (build container only: needs /root/reference)."""
import os

import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout absent")


def test_install_replaces_the_classes_the_run_scripts_import():
    ref_shim.install()                     # stubs for gym / matplotlib / gtimer (absent in this image)
    from ilswiss_b200 import adv_irl, dropin, replay_buffer, trainers

    import rlkit.core.base_algorithm as base
    import rlkit.torch.algorithms.adv_irl.adv_irl as m_irl
    import rlkit.torch.algorithms.sac.sac_alpha as m_sac
    import rlkit.torch.algorithms.td3.td3 as m_td3
    import rlkit.torch.algorithms.torch_rl_algorithm as m_alg

    ref_sac, ref_alg, ref_irl, ref_buf = m_sac.SoftActorCritic, m_alg.TorchRLAlgorithm, m_irl.AdvIRL, base.EnvReplayBuffer
    dropin.install()
    try:
        from rlkit.torch.algorithms.sac.sac_alpha import SoftActorCritic      # what sac_alpha_exp_script.py:21 does
        from rlkit.torch.algorithms.her.td3 import TD3 as HerTD3
        assert SoftActorCritic is trainers.SoftActorCritic and m_td3.TD3 is trainers.TD3 and HerTD3 is trainers.HerTD3
        assert issubclass(m_alg.TorchRLAlgorithm, adv_irl.DeviceTorchRLAlgorithmMixin) and issubclass(m_alg.TorchRLAlgorithm, ref_alg)
        assert issubclass(m_irl.AdvIRL, adv_irl.DeviceAdvIRLMixin) and issubclass(m_irl.AdvIRL, ref_irl)
        assert m_alg.TorchRLAlgorithm._do_training is adv_irl.DeviceTorchRLAlgorithmMixin._do_training
        assert base.EnvReplayBuffer is replay_buffer.DeviceEnvReplayBuffer
        # HER: our mixin in front of the reference class, which itself still derives from the ORIGINAL TorchRLAlgorithm
        import rlkit.torch.algorithms.her.her as m_her
        assert issubclass(m_her.HER, adv_irl.DeviceHERMixin) and issubclass(m_her.HER, ref_alg)
        assert not issubclass(m_her.HER, adv_irl.DeviceTorchRLAlgorithmMixin)
        assert m_her.HindsightReplayBuffer is replay_buffer.DeviceEnvHindsightReplayBuffer
        dropin.install()                   # idempotent
    finally:
        dropin.uninstall()
    assert m_sac.SoftActorCritic is ref_sac and m_alg.TorchRLAlgorithm is ref_alg and m_irl.AdvIRL is ref_irl
    assert base.EnvReplayBuffer is ref_buf
