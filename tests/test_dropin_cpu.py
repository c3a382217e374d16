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


def test_discriminator_routing_reads_the_module_structure_not_the_parameter_list():
    """simple_disc_models.py:8-93: MLPDisc(use_bn=False) with tanh or relu blocks goes to the step program with that
    activation; BatchNorm blocks, other depths and ResNetAIRLDisc do not -- including the ResNet instance whose six
    parameter tensors have exactly the shapes of the two-block MLP's."""
    ref_shim.install()
    from rlkit.torch.algorithms.adv_irl.disc_models.simple_disc_models import MLPDisc, ResNetAIRLDisc

    from ilswiss_b200 import modules
    from ilswiss_b200.adv_irl import disc_hidden_activation

    assert disc_hidden_activation(MLPDisc(14, num_layer_blocks=2, hid_dim=32, hid_act="tanh", use_bn=False)) == "tanh"
    assert disc_hidden_activation(MLPDisc(14, num_layer_blocks=2, hid_dim=32, hid_act="relu", use_bn=False)) == "relu"
    assert disc_hidden_activation(modules.MLPDisc(14, 32)) == "tanh"
    assert disc_hidden_activation(modules.MLPDisc(14, 32, hid_act="relu")) == "relu"
    look_alike = ResNetAIRLDisc(14, num_layer_blocks=2, hid_dim=32, hid_act="tanh", use_bn=False)
    assert [tuple(p.shape) for p in look_alike.parameters()] == [tuple(p.shape) for p in modules.MLPDisc(14, 32).parameters()]
    for other in (MLPDisc(14, num_layer_blocks=2, hid_dim=32, hid_act="relu", use_bn=True),       # the class defaults
                  MLPDisc(14, num_layer_blocks=3, hid_dim=32, hid_act="tanh", use_bn=False),
                  look_alike,
                  ResNetAIRLDisc(14, num_layer_blocks=3, hid_dim=32, hid_act="relu", use_bn=True)):
        with pytest.raises(NotImplementedError):
            disc_hidden_activation(other)


def test_networks_with_other_activations_are_refused_not_retrained_as_relu():
    """networks.py:23-47: hidden_activation / output_activation are constructor arguments that leave the parameter list
    unchanged; the trainers check them (trainers.check_activations) before adopting a module."""
    ref_shim.install()
    import torch
    from rlkit.torch.common.networks import FlattenMlp
    from rlkit.torch.common.policies import MlpGaussianNoisePolicy

    from ilswiss_b200 import modules
    from ilswiss_b200.trainers import check_activations, module_dims

    ok = FlattenMlp(hidden_sizes=[32, 32], input_size=14, output_size=1)
    assert module_dims(ok) == (14, 32, 1, False)
    check_activations(ok, "identity")
    check_activations(modules.FlattenMlp([32, 32], 1, 14), "identity")
    check_activations(modules.DeterministicNoisePolicy([32, 32], 11, 3), "tanh")
    check_activations(MlpGaussianNoisePolicy(hidden_sizes=[32, 32], obs_dim=11, action_dim=3, output_activation=torch.tanh), "tanh")
    with pytest.raises(NotImplementedError):
        module_dims(FlattenMlp(hidden_sizes=[32, 32], input_size=14, output_size=1, hidden_activation=torch.tanh))
    with pytest.raises(NotImplementedError):
        check_activations(FlattenMlp(hidden_sizes=[32, 32], input_size=14, output_size=1, output_activation=torch.tanh), "identity")
    with pytest.raises(NotImplementedError):      # the class default (identity) is not what td3_exp_script.py:75 builds
        check_activations(MlpGaussianNoisePolicy(hidden_sizes=[32, 32], obs_dim=11, action_dim=3), "tanh")
