"""-m gpu: the reference's REAL training loop on the B200 path.  `ilswiss_b200.dropin.install()` replaces the trainer /
algorithm / replay-buffer names inside the unmodified `rlkit` package (baseline/_ref on the GPU box, /root/reference in the
build container), then `TorchRLAlgorithm.train()` and `AdvIRL.train()` -- BaseAlgorithm.start_training,
base_algorithm.py:150-291 -- run for two short epochs on a synthetic vec-env.  The progress.csv of the device run must
have exactly the columns of the pure-reference run, in the same order, and finite values."""
import math

import numpy as np
import pytest

from oracle import ref_shim

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not importable here")]


def _finite(rows):
    for r in rows:
        for v in r:
            assert v != "" and math.isfinite(float(v)), r


def test_torch_rl_algorithm_train_runs_on_the_device_path(tmp_path):
    import ref_loop
    from ilswiss_b200 import replay_buffer, trainers

    ref = ref_loop.run_sac_loop(str(tmp_path / "ref"), device=False, epochs=2, steps_per_epoch=200)
    dev = ref_loop.run_sac_loop(str(tmp_path / "dev"), device=True, epochs=2, steps_per_epoch=200)
    assert dev["header"] == ref["header"]
    assert len(dev["rows"]) == len(ref["rows"]) == 2
    _finite(dev["rows"])
    alg = dev["algorithm"]
    assert isinstance(dev["trainer"], trainers.SoftActorCritic)                      # the script's own constructor line bound it
    assert isinstance(alg.replay_buffer, replay_buffer.DeviceEnvReplayBuffer)        # BaseAlgorithm built the device ring itself
    eng = dev["trainer"].engine
    # (200 env steps / epoch) x 2 epochs, a train call every 20 env steps after 60 warm-up steps: ONE launch per train call
    n_calls = int(dev["rows"][-1][dev["header"].index("Number of train calls total")])
    assert n_calls > 0 and eng.kernel_launches == n_calls
    assert eng.get_state().n_train_steps_total == 20 * n_calls
    # bookkeeping columns driven by the loop itself agree exactly with the reference run
    for key in ("Number of env steps total", "Number of rollouts total", "Number of train calls total", "Epoch"):
        j = ref["header"].index(key)
        assert [r[j] for r in dev["rows"]] == [r[j] for r in ref["rows"]], key


@pytest.mark.parametrize("disc_kind", ["mlp_tanh", "mlp_relu"])
def test_adv_irl_train_runs_on_the_device_path(tmp_path, disc_kind):
    import warnings

    import ref_loop

    ref = ref_loop.run_advirl_loop(str(tmp_path / "ref"), device=False, epochs=2, steps_per_epoch=200, disc_kind=disc_kind)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        dev = ref_loop.run_advirl_loop(str(tmp_path / "dev"), device=True, epochs=2, steps_per_epoch=200, disc_kind=disc_kind)
    assert not any("outside the fused AdvIRL program" in str(x.message) for x in w)      # the step program, not the eager fallback
    assert dev["header"] == ref["header"]
    _finite(dev["rows"])
    for key in ("Number of env steps total", "Number of train calls total", "Epoch"):
        j = ref["header"].index(key)
        assert [r[j] for r in dev["rows"]] == [r[j] for r in ref["rows"]], key
    acc = np.array([float(r[dev["header"].index("Disc Acc")]) for r in dev["rows"]])
    assert ((acc >= 0) & (acc <= 1)).all()


@pytest.mark.parametrize("disc_kind", ["mlp_bn_relu", "resnet"])
def test_adv_irl_with_discriminators_outside_the_fused_program(tmp_path, disc_kind):
    """simple_disc_models.py:29-38,51-93 (BatchNorm / ReLU MLPDisc, ResNetAIRLDisc; unused by the shipped yamls): the mixin
    falls back to the reference's own reward / policy training methods with the discriminator as an eager torch module on
    the device, device-resident batches and fused SAC steps.  Same log columns as the pure-reference run, and the first
    logged discriminator cross-entropy (first update of epoch 0: same buffers, same index streams, same initial
    discriminator) agrees with the CPU reference run."""
    import warnings

    import ref_loop

    ref = ref_loop.run_advirl_loop(str(tmp_path / "ref"), device=False, epochs=2, steps_per_epoch=200, disc_kind=disc_kind)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        dev = ref_loop.run_advirl_loop(str(tmp_path / "dev"), device=True, epochs=2, steps_per_epoch=200, disc_kind=disc_kind)
    assert any("outside the fused AdvIRL program" in str(x.message) for x in w)
    assert dev["header"] == ref["header"]
    _finite(dev["rows"])
    j = ref["header"].index("Disc CE Loss")
    a, b = float(ref["rows"][0][j]), float(dev["rows"][0][j])
    assert abs(a - b) <= 1e-3 * max(abs(a), 1e-3), (a, b)
    # policy updates still ran as fused SAC steps: one engine launch per policy update
    eng = dev["trainer"].engine
    assert eng.kernel_launches > 0 and eng.get_state().n_train_steps_total == eng.kernel_launches
