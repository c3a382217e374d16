"""The reference's own training loop (BaseAlgorithm.start_training) runs here with the reference's classes on a synthetic
vec-env: pins the loop harness of tests/ref_loop.py that the GPU test uses with ilswiss_b200.dropin installed."""
import pytest

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference not importable here")


def test_reference_sac_loop_writes_progress_csv(tmp_path):
    import ref_loop

    out = ref_loop.run_sac_loop(str(tmp_path), device=False, epochs=2, steps_per_epoch=120)
    h = out["header"]
    assert h[:5] == ["Reward Scale", "QF1 Loss", "QF2 Loss", "Alpha Loss", "Policy Loss"]      # sac_alpha.py:186-233
    assert "AverageReturn" in h and "Number of train calls total" in h and h[-1] == "Epoch"
    assert len(out["rows"]) == 2 and all(len(r) == len(h) for r in out["rows"])


def test_reference_advirl_loop_writes_progress_csv(tmp_path):
    import ref_loop

    out = ref_loop.run_advirl_loop(str(tmp_path), device=False, epochs=2, steps_per_epoch=120)
    h = out["header"]
    assert h[:3] == ["Disc CE Loss", "Disc Acc", "Grad Pen"] and "Disc Rew Mean" in h and "QF1 Loss" in h
    assert len(out["rows"]) == 2
