"""Pins oracle/restate.py to the golden transcripts produced by executing the unmodified
reference (oracle/make_golden.py).  CPU only; runs anywhere (no /root/reference needed)."""
import os

import numpy as np
import pytest
import torch

from oracle import configs as C
from oracle import make_golden as G

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


ALL_CASES = dict(C.CASES, **C.LOOP_CASES)


@pytest.mark.parametrize("name", list(ALL_CASES.keys()))
def test_oracle_reproduces_reference_transcript(name):
    torch.set_num_threads(1)
    case = ALL_CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    keys = [str(k) for k in gold["stat_keys"]]
    rows, final, digests = G.run_oracle(case)
    assert len(rows) == gold["stats"].shape[0] == case["steps"]
    for t, row in enumerate(rows):
        for j, k in enumerate(keys):
            ref = gold["stats"][t, j]
            if np.isnan(ref) or row.get(k) is None:
                continue
            # tolerance: fp32 round-off of a different summation order only
            assert abs(row[k] - ref) <= 2e-5 * max(abs(ref), 1e-12) + 1e-9, (name, t, k, row[k], ref)
    for k, dig in digests.items():
        np.testing.assert_allclose(dig, gold["digest_" + k], rtol=1e-5, atol=2e-6, err_msg=k)
    for k, v in final.items():
        sample = v[:: max(1, v.size // 256)][:256]
        np.testing.assert_allclose(sample, gold["sample_" + k], rtol=0, atol=2e-6, err_msg=k)


def test_replay_oracle_matches_numpy_semantics():
    """R1-R3: ring wrap-around, size saturation, uniform-with-replacement index stream."""
    from oracle.restate import ReplayOracle

    buf = ReplayOracle(8, 3, 2, random_seed=5)
    for i in range(11):
        buf.add_sample(np.full(3, i), np.full(2, -i), float(i), i % 4 == 3, np.full(3, i + 1))
    assert buf._size == 8 and buf._top == 3
    assert buf._observations[0, 0] == 8 and buf._observations[3, 0] == 3
    ref = np.random.RandomState(5).randint(0, 8, 16)
    b = buf.random_batch(16)
    assert set(b.keys()) == set(ReplayOracle.ALL_KEYS)
    np.testing.assert_array_equal(b["observations"][:, 0], buf._observations[ref, 0])
    assert b["terminals"].dtype == np.uint8 and b["rewards"].shape == (16, 1)


@pytest.mark.skipif(not os.path.isdir("/root/reference/rlkit"), reason="reference checkout absent")
def test_goldens_regenerate_from_reference():
    """In the build container: re-run the real reference for one case and compare with the
    committed fixture (guards against a stale fixture)."""
    torch.set_num_threads(1)
    name = "sac_hopper"
    rows, final, _ = G.run_reference(C.CASES[name])
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    keys = [str(k) for k in gold["stat_keys"]]
    for t, row in enumerate(rows):
        for j, k in enumerate(keys):
            if k in row:
                assert row[k] == pytest.approx(gold["stats"][t, j], rel=1e-6, abs=1e-9)


@pytest.mark.skipif(not os.path.isdir("/root/reference/rlkit"), reason="reference checkout absent")
def test_hindsight_oracle_matches_reference_buffer():
    """Pins HindsightOracle to the executed rlkit HindsightReplayBuffer: same index draws, same relabelled batches
    (bit-exact)."""
    from oracle import ref_shim
    from oracle.restate import HindsightOracle, synth_goal_episodes

    ref_shim.install()
    from rlkit.data_management.relabel_replay_buffer import HindsightReplayBuffer

    O0, G, A, N = 10, 3, 4, 400
    env = ref_shim.FakeGoalEnv(O0, G, A)
    ref = HindsightReplayBuffer(N, env, random_seed=5, relabel_type="future", her_ratio=0.8)
    ora = HindsightOracle(N, O0, G, A, random_seed=5, relabel_type="future", her_ratio=0.8)
    rs = np.random.RandomState(0)
    for ep in synth_goal_episodes(rs, 7, 50, O0, G, A):
        for (o, a, r, d, no) in ep:
            ref.add_sample(o, a, r, d, no)
            ora.add_sample(o, a, r, d, no)
        ref.terminate_episode()
        ora.terminate_episode()
    assert ref._traj_endpoints == ora._traj_endpoints and ref._top == ora._top and ref._size == ora._size
    for trial in range(3):
        np.random.seed(100 + trial)
        want = ref.random_batch(64)
        np.random.seed(100 + trial)
        got = ora.random_batch(64)
        assert set(got.keys()) == set(want.keys())
        for k in want:
            np.testing.assert_array_equal(np.asarray(got[k]), np.asarray(want[k]), err_msg=k)
        assert (want["rewards"] == 0).any() and (want["rewards"] == -1).any()          # both reward values occur


@pytest.mark.parametrize("seed", [1000, 1001, 1004, 1010, 1017])      # AIRL relu + expert mix, SAC-V, TD3, GAIL tanh, relu state_only no penalty
def test_oracle_matches_the_executed_reference_on_random_cases(seed):
    """Beyond the committed goldens: random shapes / options (tools/fuzz_hostsim.random_case) run through the UNMODIFIED
    reference classes and through the oracle in this container -- 2e-5 relative on statistics, 2e-6 absolute on parameters."""
    import sys

    from oracle import ref_shim

    if not ref_shim.reference_available():
        pytest.skip("reference checkout absent")
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_hostsim

    torch.set_num_threads(1)
    fuzz_hostsim.check_oracle_against_reference(fuzz_hostsim.random_case(seed))
