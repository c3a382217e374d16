"""Random parity cases (tools/fuzz_hostsim.py: algorithm, ragged dims and widths, AdvIRL mode / state_only / expert mix /
reward clips / discriminator activation / penalty on and off) through the host simulator of the PRODUCT's step programs
against the oracle -- a fixed handful of seeds here; the campaign of round 2 was 150 seeds without a failure."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import fuzz_hostsim  # noqa: E402
from helpers import load_hostsim  # noqa: E402

SEEDS = [1003, 1004, 1010, 1011, 1016, 1017, 1021, 1050, 1061, 1098]


@pytest.fixture(scope="module")
def lib():
    return load_hostsim()


def test_seeds_cover_every_algorithm_and_both_discriminator_activations():
    cases = [fuzz_hostsim.random_case(s) for s in SEEDS]
    assert {c["algo"] for c in cases} == {"sac_alpha", "sac_v", "td3", "adv_irl"}
    assert {c.get("disc_act") for c in cases if c["algo"] == "adv_irl"} == {"tanh", "relu"}


@pytest.mark.parametrize("seed", SEEDS)
def test_random_case_matches_oracle(lib, seed):
    torch.set_num_threads(1)
    fuzz_hostsim.check_case(lib, fuzz_hostsim.random_case(seed))


@pytest.mark.parametrize("seed", [1005, 1012, 1034, 1044])      # adv_irl relu, td3, adv_irl tanh + expert mix, sac_v
def test_random_case_program_variants_agree_bit_for_bit(lib, seed):
    fuzz_hostsim.check_variants(lib, fuzz_hostsim.random_case(seed))


@pytest.mark.parametrize("seed", [3001, 3005, 3004])      # HER-SAC, TD3 period 1..3, SAC-V at batch 512..640
def test_random_tcgen05_program_variant_case_matches_oracle(lib, seed, monkeypatch):
    torch.set_num_threads(1)
    monkeypatch.setenv("ILSW_HOSTSIM_TC5", "1")
    case = fuzz_hostsim.random_tc5_case(seed)
    import ctypes as C

    from helpers import HostSimRun

    probe, buf = HostSimRun(lib, case, precision=3), C.create_string_buffer(1 << 15)
    lib.hs_describe(probe.h, buf, 1 << 15)
    probe.close()
    assert "tcgen05" in buf.value.decode()
    fuzz_hostsim.check_case(lib, case, precision=3, frac=2.5e-3, lr=case.get("td3", {}).get("policy_lr", 3e-4))


@pytest.mark.parametrize("seed", [4100, 4101, 4102, 4103])
def test_random_relabel_at_sample_case_matches_oracle(lib, seed):
    torch.set_num_threads(1)
    fuzz_hostsim.check_her_relabel_case(lib, fuzz_hostsim.random_her_relabel_case(seed))


@pytest.mark.parametrize("seed", [9001, 9002, 9003])      # relu 2/1, tanh 3/2, state_only 3/2 updates per loop iteration
def test_random_loop_count_case_matches_oracle(lib, seed):
    torch.set_num_threads(1)
    fuzz_hostsim.check_loop_case(lib, fuzz_hostsim.random_loop_case(seed))
