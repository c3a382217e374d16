"""CPU validation of the PRODUCT's step programs (ilsw_ops.cuh + ilsw_program.h) through the
test-only host simulator: identical injected indices/eps as the oracle, per-step losses and
final parameters must agree to fp32 round-off.  No GPU, no CUDA runtime."""
import numpy as np
import pytest
import torch

from helpers import CFG, G, HostSimRun, STAT_TO_SLOT, assert_params_close, case_injection, load_hostsim, run_loop_case, her_relabel_setup, her_oracle_rows, her_desc, np_ptr

CASES = [n for n, c in CFG.CASES.items() if c["algo"] in ("sac_alpha", "sac_v", "td3", "adv_irl")]


@pytest.fixture(scope="module")
def lib():
    return load_hostsim()


@pytest.mark.parametrize("name", CASES)
def test_hostsim_matches_oracle(lib, name):
    torch.set_num_threads(1)
    case = CFG.CASES[name]
    rows, final, _ = G.run_oracle(case)
    run = HostSimRun(lib, case)
    inj = case_injection(case)
    L = run.train(case["steps"], inj)
    for t, row in enumerate(rows):
        for k, ref in row.items():
            if k not in STAT_TO_SLOT or ref is None:
                continue
            got = L[t, STAT_TO_SLOT[k]]
            if np.isnan(got):
                continue
            # bar: 1e-4 relative on losses (BASELINE.json north_star); round-off only here
            tol = 1e-4 * max(abs(ref), 1e-3) if k != "Policy Loss" else 1e-4 * max(abs(ref), 1.0)
            assert abs(got - ref) <= tol, (name, t, k, got, ref)
    for k in final:
        if k == "log_alpha":
            assert abs(lib.hs_log_alpha(run.h) - final[k][0]) < 1e-6
            continue
        assert_params_close(run.arenas[k], final[k], case["steps"], msg="%s/%s" % (name, k))
    run.close()


@pytest.mark.parametrize("name", ["sac_hopper_b512_fixed_alpha", "td3_humanoid"])
def test_hostsim_tcgen05_program_variant(lib, name, monkeypatch):
    """The program variant of the tcgen05/TMA engine (batch >= 512): 128-row tile counts and, where a first layer has
    in_dim % 4 != 0, the 16-byte aligned W0 copies that every writer of W0 (fused Adam epilogues, Polyak, the
    first-step refresh) keeps in sync.  The host executes GEMMs element-wise, so this pins the WIRING of that variant
    against the oracle; the tile itself is checked on the GPU (tools/tc5_test.cu, tests/test_gpu_engine.py)."""
    import ctypes as C
    torch.set_num_threads(1)
    monkeypatch.setenv("ILSW_HOSTSIM_TC5", "1")
    case = CFG.CASES[name]
    rows, final, _ = G.run_oracle(case)
    run = HostSimRun(lib, case, precision=3)
    buf = C.create_string_buffer(1 << 15)
    lib.hs_describe(run.h, buf, 1 << 15)
    text = buf.value.decode()
    assert "tcgen05" in text
    assert ("W0COPY" in text) == ((case["obs_dim"] + case["act_dim"]) % 4 != 0 or case["obs_dim"] % 4 != 0)
    L = run.train(case["steps"], case_injection(case))
    for t, row in enumerate(rows):
        for k, ref in row.items():
            if k not in STAT_TO_SLOT or ref is None or np.isnan(L[t, STAT_TO_SLOT[k]]):
                continue
            tol = 1e-4 * max(abs(ref), 1e-3) if k != "Policy Loss" else 1e-4 * max(abs(ref), 1.0)
            assert abs(L[t, STAT_TO_SLOT[k]] - ref) <= tol, (name, t, k, L[t, STAT_TO_SLOT[k]], ref)
    for k in final:
        if k != "log_alpha":
            assert_params_close(run.arenas[k], final[k], case["steps"], msg="%s/%s" % (name, k))
    run.close()


def test_hostsim_split_launches_equal_one_launch(lib):
    """State carried across launches (Adam step counts, alpha) == one long launch."""
    case = CFG.CASES["sac_hopper"]
    inj = case_injection(case)
    a = HostSimRun(lib, case)
    La = a.train(case["steps"], inj)
    b = HostSimRun(lib, case)
    Lb = np.concatenate([b.train(2, inj, t_offset=0), b.train(case["steps"] - 2, inj, t_offset=2)])
    np.testing.assert_array_equal(La[:, :5], Lb[:, :5])
    np.testing.assert_array_equal(a.arenas["policy"], b.arenas["policy"])
    a.close(); b.close()


def test_td3_delay_across_launches(lib):
    case = CFG.CASES["td3_hopper"]
    inj = case_injection(case)
    a = HostSimRun(lib, case)
    a.train(case["steps"], inj)
    b = HostSimRun(lib, case)
    for t in range(case["steps"]):
        b.train(1, inj, t_offset=t)
    np.testing.assert_array_equal(a.arenas["policy"], b.arenas["policy"])
    np.testing.assert_array_equal(a.arenas["target_qf1"], b.arenas["target_qf1"])
    a.close(); b.close()


def test_td3_stats_only_policy_loss(lib):
    """td3.py:126-136: when the logged step is not a policy step the reference evaluates -mean(Q1(obs, pi(obs))) for the
    log only.  The step program runs its policy / Q1 forward phases on the statistics step of a launch as well."""
    torch.set_num_threads(1)
    case = CFG.CASES["td3_hopper"]
    rows, _, _ = G.run_oracle(case)
    inj = case_injection(case)
    for t in (1, 2, 3):                     # odd steps have no policy update (period 2, starting at step 0)
        run = HostSimRun(lib, case)
        L = run.train(case["steps"], inj, stats_step=t)
        ref = rows[t]["Stats Policy Loss"]
        assert abs(L[t, STAT_TO_SLOT["Policy Loss"]] - ref) <= 1e-4 * max(abs(ref), 1.0), (t, L[t, STAT_TO_SLOT["Policy Loss"]], ref)
        for u in range(case["steps"]):      # every other non-policy step logs no policy loss
            if u != t and u % 2 == 1:
                assert np.isnan(L[u, STAT_TO_SLOT["Policy Loss"]])
        run.close()


def test_program_shape(lib):
    import ctypes as C
    run = HostSimRun(lib, CFG.CASES["gail_walker"])
    buf = C.create_string_buffer(16384)
    lib.hs_describe(run.h, buf, 16384)
    text = buf.value.decode()
    assert lib.hs_num_phases(run.h) == text.count("phase ")
    assert "replica-exchange" in text and "ADAM" in text
    run.close()


@pytest.mark.parametrize("name", ["sac_hopper", "td3_hopper", "gail_walker", "gail_hopper_relu"])
def test_fast_row_jobs_equal_generic_row_kernels(lib, name):
    """Two independent implementations of every row kernel (latency-optimised job form vs the
    generic per-row form) must agree bit for bit on the host."""
    import ctypes as C

    lib.hs_set_generic_rows.argtypes = [C.c_void_p, C.c_int]
    case = CFG.CASES[name]
    inj = case_injection(case)
    a = HostSimRun(lib, case)
    b = HostSimRun(lib, case)
    lib.hs_set_generic_rows(b.h, 1)
    La, Lb = a.train(case["steps"], inj), b.train(case["steps"], inj)
    np.testing.assert_array_equal(np.nan_to_num(La), np.nan_to_num(Lb))
    for k in a.arenas:
        np.testing.assert_array_equal(a.arenas[k], b.arenas[k], err_msg=k)
    a.close(); b.close()


@pytest.mark.parametrize("name", list(CFG.LOOP_CASES.keys()))
def test_hostsim_split_disc_policy_launches_match_oracle(lib, name):
    """adv_irl.py:126-131 with n_disc / n_policy != 1: alternating disc-only / policy-only launches of the SAME program
    (phase conditions COND_DISC_PART / COND_POLICY_PART) reproduce the oracle's nested loops."""
    torch.set_num_threads(1)
    case = CFG.LOOP_CASES[name]
    rows, final, _ = G.run_oracle(case)
    run = HostSimRun(lib, case)
    got_rows = run_loop_case(run, case, lambda m: lib.hs_set_update_mode(run.h, m))
    for t, (row, got) in enumerate(zip(rows, got_rows)):
        for k, ref in row.items():
            if k not in STAT_TO_SLOT or ref is None:
                continue
            tol = 1e-4 * max(abs(ref), 1e-3) if k != "Policy Loss" else 1e-4 * max(abs(ref), 1.0)
            assert abs(got[k] - ref) <= tol, (name, t, k, got[k], ref)
    n_updates = case["steps"] * max(case["n_disc"], case["n_policy"])
    for k in final:
        if k == "log_alpha":
            assert abs(lib.hs_log_alpha(run.h) - final[k][0]) < 1e-6
            continue
        assert_params_close(run.arenas[k], final[k], n_updates, msg="%s/%s" % (name, k))
    run.close()


def test_hostsim_fused_iteration_equals_split_launches(lib):
    """One fused disc+policy engine step == one disc-only launch followed by one policy-only launch, bit for bit."""
    case = CFG.CASES["gail_hopper"]
    inj = case_injection(case)
    a = HostSimRun(lib, case)
    La = a.train(case["steps"], inj)
    b = HostSimRun(lib, case)
    for t in range(case["steps"]):
        lib.hs_set_update_mode(b.h, 1)
        Ld = b.train(1, inj, t_offset=t)
        lib.hs_set_update_mode(b.h, 2)
        Lp = b.train(1, inj, t_offset=t)
        assert Ld[0, STAT_TO_SLOT["Disc CE Loss"]] == La[t, STAT_TO_SLOT["Disc CE Loss"]]
        assert Lp[0, STAT_TO_SLOT["QF1 Loss"]] == La[t, STAT_TO_SLOT["QF1 Loss"]]
    for k in a.arenas:
        np.testing.assert_array_equal(a.arenas[k], b.arenas[k], err_msg=k)
    a.close(); b.close()


@pytest.mark.parametrize("name", list(CFG.HER_RELABEL_CASES.keys()))
def test_hostsim_her_relabel_at_sample_matches_oracle(lib, name):
    """relabel_replay_buffer.py:63-131 inside the step program's gather: goal substitution from the future step's next
    achieved goal for the first int(her_ratio B) rows, sparse reward recomputed for every row -- against the oracle
    hindsight buffer (pinned to the reference's) feeding the HER-TD3 oracle."""
    import ctypes as C

    torch.set_num_threads(1)
    case = CFG.HER_RELABEL_CASES[name]
    setup = her_relabel_setup(case)
    rows, final = her_oracle_rows(case, setup)
    run = HostSimRun(lib, case)
    run.ring = np.ascontiguousarray(setup["ring"][:setup["ring_size"]])
    desc = her_desc(setup, lambda k: np_ptr(setup[k]))
    lib.hs_set_her(run.h, C.byref(desc))
    L = run.train(case["steps"], dict(idx=setup["idx"], eps_next=setup["eps_next"], eps_cur=setup["eps_cur"]))
    for t, row in enumerate(rows):
        for k, ref in row.items():
            got = L[t, STAT_TO_SLOT[k]]
            tol = 1e-4 * max(abs(ref), 1e-3) if k != "Policy Loss" else 1e-4 * max(abs(ref), 1.0)
            assert abs(got - ref) <= tol, (name, t, k, got, ref)
    for k in final:
        assert_params_close(run.arenas[k], final[k], case["steps"], lr=6e-4, msg="%s/%s" % (name, k))
    # speed mode: in-kernel Philox trajectory / step / future-step draws run and give finite, sparse-reward targets
    lib.hs_set_her(run.h, C.byref(desc))
    rc = run.lib.hs_train(run.h, np_ptr(run.ring), run.ring.shape[1], run.ring.shape[0], None, 0, 0, 2, None, None, 7, -1)
    assert rc == 0
    L2 = np.ctypeslib.as_array(lib.hs_losses(run.h), shape=(2, 16)).copy()
    assert np.isfinite(L2[:, :2]).all()
    run.close()


@pytest.mark.parametrize("name", ["sac_hopper", "td3_hopper", "sacv_hopper"])
@pytest.mark.parametrize("world", [2, 8])
def test_hostsim_replica_program_with_identical_replicas_equals_single_replica(lib, name, world):
    """SURVEY.md 8e on the CPU: the N-replica step program (un-fused policy weight gradients -> exchange -> flat Adam of the
    averaged gradient, COND_WORLD_N phases) run with R identical replicas is bit-identical to the single-replica program
    (Adam fused into the weight-gradient epilogues) at R = 2, where (g + g) / 2 == g exactly; at R = 8 the rank-order sum
    g + g + ... rounds (3g is not exact), so the two programs agree to round-off only."""
    case = CFG.CASES[name]
    inj = case_injection(case)
    a = HostSimRun(lib, case)
    La = a.train(case["steps"], inj)
    b = HostSimRun(lib, case)
    lib.hs_set_world(b.h, world)
    Lb = b.train(case["steps"], inj)
    if world == 2:
        np.testing.assert_array_equal(np.nan_to_num(La), np.nan_to_num(Lb))
        for k in a.arenas:
            np.testing.assert_array_equal(a.arenas[k], b.arenas[k], err_msg=k)
        for k in a.moms:
            np.testing.assert_array_equal(a.moms[k][0], b.moms[k][0], err_msg=k + ".m")
    else:
        np.testing.assert_allclose(np.nan_to_num(La), np.nan_to_num(Lb), rtol=2e-5, atol=1e-6)
        for k in a.arenas:
            assert_params_close(b.arenas[k], a.arenas[k], case["steps"], msg=k)
    a.close(); b.close()
