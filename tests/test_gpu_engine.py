"""-m gpu parity tests proper: the CUDA path, called through the C ABI, against the oracle
(oracle/restate.py) on identical injected indices / eps, and against the committed goldens
generated from the executed reference."""
import os

import numpy as np
import pytest
import torch

from helpers import CFG, G, DeviceRun, STAT_TO_SLOT, assert_params_close, case_injection, run_loop_case

pytestmark = pytest.mark.gpu
import json

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CASES = [n for n, c in CFG.CASES.items() if c["algo"] in ("sac_alpha", "sac_v", "td3", "adv_irl")]
# measured on a B200 (tools/measure_param_frac.py): per case, the worst net's fraction of parameter elements that end more
# than 1e-5 away from the oracle in the tensor-core GEMM modes.  The tests allow TWICE the measured fraction (never less
# than the 5e-4 of the exact mode -- Adam's sign-sensitive elements, see assert_params_close).
PARAM_FRAC = json.load(open(os.path.join(GOLDEN, "param_frac.json")))["cases"]


def param_frac_bar(name, precision):
    if precision == 0:
        return 5e-4
    key = "p%d" % precision
    # a case added after the last measurement run takes the largest fraction measured on any case until it has its own entry
    measured = PARAM_FRAC[name][key]["frac_beyond_1e-5"] if name in PARAM_FRAC else max(v[key]["frac_beyond_1e-5"] for v in PARAM_FRAC.values())
    return max(2.0 * measured, 5e-4)


def loss_tol(k, ref, precision=0, n_rows=512):
    if precision == 1:
        # single-pass TF32 (10-bit mantissa operands): the 1e-4 bar is a bar on LOSSES (means over the
        # batch, where rounding noise averages out).  Element-level statistics carry the per-element
        # TF32 error (~1e-3 relative), and Disc Acc is a count of sign decisions: allow 2 flips.
        if k == "Disc Acc":
            return 2.0 / n_rows + 1e-9
        if k.endswith(" Max") or k.endswith(" Min"):
            return 2e-3 * max(abs(ref), 1.0)
        return 5 * _loss_tol(k, ref)      # opt-in fast mode: 5e-4 (NOT the default; default mode 3 meets 1e-4)
    return _loss_tol(k, ref)


def _loss_tol(k, ref):
    # bar (BASELINE.json north_star): 1e-4 relative on losses; the policy loss is a cancellation
    # of O(1) terms, so its floor is 1e-4 absolute
    if k == "Policy Loss" or k.endswith(" Mean") or k.endswith(" Std"):
        return 1e-4 * max(abs(ref), 1.0)     # means of O(1)-spread vectors: 1e-4 of their scale
    return 1e-4 * max(abs(ref), 1e-2)


# GEMM precision modes of the engine: 0 = fp32 SIMT (exact gate), 1 = TF32 tensor cores,
# 3 = 3xTF32 tensor cores (fp32-level).  Loss bar is 1e-4 relative in EVERY mode.
@pytest.mark.parametrize("precision", [0, 1, 3])
@pytest.mark.parametrize("name", CASES)
def test_cuda_step_matches_oracle(name, precision):
    torch.set_num_threads(1)
    case = CFG.CASES[name]
    rows, final, _ = G.run_oracle(case)
    run = DeviceRun(case, precision=precision)
    L = run.train(case["steps"], case_injection(case))
    for t, row in enumerate(rows):
        for k, ref in row.items():
            if k not in STAT_TO_SLOT or ref is None:
                continue
            got = float(L[t, STAT_TO_SLOT[k]])
            if np.isnan(got):
                continue
            assert abs(got - ref) <= loss_tol(k, ref, precision, 2 * case["batch"]), (name, t, k, got, ref)
    for k in final:
        if k == "log_alpha":
            assert abs(run.eng.get_state().log_alpha - final[k][0]) < 1e-6
        else:
            # tensor-core modes perturb gradients (TF32: 1e-3 relative, 3xTF32: ~5e-7 + the tensor
            # core's truncating fp32 accumulation): more elements fall into Adam's sign-sensitive
            # regime (see assert_params_close); the 2*lr*steps bound holds in every mode
            assert_params_close(run.arena(k), final[k], case["steps"], msg="%s/%s" % (name, k), frac=param_frac_bar(name, precision))


@pytest.mark.parametrize("precision", [0, 3])      # 3 = the production default (bench.py)
@pytest.mark.parametrize("name", CASES)
def test_cuda_step_matches_reference_golden(name, precision):
    """Directly against the transcript of the executed reference (no oracle in between)."""
    case = CFG.CASES[name]
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    keys = [str(k) for k in gold["stat_keys"]]
    run = DeviceRun(case, precision=precision)
    L = run.train(case["steps"], case_injection(case))
    for t in range(case["steps"]):
        for j, k in enumerate(keys):
            ref = gold["stats"][t, j]
            if k not in STAT_TO_SLOT or np.isnan(ref):
                continue
            got = float(L[t, STAT_TO_SLOT[k]])
            if np.isnan(got):
                continue
            assert abs(got - ref) <= loss_tol(k, ref), (name, t, k, got, ref)
    for k in gold.files:
        if not k.startswith("sample_") or k == "sample_log_alpha":
            continue
        v = run.arena(k[len("sample_"):])
        sample = v[:: max(1, v.size // 256)][:256]
        assert np.max(np.abs(sample - gold[k])) <= 2 * 3e-4 * case["steps"] + 1e-6
        assert (np.abs(sample - gold[k]) > 1e-5).sum() <= max(1, int(np.ceil(param_frac_bar(name, precision) * sample.size)))


def test_fused_first_layer_program_matches_oracle(monkeypatch):
    """ILSW_FUSE_L0=1 (first layers produced inside the second layer's tiles; off by default since it measured slower)
    stays a supported program: same parity bar."""
    monkeypatch.setenv("ILSW_FUSE_L0", "1")
    torch.set_num_threads(1)
    for name in ("sac_hopper", "gail_walker"):
        case = CFG.CASES[name]
        rows, final, _ = G.run_oracle(case)
        run = DeviceRun(case, precision=3)
        assert "fused-L0" in run.eng.describe()
        L = run.train(case["steps"], case_injection(case))
        for t, row in enumerate(rows):
            for k in ("QF1 Loss", "QF2 Loss", "Policy Loss"):
                assert abs(float(L[t, STAT_TO_SLOT[k]]) - row[k]) <= loss_tol(k, row[k], 3), (name, t, k)
        assert_params_close(run.arena("policy"), final["policy"], case["steps"], msg=name, frac=param_frac_bar(name, 3))


def test_tma_panels_on_and_off_agree(monkeypatch):
    """The TMA panel path of the mma.sync tile (GemmOp::tma) and the cp.async path compute the same tile: bit-identical
    losses and parameters (same fragment values, same MMA order)."""
    case = CFG.CASES["sac_hopper"]
    inj = case_injection(case)
    a = DeviceRun(case, precision=3)
    La = a.train(case["steps"], inj)
    monkeypatch.setenv("ILSW_TMA_PANELS", "0")
    b = DeviceRun(case, precision=3)
    Lb = b.train(case["steps"], inj)
    np.testing.assert_array_equal(La[:, :5], Lb[:, :5])
    np.testing.assert_array_equal(a.arena("policy"), b.arena("policy"))


def test_split_launches_equal_one_launch():
    case = CFG.CASES["sac_hopper"]
    inj = case_injection(case)
    a = DeviceRun(case)
    La = a.train(case["steps"], inj)
    b = DeviceRun(case)
    Lb = np.concatenate([b.train(2, inj), b.train(case["steps"] - 2, inj, t_offset=2)])
    np.testing.assert_array_equal(La[:, :5], Lb[:, :5])   # bit-identical: deterministic reductions
    np.testing.assert_array_equal(a.arena("policy"), b.arena("policy"))


def test_philox_mode_runs_and_is_reproducible():
    case = CFG.CASES["sac_hopper"]
    a, b = DeviceRun(case), DeviceRun(case)
    La, Lb = a.train_philox(8, seed=123), b.train_philox(8, seed=123)
    np.testing.assert_array_equal(La, Lb)
    assert np.isfinite(La[:, :5]).all()
    c = DeviceRun(case)
    Lc = c.train_philox(8, seed=124)
    assert not np.array_equal(La[:, 0], Lc[:, 0])
    assert a.eng.kernel_launches == 1      # 8 gradient steps, ONE kernel launch


def test_direct_batch_equals_ring_batch():
    """Trainer.train_step(batch) path: same numbers as sampling the same rows from the ring."""
    case = dict(CFG.CASES["sac_hopper"], steps=1)
    inj = case_injection(case)
    a = DeviceRun(case)
    La = a.train(1, inj)
    b = DeviceRun(case)
    idx = torch.from_numpy(inj["idx"][0]).cuda()
    hot, _ = b.ring.gather(idx)
    O, A = case["obs_dim"], case["act_dim"]
    batch = dict(obs=hot[:, :O].contiguous(), act=hot[:, O:O + A].contiguous(), rew=hot[:, O + A].contiguous(),
                 term=hot[:, O + A + 1].contiguous(), next_obs=hot[:, O + A + 2:2 * O + A + 2].contiguous())
    dev = {k: torch.from_numpy(np.ascontiguousarray(v[:1])).cuda() for k, v in inj.items()}
    b.eng.train(None, 1, inject=dev, batch=batch)
    np.testing.assert_array_equal(La[:, :5], b.eng.losses(1)[:, :5])
    np.testing.assert_array_equal(a.arena("qf1"), b.arena("qf1"))


@pytest.mark.parametrize("precision", [0, 3])
@pytest.mark.parametrize("name", list(CFG.LOOP_CASES.keys()))
def test_cuda_split_disc_policy_launches_match_oracle_and_golden(name, precision):
    """adv_irl.py:126-131 with n_disc / n_policy != 1 (gail_humanoid.yaml): alternating disc-only / policy-only launches
    (ilsw_trainer_set_update_mode) vs the oracle's nested loops AND vs the executed reference's transcript."""
    torch.set_num_threads(1)
    case = CFG.LOOP_CASES[name]
    rows, final, _ = G.run_oracle(case)
    run = DeviceRun(case, precision=precision)
    got_rows = run_loop_case(run, case, run.eng.set_update_mode)
    gold = np.load(os.path.join(GOLDEN, name + ".npz"))
    keys = [str(k) for k in gold["stat_keys"]]
    for t, (row, got) in enumerate(zip(rows, got_rows)):
        for k, ref in row.items():
            if k in STAT_TO_SLOT and ref is not None:
                assert abs(got[k] - ref) <= loss_tol(k, ref, precision, 2 * case["batch"]), (name, t, k, got[k], ref)
        for j, k in enumerate(keys):
            ref = gold["stats"][t, j]
            if k in STAT_TO_SLOT and not np.isnan(ref):
                assert abs(got[k] - ref) <= loss_tol(k, ref, precision, 2 * case["batch"]), (name, t, k, got[k], ref)
    n_updates = case["steps"] * max(case["n_disc"], case["n_policy"])
    for k in final:
        if k == "log_alpha":
            assert abs(run.eng.get_state().log_alpha - final[k][0]) < 1e-6
        else:
            assert_params_close(run.arena(k), final[k], n_updates, msg="%s/%s" % (name, k), frac={0: 5e-4, 3: 0.1}[precision])


def test_cuda_fused_iteration_equals_split_launches():
    """One fused disc+policy step == a disc-only launch followed by a policy-only launch, bit for bit."""
    case = CFG.CASES["gail_hopper"]
    inj = case_injection(case)
    a = DeviceRun(case, precision=3)
    a.train(case["steps"], inj)
    b = DeviceRun(case, precision=3)
    for t in range(case["steps"]):
        b.eng.set_update_mode(1)
        b.train(1, inj, t_offset=t)
        b.eng.set_update_mode(2)
        b.train(1, inj, t_offset=t)
    for k in ("policy", "qf1", "qf2", "target_qf1", "disc"):
        np.testing.assert_array_equal(a.arena(k), b.arena(k), err_msg=k)
    sa, sb = a.eng.get_state(), b.eng.get_state()
    assert list(sa.adam_step) == list(sb.adam_step) and sa.log_alpha == sb.log_alpha
