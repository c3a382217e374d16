#!/usr/bin/env python
"""TEST INFRASTRUCTURE -- random parity cases (algorithm, dims, batch, hidden widths, AdvIRL mode / options, discriminator
activation) through the host simulator of the product's step programs against the oracle, at the parity tests' bars.
    python tools/fuzz_hostsim.py <first seed> <number of cases> [--variants]
--variants adds the bit-for-bit invariants between variants of one program (check_variants).
tests/test_hostsim_fuzz.py runs a fixed handful of seeds; round 2 ran seeds 1000..1149 against the oracle and
2000..2091 through the variants without a failure."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CFG, G, HostSimRun, STAT_TO_SLOT, assert_params_close, case_injection, load_hostsim  # noqa: E402


def random_case(seed):
    rs = np.random.RandomState(seed)
    algo = str(rs.choice(["sac_alpha", "sac_v", "td3", "adv_irl", "adv_irl"]))
    O, A = int(rs.randint(2, 40)), int(rs.randint(1, 10))
    B = int(rs.choice([8, 17, 33, 40, 64, 96, 130, 200]))
    n_fill = int(rs.randint(B + 5, 600))
    c = dict(algo=algo, obs_dim=O, act_dim=A, batch=B, n_fill=n_fill, steps=3, seed=int(rs.randint(1, 10000)),
             hidden=int(rs.choice([16, 24, 36, 40, 64, 100, 128])))
    if algo in ("sac_alpha", "adv_irl"):
        kw = dict(CFG.SAC_KW, alpha=float(rs.choice([0.2, 1.0])), reward_scale=float(rs.choice([1.0, 2.0, 5.0])))
        if rs.rand() < 0.3:
            kw["train_alpha"] = False
        if rs.rand() < 0.3:
            kw["target_entropy"] = -float(A)
        if rs.rand() < 0.3:
            kw["beta_1"] = 0.25
        c["sac"] = kw
    if algo == "sac_v":
        c["sac"] = dict(CFG.SAC_KW, vf_lr=3e-4, alpha=float(rs.choice([0.2, 1.0])))
    if algo == "td3":
        c["td3"] = dict(reward_scale=1.0, discount=0.99, soft_target_tau=0.005, policy_lr=3e-4, qf_lr=3e-4,
                        policy_and_target_update_period=int(rs.choice([1, 2, 3])))
        c["policy_noise"], c["policy_noise_clip"], c["steps"] = 0.2, 0.5, 4
    if algo == "adv_irl":
        c["n_expert"] = int(rs.randint(B // 2 + 2, max(B // 2 + 3, min(300, n_fill))))
        c["mode"] = str(rs.choice(["airl", "gail", "gail2", "fairl"]))
        c["disc_hid"] = int(rs.choice([16, 20, 48, 100, 128]))
        c["disc_act"] = str(rs.choice(["tanh", "relu"]))
        c["disc"] = dict(disc_lr=3e-4, disc_momentum=float(rs.choice([0.0, 0.9])), use_grad_pen=bool(rs.rand() < 0.7),
                         grad_pen_weight=float(rs.choice([4.0, 8.0, 10.0])))
        if rs.rand() < 0.3:
            c["state_only"] = True
        if rs.rand() < 0.3:
            c["from_expert"] = int(rs.randint(1, B))
        if rs.rand() < 0.3:
            c["rew_clip_min"] = -1.0
        if rs.rand() < 0.3:
            c["rew_clip_max"] = 1.0
    return c


def check_case(lib, case):
    """Losses within 1e-4 relative of the oracle's, parameters within assert_params_close: returns the worst error / bar."""
    rows, final, _ = G.run_oracle(case)
    run = HostSimRun(lib, case)
    L = run.train(case["steps"], case_injection(case))
    worst = 0.0
    for t, row in enumerate(rows):
        for k, ref in row.items():
            if k not in STAT_TO_SLOT or ref is None or np.isnan(L[t, STAT_TO_SLOT[k]]):
                continue
            got = L[t, STAT_TO_SLOT[k]]
            tol = 1e-4 * max(abs(ref), 1e-3) if k != "Policy Loss" else 1e-4 * max(abs(ref), 1.0)
            worst = max(worst, abs(got - ref) / tol)
            assert abs(got - ref) <= tol, (t, k, got, ref)
    for k in final:
        if k == "log_alpha":
            assert abs(lib.hs_log_alpha(run.h) - final[k][0]) < 1e-6
        else:
            assert_params_close(run.arenas[k], final[k], case["steps"], msg=k)
    run.close()
    return worst


def check_variants(lib, case):
    """Invariants between variants of the SAME step program, bit for bit: latency-optimised row jobs == generic row kernels;
    one launch per step == one launch; the two-replica program with identical replicas == the single-replica program
    ((g + g) / 2 == g); AdvIRL disc-only + policy-only launches == the fused iteration."""
    import ctypes as C

    lib.hs_set_generic_rows.argtypes = [C.c_void_p, C.c_int]
    inj, T = case_injection(case), case["steps"]

    def same(a, b, La, Lb, what):
        if Lb is not None:
            np.testing.assert_array_equal(np.nan_to_num(La), np.nan_to_num(Lb), err_msg=what + ": losses")
        for k in a.arenas:
            np.testing.assert_array_equal(a.arenas[k], b.arenas[k], err_msg=what + ": " + k)
        b.close()

    a = HostSimRun(lib, case)
    La = a.train(T, inj)
    b = HostSimRun(lib, case)
    lib.hs_set_generic_rows(b.h, 1)
    same(a, b, La, b.train(T, inj), "generic rows")
    b = HostSimRun(lib, case)
    same(a, b, La, np.concatenate([b.train(1, inj, t_offset=t) for t in range(T)]), "one launch per step")
    b = HostSimRun(lib, case)
    lib.hs_set_world(b.h, 2)
    same(a, b, La, b.train(T, inj), "two identical replicas")
    if case["algo"] == "adv_irl":
        b = HostSimRun(lib, case)
        for t in range(T):
            for mode in (1, 2):
                lib.hs_set_update_mode(b.h, mode)
                b.train(1, inj, t_offset=t)
        same(a, b, La, None, "disc-only + policy-only launches")
    a.close()


if __name__ == "__main__":
    torch.set_num_threads(1)
    lib = load_hostsim()
    first, n = int(sys.argv[1]), int(sys.argv[2])
    bad = 0
    for seed in range(first, first + n):
        case = random_case(seed)
        try:
            worst = check_case(lib, case)
            if "--variants" in sys.argv:
                check_variants(lib, case)
            print(seed, "ok %.3f" % worst, {k: v for k, v in case.items() if k not in ("sac", "td3", "disc")}, flush=True)
        except Exception as e:      # noqa: BLE001
            bad += 1
            print(seed, "FAIL", repr(e)[:300], case, flush=True)
    print("failures", bad)
    sys.exit(1 if bad else 0)
