"""HER relabel-at-sample on the device: DeviceHindsightReplayBuffer + HerTD3.train_from_buffer against
the oracle hindsight buffer (pinned bit-exactly to rlkit's HindsightReplayBuffer) feeding the HER-TD3 oracle."""
import numpy as np
import pytest
import torch

from helpers import CFG, G, R, STAT_TO_SLOT, assert_params_close, her_oracle_rows, her_relabel_setup

pytestmark = pytest.mark.gpu


def _fill_device_buffer(case, seed_rng):
    from ilswiss_b200.replay_buffer import DeviceHindsightReplayBuffer

    O, A, G_ = case["obs_dim"], case["act_dim"], case["her"]["goal_dim"]
    buf = DeviceHindsightReplayBuffer(case["n_fill"], O - G_, G_, A, random_seed=CFG.BUFFER_SEED, her_ratio=case["her_ratio"],
                                      distance_threshold=case["threshold"])
    rs = np.random.RandomState(seed_rng)
    for ep in R.synth_goal_episodes(rs, case["n_episodes"], case["T"], O - G_, G_, A):
        for (o, a, r, d, no) in ep:
            buf.add_sample(o, a, r, d, no)
        buf.terminate_episode()
    return buf


def _modules(case):
    from ilswiss_b200 import modules

    O, A = case["obs_dim"], case["act_dim"]
    nets = G.build_oracle_nets(case)
    mods = {"qf1": modules.FlattenMlp([256, 256], 1, O + A), "qf2": modules.FlattenMlp([256, 256], 1, O + A),
            "policy": modules.DeterministicNoisePolicy([256, 256], O, A)}
    for k, m in mods.items():
        with torch.no_grad():
            for p, v in zip(m.parameters(), nets[k].p.values()):
                p.copy_(v)
    mods["policy"].sigma, mods["policy"].min_act = case["her"]["sigma"], -1.0
    return mods


@pytest.mark.parametrize("name", [n for n, c in CFG.HER_RELABEL_CASES.items() if c["algo"] == "td3"])
def test_her_td3_relabel_at_sample_matches_oracle(name):
    from ilswiss_b200.trainers import HerTD3

    torch.set_num_threads(1)
    case = CFG.HER_RELABEL_CASES[name]
    setup = her_relabel_setup(case)
    rows, final = her_oracle_rows(case, setup)
    buf = _fill_device_buffer(case, case["seed"])
    assert buf._traj_endpoints == setup["oracle"]._traj_endpoints and buf.num_steps_can_sample() == setup["ring_size"]
    mods = _modules(case)
    tr = HerTD3(mods["policy"], mods["qf1"], mods["qf2"], batch_size=case["batch"], gemm_precision=0, **case["td3"])
    inj = {k: torch.from_numpy(setup[k]).cuda() for k in ("idx", "eps_next", "idx_her")}
    tr.train_from_buffer(buf, case["steps"], inject=inj)
    L = tr.engine.losses(case["steps"])
    for t, row in enumerate(rows):
        for k, ref in row.items():
            got = float(L[t, STAT_TO_SLOT[k]])
            tol = 1e-4 * max(abs(ref), 1e-2) if k != "Policy Loss" else 1e-4 * max(abs(ref), 1.0)
            assert abs(got - ref) <= tol, (name, t, k, got, ref)
    for k in ("policy", "qf1", "qf2"):
        got = np.concatenate([p.detach().cpu().numpy().ravel() for p in mods[k].parameters()])
        assert_params_close(got, final[k], case["steps"], lr=6e-4, msg="%s/%s" % (name, k))
    # speed mode: in-kernel Philox trajectory / step / future-step draws
    tr.end_epoch()
    tr.train_from_buffer(buf, 50)
    st = tr.get_eval_statistics()
    assert np.isfinite(st["QF1 Loss"]) and st["Q Targets Max"] <= 1e-6          # rewards in {-1, 0}, clipped returns <= 0


def test_device_hindsight_buffer_random_batch_matches_oracle():
    """Host-facing random_batch: same RNG streams, same relabelled batch as the reference's buffer (float32-rounded values)."""
    case = CFG.HER_RELABEL_CASES["her_td3_relabel"]
    setup = her_relabel_setup(case)            # consumes the same episode RNG as the device fill
    buf = _fill_device_buffer(case, case["seed"])
    ora = setup["oracle"]
    ora._np_rand_state = np.random.RandomState(CFG.BUFFER_SEED)       # rewind: her_relabel_setup already drew from it
    for trial in range(2):
        np.random.seed(7 + trial)
        want = ora.random_batch(48)
        np.random.seed(7 + trial)
        got = buf.random_batch(48)
        for k in ("observations", "next_observations", "desired_goals", "next_desired_goals", "actions", "rewards", "terminals",
                  "next_achieved_goals"):
            np.testing.assert_array_equal(np.asarray(got[k], dtype=np.float32), np.asarray(want[k], dtype=np.float32), err_msg=k)


def test_her_algorithm_mixin_and_env_buffer_constructor():
    """DeviceHERMixin in front of a stand-in of rlkit's HER (her.py:8-33) with the buffer built from a goal env's spaces."""
    from ilswiss_b200.adv_irl import DeviceHERMixin
    from ilswiss_b200.replay_buffer import DeviceEnvHindsightReplayBuffer
    from ilswiss_b200.trainers import HerTD3

    case = CFG.HER_RELABEL_CASES["her_td3_relabel"]
    O, A, G_ = case["obs_dim"], case["act_dim"], case["her"]["goal_dim"]

    class Box:
        def __init__(self, n):
            self.low, self.shape = -np.ones(n), (n,)

    class Dict:
        def __init__(self, **spaces):
            self.spaces = spaces

    class Env:
        observation_space = Dict(observation=Box(O - G_), achieved_goal=Box(G_), desired_goal=Box(G_))
        action_space = Box(A)
        distance_threshold = 0.05

    buf = DeviceEnvHindsightReplayBuffer(case["n_fill"], Env(), random_seed=3, relabel_type="future", her_ratio=0.8)
    assert buf._obs0_dim == O - G_ and buf._goal_dim == G_ and buf.distance_threshold == 0.05
    rs = np.random.RandomState(1)
    for ep in R.synth_goal_episodes(rs, 6, 40, O - G_, G_, A):
        for (o, a, r, d, no) in ep:
            buf.add_sample(o, a, r, d, no)
        buf.terminate_episode()
    mods = _modules(case)

    class RefHER:                                   # attribute names of torch_rl_algorithm.py:8-34 / her.py
        def __init__(self, trainer, replay_buffer):
            self.trainer, self.replay_buffer = trainer, replay_buffer
            self.batch_size, self.num_train_steps_per_train_call = 32, 25

        def _do_training(self, epoch):
            raise AssertionError("reference path must be overridden")

    class HER(DeviceHERMixin, RefHER):
        pass

    alg = HER(HerTD3(mods["policy"], mods["qf1"], mods["qf2"], **case["td3"]), buf)
    n0 = alg.trainer.engine.kernel_launches
    alg._do_training(0)
    assert alg.trainer._cfg.batch == 32                                       # adopted from the algorithm
    assert alg.trainer.engine.kernel_launches == n0 + 1 and alg.trainer.engine.get_state().n_train_steps_total == 25
    b = alg.get_batch()
    assert b["observations"].shape == (32, O - G_) and b["desired_goals"].shape == (32, G_) and b["rewards"].is_cuda
    alg.trainer.train_step(b)                                                 # the reference's per-step path works too
    assert np.isfinite(alg.trainer.get_eval_statistics()["QF1 Loss"])
