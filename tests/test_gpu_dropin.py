"""-m gpu: the reference-facing drop-in classes (Trainer / ReplayBuffer / AdvIRL interfaces)."""
import pickle

import numpy as np
import pytest
import torch

from helpers import CFG, G, R, STAT_TO_SLOT, case_data, case_injection, layout

pytestmark = pytest.mark.gpu


def build_modules(case):
    """Parameter containers (ilswiss_b200.modules) loaded with the seed-reproducible oracle init."""
    from ilswiss_b200 import modules

    O, A = case["obs_dim"], case["act_dim"]
    nets = G.build_oracle_nets(case)
    mods = {"qf1": modules.FlattenMlp([256, 256], 1, O + A), "qf2": modules.FlattenMlp([256, 256], 1, O + A)}
    if case["algo"] == "td3":
        mods["policy"] = modules.DeterministicNoisePolicy([256, 256], O, A, policy_noise=case["policy_noise"],
                                                          policy_noise_clip=case["policy_noise_clip"])
    else:
        mods["policy"] = modules.TanhGaussianPolicy([256, 256], O, A)
    if case["algo"] == "adv_irl":
        mods["disc"] = modules.MLPDisc(O + A, 128, hid_act=case.get("disc_act", "tanh"))
    for k, m in mods.items():
        with torch.no_grad():
            for p, v in zip(m.parameters(), nets[k].p.values()):
                assert tuple(p.shape) == tuple(v.shape), (k, p.shape, v.shape)
                p.copy_(v)
    return mods, nets


def fill(buf, data):
    buf.add_samples(**data)


def test_sac_trainer_dropin_train_step_and_stats_keys():
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import SoftActorCritic

    torch.set_num_threads(1)
    case = CFG.CASES["sac_hopper"]
    mods, nets = build_modules(case)
    policy_ref = mods["policy"]
    tr = SoftActorCritic(mods["policy"], mods["qf1"], mods["qf2"], batch_size=case["batch"], gemm_precision=3, **case["sac"])
    assert tr.policy is policy_ref and len(tr.networks) == 5
    assert all(p.is_cuda for p in tr.policy.parameters())
    data, _ = case_data(case)
    buf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=CFG.BUFFER_SEED)
    fill(buf, data)
    assert buf.num_steps_can_sample() == case["n_fill"] and buf._max_replay_buffer_size == case["n_fill"]
    # reference-side oracle on identical indices / eps
    ora = R.SacAlphaOracle(nets["policy"], nets["qf1"], nets["qf2"], case["act_dim"], **case["sac"])
    obuf = R.ReplayOracle(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=CFG.BUFFER_SEED)
    obuf.load_bulk(data)
    inj = case_injection(case)
    dev = {k: torch.from_numpy(v).cuda() for k, v in inj.items()}
    tr.train_from_buffer(buf, case["steps"], inject=dev)
    L = tr.engine.losses(case["steps"])
    for t in range(case["steps"]):
        torch.manual_seed(CFG.EPS_SEED0 + t)
        batch = R.np_to_torch_batch(obuf.random_batch(case["batch"]))
        s = ora.train_step(batch, torch.randn(case["batch"], case["act_dim"]), torch.randn(case["batch"], case["act_dim"]))
        assert abs(L[t, 0] - s["qf1_loss"]) <= 1e-4 * max(abs(s["qf1_loss"]), 1e-2)
        assert abs(L[t, 2] - s["policy_loss"]) <= 1e-4
    # the SAME nn.Module objects now hold the trained weights (exploration policy stays valid)
    got = np.concatenate([p.detach().cpu().numpy().ravel() for p in policy_ref.parameters()])
    assert np.max(np.abs(got - nets["policy"].flat())) < 5e-4
    st = tr.get_eval_statistics()
    expect = ["Reward Scale", "QF1 Loss", "QF2 Loss", "Alpha Loss", "Policy Loss"]
    for name in ("Q1 Predictions", "Q2 Predictions", "Alpha", "Log Pis", "Policy mu", "Policy log std"):
        expect += [name + s for s in (" Mean", " Std", " Max", " Min")]
    assert list(st.keys()) == expect                       # sac_alpha.py:186-233 key set and order
    tr.end_epoch()
    assert tr.get_eval_statistics() is None
    # Trainer.train_step(batch) with a dict of device tensors (np_to_pytorch_batch equivalent)
    tr.train_step(buf.random_batch_device(case["batch"]))
    assert set(tr.get_eval_statistics().keys()) == set(expect)
    # snapshot: picklable, optimizer state populated, log_alpha float64 0-dim
    snap = tr.get_snapshot()
    assert set(snap.keys()) == {"qf1", "qf2", "policy", "target_qf1", "target_qf2", "log_alpha", "policy_optimizer",
                                "qf1_optimizer", "qf2_optimizer", "alpha_optimizer"}
    assert snap["log_alpha"].dtype == torch.float64 and snap["log_alpha"].dim() == 0
    osd = snap["qf1_optimizer"].state_dict()
    assert len(osd["state"]) == 6 and float(osd["state"][0]["step"]) == case["steps"] + 1
    pickle.loads(pickle.dumps({k: v for k, v in snap.items() if "optimizer" not in k}))


def test_snapshot_round_trip_resumes_bit_identically():
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import SoftActorCritic

    case = CFG.CASES["sac_hopper"]
    inj = {k: torch.from_numpy(v).cuda() for k, v in case_injection(case).items()}
    data, _ = case_data(case)

    def mk():
        mods, _ = build_modules(case)
        tr = SoftActorCritic(mods["policy"], mods["qf1"], mods["qf2"], batch_size=case["batch"], gemm_precision=3, **case["sac"])
        buf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=1)
        fill(buf, data)
        tr.eval_statistics = {}
        return tr, buf

    a, abuf = mk()
    a.train_from_buffer(abuf, 6, inject=inj)
    b, bbuf = mk()
    b.train_from_buffer(bbuf, 3, inject={k: v[:3] for k, v in inj.items()})
    snap = b.get_snapshot()
    c, cbuf = mk()
    c.load_snapshot(snap)
    c.train_from_buffer(cbuf, 3, inject={k: v[3:] for k, v in inj.items()})
    for name in ("policy", "qf1", "target_qf2"):
        pa = torch.cat([p.flatten() for p in getattr(a, name).parameters()])
        pc = torch.cat([p.flatten() for p in getattr(c, name).parameters()])
        assert torch.equal(pa, pc), name
    assert float(a.log_alpha) == float(c.log_alpha)


def test_td3_trainer_dropin():
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import TD3

    case = CFG.CASES["td3_hopper"]
    mods, nets = build_modules(case)
    tr = TD3(mods["policy"], mods["qf1"], mods["qf2"], batch_size=case["batch"], gemm_precision=3, **case["td3"])
    assert len(tr.networks) == 6
    data, _ = case_data(case)
    buf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=1)
    fill(buf, data)
    inj = {k: torch.from_numpy(v).cuda() for k, v in case_injection(case).items()}
    tr.train_from_buffer(buf, case["steps"], inject=inj)
    rows, final, _ = G.run_oracle(case)
    L = tr.engine.losses(case["steps"])
    for t, row in enumerate(rows):
        assert abs(L[t, 0] - row["QF1 Loss"]) <= 1e-4 * max(abs(row["QF1 Loss"]), 1e-2)
    st = tr.get_eval_statistics()
    assert list(st.keys())[:3] == ["QF1 Loss", "QF2 Loss", "Policy Loss"]
    for name in ("Q1 Predictions", "Q2 Predictions", "Q Targets", "Bellman Errors 1", "Bellman Errors 2", "Policy Action"):
        assert name + " Mean" in st
    assert tr._n_train_steps_total == case["steps"]


def test_her_td3_trainer_dropin():
    """her/td3.py through the drop-in class: goal-conditioned batches via train_step (desired_goals keys), and the
    in-HBM path on a ring whose rows hold cat(obs, goal); both against the oracle (bit-identical to the reference)."""
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import HerTD3

    torch.set_num_threads(1)
    case = CFG.CASES["her_td3_reach"]
    her, O, A, B = case["her"], case["obs_dim"], case["act_dim"], case["batch"]
    G_ = her["goal_dim"]
    rows, final, _ = G.run_oracle(case)
    inj_np = case_injection(case)
    data, _ = case_data(case)

    def make():
        mods, _ = build_modules(case)
        mods["policy"].sigma, mods["policy"].min_act = her["sigma"], -1.0
        return HerTD3(mods["policy"], mods["qf1"], mods["qf2"], batch_size=B, gemm_precision=3, **case["td3"]), mods

    # (a) in-HBM ring path, injected indices / noise
    tr, mods = make()
    assert tr.clip_return_l == pytest.approx(-100.0) and tr.clip_return_r == 0.0
    buf = DeviceReplayBuffer(case["n_fill"], O, A, random_seed=1)
    fill(buf, data)
    tr.train_from_buffer(buf, case["steps"], inject={k: torch.from_numpy(v).cuda() for k, v in inj_np.items()})
    L = tr.engine.losses(case["steps"])
    for t, row in enumerate(rows):
        assert abs(L[t, 0] - row["QF1 Loss"]) <= 1e-4 * max(abs(row["QF1 Loss"]), 1e-2)
        if row.get("Policy Loss") is not None:
            assert abs(L[t, 2] - row["Policy Loss"]) <= 1e-4 * max(abs(row["Policy Loss"]), 1.0)
    got = np.concatenate([p.detach().cpu().numpy().ravel() for p in mods["policy"].parameters()])
    assert np.max(np.abs(got - final["policy"])) < 2 * 6e-4 * case["steps"] + 1e-6
    # (b) the reference's batch interface: goals handed over separately, one step per call (Philox target noise)
    tr2, _ = make()
    idx = inj_np["idx"][0]
    b = dict(observations=data["observations"][idx][:, :O - G_], desired_goals=data["observations"][idx][:, O - G_:],
             next_observations=data["next_observations"][idx][:, :O - G_], next_desired_goals=data["next_observations"][idx][:, O - G_:],
             actions=data["actions"][idx], rewards=data["rewards"][idx].reshape(B, 1), terminals=data["terminals"][idx].reshape(B, 1))
    tr2.train_step(b)
    st = tr2.get_eval_statistics()
    assert np.isfinite(st["QF1 Loss"]) and "Q Targets Mean" in st
    assert st["Q Targets Max"] <= float(np.max(data["rewards"][idx])) + 1e-5          # min target Q is clipped to <= 0


def test_her_sac_trainer_dropin():
    """her/sac.py: sac_alpha on cat(obs, goal) with target entropy -A; goal-conditioned batches through train_step."""
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import HerSAC

    torch.set_num_threads(1)
    case = CFG.CASES["her_sac_reach"]
    O, A, B, G_ = case["obs_dim"], case["act_dim"], case["batch"], case["her"]["goal_dim"]
    mods, _ = build_modules(case)
    kw = {k: v for k, v in case["sac"].items() if k != "target_entropy"}
    tr = HerSAC(mods["policy"], mods["qf1"], mods["qf2"], batch_size=B, gemm_precision=3, **kw)
    assert tr.target_entropy == -float(A)                           # her/sac.py:52, not -A/2
    data, _ = case_data(case)
    buf = DeviceReplayBuffer(case["n_fill"], O, A, random_seed=1)
    fill(buf, data)
    inj_np = case_injection(case)
    tr.train_from_buffer(buf, case["steps"], inject={k: torch.from_numpy(v).cuda() for k, v in inj_np.items()})
    rows, final, _ = G.run_oracle(case)
    L = tr.engine.losses(case["steps"])
    for t, row in enumerate(rows):
        for key, slot in (("QF1 Loss", 0), ("QF2 Loss", 1), ("Policy Loss", 2), ("Alpha Loss", 3)):
            assert abs(L[t, slot] - row[key]) <= 1e-4 * max(abs(row[key]), 1.0 if key == "Policy Loss" else 1e-2), (t, key)
    assert abs(float(tr.log_alpha) - final["log_alpha"][0]) < 1e-6
    idx = inj_np["idx"][0]
    b = dict(observations=data["observations"][idx][:, :O - G_], desired_goals=data["observations"][idx][:, O - G_:],
             next_observations=data["next_observations"][idx][:, :O - G_], next_desired_goals=data["next_observations"][idx][:, O - G_:],
             actions=data["actions"][idx], rewards=data["rewards"][idx].reshape(B, 1), terminals=data["terminals"][idx].reshape(B, 1))
    tr.end_epoch()
    tr.train_step(b)
    assert np.isfinite(tr.get_eval_statistics()["QF1 Loss"])


def test_replay_buffer_dropin_matches_reference_semantics():
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer

    rs = np.random.RandomState(0)
    O, A, cap = 5, 2, 40
    buf = DeviceReplayBuffer(cap, O, A, random_seed=11)
    ora = R.ReplayOracle(cap, O, A, random_seed=11)
    for ep in range(5):                       # episode bursts through the per-transition API
        n = 13
        path = dict(observations=rs.randn(n, O), actions=rs.uniform(-1, 1, (n, A)), rewards=rs.randn(n, 1),
                    next_observations=rs.randn(n, O), terminals=np.array([[i == n - 1] for i in range(n)]))
        buf.add_path(path)
        ora.add_path(path)
        assert buf.num_steps_can_sample() == ora.num_steps_can_sample()
        assert buf._top == ora._top and buf._traj_endpoints == ora._traj_endpoints
        got, ref = buf.random_batch(32), ora.random_batch(32)       # same RandomState stream
        assert set(got.keys()) == set(ref.keys())
        for k in ref:
            assert got[k].dtype == ref[k].dtype and got[k].shape == ref[k].shape, k
            np.testing.assert_array_equal(got[k].astype(np.float32), ref[k].astype(np.float32), err_msg=k)
        sub = buf.random_batch(8, keys=["observations", "actions"])
        ora.random_batch(8, keys=["observations", "actions"])
        assert set(sub.keys()) == {"observations", "actions"}
    clone = pickle.loads(pickle.dumps(buf))   # save_replay_buffer: true round trip
    idx = np.arange(cap)
    a, b = buf._get_batch_using_indices(idx), clone._get_batch_using_indices(idx)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    assert clone._top == buf._top and clone._size == buf._size


def test_expert_demo_ingest_and_save_data(tmp_path):
    """8f rank 3: pickled trajectory list -> normalised demos -> expert ring (adv_irl_exp_script.py:51-138), and the
    reference's on-disk dump format (simple_replay_buffer.py:110-123)."""
    from ilswiss_b200 import demos
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer

    rs = np.random.RandomState(4)
    O, A, T = 6, 2, 25
    trajs = []
    for i in range(5):
        obs = rs.randn(T + 1, O) * (i + 1)
        trajs.append(dict(observations=obs[:-1], next_observations=obs[1:], actions=rs.uniform(-1, 1, (T, A)),
                          rewards=rs.randn(T, 1), terminals=np.zeros((T, 1), dtype=bool)))
    path = str(tmp_path / "demos.pkl")
    with open(path, "wb") as f:
        pickle.dump(trajs, f)
    buf = DeviceReplayBuffer(1000, O, A, random_seed=3)
    ora = R.ReplayOracle(1000, O, A, random_seed=3)
    import random
    sel, st, name, kw = demos.ingest(path, buf, traj_num=4, scale_env_with_demo_stats=True, rng=random.Random(2))
    for t in sel:
        ora.add_path(t)
    assert name == "ScaledEnv" and buf.num_steps_can_sample() == 4 * T == ora.num_steps_can_sample()
    assert buf._traj_endpoints == ora._traj_endpoints and buf.get_traj_num() == 4
    got, ref = buf.random_batch(64), ora.random_batch(64)
    for k in ref:
        np.testing.assert_array_equal(got[k].astype(np.float32), ref[k].astype(np.float32), err_msg=k)
    allobs = buf.get_all(keys=["observations"])["observations"]
    assert abs(allobs.mean()) < 1e-6 and abs(allobs.std(0).mean() - 1.0) < 1e-4       # demos were standardised
    dump = str(tmp_path / "buffer.pkl")
    buf.save_data(dump)
    with open(dump, "rb") as f:
        d = pickle.load(f)
    assert set(d) == {"observations", "actions", "next_observations", "terminals", "timeouts", "rewards", "agent_infos", "env_infos"}
    assert d["observations"].shape == (4 * T, O) and len(d["env_infos"]) == 4 * T
    np.testing.assert_array_equal(d["observations"].astype(np.float32), ora._observations[:4 * T].astype(np.float32))


@pytest.mark.parametrize("name", ["gail_walker", "gail_hopper_relu"])      # tanh (the shipped yamls) and relu MLPDisc blocks
def test_adv_irl_engine_matches_oracle_and_stats_keys(name):
    from ilswiss_b200.adv_irl import AdvIRLEngine
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import SoftActorCritic

    torch.set_num_threads(1)
    case = CFG.CASES[name]
    mods, _ = build_modules(case)
    tr = SoftActorCritic(mods["policy"], mods["qf1"], mods["qf2"], batch_size=case["batch"], gemm_precision=3, **case["sac"])
    data, edata = case_data(case)
    buf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=1)
    ebuf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=3)
    fill(buf, data)
    fill(ebuf, edata)
    irl = AdvIRLEngine(case["mode"], mods["disc"], tr, ebuf, buf, disc_optim_batch_size=case["batch"],
                       policy_optim_batch_size=case["batch"], num_update_loops_per_train_call=case["steps"],
                       num_disc_updates_per_loop_iter=1, num_policy_updates_per_loop_iter=1, **case["disc"])
    inj = {k: torch.from_numpy(v).cuda() for k, v in case_injection(case).items()}
    irl.do_training(inject=inj)
    rows, final, _ = G.run_oracle(case)
    st = irl.disc_eval_statistics
    assert list(st.keys()) == ["Disc CE Loss", "Disc Acc", "Grad Pen", "Grad Pen W", "Disc Rew Mean", "Disc Rew Std",
                               "Disc Rew Max", "Disc Rew Min"]
    assert abs(st["Disc CE Loss"] - rows[0]["Disc CE Loss"]) < 1e-4 * abs(rows[0]["Disc CE Loss"])
    assert abs(st["Grad Pen"] - rows[0]["Grad Pen"]) < 1e-4 * max(abs(rows[0]["Grad Pen"]), 1e-2)
    assert abs(st["Disc Rew Mean"] - rows[-1]["Disc Rew Mean"]) < 1e-4 * abs(rows[-1]["Disc Rew Mean"])   # last write wins
    got = np.concatenate([p.detach().cpu().numpy().ravel() for p in mods["disc"].parameters()])
    assert np.max(np.abs(got - final["disc"])) < 1e-4
    with pytest.raises(NotImplementedError):
        AdvIRLEngine("gail", mods["disc"], tr, ebuf, buf, wrap_absorbing=True, disc_optim_batch_size=256,
                     policy_optim_batch_size=256, num_disc_updates_per_loop_iter=1, num_policy_updates_per_loop_iter=1)


def test_adv_irl_engine_multiple_updates_per_loop_iter():
    """num_disc_updates_per_loop_iter / num_policy_updates_per_loop_iter != 1 (gail_humanoid.yaml) through the drop-in
    engine: statistics follow the reference's logging rules, parameters follow the oracle's nested loops."""
    from helpers import loop_case_injection
    from ilswiss_b200.adv_irl import AdvIRLEngine
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import SoftActorCritic

    torch.set_num_threads(1)
    case = CFG.LOOP_CASES["gail_nd2_np3"]
    mods, _ = build_modules(case)
    tr = SoftActorCritic(mods["policy"], mods["qf1"], mods["qf2"], batch_size=case["batch"], gemm_precision=3, **case["sac"])
    data, edata = case_data(case)
    buf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=1)
    ebuf = DeviceReplayBuffer(case["n_fill"], case["obs_dim"], case["act_dim"], random_seed=3)
    fill(buf, data)
    fill(ebuf, edata)
    irl = AdvIRLEngine(case["mode"], mods["disc"], tr, ebuf, buf, disc_optim_batch_size=case["batch"],
                       policy_optim_batch_size=case["batch"], num_update_loops_per_train_call=case["steps"],
                       num_disc_updates_per_loop_iter=case["n_disc"], num_policy_updates_per_loop_iter=case["n_policy"],
                       **case["disc"])
    inj = {k: torch.from_numpy(v).cuda() for k, v in loop_case_injection(case).items()}
    irl.do_training(inject=inj)          # steps loop iterations in one train call: stats of the FIRST updates, rewards of the LAST
    rows, final, _ = G.run_oracle(case)
    st = irl.disc_eval_statistics
    assert abs(st["Disc CE Loss"] - rows[0]["Disc CE Loss"]) < 1e-4 * abs(rows[0]["Disc CE Loss"])
    assert abs(st["Disc Rew Mean"] - rows[-1]["Disc Rew Mean"]) < 1e-4 * max(abs(rows[-1]["Disc Rew Mean"]), 1.0)
    assert abs(tr.eval_statistics["QF1 Loss"] - rows[0]["QF1 Loss"]) < 1e-4 * abs(rows[0]["QF1 Loss"])
    got = np.concatenate([p.detach().cpu().numpy().ravel() for p in mods["disc"].parameters()])
    assert np.max(np.abs(got - final["disc"])) < 1e-4
    gotp = np.concatenate([p.detach().cpu().numpy().ravel() for p in mods["policy"].parameters()])
    assert np.max(np.abs(gotp - final["policy"])) < 2 * 3e-4 * 6 + 1e-6
    state = tr.engine.get_state()
    assert state.adam_step[4] == case["steps"] * case["n_disc"] and state.adam_step[2] == case["steps"] * case["n_policy"]
    # speed mode (in-kernel Philox) runs too
    irl.do_training(n_loops=1)
    assert np.isfinite(irl.disc_eval_statistics["Disc Rew Mean"])


def test_sac_v_trainer_dropin():
    from ilswiss_b200 import modules
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import SoftActorCriticV

    torch.set_num_threads(1)
    case = CFG.CASES["sacv_hopper"]
    O, A = case["obs_dim"], case["act_dim"]
    nets = G.build_oracle_nets(case)
    mods = {"qf1": modules.FlattenMlp([256, 256], 1, O + A), "qf2": modules.FlattenMlp([256, 256], 1, O + A),
            "vf": modules.FlattenMlp([256, 256], 1, O), "policy": modules.TanhGaussianPolicy([256, 256], O, A)}
    for k, m in mods.items():
        with torch.no_grad():
            for p, v in zip(m.parameters(), nets[k].p.values()):
                p.copy_(v)
    tr = SoftActorCriticV(mods["policy"], mods["qf1"], mods["qf2"], mods["vf"], batch_size=case["batch"],
                          gemm_precision=3, **case["sac"])
    assert len(tr.networks) == 5
    data, _ = case_data(case)
    buf = DeviceReplayBuffer(case["n_fill"], O, A, random_seed=1)
    fill(buf, data)
    inj = {k: torch.from_numpy(v).cuda() for k, v in case_injection(case).items()}
    tr.train_from_buffer(buf, case["steps"], inject=inj)
    rows, final, _ = G.run_oracle(case)
    L = tr.engine.losses(case["steps"])
    for t, row in enumerate(rows):
        for key, slot in (("QF1 Loss", 0), ("QF2 Loss", 1), ("VF Loss", 5), ("Policy Loss", 2)):
            assert abs(L[t, slot] - row[key]) <= 1e-4 * max(abs(row[key]), 1.0 if key == "Policy Loss" else 1e-2), (t, key)
    st = tr.get_eval_statistics()
    assert list(st.keys())[:5] == ["Reward Scale", "QF1 Loss", "QF2 Loss", "VF Loss", "Policy Loss"]
    assert "V Predictions Mean" in st and "Policy log std Min" in st
    snap = tr.get_snapshot()
    assert set(snap) == {"qf1", "qf2", "policy", "vf", "target_vf", "policy_optimizer", "qf1_optimizer", "qf2_optimizer", "vf_optimizer"}


def test_policy_act_kernel_matches_module_forward():
    """A1: sampler-side inference kernel == the nn.Module's own deterministic forward."""
    from ilswiss_b200 import modules
    from ilswiss_b200.trainers import TD3, SoftActorCritic

    torch.manual_seed(0)
    O, A = 17, 6
    pol = modules.TanhGaussianPolicy([256, 256], O, A)
    tr = SoftActorCritic(pol, modules.FlattenMlp([256, 256], 1, O + A), modules.FlattenMlp([256, 256], 1, O + A), batch_size=64)
    obs = torch.randn(10, O, device="cuda")
    got = tr.engine.policy_act(obs, deterministic=True)
    with torch.no_grad():
        ref = pol(obs, deterministic=True)[0]
    assert torch.allclose(got, ref, atol=2e-6, rtol=1e-5)
    sto = tr.engine.policy_act(obs, deterministic=False, seed=3)
    assert sto.abs().max() <= 1.0 and not torch.allclose(sto, ref)
    assert torch.equal(sto, tr.engine.policy_act(obs, deterministic=False, seed=3))
    # the module is still a live view of the trained parameters: get_actions works on numpy input
    act = pol.get_actions(obs.cpu().numpy(), deterministic=True)
    assert act.shape == (10, A) and np.allclose(act, ref.cpu().numpy(), atol=1e-6)
    dpol = modules.DeterministicNoisePolicy([256, 256], O, A, policy_noise=0.2, policy_noise_clip=0.5)
    t3 = TD3(dpol, modules.FlattenMlp([256, 256], 1, O + A), modules.FlattenMlp([256, 256], 1, O + A), batch_size=64)
    with torch.no_grad():
        ref3 = dpol(obs, deterministic=True)[0]
    assert torch.allclose(t3.engine.policy_act(obs, deterministic=True), ref3, atol=2e-6, rtol=1e-5)


def test_algorithm_mixins_drive_one_launch_per_train_call():
    """DeviceTorchRLAlgorithmMixin / DeviceAdvIRLMixin in front of minimal stand-ins of the reference
    algorithm classes (same attribute names as torch_rl_algorithm.py:8-34 and adv_irl.py:34-131)."""
    from ilswiss_b200 import modules
    from ilswiss_b200.adv_irl import DeviceAdvIRLMixin, DeviceTorchRLAlgorithmMixin
    from ilswiss_b200.replay_buffer import DeviceReplayBuffer
    from ilswiss_b200.trainers import SoftActorCritic

    O, A, B = 11, 3, 64
    rs = np.random.RandomState(0)

    def mk_buf(seed):
        buf = DeviceReplayBuffer(5000, O, A, random_seed=seed)
        buf.add_samples(rs.randn(3000, O), rs.uniform(-1, 1, (3000, A)), rs.randn(3000, 1), np.zeros((3000, 1)), rs.randn(3000, O))
        return buf

    def mk_trainer():
        return SoftActorCritic(modules.TanhGaussianPolicy([256, 256], O, A), modules.FlattenMlp([256, 256], 1, O + A),
                               modules.FlattenMlp([256, 256], 1, O + A), batch_size=B, max_steps_per_call=50)

    class RefRL:                                   # stand-in for rlkit TorchRLAlgorithm
        def __init__(self, trainer, replay_buffer):
            self.trainer, self.replay_buffer = trainer, replay_buffer
            self.batch_size, self.num_train_steps_per_train_call = B, 37

        def _do_training(self, epoch):
            raise AssertionError("reference path must be overridden")

    class Alg(DeviceTorchRLAlgorithmMixin, RefRL):
        pass

    # sampler coupling through the mixin: exploration_policy is the trainer's module -> DevicePolicy round trip
    class _Space:
        def sample(self):
            return np.zeros(A)

    samp = Alg(mk_trainer(), mk_buf(1))
    samp.exploration_policy, samp._n_env_steps_total, samp.action_space = samp.trainer.policy, 0, _Space()
    samp.trainer.policy.set_num_steps_total = lambda t: None
    samp._can_train = lambda: True
    obs4 = rs.randn(4, O)
    acts = samp._get_action_and_info(obs4)
    assert acts.shape == (4, A) and np.abs(acts).max() < 1.0 and samp._ilsw_dp is not False
    samp._can_train = lambda: False
    assert len(samp._get_action_and_info(obs4)) == 4            # warm-up: random actions (base_algorithm.py:376-377)

    alg = Alg(mk_trainer(), mk_buf(1))
    n0 = alg.trainer.engine.kernel_launches
    alg._do_training(0)
    assert alg.trainer.engine.kernel_launches == n0 + 1                      # 37 gradient steps, one launch
    assert alg.trainer.engine.get_state().n_train_steps_total == 37
    b = alg.get_batch()
    assert b["observations"].is_cuda and b["observations"].shape == (B, O) and b["rewards"].shape == (B, 1)
    assert "QF1 Loss" in alg.trainer.get_eval_statistics()
    # the reference gives batch_size to the ALGORITHM, not the trainer: a trainer built with the default batch adopts it
    # on first use (ensure_batch)
    tr_default = SoftActorCritic(modules.TanhGaussianPolicy([256, 256], O, A), modules.FlattenMlp([256, 256], 1, O + A),
                                 modules.FlattenMlp([256, 256], 1, O + A))
    assert tr_default._cfg.batch == 256
    alg2 = Alg(tr_default, mk_buf(1))
    alg2._do_training(0)
    assert tr_default._cfg.batch == B and tr_default.engine.get_state().n_train_steps_total == 37
    # a later batch-size change rebuilds the step program and CARRIES the optimiser state (resume path: load_snapshot,
    # then the algorithm's batch size on the first _do_training -- ADVICE r1)
    st0 = tr_default.engine.get_state()
    tr_default.ensure_batch(2 * B)
    st1 = tr_default.engine.get_state()
    assert tr_default._cfg.batch == 2 * B and st1.n_train_steps_total == 37 and list(st1.adam_step) == list(st0.adam_step)
    assert st1.log_alpha == st0.log_alpha and st1.alpha_exp_avg == st0.alpha_exp_avg and st1.alpha_step == st0.alpha_step
    tr_step = SoftActorCritic(modules.TanhGaussianPolicy([256, 256], O, A), modules.FlattenMlp([256, 256], 1, O + A),
                              modules.FlattenMlp([256, 256], 1, O + A))
    tr_step.train_step(alg.get_batch())                # Trainer.train_step(batch) with a 64-row batch
    assert tr_step._cfg.batch == B and np.isfinite(tr_step.get_eval_statistics()["QF1 Loss"])

    class RefIRL:                                  # stand-in for rlkit AdvIRL (attribute names of adv_irl.py:63-104)
        def __init__(self):
            self.mode, self.state_only = "gail2", False
            self.discriminator = modules.MLPDisc(O + A, 128)
            self.policy_trainer = mk_trainer()
            self.expert_replay_buffer, self.replay_buffer = mk_buf(3), mk_buf(1)
            self.disc_optim_batch_size = self.policy_optim_batch_size = B
            self.policy_optim_batch_size_from_expert = 0
            self.num_update_loops_per_train_call = 21
            self.num_disc_updates_per_loop_iter = self.num_policy_updates_per_loop_iter = 1
            self.disc_optimizer = torch.optim.Adam(self.discriminator.parameters(), lr=3e-4, betas=(0.9, 0.999))
            self.use_grad_pen, self.grad_pen_weight = True, 8.0
            self.rew_clip_min = self.rew_clip_max = None
            self.wrap_absorbing = False
            self.disc_eval_statistics = None

    class IRL(DeviceAdvIRLMixin, RefIRL):
        pass

    irl = IRL()
    irl._do_training(0)
    st = irl.disc_eval_statistics
    assert st is not None and "Disc CE Loss" in st and "Disc Rew Mean" in st and np.isfinite(st["Grad Pen"])
    assert irl.policy_trainer.engine.get_state().adam_step[4] == 21          # 21 discriminator updates
    eb = irl.get_batch(B, True, keys=["observations", "actions"])
    assert set(eb.keys()) == {"observations", "actions"}
