"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by EXECUTING THE UNMODIFIED
REFERENCE (/root/reference, through oracle/ref_shim.py) on the cases of oracle/configs.py.

    python -m oracle.make_golden            # (re)generates every case, and cross-checks
                                            # oracle/restate.py against the reference run

The reference nets are built by the reference's own constructors (FlattenMlp,
ReparamTanhMultivariateGaussianPolicy, MlpGaussianNoisePolicy, MLPDisc); only their
initial VALUES are overwritten with the seed-reproducible init of oracle/restate.py so the
fixture can stay small: a golden file holds the per-step losses/statistics and per-tensor
parameter digests, and everything else (buffers, indices, eps, initial params) is
regenerated from seeds by the tests, on any machine.

Step protocol (identical for the reference run and for every consumer of the goldens):
    before step t:  torch.manual_seed(EPS_SEED0 + t)
    SAC-alpha:      idx <- buffer RandomState.randint;  eps_next = randn(B,A); eps_cur = randn(B,A)
    SAC-V:          idx;  eps_cur = randn(B,A)
    TD3:            idx;  noise = randn(B,A)     (torch.normal(zeros) == randn stream, policies.py:182)
    AdvIRL iter:    idx_expert; idx_policy; gp_eps = rand(B,1) [only if use_grad_pen];
                    idx_policy2; eps_next; eps_cur          (adv_irl.py:126-131 RNG order)
"""
import os
import sys
from collections import OrderedDict

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import configs as C  # noqa: E402
from oracle import restate as R  # noqa: E402
from oracle import ref_shim  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_oracle_nets(case):
    """Seed-reproducible initial parameters (shared with tests/)."""
    rs = np.random.RandomState(case["seed"])
    O, A = case["obs_dim"], case["act_dim"]
    nets = OrderedDict()
    algo = case["algo"]
    HID = (case.get("hidden", C.HIDDEN[0]),) * 2
    DH = case.get("disc_hid", C.DISC_HID)
    if algo in ("sac_alpha", "sac_v", "adv_irl"):
        nets["qf1"] = R.Net(R.init_mlp(rs, O + A, HID, 1))
        nets["qf2"] = R.Net(R.init_mlp(rs, O + A, HID, 1))
        if algo == "sac_v":
            nets["vf"] = R.Net(R.init_mlp(rs, O, HID, 1))
        nets["policy"] = R.Net(R.init_mlp(rs, O, HID, A, init_w=1e-3, log_std_head=True))
        if algo == "adv_irl":
            nets["disc"] = R.Net(R.init_disc(rs, 2 * O if case.get("state_only") else O + A, DH))
    elif algo == "td3":
        nets["qf1"] = R.Net(R.init_mlp(rs, O + A, HID, 1))
        nets["qf2"] = R.Net(R.init_mlp(rs, O + A, HID, 1))
        nets["policy"] = R.Net(R.init_mlp(rs, O, HID, A, init_w=1e-3))
    # perturb so that nothing sits exactly on the tiny U(+-3e-3) last-layer init: makes
    # log_std clamps, relu masks and min(q1,q2) selections non-trivial in the fixtures
    for name, net in nets.items():
        if name == "disc":
            continue
        for k, v in net.p.items():
            if k.startswith("last_fc"):
                v.mul_(30.0)
    return nets


def build_buffers(case, cls):
    O, A = case["obs_dim"], case["act_dim"]
    term_p = 0.0 if case["algo"] == "adv_irl" else 0.01
    data = R.synth_transitions(case["n_fill"], O, A, C.DATA_SEED, term_p)
    buf = cls(case["n_fill"], O, A, random_seed=C.BUFFER_SEED)
    _bulk_load(buf, data)
    ebuf = None
    if case["algo"] == "adv_irl":
        edata = R.synth_transitions(case["n_expert"], O, A, C.EXPERT_DATA_SEED, 0.0)
        ebuf = cls(case["n_fill"], O, A, random_seed=C.EXPERT_SEED)
        _bulk_load(ebuf, edata)
    return buf, ebuf


def _bulk_load(buf, data):
    n = len(data["observations"])
    buf._observations[:n] = data["observations"]
    buf._actions[:n] = data["actions"]
    buf._rewards[:n] = data["rewards"]
    buf._terminals[:n] = data["terminals"]
    buf._next_obs[:n] = data["next_observations"]
    buf._top = n % buf._max_replay_buffer_size
    buf._size = n


def _load_into_module(module, net):
    with torch.no_grad():
        named = OrderedDict(module.named_parameters())
        assert list(named.keys()) == list(net.p.keys()), (list(named.keys()), list(net.p.keys()))
        for k, p in named.items():
            p.copy_(net.p[k])


def _module_flat(module):
    return np.concatenate([p.detach().numpy().ravel() for p in module.parameters()])


def _digest_module(module):
    out = []
    for p in module.parameters():
        a = p.detach().numpy().astype(np.float64).ravel()
        out.append(np.concatenate([[a.sum(), np.abs(a).sum()], a[:4] if a.size >= 4 else np.resize(a, 4)]))
    return np.stack(out)


# --------------------------------------------------------------------------------------
def run_reference(case):
    ref = ref_shim.import_reference()
    O, A, B = case["obs_dim"], case["act_dim"], case["batch"]
    nets = build_oracle_nets(case)
    buf, ebuf = build_buffers(case, ref.SimpleReplayBuffer)
    env = ref_shim.FakeEnv(O, A)
    algo = case["algo"]
    mods = OrderedDict()
    HID = [case.get("hidden", C.HIDDEN[0])] * 2
    DH = case.get("disc_hid", C.DISC_HID)

    def mk_q():
        return ref.FlattenMlp(hidden_sizes=list(HID), input_size=O + A, output_size=1)

    mods["qf1"], mods["qf2"] = mk_q(), mk_q()
    if algo == "sac_v":
        mods["vf"] = ref.FlattenMlp(hidden_sizes=list(HID), input_size=O, output_size=1)
    her = case.get("her")
    if her and algo == "sac_alpha":
        mods["policy"] = ref.ReparamTanhMultivariateGaussianConditionPolicy(
            hidden_sizes=list(HID), obs_dim=O - her["goal_dim"], condition_dim=her["goal_dim"], action_dim=A)
    elif her:
        mods["policy"] = ref.MlpGaussianAndEpsilonConditionPolicy(
            hidden_sizes=list(HID), action_space=ref_shim.FakeEnv(O, A).action_space, obs_dim=O - her["goal_dim"],
            condition_dim=her["goal_dim"], action_dim=A, output_activation=torch.tanh, max_sigma=her["sigma"],
            min_sigma=her["sigma"])
    elif algo == "td3":
        mods["policy"] = ref.MlpGaussianNoisePolicy(
            hidden_sizes=list(HID), obs_dim=O, action_dim=A, output_activation=torch.tanh,
            policy_noise=case["policy_noise"], policy_noise_clip=case["policy_noise_clip"])
    else:
        mods["policy"] = ref.ReparamTanhMultivariateGaussianPolicy(
            hidden_sizes=list(HID), obs_dim=O, action_dim=A)
    if algo == "adv_irl":
        mods["disc"] = ref.MLPDisc(2 * O if case.get("state_only") else O + A, num_layer_blocks=2, hid_dim=DH,
                                   hid_act=case.get("disc_act", "tanh"), use_bn=False, clamp_magnitude=10.0)
    for k, m in mods.items():
        _load_into_module(m, nets[k])

    if algo == "sac_alpha" and her:
        kw = {k: v for k, v in case["sac"].items() if k != "target_entropy"}     # her/sac.py takes none: -prod(action shape)
        trainer = ref.HerSAC(policy=mods["policy"], qf1=mods["qf1"], qf2=mods["qf2"], env=env, **kw)
        assert float(trainer.target_entropy) == case["sac"]["target_entropy"]
    elif algo in ("sac_alpha", "adv_irl"):
        trainer = ref.SacAlpha(policy=mods["policy"], qf1=mods["qf1"], qf2=mods["qf2"], env=env, **case["sac"])
    elif algo == "sac_v":
        trainer = ref.SacV(policy=mods["policy"], qf1=mods["qf1"], qf2=mods["qf2"], vf=mods["vf"], **case["sac"])
    elif her:
        trainer = ref.HerTD3(policy=mods["policy"], qf1=mods["qf1"], qf2=mods["qf2"], clip_return_l=her.get("clip_return_l"),
                             clip_return_r=her.get("clip_return_r"), **case["td3"])
    else:
        trainer = ref.TD3(policy=mods["policy"], qf1=mods["qf1"], qf2=mods["qf2"], **case["td3"])

    alg = None
    if algo == "adv_irl":
        alg = ref.AdvIRL(
            mode=case["mode"], discriminator=mods["disc"], policy_trainer=trainer,
            expert_replay_buffer=ebuf, state_only=case.get("state_only", False), disc_optim_batch_size=B,
            policy_optim_batch_size_from_expert=case.get("from_expert", 0),
            policy_optim_batch_size=B, num_update_loops_per_train_call=1,
            num_disc_updates_per_loop_iter=case.get("n_disc", 1), num_policy_updates_per_loop_iter=case.get("n_policy", 1),
            rew_clip_min=case.get("rew_clip_min"), rew_clip_max=case.get("rew_clip_max"),
            env=env, exploration_policy=mods["policy"], training_env=None, replay_buffer=buf,
            max_path_length=min(100, case["n_fill"] - 1), no_terminal=True, **case["disc"])

    rows = []
    for t in range(case["steps"]):
        torch.manual_seed(C.EPS_SEED0 + t)
        row = OrderedDict()
        if algo == "adv_irl":
            trainer.end_epoch()
            alg.disc_eval_statistics = None
            alg._do_training(0)
            st = dict(alg.disc_eval_statistics)
            st.update(trainer.get_eval_statistics())
            for k in ["Disc CE Loss", "Disc Acc", "Grad Pen", "Disc Rew Mean", "Disc Rew Std",
                      "Disc Rew Max", "Disc Rew Min", "QF1 Loss", "QF2 Loss", "Policy Loss",
                      "Alpha Loss", "Alpha Mean", "Q1 Predictions Mean", "Log Pis Mean"]:
                if k in st:
                    row[k] = float(st[k])
        else:
            trainer.end_epoch()
            batch = ref.np_to_pytorch_batch(buf.random_batch(B))
            if her:     # the relabel buffer hands goals over separately; the trainer concatenates them (her/td3.py:94-98)
                G_ = her["goal_dim"]
                batch["desired_goals"] = batch["observations"][:, O - G_:]
                batch["next_desired_goals"] = batch["next_observations"][:, O - G_:]
                batch["observations"] = batch["observations"][:, :O - G_]
                batch["next_observations"] = batch["next_observations"][:, :O - G_]
            trainer.train_step(batch)
            st = trainer.get_eval_statistics()
            for k in ["QF1 Loss", "QF2 Loss", "VF Loss", "Policy Loss", "Alpha Loss", "Alpha Mean",
                      "Q1 Predictions Mean", "Q1 Predictions Std", "Q2 Predictions Mean",
                      "Log Pis Mean", "Policy mu Mean", "Policy log std Mean", "Q Targets Mean",
                      "Policy Action Mean"]:
                if k in st:
                    row[k] = float(st[k])
        rows.append(row)

    final = OrderedDict((k, _module_flat(m)) for k, m in mods.items())
    if algo in ("sac_alpha", "adv_irl"):
        final["target_qf1"] = _module_flat(trainer.target_qf1)
        final["target_qf2"] = _module_flat(trainer.target_qf2)
        final["log_alpha"] = np.array([float(trainer.log_alpha.detach())])
    elif algo == "sac_v":
        final["target_vf"] = _module_flat(trainer.target_vf)
    else:
        final["target_policy"] = _module_flat(trainer.target_policy)
        final["target_qf1"] = _module_flat(trainer.target_qf1)
        final["target_qf2"] = _module_flat(trainer.target_qf2)
    digests = OrderedDict((k, _digest_module(m)) for k, m in mods.items())
    return rows, final, digests


# --------------------------------------------------------------------------------------
def run_oracle(case):
    """The same protocol through oracle/restate.py (no reference import)."""
    O, A, B = case["obs_dim"], case["act_dim"], case["batch"]
    nets = build_oracle_nets(case)
    buf, ebuf = build_buffers(case, R.ReplayOracle)
    algo = case["algo"]
    if algo in ("sac_alpha", "adv_irl"):
        tr = R.SacAlphaOracle(nets["policy"], nets["qf1"], nets["qf2"], A, **case["sac"])
    elif algo == "sac_v":
        tr = R.SacVOracle(nets["policy"], nets["qf1"], nets["qf2"], nets["vf"], **case["sac"])
    elif case.get("her"):
        h = case["her"]
        tr = R.HerTD3Oracle(nets["policy"], nets["qf1"], nets["qf2"], sigma=h["sigma"], clip_return_l=h.get("clip_return_l"),
                            clip_return_r=h.get("clip_return_r"), **case["td3"])
    else:
        tr = R.TD3Oracle(nets["policy"], nets["qf1"], nets["qf2"], policy_noise=case["policy_noise"],
                         policy_noise_clip=case["policy_noise_clip"], **case["td3"])
    disc = R.DiscOracle(nets["disc"], hid_act=case.get("disc_act", "tanh"), **case["disc"]) if algo == "adv_irl" else None
    rows = []
    for t in range(case["steps"]):
        torch.manual_seed(C.EPS_SEED0 + t)
        row = OrderedDict()
        if algo == "adv_irl":
            # adv_irl.py:126-131: n_disc reward updates, then n_policy policy updates; the logged statistics are those of
            # the FIRST update of each kind in the call (eval_statistics is filled once), reward stats of the LAST (:303-314)
            d = s = None
            second = "next_observations" if case.get("state_only") else "actions"     # adv_irl.py:139-143
            E = case.get("from_expert", 0)
            for _ in range(case.get("n_disc", 1)):
                eb = R.np_to_torch_batch(ebuf.random_batch(B, keys=["observations", second]))
                pb = R.np_to_torch_batch(buf.random_batch(B, keys=["observations", second]))
                gp_eps = torch.rand(B, 1) if case["disc"]["use_grad_pen"] else None
                di = disc.reward_step(torch.cat([eb["observations"], eb[second]], 1),
                                      torch.cat([pb["observations"], pb[second]], 1), gp_eps)
                d = d or di
            for _ in range(case.get("n_policy", 1)):
                if E:       # adv_irl.py:239-255: B - E rows from the policy buffer, then E rows from the expert buffer
                    bp = R.np_to_torch_batch(buf.random_batch(B - E))
                    be = R.np_to_torch_batch(ebuf.random_batch(E))
                    batch = {k: torch.cat([bp[k], be[k]], dim=0) for k in bp}
                else:
                    batch = R.np_to_torch_batch(buf.random_batch(B))
                rew = disc.rewards(batch["observations"], batch[second], case["mode"],
                                   case.get("rew_clip_min"), case.get("rew_clip_max"))
                batch["rewards"] = rew
                si = tr.train_step(batch, torch.randn(B, A), torch.randn(B, A))
                s = s or si
            rn = rew.numpy()
            row.update({"Disc CE Loss": d["disc_ce_loss"], "Disc Acc": d["disc_acc"]})
            if case["disc"]["use_grad_pen"]:
                row["Grad Pen"] = d["grad_pen"]
            row.update({"Disc Rew Mean": float(np.mean(rn)), "Disc Rew Std": float(np.std(rn)),
                        "Disc Rew Max": float(np.max(rn)), "Disc Rew Min": float(np.min(rn)),
                        "QF1 Loss": s["qf1_loss"], "QF2 Loss": s["qf2_loss"],
                        "Policy Loss": s["policy_loss"], "Alpha Loss": s["alpha_loss"],
                        "Alpha Mean": s["alpha"], "Q1 Predictions Mean": float(s["q1_pred"].mean()),
                        "Log Pis Mean": float(s["log_pi"].mean())})
        elif algo == "sac_alpha":
            batch = R.np_to_torch_batch(buf.random_batch(B))
            s = tr.train_step(batch, torch.randn(B, A), torch.randn(B, A))
            row.update({"QF1 Loss": s["qf1_loss"], "QF2 Loss": s["qf2_loss"],
                        "Policy Loss": s["policy_loss"]})
            if s["alpha_loss"] is not None:
                row["Alpha Loss"] = s["alpha_loss"]
            row.update({"Alpha Mean": s["alpha"], "Q1 Predictions Mean": float(s["q1_pred"].mean()),
                        "Q1 Predictions Std": float(s["q1_pred"].std()),
                        "Q2 Predictions Mean": float(s["q2_pred"].mean()),
                        "Log Pis Mean": float(s["log_pi"].mean()),
                        "Policy mu Mean": float(s["policy_mean"].mean()),
                        "Policy log std Mean": float(s["policy_log_std"].mean())})
        elif algo == "sac_v":
            batch = R.np_to_torch_batch(buf.random_batch(B))
            s = tr.train_step(batch, torch.randn(B, A))
            row.update({"QF1 Loss": s["qf1_loss"], "QF2 Loss": s["qf2_loss"],
                        "VF Loss": s["vf_loss"], "Policy Loss": s["policy_loss"]})
        else:
            batch = R.np_to_torch_batch(buf.random_batch(B))
            s = tr.train_step(batch, torch.randn(B, A))
            row.update({"QF1 Loss": s["qf1_loss"], "QF2 Loss": s["qf2_loss"]})
            if s["policy_loss"] is not None:
                row["Policy Loss"] = s["policy_loss"]
            if "stats_policy_loss" in s:
                row["Stats Policy Loss"] = s["stats_policy_loss"]     # td3.py:131-136 (what the epoch log shows)
            row.update({"Q1 Predictions Mean": float(s["q1_pred"].mean()),
                        "Q Targets Mean": float(s["q_target"].mean())})
        rows.append(row)
    final = OrderedDict((k, n.flat()) for k, n in nets.items())
    if algo in ("sac_alpha", "adv_irl"):
        final["target_qf1"], final["target_qf2"] = tr.target_qf1.flat(), tr.target_qf2.flat()
        final["log_alpha"] = np.array([tr.log_alpha])
    elif algo == "sac_v":
        final["target_vf"] = tr.target_vf.flat()
    else:
        final["target_policy"] = tr.target_policy.flat()
        final["target_qf1"], final["target_qf2"] = tr.target_qf1.flat(), tr.target_qf2.flat()
    digests = OrderedDict((k, R.param_digest(n)) for k, n in nets.items())
    return rows, final, digests


def compare_rows(ref_rows, ora_rows, rtol):
    worst = 0.0
    for t, (a, b) in enumerate(zip(ref_rows, ora_rows)):
        for k, va in a.items():
            if k not in b or b[k] is None:
                # TD3 logs a stats-only policy loss on odd steps (td3.py:131-136); skip
                continue
            err = abs(va - b[k]) / max(abs(va), 1e-12)
            worst = max(worst, err)
            assert err <= rtol or abs(va - b[k]) < 1e-9, (t, k, va, b[k], err)
    return worst


def main():
    torch.set_num_threads(1)
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    all_cases = dict(C.CASES, **C.LOOP_CASES)
    names = sys.argv[1:] or list(all_cases.keys())
    for name in names:
        case = all_cases[name]
        ref_rows, ref_final, ref_dig = run_reference(case)
        ora_rows, ora_final, ora_dig = run_oracle(case)
        worst = compare_rows(ref_rows, ora_rows, 2e-5)
        pmax = max(float(np.max(np.abs(ref_final[k] - ora_final[k]))) for k in ref_final)
        print("%-30s steps=%d  worst stat rel err oracle-vs-reference %.2e, params max abs diff %.2e"
              % (name, case["steps"], worst, pmax))
        assert pmax < 2e-6, (name, pmax)
        keys = sorted({k for r in ref_rows for k in r})
        table = np.full((len(ref_rows), len(keys)), np.nan)
        for t, r in enumerate(ref_rows):
            for j, k in enumerate(keys):
                if k in r:
                    table[t, j] = r[k]
        out = dict(stat_keys=np.array(keys), stats=table)
        for k, v in ref_dig.items():
            out["digest_" + k] = v
        for k, v in ref_final.items():  # small strided sample of every final tensor
            out["sample_" + k] = v[:: max(1, v.size // 256)][:256].astype(np.float64)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)


if __name__ == "__main__":
    main()
