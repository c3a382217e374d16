"""Shared test helpers: seed-reproducible inputs for the parity cases (oracle/configs.py)
following the step protocol documented in oracle/make_golden.py."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import configs as CFG  # noqa: E402
from oracle import make_golden as G  # noqa: E402
from oracle import restate as R  # noqa: E402
from ilswiss_b200 import _abi, layout  # noqa: E402

STAT_TO_SLOT = {
    "QF1 Loss": _abi.L_QF1, "QF2 Loss": _abi.L_QF2, "Policy Loss": _abi.L_POLICY,
    "Alpha Loss": _abi.L_ALPHA_LOSS, "Alpha Mean": _abi.L_ALPHA, "Disc CE Loss": _abi.L_DISC_CE,
    "Disc Acc": _abi.L_DISC_ACC, "Grad Pen": _abi.L_GRAD_PEN, "Disc Rew Mean": _abi.L_REW_MEAN,
    "Disc Rew Std": _abi.L_REW_STD, "Disc Rew Max": _abi.L_REW_MAX, "Disc Rew Min": _abi.L_REW_MIN,
    "Q1 Predictions Mean": _abi.L_Q1_MEAN, "Log Pis Mean": _abi.L_LOGPI_MEAN,
    "Q Targets Mean": _abi.L_QT_MEAN, "VF Loss": _abi.L_VF,
}


def case_data(case):
    O, A = case["obs_dim"], case["act_dim"]
    term_p = 0.0 if case["algo"] == "adv_irl" else 0.01
    data = R.synth_transitions(case["n_fill"], O, A, CFG.DATA_SEED, term_p)
    edata = None
    if case["algo"] == "adv_irl":
        edata = R.synth_transitions(case["n_expert"], O, A, CFG.EXPERT_DATA_SEED, 0.0)
    return data, edata


def case_injection(case, steps=None):
    """idx / eps streams exactly as the reference run consumed them."""
    B, A = case["batch"], case["act_dim"]
    T = steps or case["steps"]
    rs = np.random.RandomState(CFG.BUFFER_SEED)
    ers = np.random.RandomState(CFG.EXPERT_SEED)
    algo = case["algo"]
    out = dict(idx=np.zeros((T, B), np.int32), eps_next=np.zeros((T, B, A), np.float32),
               eps_cur=np.zeros((T, B, A), np.float32), idx_expert=np.zeros((T, B), np.int32),
               idx_policy_d=np.zeros((T, B), np.int32), gp_eps=np.zeros((T, B), np.float32))
    for t in range(T):
        torch.manual_seed(CFG.EPS_SEED0 + t)
        if algo == "adv_irl":
            out["idx_expert"][t] = ers.randint(0, case["n_expert"], B)
            out["idx_policy_d"][t] = rs.randint(0, case["n_fill"], B)
            if case["disc"]["use_grad_pen"]:
                out["gp_eps"][t] = torch.rand(B, 1).numpy().ravel()
        E = case.get("from_expert", 0)
        out["idx"][t, :B - E] = rs.randint(0, case["n_fill"], B - E)
        if E:       # adv_irl.py:239-255: policy-buffer draw first, then the expert-buffer draw
            out["idx"][t, B - E:] = ers.randint(0, case["n_expert"], E)
        out["eps_next"][t] = torch.randn(B, A).numpy()
        if algo in ("sac_alpha", "adv_irl"):
            out["eps_cur"][t] = torch.randn(B, A).numpy()
        elif algo == "sac_v":
            out["eps_cur"][t] = out["eps_next"][t]
    return out


def loop_case_injection(case):
    """Injection streams of a LOOP_CASES case (n_disc / n_policy updates per _do_training call), in the order the
    reference consumes its RNGs: per call t (torch.manual_seed(EPS_SEED0 + t)): n_disc x [expert randint, policy
    randint, torch.rand(B,1)], then n_policy x [policy randint, randn(B,A), randn(B,A)].  Disc arrays have
    steps*n_disc rows, policy arrays steps*n_policy rows."""
    B, A, T = case["batch"], case["act_dim"], case["steps"]
    nd, npol = case["n_disc"], case["n_policy"]
    rs = np.random.RandomState(CFG.BUFFER_SEED)
    ers = np.random.RandomState(CFG.EXPERT_SEED)
    out = dict(idx=np.zeros((T * npol, B), np.int32), eps_next=np.zeros((T * npol, B, A), np.float32),
               eps_cur=np.zeros((T * npol, B, A), np.float32), idx_expert=np.zeros((T * nd, B), np.int32),
               idx_policy_d=np.zeros((T * nd, B), np.int32), gp_eps=np.zeros((T * nd, B), np.float32))
    for t in range(T):
        torch.manual_seed(CFG.EPS_SEED0 + t)
        for i in range(nd):
            out["idx_expert"][t * nd + i] = ers.randint(0, case["n_expert"], B)
            out["idx_policy_d"][t * nd + i] = rs.randint(0, case["n_fill"], B)
            if case["disc"]["use_grad_pen"]:
                out["gp_eps"][t * nd + i] = torch.rand(B, 1).numpy().ravel()
        for j in range(npol):
            out["idx"][t * npol + j] = rs.randint(0, case["n_fill"], B)
            out["eps_next"][t * npol + j] = torch.randn(B, A).numpy()
            out["eps_cur"][t * npol + j] = torch.randn(B, A).numpy()
    return out


def her_relabel_setup(case):
    """A HER_RELABEL_CASES case: the oracle hindsight buffer filled with synthetic goal episodes, the device-side
    views of it (ring rows with obs = cat(observation, desired_goal), next-achieved-goal side array, trajectory table) and,
    per step, the index draws + the relabelled oracle batch (networks see cat(obs, goal), her/td3.py:94-98)."""
    O, A, B, G_ = case["obs_dim"], case["act_dim"], case["batch"], case["her"]["goal_dim"]
    N, T = case["n_fill"], case["steps"]
    ora = R.HindsightOracle(N, O - G_, G_, A, random_seed=CFG.BUFFER_SEED, her_ratio=case["her_ratio"],
                            distance_threshold=case["threshold"])
    rs = np.random.RandomState(case["seed"])
    for ep in R.synth_goal_episodes(rs, case["n_episodes"], case["T"], O - G_, G_, A):
        for (o, a, r, d, no) in ep:
            ora.add_sample(o, a, r, d, no)
        ora.terminate_episode()
    cat = lambda d: np.concatenate([d["observation"], d["desired_goal"]], axis=1)
    ring = layout.pack_hot_rows(cat(ora._observations), ora._actions, ora._rewards, ora._terminals, cat(ora._next_obs))
    starts = np.array(list(ora._traj_endpoints.keys()), dtype=np.int32)
    lens = np.array([(ora._traj_endpoints[int(s)] - int(s)) % ora._size for s in starts], dtype=np.int32)
    out = dict(oracle=ora, ring=ring, ring_size=ora._size, ag_next=np.ascontiguousarray(ora._next_obs["achieved_goal"], dtype=np.float32),
               traj_start=starts, traj_len=lens, relabel_num=int(case["her_ratio"] * B), threshold=case["threshold"],
               idx=np.zeros((T, B), np.int32), idx_her=np.zeros((T, B), np.int32),
               eps_next=np.zeros((T, B, A), np.float32), eps_cur=np.zeros((T, B, A), np.float32), batches=[])
    for t in range(T):
        np.random.seed(CFG.EPS_SEED0 + t)               # the future-step draw uses the GLOBAL numpy RNG (:88)
        torch.manual_seed(CFG.EPS_SEED0 + t)
        i, ih = ora.sample_indices(B)
        out["idx"][t], out["idx_her"][t] = i, ih
        b = ora.batch_from_indices(i, ih)
        out["batches"].append(dict(
            observations=np.concatenate([b["observations"], b["desired_goals"]], axis=1),
            next_observations=np.concatenate([b["next_observations"], b["next_desired_goals"]], axis=1),
            actions=b["actions"], rewards=b["rewards"], terminals=b["terminals"]))
        out["eps_next"][t] = torch.randn(B, A).numpy()
        out["eps_cur"][t] = torch.randn(B, A).numpy()
    return out


def her_oracle_rows(case, setup):
    """HerTD3Oracle driven by the relabelled oracle batches: per-step statistics + final parameters."""
    nets = G.build_oracle_nets(case)
    h = case["her"]
    if case["algo"] == "sac_alpha":          # her/sac.py == sac_alpha on cat(obs, goal) with target entropy -A
        tr = R.SacAlphaOracle(nets["policy"], nets["qf1"], nets["qf2"], case["act_dim"], **case["sac"])
        rows = []
        for t, b in enumerate(setup["batches"]):
            s = tr.train_step(R.np_to_torch_batch(b), torch.from_numpy(setup["eps_next"][t]), torch.from_numpy(setup["eps_cur"][t]))
            rows.append({"QF1 Loss": s["qf1_loss"], "QF2 Loss": s["qf2_loss"], "Policy Loss": s["policy_loss"],
                         "Alpha Loss": s["alpha_loss"]})
        final = {k: n.flat() for k, n in nets.items()}
        final["target_qf1"] = tr.target_qf1.flat()
        return rows, final
    tr = R.HerTD3Oracle(nets["policy"], nets["qf1"], nets["qf2"], sigma=h["sigma"], clip_return_l=h.get("clip_return_l"),
                        clip_return_r=h.get("clip_return_r"), **case["td3"])
    rows = []
    for t, b in enumerate(setup["batches"]):
        s = tr.train_step(R.np_to_torch_batch(b), torch.from_numpy(setup["eps_next"][t]))
        row = {"QF1 Loss": s["qf1_loss"], "QF2 Loss": s["qf2_loss"], "Q Targets Mean": float(s["q_target"].mean())}
        if s["policy_loss"] is not None:
            row["Policy Loss"] = s["policy_loss"]
        rows.append(row)
    final = {k: n.flat() for k, n in nets.items()}
    final["target_qf1"], final["target_policy"] = tr.target_qf1.flat(), tr.target_policy.flat()
    return rows, final


def her_desc(setup, ptr, inj_offset=0, n=None):
    """_abi.HerSamplingDesc over arrays / tensors addressed through `ptr` (np_ptr on the host, data_ptr on the device)."""
    d = _abi.HerSamplingDesc()
    d.enabled, d.n_traj = 1, len(setup["traj_start"])
    d.traj_start, d.traj_len, d.next_achieved_goal = ptr("traj_start"), ptr("traj_len"), ptr("ag_next")
    d.goal_dim, d.relabel_num, d.distance_threshold = setup["ag_next"].shape[1], setup["relabel_num"], setup["threshold"]
    d.inj_idx_her = ptr("idx_her")
    return d


DISC_STATS = ("Disc CE Loss", "Disc Acc", "Grad Pen")
REW_STATS = ("Disc Rew Mean", "Disc Rew Std", "Disc Rew Max", "Disc Rew Min")


def run_loop_case(run, case, set_mode):
    """Drives a HostSimRun / DeviceRun through a LOOP_CASES case with alternating disc-only / policy-only launches
    (set_mode(1|2|0)) and returns one statistics row per _do_training call with the reference's logging rules: first
    disc update, first policy update, reward statistics of the last policy update."""
    inj = loop_case_injection(case)
    nd, npol = case["n_disc"], case["n_policy"]
    disc_keys, pol_keys = ("idx_expert", "idx_policy_d", "gp_eps"), ("idx", "eps_next", "eps_cur")
    rows = []
    for t in range(case["steps"]):
        set_mode(1)
        Ld = run.train(nd, {k: inj[k] for k in disc_keys}, t_offset=t * nd)
        set_mode(2)
        Lp = run.train(npol, {k: inj[k] for k in pol_keys}, t_offset=t * npol)
        row = {}
        for k, slot in STAT_TO_SLOT.items():
            if k in DISC_STATS:
                row[k] = float(Ld[0, slot])
            elif k in REW_STATS:
                row[k] = float(Lp[-1, slot])
            else:
                row[k] = float(Lp[0, slot])
        rows.append(row)
    set_mode(0)
    return rows


def trainer_config(case, max_steps=64, precision=0):
    algo = case["algo"]
    cfg = _abi.TrainerConfig()
    cfg.gemm_precision = precision
    cfg.obs_dim, cfg.act_dim, cfg.batch = case["obs_dim"], case["act_dim"], case["batch"]
    cfg.max_steps_per_call = max_steps
    cfg.beta_2, cfg.adam_eps = 0.999, 1e-8
    cfg.max_act = 1.0
    if algo in ("sac_alpha", "adv_irl"):
        kw = case["sac"]
        cfg.algo = _abi.ALGO_SAC_ALPHA
        cfg.reward_scale, cfg.discount = kw["reward_scale"], kw["discount"]
        cfg.soft_target_tau = kw["soft_target_tau"]
        cfg.policy_lr, cfg.qf_lr = kw["policy_lr"], kw["qf_lr"]
        cfg.alpha_lr = kw.get("alpha_lr", 3e-4)
        cfg.beta_1 = kw.get("beta_1", 0.9)
        cfg.alpha = kw.get("alpha", 0.2)
        cfg.train_alpha = int(kw.get("train_alpha", True))
        te = kw.get("target_entropy")
        cfg.target_entropy = (-case["act_dim"] / 2.0) if te is None else te
        cfg.policy_mean_reg_weight = kw["policy_mean_reg_weight"]
        cfg.policy_std_reg_weight = kw["policy_std_reg_weight"]
    elif algo == "td3":
        kw = case["td3"]
        cfg.algo = _abi.ALGO_TD3
        cfg.reward_scale, cfg.discount = kw["reward_scale"], kw["discount"]
        cfg.soft_target_tau = kw["soft_target_tau"]
        cfg.policy_lr, cfg.qf_lr = kw["policy_lr"], kw["qf_lr"]
        cfg.beta_1 = 0.9
        cfg.alpha = 1.0
        cfg.policy_and_target_update_period = kw["policy_and_target_update_period"]
        cfg.policy_noise, cfg.policy_noise_clip = case["policy_noise"], case["policy_noise_clip"]
        if case.get("her"):
            h = case["her"]
            cfg.her, cfg.her_sigma, cfg.min_act = 1, h["sigma"], -1.0
            cl, cr = h.get("clip_return_l"), h.get("clip_return_r")
            cfg.clip_return_l = -1.0 / (1.0 - kw["discount"]) if cl is None else cl
            cfg.clip_return_r = 0.0 if cr is None else cr
    elif algo == "sac_v":
        kw = case["sac"]
        cfg.algo = _abi.ALGO_SAC_V
        cfg.reward_scale, cfg.discount = kw["reward_scale"], kw["discount"]
        cfg.soft_target_tau = kw["soft_target_tau"]
        cfg.policy_lr, cfg.qf_lr, cfg.vf_lr = kw["policy_lr"], kw["qf_lr"], kw["vf_lr"]
        cfg.beta_1 = kw.get("beta_1", 0.9)
        cfg.alpha = kw.get("alpha", 1.0)
        cfg.train_alpha = 0
        cfg.policy_mean_reg_weight = kw["policy_mean_reg_weight"]
        cfg.policy_std_reg_weight = kw["policy_std_reg_weight"]
    else:
        raise NotImplementedError(algo)
    return cfg


def disc_config(case):
    d = case["disc"]
    dc = _abi.DiscConfig()
    dc.mode = _abi.DISC_MODES[case["mode"]]
    dc.batch = case["batch"]
    dc.disc_lr, dc.disc_momentum = d["disc_lr"], d["disc_momentum"]
    dc.use_grad_pen, dc.grad_pen_weight = int(d["use_grad_pen"]), d["grad_pen_weight"]
    dc.clamp_magnitude = 10.0
    dc.hid_act = _abi.DISC_ACTS[case.get("disc_act", "tanh")]
    dc.rew_clip_min_on = int(case.get("rew_clip_min") is not None)
    dc.rew_clip_max_on = int(case.get("rew_clip_max") is not None)
    dc.rew_clip_min = case.get("rew_clip_min") or 0.0
    dc.rew_clip_max = case.get("rew_clip_max") or 0.0
    dc.state_only, dc.policy_batch_from_expert = int(case.get("state_only", False)), case.get("from_expert", 0)
    return dc


def net_order(case):
    """Order of networks expected by ilsw_trainer_create."""
    if case["algo"] == "td3":
        return ["policy", "qf1", "qf2", "target_qf1", "target_qf2", "target_policy"]
    if case["algo"] == "sac_v":
        return ["policy", "qf1", "qf2", "vf", "target_vf"]
    return ["policy", "qf1", "qf2", "target_qf1", "target_qf2"]


def initial_arenas(case):
    """name -> float32 flat parameter arena (targets are copies, like PyTorchModule.copy())."""
    nets = G.build_oracle_nets(case)
    flat = {k: n.flat().astype(np.float32) for k, n in nets.items()}
    flat["target_qf1"], flat["target_qf2"] = flat["qf1"].copy(), flat["qf2"].copy()
    if case["algo"] == "sac_v":
        flat["target_vf"] = flat["vf"].copy()
    if case["algo"] == "td3":
        flat["target_policy"] = flat["policy"].copy()
    return flat


def mlp_dims(case, name):
    O, A = case["obs_dim"], case["act_dim"]
    H = case.get("hidden", CFG.HIDDEN[0])
    if name in ("policy", "target_policy"):
        return O, H, A, int(case["algo"] != "td3")
    if name == "disc":
        return (2 * O if case.get("state_only") else O + A), case.get("disc_hid", CFG.DISC_HID), 1, 0
    if name in ("vf", "target_vf"):
        return O, H, 1, 0
    return O + A, H, 1, 0


# ---------------------------------------------------------------------------------------------
# host simulator (tests/hostsim) loader
# ---------------------------------------------------------------------------------------------
def load_hostsim():
    d = os.path.join(ROOT, "tests", "hostsim")
    so = os.path.join(d, "libilsw_hostsim.so")
    srcs = [os.path.join(d, "hostsim.cpp")] + [
        os.path.join(ROOT, "ilswiss_b200", "csrc", f) for f in ("ilsw_ops.cuh", "ilsw_rows_fast.cuh", "ilsw_program.h", "ilsw_types.h")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so, srcs[0]])
    lib = C.CDLL(so)
    lib.hs_create.restype = C.c_void_p
    lib.hs_create.argtypes = [C.POINTER(_abi.TrainerConfig), C.POINTER(_abi.Mlp), C.c_int,
                              C.POINTER(_abi.DiscConfig), C.POINTER(_abi.Mlp), C.c_char_p, C.c_int]
    lib.hs_destroy.argtypes = [C.c_void_p]
    lib.hs_train.restype = C.c_int
    lib.hs_train.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                             C.POINTER(_abi.Inject), C.POINTER(_abi.Batch), C.c_uint64, C.c_int]
    lib.hs_losses.restype = C.POINTER(C.c_float)
    lib.hs_losses.argtypes = [C.c_void_p]
    lib.hs_stats.restype = C.POINTER(C.c_float)
    lib.hs_stats.argtypes = [C.c_void_p]
    lib.hs_stats_floats.restype = C.c_int
    lib.hs_stats_floats.argtypes = [C.c_void_p]
    lib.hs_log_alpha.restype = C.c_double
    lib.hs_log_alpha.argtypes = [C.c_void_p]
    lib.hs_num_phases.restype = C.c_int
    lib.hs_num_phases.argtypes = [C.c_void_p]
    lib.hs_describe.restype = C.c_int
    lib.hs_describe.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.hs_set_update_mode.argtypes = [C.c_void_p, C.c_int]
    lib.hs_set_world.argtypes = [C.c_void_p, C.c_int]
    lib.hs_set_her.argtypes = [C.c_void_p, C.POINTER(_abi.HerSamplingDesc)]
    lib.hs_grad.restype = C.POINTER(C.c_float)
    lib.hs_grad.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    return lib


def np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class HostSimRun:
    """Runs one parity case through the host simulator with injected randomness."""

    def __init__(self, lib, case, max_steps=64, precision=0):
        self.lib, self.case = lib, case
        self.arenas = initial_arenas(case)
        self.moms = {}
        names = net_order(case)
        mlps = (_abi.Mlp * len(names))()
        for i, n in enumerate(names):
            i_d, h_d, o_d, ls = mlp_dims(case, n)
            a = self.arenas[n]
            assert a.size == layout.mlp_num_params(i_d, h_d, o_d, ls), (n, a.size)
            m, v = np.zeros_like(a), np.zeros_like(a)
            self.moms[n] = (m, v)
            mlps[i] = _abi.Mlp(np_ptr(a), np_ptr(m), np_ptr(v), i_d, h_d, o_d, ls)
        cfg = trainer_config(case, max_steps, precision)
        dcfg_p, disc_p = None, None
        if case["algo"] == "adv_irl":
            self.dcfg = disc_config(case)
            i_d, h_d, o_d, ls = mlp_dims(case, "disc")
            a = self.arenas["disc"]
            m, v = np.zeros_like(a), np.zeros_like(a)
            self.moms["disc"] = (m, v)
            self.disc = _abi.Mlp(np_ptr(a), np_ptr(m), np_ptr(v), i_d, h_d, o_d, ls)
            dcfg_p, disc_p = C.byref(self.dcfg), C.byref(self.disc)
        err = C.create_string_buffer(256)
        self.h = lib.hs_create(C.byref(cfg), mlps, len(names), dcfg_p, disc_p, err, 256)
        assert self.h, err.value
        data, edata = case_data(case)
        self.ring = layout.pack_hot_rows(**data)
        self.ering = layout.pack_hot_rows(**edata) if edata is not None else None

    def train(self, n_steps, inj, stats_step=-1, t_offset=0, seed=0):
        keep = {k: np.ascontiguousarray(v[t_offset:t_offset + n_steps]) for k, v in inj.items()}
        ptr = lambda k: np_ptr(keep[k]) if k in keep else None       # split launches inject only their own streams
        ij = _abi.Inject(ptr("idx"), ptr("eps_next"), ptr("eps_cur"), ptr("idx_expert"), ptr("idx_policy_d"), ptr("gp_eps"))
        er = self.ering
        rc = self.lib.hs_train(self.h, np_ptr(self.ring), self.ring.shape[1], self.ring.shape[0],
                               np_ptr(er) if er is not None else None, er.shape[1] if er is not None else 0,
                               er.shape[0] if er is not None else 0, n_steps, C.byref(ij), None, seed, stats_step)
        assert rc == 0
        L = np.ctypeslib.as_array(self.lib.hs_losses(self.h), shape=(n_steps, _abi.LOSS_SLOTS)).copy()
        return L

    def close(self):
        self.lib.hs_destroy(self.h)


def assert_params_close(got, ref, steps, lr=3e-4, msg="", frac=5e-4):
    """Parameter parity bar.  <=1e-5 abs for all but a vanishing fraction of
    elements: Adam's normalised update m/(sqrt(v)+eps) is +-lr in the first steps REGARDLESS of
    |g|, so an element whose true gradient is ~0 (|g| ~ 1e-9, pure summation-order noise) can
    legitimately differ by up to 2*lr per step between two fp32 implementations.  Those
    elements are bounded by 2*lr*steps and must be rarer than 5e-4 of the tensor."""
    got = np.asarray(got, dtype=np.float64).ravel()
    ref = np.asarray(ref, dtype=np.float64).ravel()
    assert got.shape == ref.shape, (msg, got.shape, ref.shape)
    diff = np.abs(got - ref)
    bad = int((diff > 1e-5).sum())
    assert bad <= max(2, int(frac * diff.size)), (msg, bad, diff.size, float(diff.max()))
    assert float(diff.max()) <= 2.0 * lr * steps + 1e-6, (msg, float(diff.max()))


# ---------------------------------------------------------------------------------------------
# device runner (the product path, through the C ABI)
# ---------------------------------------------------------------------------------------------
class DeviceRun:
    """Runs one parity case on the GPU through ilswiss_b200.engine (C ABI) with injected randomness."""

    def __init__(self, case, max_steps=64, precision=0):
        from ilswiss_b200 import engine

        self.case = case
        self.engine_mod = engine
        arenas = initial_arenas(case)
        self.nets = {}
        order = net_order(case)
        for n in order:
            i_d, h_d, o_d, ls = mlp_dims(case, n)
            self.nets[n] = engine.NetArena(i_d, h_d, o_d, ls, trainable=not n.startswith("target"), init=arenas[n])
        dcfg = dnet = None
        if case["algo"] == "adv_irl":
            i_d, h_d, o_d, ls = mlp_dims(case, "disc")
            dnet = engine.NetArena(i_d, h_d, o_d, ls, init=arenas["disc"])
            self.nets["disc"] = dnet
            dcfg = disc_config(case)
        self.eng = engine.StepEngine(trainer_config(case, max_steps, precision), [self.nets[n] for n in order], dcfg, dnet)
        data, edata = case_data(case)
        O, A = case["obs_dim"], case["act_dim"]
        self.ring = engine.ReplayRing(case["n_fill"], O, A)
        self.ring.load_device(torch.from_numpy(layout.pack_hot_rows(**data)).cuda())
        self.ering = None
        if edata is not None:
            self.ering = engine.ReplayRing(case["n_fill"], O, A)
            self.ering.load_device(torch.from_numpy(layout.pack_hot_rows(**edata)).cuda())

    def train(self, n_steps, inj, stats_step=-1, t_offset=0, seed=0):
        dev = {k: torch.from_numpy(np.ascontiguousarray(v[t_offset:t_offset + n_steps])).cuda() for k, v in inj.items()}
        self.eng.train(self.ring, n_steps, expert_ring=self.ering, inject=dev, stats_step=stats_step, seed=seed)
        return self.eng.losses(n_steps)

    def train_philox(self, n_steps, seed=1, stats_step=-1):
        self.eng.train(self.ring, n_steps, expert_ring=self.ering, seed=seed, stats_step=stats_step)
        return self.eng.losses(n_steps)

    def arena(self, name):
        return self.nets[name].p.detach().cpu().numpy()
