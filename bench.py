#!/usr/bin/env python
"""Benchmark of the hot path: gradient-steps/sec of the fused step engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sac_hopper|gail_walker|td3_humanoid|sac_ant]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's own classes (rlkit, via oracle/ref_shim) or the oracle port on the host cores

A "step" = ONE gradient step (SAC/TD3: random_batch + train_step; GAIL: one AdvIRL loop iteration,
SURVEY.md 8d), on synthetic MuJoCo-shaped transitions with random-init nets.  The headline line is
BASELINE.json configs[1]: SAC Hopper (obs 11, act 3), 1M-transition HBM ring, batch 256; the same
line carries `workloads` sub-records measured the same way for the other BASELINE configs (GAIL
Walker2d, TD3 Humanoid B1024, SAC Ant -- at N > 1 only SAC Ant, BASELINE config 5) and, at N > 1,
`replica_check` (policies bit-equal across ranks; one step against the oracle's R-replica emulation).
Prints ONE JSON line (see the task contract for the keys).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: algo, O, A, B, ring capacity, extra
    "sac_hopper": dict(algo="sac", O=11, A=3, B=256, N=1_000_000, target_entropy=None, reward_scale=1.0, beta_1=0.9, steps_per_train_call=1000),
    "sac_ant": dict(algo="sac", O=111, A=8, B=256, N=1_000_000, target_entropy=-4.0, reward_scale=1.0, beta_1=0.9),
    "gail_walker": dict(algo="gail", O=17, A=6, B=256, N=20_000, NE=4000, target_entropy=None, reward_scale=2.0, beta_1=0.25, steps_per_train_call=1000),
    "td3_humanoid": dict(algo="td3", O=376, A=17, B=1024, N=2_000_000),
    # exp_specs/her/her_pick_td3.yaml (her/td3.py): FetchPickAndPlace observation 25 + desired_goal 3, act 4, batch 4096, net_size 300
    "her_td3_pick": dict(algo="td3", her=True, O=28, A=4, B=4096, N=1_000_000, H=300),
}
H, DH = 256, 128
LAUNCH = 1000   # gradient steps per kernel launch (num_train_steps_per_train_call of sac_hopper.yaml:20)


def algorithmic_bytes_per_step(w):
    """24*P_trainable_updated + 8*P_target_updated + 4*B*(2O+A+2) (+ disc)."""
    O, A, B = w["O"], w["A"], w["B"]
    H = w.get("H", 256)
    Pq = (O + A) * H + H + H * H + H + H + 1
    if w["algo"] == "td3":
        Pp = O * H + H + H * H + H + H * A + A
        return 24 * (2 * Pq + 0.5 * Pp) + 8 * 0.5 * (2 * Pq + Pp) + 4 * B * (2 * O + A + 2)
    Pp = O * H + H + H * H + H + 2 * (H * A + A)
    b = 24 * (2 * Pq + Pp + 1) + 8 * 2 * Pq + 4 * B * (2 * O + A + 2)
    if w["algo"] == "gail":
        D = O + A
        Pd = D * DH + DH + DH * DH + DH + DH + 1
        b += 24 * Pd + 2 * B * D * 4
    return b


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clock / throttle reasons DURING the timed regions.  The driver times 20-step runs (a few ms), far below the
    200 ms period of `nvidia-smi -lms`, so the clocks are polled through NVML (the library nvidia-smi itself uses) every
    ~1 ms, and only samples that fall inside a timed region (mark()/unmark()) are reported."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.samples, self._stop_flag, self._active = gpu_index, [], False, False
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_flag:
            if not self._active:
                time.sleep(0.0005)
                continue
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                rs = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                self.samples.append((sm, rs))
            except Exception:
                break

    def mark(self):
        self._active = True

    def unmark(self):
        self._active = False

    def stop(self):
        self._stop_flag = True
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_sm", None), "reasons": [], "samples": 0}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(k for k, bit in names.items() if any(rs & bit for _, rs in self.samples))
        return {"sm_mhz": float(np.median([sm for sm, _ in self.samples])), "sm_max_mhz": self.max_sm, "reasons": reasons,
                "samples": len(self.samples), "source": "NVML polled inside the timed regions"}


# ------------------------------------------------------------------------------------------------
def build_ours(w, seed, steps_per_launch):
    from ilswiss_b200 import adv_irl, modules, replay_buffer, trainers

    torch.manual_seed(seed)
    O, A, B = w["O"], w["A"], w["B"]
    H = w.get("H", 256)
    qf1 = modules.FlattenMlp([H, H], 1, O + A)
    qf2 = modules.FlattenMlp([H, H], 1, O + A)
    if w["algo"] == "td3":
        policy = modules.DeterministicNoisePolicy([H, H], O, A, policy_noise=0.2, policy_noise_clip=0.5)
        if w.get("her"):
            policy.sigma, policy.min_act = 0.2, -1.0
            tr = trainers.HerTD3(policy, qf1, qf2, reward_scale=1.0, discount=0.99, policy_lr=6e-4, qf_lr=3e-4,
                                 policy_and_target_update_period=2, soft_target_tau=0.005, batch_size=B,
                                 max_steps_per_call=steps_per_launch)
        else:
            tr = trainers.TD3(policy, qf1, qf2, reward_scale=1.0, discount=0.99, policy_lr=3e-4, qf_lr=3e-4,
                              policy_and_target_update_period=2, soft_target_tau=0.005, batch_size=B,
                              max_steps_per_call=steps_per_launch)
    else:
        policy = modules.TanhGaussianPolicy([H, H], O, A)
        tr = trainers.SoftActorCritic(policy, qf1, qf2, reward_scale=w["reward_scale"], discount=0.99, policy_lr=3e-4,
                                      qf_lr=3e-4, alpha_lr=3e-4, soft_target_tau=0.005, alpha=0.2, train_alpha=True,
                                      policy_mean_reg_weight=1e-3, policy_std_reg_weight=1e-3, beta_1=w["beta_1"],
                                      target_entropy=w["target_entropy"], batch_size=B, max_steps_per_call=steps_per_launch)
    buf = replay_buffer.DeviceReplayBuffer(w["N"], O, A, random_seed=1)
    fill_synthetic(buf, w["N"], O, A, seed=7, term_p=0.0 if w["algo"] == "gail" else 0.01)
    irl = None
    if w["algo"] == "gail":
        ebuf = replay_buffer.DeviceReplayBuffer(w["N"], O, A, random_seed=3)
        fill_synthetic(ebuf, w["NE"], O, A, seed=8, term_p=0.0)
        disc = modules.MLPDisc(O + A, DH)
        irl = adv_irl.AdvIRLEngine("gail2", disc, tr, ebuf, buf, disc_optim_batch_size=B, policy_optim_batch_size=B,
                                   num_update_loops_per_train_call=steps_per_launch, num_disc_updates_per_loop_iter=1,
                                   num_policy_updates_per_loop_iter=1, disc_lr=3e-4, disc_momentum=0.9,
                                   use_grad_pen=True, grad_pen_weight=8.0)
    return tr, buf, irl


def fill_synthetic(buf, n, O, A, seed, term_p):
    """obs,next_obs ~ N(0,1), act ~ U(-1,1), rew ~ N(0,1), term ~ Bern(p); generated on the device in
    chunks, written straight into the HBM ring (no host copy)."""
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    S = buf.ring.stride
    chunk = 250_000
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        rows = torch.zeros((m, S), dtype=torch.float32, device="cuda")
        rows[:, :O] = torch.randn((m, O), generator=g, device="cuda")
        rows[:, O:O + A] = torch.rand((m, A), generator=g, device="cuda") * 2 - 1
        rows[:, O + A] = torch.randn((m,), generator=g, device="cuda")
        rows[:, O + A + 1] = (torch.rand((m,), generator=g, device="cuda") < term_p).float()
        rows[:, O + A + 2:2 * O + A + 2] = torch.randn((m, O), generator=g, device="cuda")
        buf.load_device_rows(rows)
    torch.cuda.synchronize()


def run_steps(tr, buf, irl, n):
    if irl is not None:
        irl.do_training(n)
    else:
        tr.train_from_buffer(buf, n)


# ------------------------------------------------------------------------------------------------
def time_cpu_port(w, steps, warmup, threads):
    """The reference's algorithm for this path on the host cores: oracle/restate.py (a CPU torch
    restatement pinned to the executed reference; the reference itself is Python and cannot travel)."""
    from oracle import restate as R

    torch.set_num_threads(threads)
    O, A, B = w["O"], w["A"], w["B"]
    H = w.get("H", 256)
    n = min(w["N"], 1_000_000)
    rs = np.random.RandomState(0)
    nets = dict(qf1=R.Net(R.init_mlp(rs, O + A, (H, H), 1)), qf2=R.Net(R.init_mlp(rs, O + A, (H, H), 1)))
    data = R.synth_transitions(n, O, A, 7, 0.0 if w["algo"] == "gail" else 0.01)
    buf = R.ReplayOracle(n, O, A, random_seed=1)
    buf.load_bulk(data)
    disc = ebuf = None
    if w["algo"] == "td3":
        nets["policy"] = R.Net(R.init_mlp(rs, O, (H, H), A, init_w=1e-3))
        if w.get("her"):
            tr = R.HerTD3Oracle(nets["policy"], nets["qf1"], nets["qf2"], sigma=0.2, policy_lr=6e-4, qf_lr=3e-4)
        else:
            tr = R.TD3Oracle(nets["policy"], nets["qf1"], nets["qf2"], policy_lr=3e-4, qf_lr=3e-4)
    else:
        nets["policy"] = R.Net(R.init_mlp(rs, O, (H, H), A, init_w=1e-3, log_std_head=True))
        tr = R.SacAlphaOracle(nets["policy"], nets["qf1"], nets["qf2"], A, reward_scale=w["reward_scale"], policy_lr=3e-4,
                              qf_lr=3e-4, soft_target_tau=0.005, beta_1=w["beta_1"], target_entropy=w["target_entropy"])
        if w["algo"] == "gail":
            disc = R.DiscOracle(R.Net(R.init_disc(rs, O + A, DH)), disc_lr=3e-4, disc_momentum=0.9, grad_pen_weight=8.0)
            ebuf = R.ReplayOracle(w["NE"], O, A, random_seed=3)
            ebuf.load_bulk(R.synth_transitions(w["NE"], O, A, 8, 0.0))

    def one():
        if disc is not None:
            eb = R.np_to_torch_batch(ebuf.random_batch(B, keys=["observations", "actions"]))
            pb = R.np_to_torch_batch(buf.random_batch(B, keys=["observations", "actions"]))
            disc.reward_step(torch.cat([eb["observations"], eb["actions"]], 1),
                             torch.cat([pb["observations"], pb["actions"]], 1), torch.rand(B, 1))
        batch = R.np_to_torch_batch(buf.random_batch(B))
        if disc is not None:
            batch["rewards"] = disc.rewards(batch["observations"], batch["actions"], "gail2")
        if w["algo"] == "td3":
            tr.train_step(batch, torch.randn(B, A))
        else:
            tr.train_step(batch, torch.randn(B, A), torch.randn(B, A))

    for _ in range(warmup):
        one()
    t0 = time.perf_counter()
    for _ in range(steps):
        one()
    dt = time.perf_counter() - t0
    return steps / dt, dt


def best_cpu_threads(w):
    """torch CPU intra-op scaling of this tiny-GEMM workload saturates at a few threads (SURVEY.md
    section 6) and collapses when oversubscribed: calibrate, then time with the best setting."""
    cores = os.cpu_count() or 1
    best, best_sps = 1, 0.0
    for th in (1, 2, 4, 8, 16, 32):
        if th > cores:
            break
        sps, _ = time_cpu_port(w, 12, 3, th)
        if sps > best_sps:
            best, best_sps = th, sps
    return best


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


# ------------------------------------------------------------------------------------------------
def time_reference_classes(w, steps, warmup, threads):
    """The UNMODIFIED reference classes (rlkit SimpleReplayBuffer + trainers + AdvIRL) on the host cores, when the reference
    is installed next to the repo (baseline/_ref, shipped with the lease) -- through oracle/ref_shim (checker infra).
    Returns None when the reference is not importable here."""
    try:
        from oracle import ref_shim
        if not ref_shim.reference_available():
            return None
        return ref_shim.time_reference(w, steps, warmup, threads)
    except Exception as e:      # fall back to the port, say why
        sys.stderr.write("reference classes unavailable (%s): timing the oracle port\n" % (e,))
        return None


def cpu_arm(w, steps, warmup):
    """(steps/s, seconds, cores, kind) of the reference's CPU path for workload `w`."""
    cores = best_cpu_threads(w)
    got = time_reference_classes(w, steps, warmup, cores)
    if got is not None:
        return got[0], got[1], cores, "reference"
    sps, dt = time_cpu_port(w, steps, warmup, cores)
    return sps, dt, cores, "port"


def metric_name(w):
    return "%s gradient-steps/sec at batch %d" % ("HER-TD3" if w.get("her") else {"sac": "SAC", "gail": "GAIL (adv_irl)", "td3": "TD3"}[w["algo"]], w["B"])


def workload_desc(name, w, launch):
    return ("%s: obs=%d act=%d batch=%d, %d-transition HBM replay ring, 2x%d MLPs, %d gradient steps per kernel launch"
            % (name, w["O"], w["A"], w["B"], w["N"], w.get("H", 256), launch))


def measure(name, args, K, W, rank, world, dist, flush, clock, peak, peak_src, extras):
    """One workload, measured the contract's way: W warm-up steps, EXACTLY K timed steps (CUDA events around every launch,
    L2 flushed between launches, max over ranks), then the e2e legs through the public API with host buffers."""
    from ilswiss_b200 import replicas
    w = WORKLOADS[name]
    launch = min(LAUNCH, K)
    tr, buf, irl = build_ours(w, seed=100 + rank, steps_per_launch=launch)
    tr._seed = replicas.replica_seed(12345, rank)
    solo = None
    if world > 1:
        # diagnostic (untimed region): every rank's step time BEFORE the replicas are connected.  Lock-step replicas run at
        # the pace of the slowest GPU, so mean(solo) / max(solo) bounds the scaling efficiency from above.
        run_steps(tr, buf, irl, launch)
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        run_steps(tr, buf, irl, launch)
        s1.record()
        torch.cuda.synchronize()
        t = torch.tensor([s0.elapsed_time(s1) * 1000.0 / launch], dtype=torch.float64, device="cuda")
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        solo = [float(x.item()) for x in allt]
        replicas.connect_replicas(tr)
    tr.eval_statistics = {}          # no per-epoch stats read-back inside the timed region
    if irl is not None:
        irl.disc_eval_statistics = {}

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    done = 0
    while done < W:
        k = min(launch, W - done)
        run_steps(tr, buf, irl, k)
        done += k
    sync_all()
    launches0 = tr.engine.kernel_launches
    evs = []
    sync_all()
    clock.mark()
    done = 0
    while done < K:                  # EXACTLY K gradient steps
        k = min(launch, K - done)
        flush.zero_()                # L2 flush between timed launches (not timed)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run_steps(tr, buf, irl, k)
        e1.record()
        evs.append((e0, e1, k))
        done += k
    sync_all()
    clock.unmark()
    ms = sum(a.elapsed_time(b) for a, b, _ in evs)
    launches = tr.engine.kernel_launches - launches0
    full = [a.elapsed_time(b) for a, b, k in evs if k == launch]
    ms_per_launch = float(np.mean(full)) if full else ms
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K / (ms / 1000.0)

    # ---- e2e through the public API with HOST buffers: per gradient step one transition is appended from (pinned) host
    # memory, one fused step runs, the step's losses are read back to the host.  At least 2000 steps: the per-call
    # latency of a 1-step launch is what sac_ant.yaml / td3_humanoid.yaml (1 gradient step per train call) users see.
    Ke = max(min(args.e2e_steps, 4000), 1)
    O, A = w["O"], w["A"]
    rs = np.random.RandomState(5)
    host_obs, host_act = rs.randn(Ke, O), rs.uniform(-1, 1, (Ke, A))
    host_nobs, host_rew = rs.randn(Ke, O), rs.randn(Ke)

    def e2e_loop(pipelined):
        sync_all()
        clock.mark()
        t0 = time.perf_counter()
        for i in range(Ke):
            buf.add_sample(host_obs[i], host_act[i], host_rew[i], False, host_nobs[i])
            buf.flush()                          # pinned H2D of this step's transition (side stream)
            run_steps(tr, buf, irl, 1)           # one fused gradient step (one kernel launch)
            if pipelined:
                tr.engine.losses_async(1)        # D2H of this step's losses into a pinned ring, no host sync
            else:
                tr.engine.losses(1)              # D2H + host sync every step
        if pipelined:
            got = tr.engine.losses_collect()     # one sync; every step's losses are on the host now
            assert got.shape[0] == Ke and np.isfinite(got[:, :2]).all(), got.shape
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        clock.unmark()
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    e2e_sync_dt = e2e_loop(False)
    e2e_dt = e2e_loop(True)
    # train-call granularity (what _do_training does): burst of `launch` transitions + `launch` steps + loss log D2H
    reps = max(3, -(-3000 // launch))          # >= 3000 gradient steps: a 20-step launch (driver default) alone is a 3 ms sample
    burst = dict(observations=rs.randn(launch, O), actions=rs.uniform(-1, 1, (launch, A)), rewards=rs.randn(launch, 1),
                 terminals=np.zeros((launch, 1)), next_observations=rs.randn(launch, O))
    sync_all()
    t0 = time.perf_counter()
    for _ in range(reps):
        buf.add_samples(**burst)
        run_steps(tr, buf, irl, launch)
        tr.engine.losses(launch)
    torch.cuda.synchronize()
    call_dt = time.perf_counter() - t0
    e2e_call = {"value": world * reps * launch / call_dt, "unit": "gradient-steps/s",
                "h2d_bytes_per_step": buf.ring.host_w * 4, "d2h_bytes_per_step": 16 * 4,
                "mode": "per train call: %d-transition burst H2D -> %d-step launch -> loss log D2H" % (launch, launch)}
    # The headline e2e is measured at the call granularity the reference's OWN yaml of this workload uses
    # (num_train_steps_per_train_call): sac_hopper.yaml:19-20 / gail_walker.yaml:39,51 train 1000 steps per train call,
    # sac_ant.yaml:19-20 / td3_humanoid.yaml:21-22 ONE step per env step with the next get_actions waiting on it -- there
    # the figure a user sees is the per-step loop WITH a host sync every step.
    per_call = w.get("steps_per_train_call", 1)
    per_step_sync = world * Ke / e2e_sync_dt
    per_step_pipe = world * Ke / e2e_dt
    if per_call > 1:
        e2e = {"value": e2e_call["value"], "granularity": "train call of %d gradient steps (the workload's yaml: %d per call)" % (launch, per_call),
               "mode": e2e_call["mode"]}
    else:
        e2e = {"value": per_step_sync, "granularity": "one gradient step per call, host sync after every step (the workload's yaml: 1 per call)",
               "mode": "per step: add_sample + flush (pinned H2D, side stream) -> 1-step launch -> losses read on the host (mailbox poll)"}
    e2e.update({"unit": "gradient-steps/s", "h2d_bytes_per_step": buf.ring.host_w * 4, "d2h_bytes_per_step": 16 * 4, "steps_per_step_loops": Ke,
                "per_step_host_sync": per_step_sync, "per_step_pipelined": per_step_pipe, "per_train_call": e2e_call["value"],
                "value_with_host_sync_every_step": per_step_sync})

    bytes_step = algorithmic_bytes_per_step(w)
    achieved = bytes_step * launch / (ms_per_launch / 1000.0) / 1e9
    traffic = None
    try:
        per_step = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(name + "_per_step")
        traffic = None if per_step is None else per_step * launch      # same launch the algorithmic bytes are for
    except Exception:
        pass
    rec = {"metric": metric_name(w), "value": value, "unit": "gradient-steps/s", "steps": K, "warmup": W, "ms_per_step": ms / K,
           "timing": "device-timed: sum of CUDA-event times around the %d launch(es), max over ranks" % len(evs),
           "config": {"workload": workload_desc(name, w, launch), "parallelism": "replicas x%d" % world, "global_batch": w["B"] * world,
                      "l2": "flushed between timed launches (256 MiB write, untimed)", "sampling": "in-kernel Philox, uniform with replacement",
                      "engine": "tcgen05/TMA GEMM tiles" if getattr(tr.engine, "uses_tc5", lambda: False)() else "mma.sync 32x32 tiles (TMA panels)",
                      "exchange": ("none (1 replica)" if world == 1 else
                                   ("in-kernel push from the weight-gradient epilogues, NVLS multicast stores" if getattr(tr.engine, "replica_multicast", False)
                                    else "in-kernel push from the weight-gradient epilogues, per-peer NVLink stores"))},
           "e2e": e2e, "e2e_train_call": e2e_call, "gpu_launches": int(launches),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                        "traffic": traffic, "traffic_note": "ncu dram bytes per step (profiles/) x steps of this launch", "peak_source": peak_src,
                        "kernel": "ilsw_engine_kernel", "launch_steps": launch, "algorithmic_bytes_per_launch": bytes_step * launch,
                        "launch_ms": ms_per_launch}}

    if solo is not None:
        rec["replicas_unconnected_us_per_step"] = {"per_rank": solo, "min": min(solo), "max": max(solo), "mean": float(np.mean(solo)),
                                                   "note": "each GPU alone, same program without the exchange, measured before connect_replicas"}
    if world == 1 and extras:
        # ---- sampler coupling: the per-env-step get_actions round trip with host buffers, env_num = 4
        from ilswiss_b200.sampler import DevicePolicy
        dp = DevicePolicy(tr, seed=1)
        obs4 = rs.randn(4, O)
        for _ in range(20):
            dp.get_actions(obs4)
            tr.policy.get_actions(obs4)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(500):
            dp.get_actions(obs4)
        ours_us = (time.perf_counter() - t0) / 500 * 1e6
        t0 = time.perf_counter()
        with torch.no_grad():
            for _ in range(500):
                tr.policy.get_actions(obs4)
        eager_us = (time.perf_counter() - t0) / 500 * 1e6
        rec["sampler"] = {"get_actions_us": ours_us, "eager_module_get_actions_us": eager_us, "env_num": 4,
                          "note": "host numpy obs -> host numpy actions, stochastic; eager = the nn.Module forward the reference's eval_np runs on the same GPU"}
    if world == 1:
        # ---- the replay ring's own kernel against the HBM roofline: in-kernel Philox sample + gather of a LARGE batch
        rows_n = (1 << 20) if buf.ring.stride <= 64 else (1 << 18)
        for _ in range(3):
            buf.ring.sample(rows_n, 7, 1)
        ts = []
        for i in range(10):
            flush.zero_()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record(); out = buf.ring.sample(rows_n, 7, 2 + i); g1.record()
            torch.cuda.synchronize()
            ts.append(g0.elapsed_time(g1))
            del out
        gms = float(np.median(ts))
        row_bytes = (2 * O + A + 2) * 4
        gby = rows_n * row_bytes * 2 + rows_n * 4
        ring_bytes = w["N"] * buf.ring.stride * 4
        in_l2 = ring_bytes < (64 << 20)           # a ring far below the 126 MB L2 (GAIL: 20 000 rows = 3.8 MB, every row sampled ~50 times) is not an HBM measurement
        gby_stride = rows_n * buf.ring.stride * 4 * 2 + rows_n * 4
        rec["replay_roofline"] = {"kernel": "%s (Philox sample + gather)" % ("rb_gather_bulk_kernel: one TMA bulk copy per row" if buf.ring.stride * 4 <= 512 else "rb_gather_kernel: 128-bit loads"),
                                  "rows": rows_n, "row_bytes": row_bytes,
                                  "ring_row_stride_bytes": buf.ring.stride * 4, "ms": gms, "algorithmic_bytes": gby,
                                  "achieved": gby / gms / 1e6, "unit": "GB/s", "peak": peak,
                                  "frac": None if in_l2 else gby / gms / 1e6 / peak,
                                  "frac_of_padded_rows": None if in_l2 else gby_stride / gms / 1e6 / peak,
                                  "note": ("ring of %.0f MB is L2 resident: not an HBM-roofline measurement" % (ring_bytes / 1e6)) if in_l2 else
                                          "frac counts the payload bytes of a row (2O+A+2 floats, read + written); rows are padded to 64-byte multiples in the ring "
                                          "(frac_of_padded_rows counts the stride); sampling is with replacement, so repeated rows of a "
                                          "batch are L2 hits (%d samples from %d ring rows)" % (rows_n, w["N"]),
                                  "l2": "flushed before every timed launch"}
    if world == 1 and not args.no_cpu_baseline:
        n_cpu = {"td3": 120, "gail": 250}.get(w["algo"], 400) if not extras else (600 if w["algo"] != "td3" else 250)
        sps, dt, cores, kind = cpu_arm(w, n_cpu, 20)
        rec["cpu_baseline"] = {"value": sps, "unit": "gradient-steps/s", "cores": cores, "kind": kind,
                               "sample": "%d gradient steps of the same workload (%.1f s), %s on torch CPU with %d threads (best of a 1..32 calibration; host has %d cores), %s"
                                         % (n_cpu, dt, "the reference's rlkit classes (oracle/ref_shim)" if kind == "reference" else "oracle/restate.py",
                                            cores, os.cpu_count() or 1, cpu_model())}
    check = None
    if world > 1:
        check = replica_check(tr, buf, w, rank, world, dist)
    return rec, check, tr, buf, irl


def replica_check(tr, buf, w, rank, world, dist):
    """CHECKER (after the timed region; oracle/ is test infrastructure): (1) the policy arenas of all ranks are bit-equal
    after the timed launches; (2) ONE more injected step on every rank against the oracle's R-replica emulation -- each
    rank restates its own step on the CPU from its own critics and batch, the per-rank policy gradients are summed in rank
    order / R (replicas.emulate_replica_average) and fed to the oracle's Adam; the device policy must match."""
    from ilswiss_b200 import replicas
    pol = tr._arenas["policy"]
    gathered = [torch.empty_like(pol.p) for _ in range(world)]
    dist.all_gather(gathered, pol.p.contiguous())
    equal = all(torch.equal(gathered[0], g) for g in gathered[1:])
    out = {"equal": bool(equal), "world": world}
    if w["algo"] != "sac":
        return out
    try:
        from oracle import restate as R
        O, A, B = w["O"], w["A"], w["B"]
        names = R.mlp_param_names(2)

        def net_from(arena, extra=()):
            shapes = [tuple(v.shape) for v in arena.views("p")]
            flat = arena.p.detach().cpu().numpy()
            keys = R.mlp_param_names(2, extra)
            d, off = {}, 0
            for k, shp in zip(keys, shapes):
                n = int(np.prod(shp)); d[k] = flat[off:off + n].reshape(shp).copy(); off += n
            net = R.Net(d)
            for name, store in (("m", net.m), ("v", net.v)):
                t = getattr(arena, name, None)
                if t is not None:
                    f, off = t.detach().cpu().numpy(), 0
                    for k, shp in zip(keys, shapes):
                        n = int(np.prod(shp)); store[k] = torch.from_numpy(f[off:off + n].reshape(shp).copy()); off += n
            return net

        st = tr.engine.get_state()
        nets = {k: net_from(tr._arenas[k], ("last_fc_log_std",) if k == "policy" else ()) for k in ("policy", "qf1", "qf2", "target_qf1", "target_qf2")}
        nets["policy"].t, nets["qf1"].t, nets["qf2"].t = int(st.adam_step[2]), int(st.adam_step[0]), int(st.adam_step[1])
        ora = R.SacAlphaOracle(nets["policy"], nets["qf1"], nets["qf2"], A, reward_scale=w["reward_scale"], policy_lr=3e-4, qf_lr=3e-4,
                               soft_target_tau=0.005, beta_1=w["beta_1"], target_entropy=w["target_entropy"])
        ora.target_qf1, ora.target_qf2 = nets["target_qf1"], nets["target_qf2"]
        ora.log_alpha = float(st.log_alpha)
        rs = np.random.RandomState(1000 + rank)
        idx = rs.randint(0, buf.num_steps_can_sample(), B).astype(np.int32)
        eps_n, eps_c = rs.randn(B, A).astype(np.float32), rs.randn(B, A).astype(np.float32)
        batch = buf.random_batch_device(B, indices=idx)
        cpu_batch = {k: v.detach().cpu().clone() for k, v in batch.items()}
        cpu_batch["rewards"] = cpu_batch["rewards"].reshape(B, 1); cpu_batch["terminals"] = cpu_batch["terminals"].reshape(B, 1)
        pol0 = nets["policy"].clone()
        res = ora.train_step(cpu_batch, torch.from_numpy(eps_n), torch.from_numpy(eps_c))
        g_local = torch.from_numpy(np.concatenate([g.ravel() for g in res["grads"]["policy"]])).cuda()
        gs = [torch.empty_like(g_local) for _ in range(world)]
        dist.all_gather(gs, g_local)
        g_avg = replicas.emulate_replica_average([g.cpu().numpy() for g in gs])
        parts, off = [], 0
        for v in pol0.p.values():
            n = v.numel(); parts.append(torch.from_numpy(g_avg[off:off + n].reshape(tuple(v.shape)).copy())); off += n
        R.adam_update(pol0, parts, 3e-4, w["beta_1"])
        inj = dict(idx=torch.from_numpy(idx[None]).cuda(), eps_next=torch.from_numpy(eps_n[None]).cuda(), eps_cur=torch.from_numpy(eps_c[None]).cuda())
        tr.train_from_buffer(buf, 1, inject=inj)
        torch.cuda.synchronize()
        got = tr._arenas["policy"].p.detach().cpu().numpy()
        want = pol0.flat()
        diff = np.abs(got - want)
        out.update({"oracle_max_abs_err": float(diff.max()), "oracle_rel_err": float(diff.max() / max(np.abs(want).max(), 1e-12)),
                    "frac_beyond_1e-5": float((diff > 1e-5).mean()), "qf1_loss_rel_err": None})
        L = tr.engine.losses(1)
        out["qf1_loss_rel_err"] = float(abs(L[0, 0] - res["qf1_loss"]) / max(abs(res["qf1_loss"]), 1e-12))
        again = [torch.empty_like(pol.p) for _ in range(world)]
        dist.all_gather(again, pol.p.contiguous())
        out["equal_after_check_step"] = bool(all(torch.equal(again[0], g) for g in again[1:]))
        out["ok"] = bool(out["equal"] and out["equal_after_check_step"] and out["frac_beyond_1e-5"] < 5e-3 and out["oracle_max_abs_err"] < 7e-4
                         and out["qf1_loss_rel_err"] < 1e-4)
    except Exception as e:
        out["oracle_error"] = repr(e)
        out["ok"] = False
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=2000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sac_hopper", choices=sorted(WORKLOADS))
    ap.add_argument("--sub", default=None, help="comma-separated sub-record workloads (default: the other BASELINE configs; 'none' to skip)")
    ap.add_argument("--e2e-steps", type=int, default=2000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--precision", type=int, default=None, help="GEMM mode: 0 fp32 SIMT, 1 TF32, 3 3xTF32 (default: library default)")
    args = ap.parse_args()
    if args.precision is not None:
        os.environ["ILSW_GEMM_PRECISION"] = str(args.precision)
    w = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    K, W = args.steps, max(args.warmup, 3)
    launch = min(LAUNCH, K)
    config = {"workload": workload_desc(args.workload, w, launch), "parallelism": "replicas x%d" % max(world, args.gpus),
              "global_batch": w["B"] * max(world, 1)}

    if args.impl == "reference":
        if rank != 0:
            return
        steps = min(K, 1500 if w["algo"] != "td3" else 600)
        sps, dt, cores, kind = cpu_arm(w, steps, min(W, 30))
        line = {"impl": "reference", "metric": metric_name(w), "value": sps, "unit": "gradient-steps/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": min(W, 30), "ms_per_step": 1000.0 / sps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": sps, "unit": "gradient-steps/s", "cores": cores, "kind": kind,
                                 "sample": "%d gradient steps (of the %d requested) of the same workload incl. random_batch + "
                                           "np_to_pytorch_batch, %s, torch CPU %d threads (best of a 1..32 thread calibration; host has %d cores), %s"
                                           % (steps, K, "the reference's own rlkit classes (baseline/_ref through oracle/ref_shim)" if kind == "reference"
                                              else "oracle/restate.py (CPU port pinned to the executed reference)", cores, os.cpu_count() or 1, cpu_model())},
                "e2e": {"value": sps, "unit": "gradient-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    clock = ClockSampler(local_rank)
    if rank == 0:
        clock.start()

    rec, check, tr, buf, irl = measure(args.workload, args, K, W, rank, world, dist, flush, clock, peak, peak_src, extras=True)

    # ---- the same metric in the other GEMM precision modes (short runs, same timing method)
    by_prec = {}
    if world == 1 and args.precision is None:
        cur = int(os.environ.get("ILSW_GEMM_PRECISION", "3"))
        by_prec[{0: "fp32_simt", 1: "tf32", 3: "tf32x3"}[cur]] = rec["value"]
        for pm in (0, 3, 1):
            if pm == cur:
                continue
            os.environ["ILSW_GEMM_PRECISION"] = str(pm)
            tr2, buf2, irl2 = build_ours(w, seed=100, steps_per_launch=launch)
            tr2.eval_statistics = {}
            if irl2 is not None:
                irl2.disc_eval_statistics = {}
            run_steps(tr2, buf2, irl2, launch)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(3):
                run_steps(tr2, buf2, irl2, launch)
            a1.record()
            torch.cuda.synchronize()
            by_prec[{0: "fp32_simt", 1: "tf32", 3: "tf32x3"}[pm]] = 3 * launch / (a0.elapsed_time(a1) / 1000.0)
            del tr2, buf2, irl2
        os.environ["ILSW_GEMM_PRECISION"] = str(cur)
    del tr, buf, irl
    torch.cuda.empty_cache()

    # ---- the other BASELINE.json configs, measured the same way (sub-records of the same line)
    if args.sub is None:
        subs = [n for n in (("gail_walker", "td3_humanoid", "sac_ant") if world == 1 else ("sac_ant",)) if n != args.workload]
    else:
        subs = [n for n in args.sub.split(",") if n and n != "none"]
    workloads, checks = {}, {}
    if check is not None:
        checks[args.workload] = check
    for name in subs:
        r2, c2, t2, b2, i2 = measure(name, args, K, W, rank, world, dist, flush, clock, peak, peak_src, extras=False)
        workloads[name] = r2
        if c2 is not None:
            checks[name] = c2
        del t2, b2, i2
        torch.cuda.empty_cache()
    clocks = clock.stop() if rank == 0 else None
    if rank != 0:
        return
    line = {"metric": rec["metric"], "value": rec["value"], "unit": "gradient-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": rec["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "tf32 (tensor-core GEMM operands, f32 accumulate; f32 everywhere else)",
                      3: "f32 via 3xTF32 split on tensor cores"}[int(os.environ.get("ILSW_GEMM_PRECISION", "3"))],
            "value_by_gemm_precision": by_prec, "data": "synthetic", "config": dict(config, **{k: v for k, v in rec["config"].items() if k not in config}),
            "timing": rec["timing"], "e2e": rec["e2e"], "e2e_train_call": rec["e2e_train_call"], "sampler": rec.get("sampler"),
            "gpu_launches": rec["gpu_launches"], "roofline": rec["roofline"], "clocks": clocks}
    for k in ("replay_roofline", "cpu_baseline", "replicas_unconnected_us_per_step"):
        if k in rec:
            line[k] = rec[k]
    if workloads:
        line["workloads"] = workloads
    if checks:
        line["replica_check"] = checks
    print(json.dumps(line))


if __name__ == "__main__":
    main()
