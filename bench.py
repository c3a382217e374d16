#!/usr/bin/env python
"""Benchmark of the hot path: gradient-steps/sec of the fused step engine.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload sac_hopper|gail_walker|td3_humanoid|sac_ant]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the reference's CPU algorithm (oracle port) on the host cores

A "step" = ONE gradient step (SAC/TD3: random_batch + train_step; GAIL: one AdvIRL loop iteration,
SURVEY.md 8d), on synthetic MuJoCo-shaped transitions with random-init nets.  Default workload =
BASELINE.json configs[1]: SAC Hopper (obs 11, act 3), 1M-transition HBM ring, batch 256.
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
    "sac_hopper": dict(algo="sac", O=11, A=3, B=256, N=1_000_000, target_entropy=None, reward_scale=1.0, beta_1=0.9),
    "sac_ant": dict(algo="sac", O=111, A=8, B=256, N=1_000_000, target_entropy=-4.0, reward_scale=1.0, beta_1=0.9),
    "gail_walker": dict(algo="gail", O=17, A=6, B=256, N=20_000, NE=4000, target_entropy=None, reward_scale=2.0, beta_1=0.25),
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
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu_index, self.rows, self.proc = gpu_index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


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
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=2000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="sac_hopper", choices=sorted(WORKLOADS))
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
    metric = "%s gradient-steps/sec at batch %d" % ("HER-TD3" if w.get("her") else {"sac": "SAC", "gail": "GAIL (adv_irl)", "td3": "TD3"}[w["algo"]], w["B"])
    config = {"workload": "%s: obs=%d act=%d batch=%d, %d-transition HBM replay ring, 2x%d MLPs, %d gradient steps per kernel launch"
              % (args.workload, w["O"], w["A"], w["B"], w["N"], w.get("H", 256), LAUNCH), "parallelism": "replicas x%d" % max(world, args.gpus),
              "global_batch": w["B"] * max(world, 1)}

    if args.impl == "reference":
        if rank != 0:
            return
        cores = best_cpu_threads(w)
        steps = min(K, 1500 if w["algo"] != "td3" else 600)
        sps, dt = time_cpu_port(w, steps, min(W, 30), cores)
        line = {"impl": "reference", "metric": metric, "value": sps, "unit": "gradient-steps/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": min(W, 30), "ms_per_step": 1000.0 / sps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": sps, "unit": "gradient-steps/s", "cores": cores, "kind": "port",
                                 "sample": "%d gradient steps (of the %d requested) of the same workload incl. random_batch + "
                                           "np_to_pytorch_batch, torch CPU %d threads (best of a 1..32 thread calibration; host has %d cores), %s"
                                           % (steps, K, cores, os.cpu_count() or 1, cpu_model())},
                "e2e": {"value": sps, "unit": "gradient-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from ilswiss_b200 import replicas

    launch = min(LAUNCH, K)
    tr, buf, irl = build_ours(w, seed=100 + rank, steps_per_launch=launch)
    tr._seed = replicas.replica_seed(12345, rank)
    if world > 1:
        replicas.connect_replicas(tr)
    tr.eval_statistics = {}          # no per-epoch stats read-back inside the timed region
    if irl is not None:
        irl.disc_eval_statistics = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")   # > 126 MB L2

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
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = tr.engine.kernel_launches
    evs = []
    sync_all()
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
    clocks = sampler.stop() if rank == 0 else None
    ms = sum(a.elapsed_time(b) for a, b, _ in evs)
    launches = tr.engine.kernel_launches - launches0
    full = [a.elapsed_time(b) for a, b, k in evs if k == launch]
    ms_per_launch = float(np.mean(full)) if full else ms
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = world * K / (ms / 1000.0)

    # ---- e2e through the public API with HOST buffers: per gradient step one transition is appended
    # from (pinned) host memory, one fused step runs, the step's losses are read back to the host.
    Ke = min(args.e2e_steps, K)
    O, A = w["O"], w["A"]
    rs = np.random.RandomState(5)
    host_obs, host_act = rs.randn(Ke, O), rs.uniform(-1, 1, (Ke, A))
    host_nobs, host_rew = rs.randn(Ke, O), rs.randn(Ke)
    def e2e_loop(pipelined):
        sync_all()
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
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        return dt

    e2e_sync_dt = e2e_loop(False)
    e2e_dt = e2e_loop(True)
    e2e = {"value": world * Ke / e2e_dt, "unit": "gradient-steps/s", "h2d_bytes_per_step": buf.ring.host_w * 4,
           "d2h_bytes_per_step": 16 * 4, "steps": Ke,
           "mode": "per-step API calls with host buffers: add_sample+flush (pinned H2D, side stream) -> 1-step launch -> "
                   "loss D2H into a pinned ring (stream-ordered, collected once at the end)",
           "value_with_host_sync_every_step": world * Ke / e2e_sync_dt}
    # train-call granularity (what _do_training does): burst of `launch` transitions + `launch` steps + loss log D2H
    reps = 3
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

    # ---- sampler coupling: the per-env-step get_actions round trip with host buffers, env_num = 4,
    # through ilsw_policy_act_host (one kernel, one sync) next to the eager torch module path the reference uses
    sampler = None
    if world == 1:
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
        sampler = {"get_actions_us": ours_us, "eager_module_get_actions_us": eager_us, "env_num": 4,
                   "note": "host numpy obs -> host numpy actions, stochastic; eager = the nn.Module forward the reference's eval_np runs on the same GPU"}

    # ---- the replay ring's own kernel against the HBM roofline: in-kernel Philox sample + gather of a LARGE batch from this
    # workload's ring (the per-step batch of B rows is gathered inside the engine kernel; this is what the kernel sustains)
    replay = None
    if world == 1:
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
        gby = rows_n * buf.ring.stride * 4 * 2 + rows_n * 4
        replay = {"kernel": "rb_gather_kernel (Philox sample + gather)", "rows": rows_n, "row_bytes": buf.ring.stride * 4, "ms": gms,
                  "algorithmic_bytes": gby, "achieved": gby / gms / 1e6, "unit": "GB/s", "l2": "flushed before every timed launch"}

    # ---- the same metric in the other GEMM precision modes (short runs, same timing method)
    by_prec = {}
    if world == 1 and args.precision is None:
        cur = int(os.environ.get("ILSW_GEMM_PRECISION", "3"))
        by_prec[{0: "fp32_simt", 1: "tf32", 3: "tf32x3"}[cur]] = value
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
    if rank != 0:
        return
    peaks, peak_src = {}, "fallback"
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_step = algorithmic_bytes_per_step(w)
    achieved = bytes_step * launch / (ms_per_launch / 1000.0) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json"))).get(args.workload)
    except Exception:
        pass
    line = {"metric": metric, "value": value, "unit": "gradient-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {0: "f32", 1: "tf32 (tensor-core GEMM operands, f32 accumulate; f32 everywhere else)",
                      3: "f32 via 3xTF32 split on tensor cores"}[int(os.environ.get("ILSW_GEMM_PRECISION", "3"))],
            "value_by_gemm_precision": by_prec,
            "data": "synthetic", "config": dict(config, l2="flushed between timed launches (256 MiB write, untimed)",
                                                sampling="in-kernel Philox, uniform with replacement"),
            "e2e": e2e, "e2e_train_call": e2e_call, "sampler": sampler, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "ilsw_engine_kernel",
                         "algorithmic_bytes_per_launch": bytes_step * launch, "launch_ms": ms_per_launch},
            "clocks": clocks}
    if replay is not None:
        replay["peak"], replay["frac"] = peak, replay["achieved"] / peak
        line["replay_roofline"] = replay
    if world == 1 and not args.no_cpu_baseline:
        cores = best_cpu_threads(w)
        n_cpu = 600 if w["algo"] != "td3" else 250
        sps, dt = time_cpu_port(w, n_cpu, 20, cores)
        line["cpu_baseline"] = {"value": sps, "unit": "gradient-steps/s", "cores": cores, "kind": "port",
                                "sample": "%d gradient steps of the same workload (%.1f s), oracle/restate.py on torch CPU with %d threads (best of a 1..32 calibration; host has %d cores), %s"
                                          % (n_cpu, dt, cores, os.cpu_count() or 1, cpu_model())}
    print(json.dumps(line))


if __name__ == "__main__":
    main()
