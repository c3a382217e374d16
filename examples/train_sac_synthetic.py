#!/usr/bin/env python
"""Minimal end-to-end use of ilswiss_b200 WITHOUT an ILSwiss checkout (needs a B200 + the built library):
a synthetic vec-env loop that queries actions through DevicePolicy, appends transitions to the HBM replay ring
and runs SAC train calls of 1000 fused gradient steps each -- the structure of BaseAlgorithm.start_training
(rlkit/core/base_algorithm.py:150-260) with the drop-in classes in the places INTEGRATION.md names.

    python examples/train_sac_synthetic.py [--epochs 3] [--env-num 4]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ilswiss_b200 import modules  # noqa: E402
from ilswiss_b200.replay_buffer import DeviceReplayBuffer  # noqa: E402
from ilswiss_b200.sampler import DevicePolicy, MakeDeterministic  # noqa: E402
from ilswiss_b200.trainers import SoftActorCritic  # noqa: E402


class SyntheticVecEnv:
    """Hopper-shaped linear system with a quadratic cost: just enough dynamics for the losses to move."""

    def __init__(self, n, obs_dim=11, act_dim=3, seed=0):
        self.rs = np.random.RandomState(seed)
        self.n, self.O, self.A = n, obs_dim, act_dim
        self.Amat = 0.8 * np.eye(obs_dim) + 0.02 * self.rs.randn(obs_dim, obs_dim)       # spectral radius < 1: stable
        self.Bmat = 0.3 * self.rs.randn(obs_dim, act_dim)
        self.state = self.rs.randn(n, obs_dim)

    def reset(self):
        self.state = self.rs.randn(self.n, self.O)
        return self.state.copy()

    def step(self, act):
        nxt = np.clip(self.state @ self.Amat.T + act @ self.Bmat.T + 0.05 * self.rs.randn(self.n, self.O), -10.0, 10.0)
        rew = -(nxt ** 2).sum(1) * 0.1 - 0.01 * (act ** 2).sum(1)
        self.state = nxt
        return nxt.copy(), rew, np.zeros(self.n, dtype=bool)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--env-num", type=int, default=4)
    ap.add_argument("--steps-per-epoch", type=int, default=2000)
    args = ap.parse_args()
    O, A, B = 11, 3, 256
    policy = modules.TanhGaussianPolicy([256, 256], O, A)
    trainer = SoftActorCritic(policy, modules.FlattenMlp([256, 256], 1, O + A), modules.FlattenMlp([256, 256], 1, O + A),
                              reward_scale=1.0, discount=0.99, policy_lr=3e-4, qf_lr=3e-4, soft_target_tau=0.005,
                              alpha=0.2, train_alpha=True, batch_size=B, max_steps_per_call=1000)
    buf = DeviceReplayBuffer(1_000_000, O, A, random_seed=1)
    explore, evaluate = DevicePolicy(trainer, seed=1), MakeDeterministic(DevicePolicy(trainer, seed=2))
    env = SyntheticVecEnv(args.env_num, O, A)
    obs = env.reset()
    for epoch in range(args.epochs):
        t0, n_grad = time.time(), 0
        for step in range(args.steps_per_epoch):
            act = explore.get_actions(obs)                              # one C-ABI call: H2D, kernel, D2H
            nxt, rew, done = env.step(act)
            for i in range(args.env_num):
                buf.add_sample(obs[i], act[i], rew[i], done[i], nxt[i])  # pinned, staged on a side stream
            obs = nxt
            if (step + 1) % 1000 == 0 and buf.num_steps_can_sample() >= 1000:
                trainer.train_from_buffer(buf, 1000)                     # ONE kernel launch: 1000 gradient steps
                n_grad += 1000
        st = trainer.get_eval_statistics() or {}
        ret = float(np.mean([env.step(evaluate.get_actions(env.state))[1].mean() for _ in range(20)]))
        print("epoch %d: %d env steps, %d gradient steps in %.2f s | QF1 loss %.4f policy loss %.4f alpha %.4f | eval reward %.3f"
              % (epoch, args.steps_per_epoch * args.env_num, n_grad, time.time() - t0, st.get("QF1 Loss", float("nan")),
                 st.get("Policy Loss", float("nan")), float(trainer.alpha), ret))
        trainer.end_epoch()


if __name__ == "__main__":
    main()
