"""Expert-demonstration ingest: the host half of run_scripts/adv_irl_exp_script.py:48-138.

On-disk format (demos_listing.yaml -> *.pkl): a pickled LIST of trajectory dicts, each with per-step arrays
`observations, actions, rewards, next_observations, terminals` (the output of run_scripts/gen_expert_demos.py).
The script samples `traj_num` trajectories, derives normalisation statistics from their observations, optionally
rescales the demos with them (ScaledEnv / MinmaxEnv use the same statistics for the live environment,
rlkit/envs/wrappers.py:9,58-129) and add_path()s every trajectory into the expert replay buffer.

Here the same steps end in ONE pinned host->device copy per trajectory into the HBM-resident expert ring
(DeviceReplayBuffer.add_path), from which the discriminator step samples in-kernel.
"""
import pickle
import random

import numpy as np

EPS = np.finfo(np.float32).eps.item()       # rlkit/envs/wrappers.py:9


def load_demos(demos_path, traj_num=None, rng=random):
    """adv_irl_exp_script.py:51-53: unpickle the trajectory list and sample traj_num of them (python `random`)."""
    with open(demos_path, "rb") as f:
        traj_list = pickle.load(f)
    if traj_num is not None:
        traj_list = rng.sample(traj_list, traj_num)
    return traj_list


def demo_stats(traj_list):
    """:55-60 -- statistics over all observations of the selected trajectories (actions are left unscaled)."""
    obs = np.vstack([t["observations"] for t in traj_list])
    return dict(obs_mean=np.mean(obs, axis=0), obs_std=np.std(obs, axis=0), acts_mean=None, acts_std=None,
                obs_min=np.min(obs, axis=0), obs_max=np.max(obs, axis=0))


def normalize_demos(traj_list, stats, scale_env_with_demo_stats=False, minmax_env_with_demo_stats=False):
    """:86-113 -- rescales observations / next_observations IN PLACE exactly as the script does and returns the
    (wrapper_name, wrapper_kwargs) the script hands to the env constructor."""
    if scale_env_with_demo_stats:
        m, sd = stats["obs_mean"], stats["obs_std"]
        for t in traj_list:
            t["observations"] = (t["observations"] - m) / (sd + EPS)
            t["next_observations"] = (t["next_observations"] - m) / (sd + EPS)
        return "ScaledEnv", dict(obs_mean=m, obs_std=sd, acts_mean=None, acts_std=None)
    if minmax_env_with_demo_stats:
        lo, hi = stats["obs_min"], stats["obs_max"]
        for t in traj_list:
            t["observations"] = (t["observations"] - lo) / (hi - lo + EPS)
            t["next_observations"] = (t["next_observations"] - lo) / (hi - lo + EPS)
        return "MinmaxEnv", dict(obs_min=lo, obs_max=hi)
    return "ProxyEnv", {}


def fill_expert_buffer(expert_replay_buffer, traj_list, absorbing=False, env=None):
    """:135-138 -- one add_path per trajectory (one pinned H2D copy each)."""
    for t in traj_list:
        expert_replay_buffer.add_path(t, absorbing=absorbing, env=env)
    return expert_replay_buffer


def ingest(demos_path, expert_replay_buffer, traj_num=None, scale_env_with_demo_stats=False,
           minmax_env_with_demo_stats=False, rng=random):
    """The whole sequence; returns (traj_list, stats, wrapper_name, wrapper_kwargs)."""
    traj_list = load_demos(demos_path, traj_num, rng)
    stats = demo_stats(traj_list)
    name, kw = normalize_demos(traj_list, stats, scale_env_with_demo_stats, minmax_env_with_demo_stats)
    fill_expert_buffer(expert_replay_buffer, traj_list)
    return traj_list, stats, name, kw
