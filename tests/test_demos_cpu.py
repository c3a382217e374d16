"""Expert-demo ingest: host logic of ilswiss_b200/demos.py against the arithmetic of
run_scripts/adv_irl_exp_script.py:51-113 (restated inline) -- no GPU needed."""
import pickle
import random

import numpy as np

from ilswiss_b200 import demos


def _make_demos(path, n_traj=6, T=20, O=5, A=2, seed=0):
    rs = np.random.RandomState(seed)
    trajs = []
    for i in range(n_traj):
        obs = rs.randn(T + 1, O) * (1 + i)
        trajs.append(dict(observations=obs[:-1], next_observations=obs[1:], actions=rs.uniform(-1, 1, (T, A)),
                          rewards=rs.randn(T, 1), terminals=np.zeros((T, 1), dtype=bool)))
    with open(path, "wb") as f:
        pickle.dump(trajs, f)
    return trajs


def test_load_stats_and_normalisation_follow_the_script(tmp_path):
    path = str(tmp_path / "demos.pkl")
    all_trajs = _make_demos(path)
    random.seed(3)
    expect = random.sample(all_trajs, 4)
    random.seed(3)
    got = demos.load_demos(path, 4)
    assert len(got) == 4
    for a, b in zip(got, expect):
        np.testing.assert_array_equal(a["observations"], b["observations"])
    st = demos.demo_stats(got)
    obs = np.vstack([t["observations"] for t in expect])
    np.testing.assert_array_equal(st["obs_mean"], obs.mean(0))
    np.testing.assert_array_equal(st["obs_std"], obs.std(0))
    np.testing.assert_array_equal(st["obs_min"], obs.min(0))
    eps = np.finfo(np.float32).eps.item()
    before = [t["observations"].copy() for t in got]
    name, kw = demos.normalize_demos(got, st, scale_env_with_demo_stats=True)
    assert name == "ScaledEnv" and set(kw) == {"obs_mean", "obs_std", "acts_mean", "acts_std"}
    np.testing.assert_array_equal(got[0]["observations"], (before[0] - obs.mean(0)) / (obs.std(0) + eps))
    got2 = demos.load_demos(path, None)
    st2 = demos.demo_stats(got2)
    b2 = got2[1]["next_observations"].copy()
    name, kw = demos.normalize_demos(got2, st2, minmax_env_with_demo_stats=True)
    assert name == "MinmaxEnv"
    np.testing.assert_array_equal(got2[1]["next_observations"], (b2 - st2["obs_min"]) / (st2["obs_max"] - st2["obs_min"] + eps))
    assert demos.normalize_demos(got2, st2) == ("ProxyEnv", {})


def test_fill_expert_buffer_calls_add_path_per_trajectory(tmp_path):
    path = str(tmp_path / "demos.pkl")
    _make_demos(path, n_traj=3)

    class Buf:
        def __init__(self):
            self.paths = []

        def add_path(self, path, absorbing=False, env=None):
            self.paths.append((len(path["observations"]), absorbing))

    buf = Buf()
    trajs, st, name, kw = demos.ingest(path, buf, traj_num=2, rng=random.Random(1))
    assert buf.paths == [(20, False), (20, False)] and name == "ProxyEnv" and len(trajs) == 2
