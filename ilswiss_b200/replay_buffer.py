"""Device replay buffer with the reference's ReplayBuffer interface.

Drop-in for rlkit.data_management.simple_replay_buffer.SimpleReplayBuffer /
env_replay_buffer.EnvReplayBuffer (flat observations): same constructor arguments, same
method names, argument meaning and return types -- but the transitions live in HBM
(ilswiss_b200.engine.ReplayRing) and `random_batch` is one gather kernel.

Inject it through the reference's own, supported constructor argument
`BaseAlgorithm(replay_buffer=...)` (rlkit/core/base_algorithm.py:36,116-123).
"""
import numpy as np
import torch

from . import layout
from .engine import ReplayRing

ALL_KEYS = ("observations", "actions", "rewards", "terminals", "next_observations", "absorbing")


class DeviceReplayBuffer:
    """simple_replay_buffer.py:17-323 (int observation_dim only; dict/image observations are
    outside the hot path and raise)."""

    def __init__(self, max_replay_buffer_size, observation_dim, action_dim, random_seed=1995,
                 flush_threshold=4096):
        if not isinstance(observation_dim, (int, np.integer)):
            raise NotImplementedError("DeviceReplayBuffer supports flat (int) observation_dim only")
        self._np_rand_state = np.random.RandomState(random_seed)   # :20 -- same index stream
        self._observation_dim = int(observation_dim)
        self._action_dim = int(action_dim)
        self._max_replay_buffer_size = int(max_replay_buffer_size)
        self.ring = ReplayRing(self._max_replay_buffer_size, self._observation_dim, self._action_dim)
        self._pending = []            # host rows not yet staged (episode-burst appends)
        self._flush_threshold = max(1, min(int(flush_threshold), self._max_replay_buffer_size))   # a burst never exceeds the ring
        self._top = 0
        self._size = 0
        self._trajs = 0
        self._cur_start = 0
        self._traj_endpoints = {}
        self._philox_counter = 0

    # -- append side (R2) ----------------------------------------------------------------
    def add_sample(self, observation, action, reward, terminal, next_observation, timeout=False, **kwargs):
        """:78-108.  Values are rounded to float32 here exactly as the reference rounds them at
        sample time (np_to_pytorch_batch, rlkit/torch/core.py:124-143)."""
        O, A = self._observation_dim, self._action_dim
        row = np.empty(layout.host_row_floats(O, A), dtype=np.float32)
        row[:O] = np.asarray(observation, dtype=np.float64).reshape(O)
        row[O:O + A] = np.asarray(action, dtype=np.float64).reshape(A)
        row[O + A] = float(np.asarray(reward).reshape(-1)[0])
        term = bool(np.asarray(terminal).reshape(-1)[0])
        row[O + A + 1] = 1.0 if term else 0.0
        row[O + A + 2:2 * O + A + 2] = np.asarray(next_observation, dtype=np.float64).reshape(O)
        ab = kwargs.get("absorbing")
        row[2 * O + A + 2:2 * O + A + 4] = 0.0 if ab is None else np.asarray(ab, dtype=np.float64).reshape(2)
        row[2 * O + A + 4] = 1.0 if timeout else 0.0
        self._pending.append(row)
        if term:
            next_start = (self._top + 1) % self._max_replay_buffer_size
            self._traj_endpoints[self._cur_start] = next_start
            self._cur_start = next_start
        self._advance()
        if len(self._pending) >= self._flush_threshold:
            self.flush()

    def add_samples(self, observations, actions, rewards, terminals, next_observations, absorbing=None, timeouts=None):
        """Vectorised append of n transitions (not in the reference; same result as n add_sample
        calls without terminal bookkeeping side effects other than sizes)."""
        self.flush()
        rows = layout.pack_host_rows(observations, actions, rewards, terminals, next_observations, absorbing, timeouts)
        n = rows.shape[0]
        self.ring.append_host(rows)
        self._top = (self._top + n) % self._max_replay_buffer_size
        self._size = min(self._size + n, self._max_replay_buffer_size)

    def load_device_rows(self, hot_rows):
        """Bulk fill from a CUDA tensor already in the hot-row layout [n, stride] (synthetic
        benchmarks, snapshot restore): one D2D copy, no host involvement."""
        self.flush()
        n = hot_rows.shape[0]
        self.ring.load_device(hot_rows)
        self._top = (self._top + n) % self._max_replay_buffer_size
        self._size = min(self._size + n, self._max_replay_buffer_size)

    def _advance(self):
        # :228-237
        if self._top in self._traj_endpoints:
            del self._traj_endpoints[self._top]
        self._top = (self._top + 1) % self._max_replay_buffer_size
        if self._size < self._max_replay_buffer_size:
            self._size += 1

    def terminate_episode(self):
        # :125-132, plus: the finished episode is staged to the GPU on the side stream
        if self._cur_start != self._top:
            self._traj_endpoints[self._cur_start] = self._top
            self._cur_start = self._top
        self.flush()

    def add_path(self, path, absorbing=False, env=None):
        # :134-216
        if absorbing:
            raise NotImplementedError("wrap_absorbing is rejected by the reference itself (base_algorithm.py:137-139)")
        # same result as the reference's per-transition add_sample loop; the rows of the whole path are packed in one
        # vectorised pass and go to the GPU as ONE pinned copy when the episode is terminated
        n = len(path["observations"])
        if n:
            terms = np.asarray(path["terminals"]).reshape(n).astype(bool)
            rows = layout.pack_host_rows(np.asarray(path["observations"]).reshape(n, -1), np.asarray(path["actions"]).reshape(n, -1),
                                         np.asarray(path["rewards"]).reshape(n), terms,
                                         np.asarray(path["next_observations"]).reshape(n, -1))
            self._pending.extend(rows)
            for t in range(n):
                if terms[t]:
                    next_start = (self._top + 1) % self._max_replay_buffer_size
                    self._traj_endpoints[self._cur_start] = next_start
                    self._cur_start = next_start
                self._advance()
        self.terminate_episode()
        self._trajs += 1

    def save_data(self, save_name):
        """:110-123 -- the reference's on-disk dump (fields up to _top), readable by its own tools."""
        import pickle

        d = self.__getstate__()
        top = self._top
        save_dict = {
            "observations": d["_observations"][:top], "actions": d["_actions"][:top],
            "next_observations": d["_next_obs"][:top], "terminals": d["_terminals"][:top],
            "timeouts": np.zeros((top, 1), dtype="uint8"), "rewards": d["_rewards"][:top],
            "agent_infos": [None] * top, "env_infos": [None] * top,
        }
        with open(save_name, "wb") as f:
            pickle.dump(save_dict, f)

    def get_traj_num(self):
        return self._trajs

    def flush(self):
        """Stage pending host rows: pinned cudaMemcpyAsync on the ring's side stream; they enter
        the ring (scatter kernel on the compute stream) right before the next sample / train call."""
        if self._pending:
            rows = np.stack(self._pending)
            self._pending = []
            cap = self._max_replay_buffer_size
            for s0 in range(0, rows.shape[0], cap):     # a burst never exceeds the ring (ilsw_rb_append)
                self.ring.append_host(rows[s0:s0 + cap])

    # -- sample side (R3/R4) -------------------------------------------------------------
    def num_steps_can_sample(self):
        return self._size

    def sample_indices(self, batch_size):
        """:242 -- RandomState.randint(0, size, B): uniform WITH replacement, same stream as the reference."""
        return self._np_rand_state.randint(0, self._size, batch_size)

    def random_batch(self, batch_size, keys=None, multi_step=False, step_num=1, **kwargs):
        """:239-293.  Returns the reference's dict of numpy arrays (float64 / uint8 dtypes); the
        values are the float32-rounded ones the reference feeds to its nets."""
        if multi_step:
            raise NotImplementedError("multi_step sampling is not on the hot path")
        return self._get_batch_using_indices(self.sample_indices(batch_size), keys=keys)

    def _get_batch_using_indices(self, indices, keys=None, **kwargs):
        if keys is None:
            keys = set(ALL_KEYS)
        self.flush()
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int32), device="cuda")
        hot, cold = self.ring.gather(idx)
        full = layout.unpack_hot_rows(hot.cpu().numpy(), self._observation_dim, self._action_dim)
        full["absorbing"] = cold[:, :2].cpu().numpy().astype(np.float64)
        return {k: v for k, v in full.items() if k in keys}

    def random_batch_device(self, batch_size, indices=None):
        """Device-resident batch (dict of CUDA float32 tensors, reference key names): what
        np_to_pytorch_batch(random_batch(B)) yields in the reference, without leaving HBM."""
        self.flush()
        if indices is None:
            indices = self.sample_indices(batch_size)
        idx = torch.as_tensor(np.asarray(indices, dtype=np.int32), device="cuda")
        hot, cold = self.ring.gather(idx)
        O, A = self._observation_dim, self._action_dim
        return {
            "observations": hot[:, :O], "actions": hot[:, O:O + A], "rewards": hot[:, O + A:O + A + 1],
            "terminals": hot[:, O + A + 1:O + A + 2], "next_observations": hot[:, O + A + 2:2 * O + A + 2],
            "absorbing": cold[:, :2],
        }

    def get_all(self, keys=None, **kwargs):
        return self._get_batch_using_indices(np.arange(self._size), keys=keys)

    def clear(self):
        self._pending = []
        self.ring.clear()
        self._top = self._size = self._cur_start = 0
        self._traj_endpoints = {}

    # -- snapshots (base_algorithm.py:562-580 save_replay_buffer) ---------------------------
    def __getstate__(self):
        """Downloads to the reference's numpy field layout so `extra_data.pkl` stays readable."""
        self.flush()
        self.ring.commit()
        nrows = self._max_replay_buffer_size
        rows = self.ring.rows_view()[:nrows].cpu().numpy()
        cold = (self.ring.gather(torch.arange(nrows, dtype=torch.int32, device="cuda"))[1].cpu().numpy()
                if self.ring.committed_size > 0 else np.zeros((nrows, 4), np.float32))
        d = layout.unpack_hot_rows(rows, self._observation_dim, self._action_dim)
        return dict(
            _observation_dim=self._observation_dim, _action_dim=self._action_dim,
            _max_replay_buffer_size=self._max_replay_buffer_size, _observations=d["observations"],
            _next_obs=d["next_observations"], _actions=d["actions"], _rewards=d["rewards"],
            _terminals=d["terminals"], _absorbing=cold[:, :2].astype(np.float64), _timeouts=(cold[:, 2:3] != 0).astype("uint8"),
            _top=self._top, _size=self._size, _trajs=self._trajs,
            _cur_start=self._cur_start, _traj_endpoints=dict(self._traj_endpoints),
            _np_rand_state=self._np_rand_state, _flush_threshold=self._flush_threshold,
        )

    def __setstate__(self, d):
        # the BASE constructor explicitly: subclasses have other signatures (env / goal dims) and restore their own fields
        DeviceReplayBuffer.__init__(self, d["_max_replay_buffer_size"], d["_observation_dim"], d["_action_dim"],
                                    flush_threshold=d.get("_flush_threshold", 4096))
        self._np_rand_state = d["_np_rand_state"]
        n = d["_size"]
        cap = self._max_replay_buffer_size
        # restore the physical layout (slot i holds row i, cold side array included), then the ring cursor
        rows = layout.pack_host_rows(d["_observations"][:cap], d["_actions"][:cap], np.asarray(d["_rewards"][:cap]).reshape(-1),
                                     np.asarray(d["_terminals"][:cap]).reshape(-1), d["_next_obs"][:cap],
                                     d.get("_absorbing", None) if d.get("_absorbing", None) is None else d["_absorbing"][:cap],
                                     None if d.get("_timeouts", None) is None else np.asarray(d["_timeouts"][:cap]).reshape(-1))
        self.ring.clear()
        step = max(1, min(1 << 16, cap))
        for s0 in range(0, rows.shape[0], step):
            self.ring.append_host(rows[s0:s0 + step])
            self.ring.commit()
        self.ring.set_cursor(d["_top"], n)
        self._top, self._size, self._trajs = d["_top"], n, d["_trajs"]
        self._cur_start, self._traj_endpoints = d["_cur_start"], dict(d["_traj_endpoints"])


class DeviceEnvReplayBuffer(DeviceReplayBuffer):
    """env_replay_buffer.py:7-49: dims taken from env.observation_space / env.action_space."""

    def __init__(self, max_replay_buffer_size, env, random_seed=1995):
        self._ob_space = env.observation_space
        self._action_space = env.action_space
        super().__init__(max_replay_buffer_size, get_dim(self._ob_space), get_dim(self._action_space), random_seed)

    def __getstate__(self):
        d = super().__getstate__()
        d["_ob_space"], d["_action_space"] = self._ob_space, self._action_space
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self._ob_space, self._action_space = d.get("_ob_space"), d.get("_action_space")


def get_dim(space):
    """env_replay_buffer.py:35-49 for Box / Discrete spaces (duck-typed, no gym import)."""
    name = type(space).__name__
    if name == "Discrete":
        return 1
    if hasattr(space, "low"):
        low = np.asarray(space.low)
        if low.ndim > 1:
            raise NotImplementedError("image observations are not on the hot path")
        return int(low.size)
    if hasattr(space, "flat_dim"):
        return int(space.flat_dim)
    raise TypeError("Unknown space: {}".format(space))


class DeviceHindsightReplayBuffer(DeviceReplayBuffer):
    """rlkit/data_management/relabel_replay_buffer.py:12-131 (HindsightReplayBuffer, relabel_type "future" / her_ratio) with
    the transitions in HBM: ring rows hold obs = cat(observation, desired_goal) / next_obs likewise, a device side array
    holds every slot's next achieved goal, and the trajectory table of finished episodes is mirrored on the device, so that
    the step program samples trajectory -> step -> future step, relabels and recomputes the sparse goal reward INSIDE its
    gather phase (`HerTD3.train_from_buffer` / `ilsw_trainer_set_her`).

    Observations are the goal environments' dicts (keys observation / achieved_goal / desired_goal).  The reward rule is
    the sparse one every shipped HER yaml's environment uses, -(||achieved - desired|| > distance_threshold)
    (gym robotics `compute_reward`, which the reference takes from `env.compute_reward`, :36-37).
    random_batch() keeps the reference's host semantics (same RNG streams: the buffer's RandomState for the trajectory
    shuffle / trajectory / step draws, the global numpy RNG for the future step, :70-95)."""

    def __init__(self, max_replay_buffer_size, obs_dim, goal_dim, action_dim, random_seed=1995, relabel_type="future",
                 her_ratio=0.8, distance_threshold=0.05, observation_key="observation", desired_goal_key="desired_goal",
                 achieved_goal_key="achieved_goal"):
        if relabel_type not in ("future", None):
            raise NotImplementedError("relabel_type %r (the shipped yamls use 'future')" % (relabel_type,))
        super().__init__(max_replay_buffer_size, int(obs_dim) + int(goal_dim), action_dim, random_seed)
        self._obs0_dim, self._goal_dim = int(obs_dim), int(goal_dim)
        self.relabel_type, self.her_ratio, self.distance_threshold = relabel_type, float(her_ratio), float(distance_threshold)
        self.observation_key, self.desired_goal_key, self.achieved_goal_key = observation_key, desired_goal_key, achieved_goal_key
        self._ag_next = torch.zeros((self._max_replay_buffer_size, self._goal_dim), dtype=torch.float32, device="cuda")
        self._ag_pending = []         # (slot, next achieved goal) of rows not yet staged
        self._traj_dev = None         # (starts, lens) int32 CUDA tensors, rebuilt when the table changes
        self._traj_dirty = True

    # the flat bulk-append paths of the base class know nothing about goal dicts, next achieved goals or the trajectory table
    def add_samples(self, *a, **k):
        raise NotImplementedError("DeviceHindsightReplayBuffer: append goal-dict transitions through add_sample")

    def load_device_rows(self, *a, **k):
        raise NotImplementedError("DeviceHindsightReplayBuffer: append goal-dict transitions through add_sample")

    def add_path(self, path, absorbing=False, env=None):
        """simple_replay_buffer.py:134-216 on goal-dict observations: the reference's per-transition add_sample loop."""
        if absorbing:
            raise NotImplementedError("wrap_absorbing is rejected by the reference itself (base_algorithm.py:137-139)")
        for ob, ac, rw, nob, tm in zip(path["observations"], path["actions"], path["rewards"], path["next_observations"], path["terminals"]):
            self.add_sample(ob, ac, rw, tm, nob)
        self.terminate_episode()
        self._trajs += 1

    def __getstate__(self):
        d = super().__getstate__()
        d.update(_obs0_dim=self._obs0_dim, _goal_dim=self._goal_dim, relabel_type=self.relabel_type, her_ratio=self.her_ratio,
                 distance_threshold=self.distance_threshold, observation_key=self.observation_key,
                 desired_goal_key=self.desired_goal_key, achieved_goal_key=self.achieved_goal_key,
                 _ag_next=self._ag_next.cpu().numpy())
        return d

    def __setstate__(self, d):
        DeviceReplayBuffer.__setstate__(self, d)
        self._obs0_dim, self._goal_dim = d["_obs0_dim"], d["_goal_dim"]
        self.relabel_type, self.her_ratio, self.distance_threshold = d["relabel_type"], d["her_ratio"], d["distance_threshold"]
        self.observation_key, self.desired_goal_key, self.achieved_goal_key = d["observation_key"], d["desired_goal_key"], d["achieved_goal_key"]
        self._ag_next = torch.from_numpy(np.ascontiguousarray(d["_ag_next"], dtype=np.float32)).cuda()
        self._ag_pending, self._traj_dev, self._traj_dirty = [], None, True

    def _cat(self, obs):
        return np.concatenate([np.asarray(obs[self.observation_key], dtype=np.float64).ravel(),
                               np.asarray(obs[self.desired_goal_key], dtype=np.float64).ravel()])

    def add_sample(self, observation, action, reward, terminal, next_observation, **kwargs):
        assert isinstance(observation, dict), "Observation should be dict!"           # :53
        self._ag_pending.append((self._top, np.asarray(next_observation[self.achieved_goal_key], dtype=np.float32).ravel()))
        n_before = len(self._traj_endpoints)
        super().add_sample(self._cat(observation), action, reward, terminal, self._cat(next_observation), **kwargs)
        self._traj_dirty = self._traj_dirty or len(self._traj_endpoints) != n_before or bool(np.asarray(terminal).reshape(-1)[0])

    def terminate_episode(self):
        super().terminate_episode()
        self._traj_dirty = True

    def flush(self):
        DeviceReplayBuffer.flush(self)
        if self._ag_pending:
            slots = torch.as_tensor(np.array([s for s, _ in self._ag_pending], dtype=np.int64), device="cuda")
            vals = torch.as_tensor(np.stack([v for _, v in self._ag_pending]), device="cuda")
            self._ag_next[slots] = vals
            self._ag_pending = []

    def trajectory_table(self):
        """(starts, lens) of the finished trajectories as int32 CUDA tensors (:68-75: [start, end) modulo the fill level)."""
        if self._traj_dirty or self._traj_dev is None:
            starts = np.array(list(self._traj_endpoints.keys()), dtype=np.int32)
            lens = np.array([(self._traj_endpoints[int(s)] - int(s)) % self._size for s in starts], dtype=np.int32)
            keep = lens > 0
            self._traj_dev = (torch.as_tensor(starts[keep], device="cuda"), torch.as_tensor(lens[keep], device="cuda"))
            self._traj_dirty = False
        return self._traj_dev

    def her_desc(self, batch_size, inj_idx_her=None):
        """The ilsw_her_sampling descriptor of this buffer for a trainer with `batch_size` rows per step."""
        from . import _abi

        self.flush()
        starts, lens = self.trajectory_table()
        d = _abi.HerSamplingDesc()
        d.enabled, d.n_traj = 1, int(starts.numel())
        d.traj_start, d.traj_len, d.next_achieved_goal = starts.data_ptr(), lens.data_ptr(), self._ag_next.data_ptr()
        d.goal_dim, d.distance_threshold = self._goal_dim, self.distance_threshold
        relabel = (self.relabel_type is not None) and (self.her_ratio > 0)
        d.relabel_num = int(self.her_ratio * batch_size) if relabel else 0
        d.inj_idx_her = inj_idx_her.data_ptr() if inj_idx_her is not None else None
        self._desc_keep = (starts, lens, inj_idx_her)
        return d

    # -- host-facing sampling with the reference's semantics ---------------------------------
    def sample_indices(self, batch_size):
        relabel = (self.relabel_type is not None) and (self.her_ratio > 0)
        keys_list = list(self._traj_endpoints.keys())
        starts = self._np_rand_state.choice(keys_list, size=len(keys_list), replace=False)
        ends = [self._traj_endpoints[k] for k in starts]
        traj_indice = self._np_rand_state.randint(0, len(starts), batch_size)
        indices, indices_relabel = [], []
        for i in traj_indice:
            traj_len = (ends[i] - starts[i]) % self._size
            step = (self._np_rand_state.randint(0, traj_len, 1)[0] + starts[i]) % self._size
            indices.append(step)
            if relabel:
                indices_relabel.append(np.random.randint(step, (traj_len + starts[i])) % self._size)
        return np.asarray(indices), np.asarray(indices_relabel)

    def random_batch(self, batch_size, keys=None, **kwargs):
        """:63-131 -- the reference's goal-conditioned batch dict (observations / desired_goals / achieved_goals / next_* /
        actions / rewards / terminals), relabelled."""
        idx, idx_her = self.sample_indices(batch_size)
        self.flush()
        O0 = self._obs0_dim
        full = self._get_batch_using_indices(idx, keys=None)
        ag_next = self._ag_next[torch.as_tensor(idx, device="cuda", dtype=torch.long)].cpu().numpy().astype(np.float64)
        out = dict(actions=full["actions"], terminals=full["terminals"], rewards=full["rewards"],
                   observations=full["observations"][:, :O0], next_observations=full["next_observations"][:, :O0],
                   desired_goals=full["observations"][:, O0:].copy(), next_desired_goals=full["next_observations"][:, O0:].copy(),
                   next_achieved_goals=ag_next)
        # :116-118 achieved_goals = the achieved goal of the CURRENT observation: the next achieved goal of the previous slot
        # of the same trajectory (the ring keeps only next achieved goals); the first step of a trajectory has no
        # predecessor in the ring, so the key is only provided on request
        if keys is not None and "achieved_goals" in keys:
            raise NotImplementedError("achieved_goals of the current observation are not stored in the device ring")
        if len(idx_her):
            n = int(self.her_ratio * batch_size)
            src = self._ag_next[torch.as_tensor(idx_her, device="cuda", dtype=torch.long)].cpu().numpy().astype(np.float64)
            out["desired_goals"][:n] = src[:n]
            out["next_desired_goals"][:n] = src[:n]
            d = np.linalg.norm(out["next_achieved_goals"] - out["desired_goals"], axis=-1)
            out["rewards"] = (-(d > self.distance_threshold).astype(np.float32)).reshape(-1, 1)
        return out


def _check_sparse_reward_rule(env, goal_dim, thr):
    """relabel_replay_buffer.py:36-37,139 recomputes rewards with env.compute_reward; the device path hard-wires the sparse
    rule of the gym robotics environments the shipped HER yamls use.  Probe the environment's function on a few points and
    refuse anything else (a dense or custom reward would otherwise be relabelled silently wrong)."""
    fn = getattr(env, "compute_reward", None)
    if fn is None:
        return
    rs = np.random.RandomState(0)
    ag = rs.uniform(-1, 1, (16, goal_dim))
    g = ag + rs.uniform(-2 * thr, 2 * thr, (16, goal_dim)) / np.sqrt(goal_dim)
    try:
        got = np.asarray(fn(ag, g, None), dtype=np.float64).reshape(-1)
    except Exception:
        return                      # cannot be probed without an episode context: trust the threshold attribute
    want = -(np.linalg.norm(ag - g, axis=-1) > thr).astype(np.float64)
    if got.shape != want.shape or not np.array_equal(got, want):
        raise NotImplementedError("env.compute_reward is not the sparse goal reward -(||ag - g|| > distance_threshold): use the "
                                  "reference's host HindsightReplayBuffer with HerTD3.train_step(batch) for this environment")


class DeviceEnvHindsightReplayBuffer(DeviceHindsightReplayBuffer):
    """HindsightReplayBuffer's own constructor signature (relabel_replay_buffer.py:13-48): dims from the goal environment's
    Dict observation space, the sparse-reward threshold from `env.distance_threshold` (gym robotics; default 0.05)."""

    def __init__(self, max_replay_buffer_size, env, random_seed=1995, relabel_type="future", her_ratio=0.8,
                 observation_key="observation", desired_goal_key="desired_goal", achieved_goal_key="achieved_goal"):
        spaces = env.observation_space.spaces
        self._ob_space, self._action_space = env.observation_space, env.action_space
        base_env = getattr(env, "unwrapped", env)
        if not hasattr(base_env, "distance_threshold") and not hasattr(env, "distance_threshold"):
            raise NotImplementedError("DeviceEnvHindsightReplayBuffer: the environment exposes no distance_threshold; the device "
                                      "relabel implements the sparse goal reward -(||ag - g|| > threshold) only")
        thr = float(getattr(base_env, "distance_threshold", getattr(env, "distance_threshold", 0.05)))
        _check_sparse_reward_rule(env, get_dim(spaces[desired_goal_key]), thr)
        super().__init__(max_replay_buffer_size, get_dim(spaces[observation_key]), get_dim(spaces[desired_goal_key]),
                         get_dim(self._action_space), random_seed=random_seed, relabel_type=relabel_type, her_ratio=her_ratio,
                         distance_threshold=thr, observation_key=observation_key, desired_goal_key=desired_goal_key,
                         achieved_goal_key=achieved_goal_key)

    def __getstate__(self):
        d = super().__getstate__()
        d["_ob_space"], d["_action_space"] = self._ob_space, self._action_space
        return d

    def __setstate__(self, d):
        super().__setstate__(d)
        self._ob_space, self._action_space = d.get("_ob_space"), d.get("_action_space")
