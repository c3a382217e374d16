"""The product's DeviceReplayBuffer HOST logic (cursor, size, trajectory table, burst staging and its chunking, row packing)
against the reference's own SimpleReplayBuffer under random operation sequences -- the device ring is replaced by a numpy
ring with the C ABI's append semantics (slot = (top + i) % capacity), so no GPU is needed.  The gather / sample side is
covered on the GPU (tests/test_gpu_replay.py)."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="reference checkout absent")


class NumpyRing:
    """What engine.ReplayRing does to host rows (ilsw_rb_append + ilsw_rb_commit), in numpy."""

    def __init__(self, capacity, obs_dim, act_dim):
        self.cap, self.W = capacity, 2 * obs_dim + act_dim + 5
        self.rows = np.zeros((capacity, self.W), np.float32)
        self.top = self.size = 0
        self.bursts = []

    def append_host(self, rows):
        assert rows.dtype == np.float32 and rows.shape[1] == self.W and 0 < rows.shape[0] <= self.cap     # ilsw_rb_append
        self.bursts.append(rows.shape[0])
        for r in rows:
            self.rows[self.top] = r
            self.top = (self.top + 1) % self.cap
        self.size = min(self.size + rows.shape[0], self.cap)

    def clear(self):
        self.top = self.size = 0

    def commit(self):
        pass

    def set_cursor(self, top, size):
        assert 0 <= top < self.cap and 0 <= size <= self.cap      # ilsw_rb_set_cursor
        self.top, self.size = top, size

    # the sample side, as ilsw_rb_gather lays it out: hot rows [n, stride] + cold rows [n, 4] (absorbing x2, timeout, 0)
    @property
    def committed_size(self):
        return self.size

    def gather(self, idx):
        import torch

        i = np.asarray(idx, dtype=np.int64)
        hot_w = self.W - 3
        hot = np.zeros((len(i), (hot_w + 15) // 16 * 16), np.float32)
        hot[:, :hot_w] = self.rows[i, :hot_w]
        cold = np.zeros((len(i), 4), np.float32)
        cold[:, :3] = self.rows[i, hot_w:]
        return torch.from_numpy(hot), torch.from_numpy(cold)

    def rows_view(self):
        import torch

        return self.gather(np.arange(self.cap))[0]


class _TorchOnCpu:
    """`torch` as replay_buffer.py sees it, with device="cuda" requests served on the CPU (no GPU in this container)."""

    def __getattr__(self, name):
        import torch

        return getattr(torch, name)

    @staticmethod
    def _strip(kw):
        kw.pop("device", None)
        return kw

    def zeros(self, *a, **kw):
        import torch

        return torch.zeros(*a, **self._strip(kw))

    def as_tensor(self, *a, **kw):
        import torch

        return torch.as_tensor(*a, **self._strip(kw))

    def arange(self, *a, **kw):
        import torch

        return torch.arange(*a, **self._strip(kw))


def _drive(cap, O, A, ops, flush_threshold, monkeypatch):
    ref_shim.install()
    from rlkit.data_management.simple_replay_buffer import SimpleReplayBuffer

    import ilswiss_b200.replay_buffer as rb

    monkeypatch.setattr(rb, "ReplayRing", NumpyRing)
    monkeypatch.setattr(rb, "torch", _TorchOnCpu())
    ref = SimpleReplayBuffer(cap, O, A, random_seed=5)
    dev = rb.DeviceReplayBuffer(cap, O, A, random_seed=5, flush_threshold=flush_threshold)
    rs = np.random.RandomState(len(ops) + cap)
    for op in ops:
        if op == "term":
            ref.terminate_episode(); dev.terminate_episode()
        elif isinstance(op, tuple):                      # ("path", T): a whole trajectory, last transition terminal or not
            T, last_terminal = op[1], op[2]
            terms = np.zeros((T, 1)); terms[-1, 0] = float(last_terminal)
            path = dict(observations=rs.randn(T, O), actions=rs.uniform(-1, 1, (T, A)), rewards=rs.randn(T, 1),
                        next_observations=rs.randn(T, O), terminals=terms, absorbings=np.zeros((T, 2)),
                        env_infos=[{}] * T, agent_infos=[{}] * T)
            ref.add_path(path); dev.add_path(path)
        else:                                            # one transition, terminal with probability op
            o, a, r, no = rs.randn(O), rs.uniform(-1, 1, A), rs.randn(), rs.randn(O)
            d = bool(rs.rand() < op)
            ref.add_sample(o, a, r, d, no); dev.add_sample(o, a, r, d, no)
        assert (dev._top, dev._size, dev._cur_start) == (ref._top, ref._size, ref._cur_start)
        assert dev._traj_endpoints == ref._traj_endpoints
        assert dev.num_steps_can_sample() == ref.num_steps_can_sample() and dev.get_traj_num() == ref.get_traj_num()
    dev.flush()
    return ref, dev


@settings(max_examples=60, deadline=None)
@given(cap=st.integers(3, 40), O=st.integers(1, 6), A=st.integers(1, 3), thr=st.sampled_from([1, 4, 4096]),
       ops=st.lists(st.one_of(st.sampled_from([0.0, 0.1, 0.5]), st.just("term"),
                              st.tuples(st.just("path"), st.integers(1, 12), st.booleans())), min_size=1, max_size=60))
def test_bookkeeping_and_staged_rows_equal_the_reference_buffer(cap, O, A, thr, ops):
    with pytest.MonkeyPatch.context() as mp:
        ops = [op for op in ops if not (isinstance(op, tuple) and op[1] >= cap)]      # the reference's own precondition (:130-131)
        if not ops:
            return
        ref, dev = _drive(cap, O, A, ops, thr, mp)
        _compare_with_reference(ref, dev, cap, O, A)


def _compare_with_reference(ref, dev, cap, O, A):
    ring = dev.ring
    assert (ring.top, ring.size) == (ref._top, ref._size)
    assert all(n <= cap for n in ring.bursts)
    # every slot the reference holds, rounded to float32 as np_to_pytorch_batch does (rlkit/torch/core.py:124-143)
    n = ref._size
    f32 = lambda x: np.asarray(x, dtype=np.float32)
    np.testing.assert_array_equal(ring.rows[:n, :O], f32(ref._observations[:n]))
    np.testing.assert_array_equal(ring.rows[:n, O:O + A], f32(ref._actions[:n]))
    np.testing.assert_array_equal(ring.rows[:n, O + A], f32(ref._rewards[:n, 0]))
    np.testing.assert_array_equal(ring.rows[:n, O + A + 1], f32(ref._terminals[:n, 0]))
    np.testing.assert_array_equal(ring.rows[:n, O + A + 2:2 * O + A + 2], f32(ref._next_obs[:n]))
    # and the index stream of the next random_batch is the reference's
    if ref._size:
        np.testing.assert_array_equal(dev.sample_indices(7), ref._np_randint(0, ref._size, 7))
        # random_batch: the reference's dict (simple_replay_buffer.py:239-293) -- keys, dtypes, shapes, float32-rounded values
        for keys in (None, ["observations", "actions"]):
            want, got = ref.random_batch(9, keys=keys), dev.random_batch(9, keys=keys)
            assert set(got.keys()) == set(want.keys())
            for k in want:
                assert got[k].dtype == want[k].dtype and got[k].shape == want[k].shape, k
                np.testing.assert_array_equal(got[k].astype(np.float32), want[k].astype(np.float32), err_msg=k)
        want, got = ref.get_all(), dev.get_all()
        for k in want:
            np.testing.assert_array_equal(got[k].astype(np.float32), want[k].astype(np.float32), err_msg=k)


@pytest.mark.parametrize("n_steps", [17, 40, 95])          # partly filled, exactly full, wrapped twice
def test_reference_buffer_state_restores_into_the_device_classes(n_steps, monkeypatch):
    """base_algorithm.py:562-580 save_replay_buffer -> extra_data.pkl -> resume: the state of the reference's own
    EnvReplayBuffer (its attribute dict is what it pickles) loads into DeviceEnvReplayBuffer / DeviceReplayBuffer through
    __setstate__ -- which must not go through the subclass constructor (env-based signature) -- with every slot, the cursor,
    the trajectory table, the spaces and the RandomState stream carried over."""
    import pickle

    ref_shim.install()
    from rlkit.data_management.env_replay_buffer import EnvReplayBuffer

    import ilswiss_b200.replay_buffer as rb

    monkeypatch.setattr(rb, "ReplayRing", NumpyRing)
    O, A, cap = 4, 2, 40
    env = ref_shim.FakeEnv(O, A)
    ref = EnvReplayBuffer(cap, env, random_seed=9)
    rs = np.random.RandomState(1)
    for i in range(n_steps):
        ref.add_sample(rs.randn(O), rs.uniform(-1, 1, A), rs.randn(), bool(i % 11 == 10), rs.randn(O), timeout=bool(i % 11 == 10),
                       absorbing=np.array([0.0, float(i % 2)]))
    # the attribute dict is what extra_data.pkl holds (the gym stub spaces of the shim are not picklable: copy field by field)
    state = {k: (pickle.loads(pickle.dumps(v)) if k not in ("_ob_space", "_action_space") else v) for k, v in ref.__dict__.items()}
    for cls in (rb.DeviceEnvReplayBuffer, rb.DeviceReplayBuffer):
        dev = cls.__new__(cls)
        dev.__setstate__({k: (pickle.loads(pickle.dumps(v)) if k == "_np_rand_state" else v) for k, v in state.items()})
        assert type(dev.ring) is NumpyRing and (dev.ring.top, dev.ring.size) == (ref._top, ref._size)
        assert (dev._top, dev._size, dev._cur_start, dev._trajs) == (ref._top, ref._size, ref._cur_start, ref._trajs)
        assert dev._traj_endpoints == ref._traj_endpoints and dev._max_replay_buffer_size == cap
        n = ref._size
        np.testing.assert_array_equal(dev.ring.rows[:n, :O], ref._observations[:n].astype(np.float32))
        np.testing.assert_array_equal(dev.ring.rows[:n, O + A + 1], ref._terminals[:n, 0].astype(np.float32))
        np.testing.assert_array_equal(dev.ring.rows[:n, O + A + 2:2 * O + A + 2], ref._next_obs[:n].astype(np.float32))
        np.testing.assert_array_equal(dev.ring.rows[:n, 2 * O + A + 2:2 * O + A + 4], ref._absorbing[:n].astype(np.float32))
        np.testing.assert_array_equal(dev.ring.rows[:n, 2 * O + A + 4], ref._timeouts[:n, 0].astype(np.float32))
        np.testing.assert_array_equal(dev.sample_indices(5), pickle.loads(pickle.dumps(ref._np_rand_state)).randint(0, n, 5))
        if cls is rb.DeviceEnvReplayBuffer:
            assert dev._ob_space is not None and dev._action_space is not None
        # appending goes on where the reference stopped
        dev.add_sample(np.zeros(O), np.zeros(A), 1.5, False, np.ones(O))
        dev.flush()
        assert dev._top == (ref._top + 1) % cap and dev.ring.rows[ref._top, O + A] == 1.5


@settings(max_examples=25, deadline=None)
@given(cap=st.integers(20, 200), n_ep=st.integers(2, 12), T=st.integers(2, 19), ratio=st.sampled_from([0.3, 0.8, 1.0]),
       G=st.integers(1, 4), seed=st.integers(0, 1000))
def test_hindsight_buffer_random_episode_tables(cap, n_ep, T, ratio, G, seed):
    """The same comparison over random ring sizes / episode counts and lengths (rings that wrapped several times included),
    goal dims and her_ratio."""
    with pytest.MonkeyPatch.context() as mp:
        _hindsight_vs_reference(cap, n_ep, T, mp, ratio=ratio, G=G, seed=seed, round_trip=False)


@pytest.mark.parametrize("cap, n_ep, T", [(400, 7, 50), (130, 9, 20), (61, 12, 7)])      # no wrap, wrapped once, wrapped often
def test_hindsight_buffer_host_path_equals_the_reference_buffer(cap, n_ep, T, monkeypatch):
    _hindsight_vs_reference(cap, n_ep, T, monkeypatch)


def _hindsight_vs_reference(cap, n_ep, T, monkeypatch, ratio=0.8, G=3, seed=0, round_trip=True):
    """DeviceEnvHindsightReplayBuffer.random_batch (relabel_replay_buffer.py:63-131 on the host: trajectory shuffle,
    trajectory / step / future-step draws from the reference's two RNG streams, goal substitution, sparse reward) against the
    executed HindsightReplayBuffer, including rings that wrapped over old episodes -- and the same after a pickle-state
    round trip of the device class (its own __getstate__ / __setstate__, env-based constructor not re-run)."""
    import torch

    from oracle.restate import synth_goal_episodes

    ref_shim.install()
    from rlkit.data_management.relabel_replay_buffer import HindsightReplayBuffer

    import ilswiss_b200.replay_buffer as rb

    monkeypatch.setattr(rb, "ReplayRing", NumpyRing)
    monkeypatch.setattr(rb, "torch", _TorchOnCpu())
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)
    O0, A = 6, 2
    env = ref_shim.FakeGoalEnv(O0, G, A)
    ref = HindsightReplayBuffer(cap, env, random_seed=5, relabel_type="future", her_ratio=ratio)
    dev = rb.DeviceEnvHindsightReplayBuffer(cap, env, random_seed=5, relabel_type="future", her_ratio=ratio)
    rs = np.random.RandomState(seed)
    for ep in synth_goal_episodes(rs, n_ep, T, O0, G, A):
        for (o, a, r, d, no) in ep:
            ref.add_sample(o, a, r, d, no)
            dev.add_sample(o, a, r, d, no)
        ref.terminate_episode()
        dev.terminate_episode()
    assert (dev._top, dev._size, dev._traj_endpoints) == (ref._top, ref._size, ref._traj_endpoints)

    def compare(buf, trial):
        np.random.seed(100 + trial)
        want = ref.random_batch(48)
        np.random.seed(100 + trial)
        got = buf.random_batch(48)
        for k in got:                       # every key the device class returns (achieved_goals of the current step is not stored)
            np.testing.assert_array_equal(np.asarray(got[k], dtype=np.float32), np.asarray(want[k], dtype=np.float32), err_msg=k)
        assert set(want.keys()) - set(got.keys()) <= {"achieved_goals"}

    for trial in range(3):
        compare(dev, trial)
    starts, lens = dev.trajectory_table()
    assert sorted(zip(starts.tolist(), lens.tolist())) == sorted(
        (s, (e - s) % ref._size) for s, e in ref._traj_endpoints.items() if (e - s) % ref._size > 0)
    if not round_trip:
        return
    # snapshot round trip: same class, same state, same RandomState position as the original
    clone = rb.DeviceEnvHindsightReplayBuffer.__new__(rb.DeviceEnvHindsightReplayBuffer)
    clone.__setstate__(dev.__getstate__())
    assert (clone._top, clone._size, clone._goal_dim, clone.her_ratio, clone.distance_threshold) == (
        dev._top, dev._size, G, ratio, env.distance_threshold)
    np.testing.assert_array_equal(clone._ag_next.numpy(), dev._ag_next.numpy())
    np.testing.assert_array_equal(clone.ring.rows[:ref._size], dev.ring.rows[:ref._size])
    # the RandomState object is SHARED through the state dict here (no pickling): advance the clone only
    compare(clone, 7)
