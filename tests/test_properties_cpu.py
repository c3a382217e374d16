"""Size-independent properties of the host-side logic (hypothesis; CPU only)."""
import numpy as np
from hypothesis import given, settings, strategies as st

from ilswiss_b200 import layout
from oracle.restate import HindsightOracle, ReplayOracle, synth_goal_episodes


@settings(max_examples=40, deadline=None)
@given(O=st.integers(1, 40), A=st.integers(1, 9), n=st.integers(1, 33), seed=st.integers(0, 10_000))
def test_hot_row_layout_round_trip(O, A, n, seed):
    """pack -> unpack is the identity on float32-rounded values for every shape; rows are 16-byte aligned and
    cat(obs, act) is contiguous at the start of a row (what the critic / discriminator tiles read)."""
    rs = np.random.RandomState(seed)
    obs, act, nobs = rs.randn(n, O), rs.uniform(-1, 1, (n, A)), rs.randn(n, O)
    rew, term = rs.randn(n, 1), (rs.rand(n, 1) < 0.3).astype(np.uint8)
    hot = layout.pack_hot_rows(obs, act, rew, term, nobs)
    stride = layout.hot_row_stride(O, A)
    assert hot.shape == (n, stride) and stride % 4 == 0 and stride >= 2 * O + A + 2
    np.testing.assert_array_equal(hot[:, :O + A], np.concatenate([obs, act], 1).astype(np.float32))
    back = layout.unpack_hot_rows(hot, O, A)
    np.testing.assert_array_equal(back["next_observations"], nobs.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(back["rewards"], rew.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(back["terminals"], term)
    assert (hot[:, 2 * O + A + 2:] == 0).all()                    # padding


@settings(max_examples=25, deadline=None)
@given(cap=st.integers(4, 60), n=st.integers(1, 150), seed=st.integers(0, 1000))
def test_replay_oracle_ring_invariants(cap, n, seed):
    """R1/R2: after n appends the ring holds the last min(n, cap) transitions, _top = n mod cap, and the sampled
    index stream is RandomState(seed).randint(0, size, B)."""
    buf = ReplayOracle(cap, 2, 1, random_seed=seed)
    for i in range(n):
        buf.add_sample(np.full(2, i), np.full(1, i), float(i), False, np.full(2, i + 1))
    assert buf._size == min(n, cap) and buf._top == n % cap
    held = sorted(buf._rewards[:buf._size, 0].astype(int).tolist())
    assert held == list(range(max(0, n - cap), n))
    idx = np.random.RandomState(seed).randint(0, buf._size, 16)
    np.testing.assert_array_equal(buf.random_batch(16)["rewards"][:, 0], buf._rewards[idx, 0])


@settings(max_examples=15, deadline=None)
@given(n_ep=st.integers(2, 6), T=st.integers(2, 30), ratio=st.sampled_from([0.0, 0.5, 0.8, 1.0]), seed=st.integers(0, 1000))
def test_hindsight_sampling_invariants(n_ep, T, ratio, seed):
    """relabel_replay_buffer.py:63-131: the future step lies in the SAME trajectory at or after the sampled step; only the
    first int(her_ratio B) rows get a new goal; recomputed rewards are the sparse goal reward of the (new) goal."""
    O0, G, A, B = 4, 2, 2, 24
    buf = HindsightOracle(n_ep * T + 5, O0, G, A, random_seed=seed, her_ratio=ratio, distance_threshold=0.05)
    rs = np.random.RandomState(seed)
    for ep in synth_goal_episodes(rs, n_ep, T, O0, G, A):
        for (o, a, r, d, no) in ep:
            buf.add_sample(o, a, r, d, no)
        buf.terminate_episode()
    np.random.seed(seed)
    idx, idx_her = buf.sample_indices(B)
    starts = sorted(buf._traj_endpoints)
    traj_of = lambda i: max(s for s in starts if s <= i)
    if ratio > 0:
        assert len(idx_her) == B
        for i, j in zip(idx, idx_her):
            assert traj_of(i) == traj_of(j) and i <= j < buf._traj_endpoints[traj_of(i)]
    else:
        assert len(idx_her) == 0
    batch = buf.batch_from_indices(idx, idx_her)
    n = int(ratio * B)
    orig = buf._observations["desired_goal"][idx]
    np.testing.assert_array_equal(batch["desired_goals"][n:], orig[n:])
    np.testing.assert_array_equal(batch["desired_goals"], batch["next_desired_goals"])
    if ratio > 0:
        np.testing.assert_array_equal(batch["desired_goals"][:n], buf._next_obs["achieved_goal"][idx_her][:n])
        d = np.linalg.norm(batch["next_achieved_goals"] - batch["desired_goals"], axis=-1)
        np.testing.assert_array_equal(batch["rewards"][:, 0], -(d > 0.05).astype(np.float32))
        assert set(np.unique(batch["rewards"])) <= {-1.0, 0.0}
