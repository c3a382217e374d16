"""-m gpu: the HBM replay ring (R1-R4) against the numpy restatement of SimpleReplayBuffer."""
import numpy as np
import pytest
import torch

from helpers import R, layout

pytestmark = pytest.mark.gpu


def make_rows(rs, n, O, A):
    return dict(observations=rs.randn(n, O), actions=rs.uniform(-1, 1, (n, A)), rewards=rs.randn(n, 1),
                terminals=(rs.rand(n, 1) < 0.2).astype(np.uint8), next_observations=rs.randn(n, O))


@pytest.mark.parametrize("O,A,cap", [(11, 3, 50), (17, 6, 64), (376, 17, 40), (1, 1, 7)])
def test_append_wraps_like_reference_and_gather_is_bit_exact(O, A, cap):
    from ilswiss_b200.engine import ReplayRing

    rs = np.random.RandomState(3)
    ring = ReplayRing(cap, O, A)
    ora = R.ReplayOracle(cap, O, A, random_seed=9)
    total = 0
    for burst in (1, cap // 2, cap - 3, 5, cap):      # episode-sized bursts, wraps several times
        d = make_rows(rs, burst, O, A)
        absorbing = rs.rand(burst, 2).round()
        for i in range(burst):
            ora.add_sample(d["observations"][i], d["actions"][i], d["rewards"][i], d["terminals"][i],
                           d["next_observations"][i], absorbing=absorbing[i], timeout=bool(i % 2))
        ring.append_host(layout.pack_host_rows(absorbing=absorbing, timeouts=np.arange(burst) % 2, **d))
        ring.commit()
        total += burst
        assert ring.committed_size == ora._size == min(total, cap)
        assert ring.top == ora._top
        idx = ora.sample_indices(33)
        ref = ora.get_batch_using_indices(idx)
        hot, cold = ring.gather(torch.from_numpy(idx.astype(np.int32)).cuda())
        got = layout.unpack_hot_rows(hot.cpu().numpy(), O, A)
        for k in got:   # bit-exact vs the reference's float32 device batch (np_to_pytorch_batch)
            np.testing.assert_array_equal(got[k].astype(np.float32), ref[k].astype(np.float32), err_msg=k)
        cold = cold.cpu().numpy()
        np.testing.assert_array_equal(cold[:, :2], ora._absorbing[idx].astype(np.float32))
        np.testing.assert_array_equal(cold[:, 2], ora._timeouts[idx, 0].astype(np.float32))
        assert (hot.cpu().numpy()[:, 2 * O + A + 2:] == 0).all()    # padding stays zero


def test_pending_appends_are_visible_to_the_next_sample_without_explicit_commit():
    from ilswiss_b200.engine import ReplayRing

    rs = np.random.RandomState(1)
    ring = ReplayRing(32, 4, 2)
    d = make_rows(rs, 10, 4, 2)
    ring.append_host(layout.pack_host_rows(**d))
    assert ring.size == 10 and ring.committed_size == 0
    hot, _ = ring.gather(torch.arange(10, dtype=torch.int32, device="cuda"))   # gather commits first
    np.testing.assert_array_equal(hot.cpu().numpy()[:, :4], d["observations"].astype(np.float32))


def test_philox_sample_is_uniform_with_replacement_and_in_range():
    from ilswiss_b200.engine import ReplayRing

    ring = ReplayRing(1000, 3, 1)
    rows = np.zeros((700, ring.stride), np.float32)
    rows[:, 0] = np.arange(700)
    ring.load_device(torch.from_numpy(rows).cuda())
    idx, hot = ring.sample(200000, seed=5, counter=0)
    idx = idx.cpu().numpy()
    assert idx.min() >= 0 and idx.max() < 700           # only the filled part [0,size)
    np.testing.assert_array_equal(hot[:, 0].cpu().numpy(), idx.astype(np.float32))
    counts = np.bincount(idx, minlength=700)
    chi2 = ((counts - 200000 / 700) ** 2 / (200000 / 700)).sum()
    assert 550 < chi2 < 860                               # 699 dof, +-4 sigma
    idx2, _ = ring.sample(200000, seed=5, counter=0)
    np.testing.assert_array_equal(idx, idx2.cpu().numpy())
    idx3, _ = ring.sample(200000, seed=5, counter=1)
    assert not np.array_equal(idx, idx3.cpu().numpy())


def test_full_size_ring_checksum_of_gathers():
    """BASELINE config 2 size (1M x Hopper): size-independent property -- sum over a gathered
    batch equals the sum computed from the generating formula of the rows."""
    from ilswiss_b200.engine import ReplayRing

    N, O, A = 1_000_000, 11, 3
    ring = ReplayRing(N, O, A)
    ar = torch.arange(N, dtype=torch.float32, device="cuda")
    rows = torch.zeros((N, ring.stride), dtype=torch.float32, device="cuda")
    rows[:, 0] = ar
    rows[:, O + A] = -ar
    rows[:, O + A + 2] = 2 * ar
    ring.load_device(rows)
    idx = torch.randint(0, N, (4096,), dtype=torch.int32, device="cuda")
    hot, _ = ring.gather(idx)
    f = idx.float()
    assert torch.equal(hot[:, 0], f) and torch.equal(hot[:, O + A], -f) and torch.equal(hot[:, O + A + 2], 2 * f)
    # ring is full: one more append overwrites slot 0 (FIFO), size stays N
    ring.append_host(np.full((1, ring.host_w), 7.0, np.float32))
    ring.commit()
    assert ring.committed_size == N and ring.top == 1
    hot, _ = ring.gather(torch.zeros(1, dtype=torch.int32, device="cuda"))
    assert float(hot[0, 0]) == 7.0
