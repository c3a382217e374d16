"""Host-side layout helpers (pure numpy; no CUDA needed): replay-row packing and the flat
parameter-arena layout shared with the kernels (include/ilswiss_b200.h)."""
import numpy as np


def hot_row_stride(obs_dim, act_dim):
    """Padded stride (floats) of one transition in the HBM ring:
    [obs | act | reward | terminal | next_obs | pad->multiple of 16 floats = 64 bytes, the DRAM access granule]."""
    return (2 * obs_dim + act_dim + 2 + 15) // 16 * 16


def host_row_floats(obs_dim, act_dim):
    """Packed staging row for ilsw_rb_append: hot fields + absorbing0, absorbing1, timeout."""
    return 2 * obs_dim + act_dim + 5


def derive_seed(salt=0):
    """A Philox seed that is reproducible under np.random.seed(...) but does NOT advance numpy's global stream: that stream
    belongs to the reference's own code (buffer seeds drawn by the run scripts, exploration noise, hindsight sampling), and
    its call order must stay what it is without this package."""
    st = np.random.get_state()
    key, pos = st[1], int(st[2])
    h = int(key[pos % 624]) ^ (int(key[(pos + 1) % 624]) << 1) ^ pos ^ int(salt)
    return h % (2 ** 31 - 2) + 1


def pack_host_rows(observations, actions, rewards, terminals, next_observations, absorbing=None,
                   timeouts=None, out=None):
    """float64/uint8 reference-typed fields (simple_replay_buffer.py:48-60) -> float32 staging
    rows.  The float32 cast here is the same rounding the reference applies at sample time in
    np_to_pytorch_batch (rlkit/torch/core.py:124-143, pytorch_util.py:84-88), so gathered
    batches are bit-identical to the reference's device batches."""
    obs = np.asarray(observations)
    n, O = obs.shape
    act = np.asarray(actions).reshape(n, -1)
    A = act.shape[1]
    W = host_row_floats(O, A)
    if out is None:
        out = np.empty((n, W), dtype=np.float32)
    out[:, :O] = obs
    out[:, O:O + A] = act
    out[:, O + A] = np.asarray(rewards).reshape(n)
    out[:, O + A + 1] = np.asarray(terminals).reshape(n)
    out[:, O + A + 2:2 * O + A + 2] = np.asarray(next_observations)
    if absorbing is None:
        out[:, 2 * O + A + 2:2 * O + A + 4] = 0.0
    else:
        out[:, 2 * O + A + 2:2 * O + A + 4] = np.asarray(absorbing).reshape(n, 2)
    out[:, 2 * O + A + 4] = 0.0 if timeouts is None else np.asarray(timeouts).reshape(n)
    return out


def pack_hot_rows(observations, actions, rewards, terminals, next_observations):
    """Device hot-row image ([n, stride] float32, zero padded) for ilsw_rb_load_device."""
    obs = np.asarray(observations)
    n, O = obs.shape
    act = np.asarray(actions).reshape(n, -1)
    A = act.shape[1]
    S = hot_row_stride(O, A)
    out = np.zeros((n, S), dtype=np.float32)
    out[:, :O] = obs
    out[:, O:O + A] = act
    out[:, O + A] = np.asarray(rewards).reshape(n)
    out[:, O + A + 1] = np.asarray(terminals).reshape(n)
    out[:, O + A + 2:2 * O + A + 2] = np.asarray(next_observations)
    return out


def unpack_hot_rows(rows, obs_dim, act_dim):
    """Inverse of the hot-row layout -> dict with the reference's random_batch keys/dtypes
    (simple_replay_buffer.py:278-293): float64 arrays, uint8 terminals."""
    O, A = obs_dim, act_dim
    rows = np.asarray(rows)
    return {
        "observations": rows[:, :O].astype(np.float64),
        "actions": rows[:, O:O + A].astype(np.float64),
        "rewards": rows[:, O + A:O + A + 1].astype(np.float64),
        "terminals": rows[:, O + A + 1:O + A + 2].astype(np.uint8),
        "next_observations": rows[:, O + A + 2:2 * O + A + 2].astype(np.float64),
    }


def mlp_num_params(in_dim, hidden, out_dim, log_std_head=False):
    n = hidden * in_dim + hidden + hidden * hidden + hidden + out_dim * hidden + out_dim
    if log_std_head:
        n += out_dim * hidden + out_dim
    return n


def mlp_param_shapes(in_dim, hidden, out_dim, log_std_head=False):
    """Shapes in nn.Module.parameters() order of the reference nets (networks.py:23-83,
    policies.py:191-243; MLPDisc de-duplicated order simple_disc_models.py:29-41)."""
    shapes = [(hidden, in_dim), (hidden,), (hidden, hidden), (hidden,), (out_dim, hidden), (out_dim,)]
    if log_std_head:
        shapes += [(out_dim, hidden), (out_dim,)]
    return shapes


def flatten_params(param_arrays):
    return np.concatenate([np.asarray(p, dtype=np.float32).ravel() for p in param_arrays])
