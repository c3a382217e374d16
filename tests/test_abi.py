"""CPU checks of the boundary: the in-tree CUDA library loads and exports every symbol that
include/ilswiss_b200.h declares, the ctypes mirror agrees with the header, and the host-side
layout helpers do what the kernels assume.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from ilswiss_b200 import _abi, _lib, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ilswiss_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ilsw_[a-z0-9_]+)\s*\(", text)))


def test_header_and_ctypes_mirror_agree():
    syms = header_symbols()
    assert len(syms) >= 30
    assert sorted(_abi.PROTOTYPES.keys()) == syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__ as g

        g.build()
    lib = _lib.load()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.ilsw_abi_version() == _abi.ABI_VERSION
    # pure host helper, no GPU needed
    assert lib.ilsw_mlp_num_params(14, 256, 1, 0) == 69889      # P_q Hopper
    assert lib.ilsw_mlp_num_params(11, 256, 3, 1) == 70406      # P_pi Hopper
    assert lib.ilsw_mlp_num_params(376, 256, 17, 0) == 166673   # P_pi TD3 Humanoid


def test_fails_loudly_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from ilswiss_b200 import engine

    with pytest.raises(_lib.IlswError):
        engine.ReplayRing(16, 3, 2)
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ilsw_rb_create(C.byref(h), 16, 3, 2) < 0   # CUDA error surfaced as a status, not a crash
    assert len(lib.ilsw_last_error()) > 0


def test_row_layout_roundtrip():
    rs = np.random.RandomState(0)
    O, A, n = 11, 3, 17
    obs, act = rs.randn(n, O), rs.uniform(-1, 1, (n, A))
    rew, term, nobs = rs.randn(n, 1), (rs.rand(n, 1) < 0.3).astype(np.uint8), rs.randn(n, O)
    hot = layout.pack_hot_rows(obs, act, rew, term, nobs)
    assert hot.shape == (n, layout.hot_row_stride(O, A)) == (n, 32)
    back = layout.unpack_hot_rows(hot, O, A)
    np.testing.assert_array_equal(back["observations"], obs.astype(np.float32).astype(np.float64))
    np.testing.assert_array_equal(back["terminals"], term)
    assert back["terminals"].dtype == np.uint8 and back["rewards"].shape == (n, 1)
    host = layout.pack_host_rows(obs, act, rew, term, nobs, absorbing=np.ones((n, 2)), timeouts=np.ones(n))
    assert host.shape == (n, layout.host_row_floats(O, A))
    np.testing.assert_array_equal(host[:, :2 * O + A + 2], hot[:, :2 * O + A + 2])
    assert (host[:, -3:] == 1).all()
    assert layout.hot_row_stride(376, 17) == 784 and layout.hot_row_stride(17, 6) == 48      # 64-byte row alignment
