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


def test_ctypes_struct_layouts_equal_the_c_header(tmp_path):
    """Every struct that crosses the boundary: sizeof and each field's offsetof, as gcc lays include/ilswiss_b200.h out,
    against the ctypes mirror (a field added on one side only would shift everything behind it silently)."""
    import subprocess

    pairs = [("ilsw_mlp", _abi.Mlp), ("ilsw_trainer_config", _abi.TrainerConfig), ("ilsw_disc_config", _abi.DiscConfig),
             ("ilsw_her_sampling", _abi.HerSamplingDesc), ("ilsw_inject", _abi.Inject), ("ilsw_batch", _abi.Batch),
             ("ilsw_state", _abi.State)]
    lines = ["#include <stdio.h>", "#include <stddef.h>", '#include "ilswiss_b200.h"', "int main(void) {"]
    for cname, cls in pairs:
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    lines += ["  return 0;", "}"]
    src, exe = tmp_path / "layout.c", tmp_path / "layout"
    src.write_text("\n".join(lines) + "\n")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)
    # and the header declares no struct field the mirror lacks
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "ilswiss_b200.h")).read(), flags=re.S)
    bodies = {name: body for body, name in re.findall(r"typedef struct \{([^}]*)\} (\w+);", text)}
    for cname, cls in pairs:
        body = bodies[cname]
        declared = []
        for decl in body.split(";"):
            decl = decl.strip()
            if decl:
                declared += [re.sub(r"\[.*?\]", "", d).replace("*", " ").split()[-1] for d in decl.split(",")]
        assert declared == [f for f, _ in cls._fields_], cname
