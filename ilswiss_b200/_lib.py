"""Loader of the in-tree CUDA library (ilswiss_b200/csrc/libilswiss_b200.so).

The product path has NO CPU fallback: if the library is missing or does not export the full
ABI this module raises, and every op in ilswiss_b200 fails loudly."""
import ctypes as C
import os

from . import _abi

_HERE = os.path.dirname(os.path.abspath(__file__))
# ILSW_LIB: development override used to A/B kernel variants on one GPU box (tools/ab.sh)
LIB_PATH = os.environ.get("ILSW_LIB") or os.path.join(_HERE, "csrc", "libilswiss_b200.so")
_lib = None


class IlswError(RuntimeError):
    pass


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise IlswError(
            "ilswiss_b200: %s not found -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C ilswiss_b200/csrc`).  There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    missing = [n for n in _abi.PROTOTYPES if not hasattr(lib, n)]
    if missing:
        raise IlswError("ilswiss_b200: library does not export %s" % missing)
    _abi.declare(lib)
    if lib.ilsw_abi_version() != _abi.ABI_VERSION:
        raise IlswError("ilswiss_b200: ABI version mismatch")
    _lib = lib
    return lib


def check(rc, what=""):
    if rc < 0:
        msg = load().ilsw_last_error().decode(errors="replace")
        raise IlswError("ilswiss_b200 %s failed (%d): %s" % (what, rc, msg))
    return rc
