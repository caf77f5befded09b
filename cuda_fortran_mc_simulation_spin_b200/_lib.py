"""ctypes loader for libb200mc.so (the C ABI declared in include/b200mc.h).

There is deliberately no fallback: if the CUDA library is missing or cannot be
loaded, importing a model module raises.  Nothing here touches ``oracle/``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("B200MC_SO") or os.path.join(_PKG, "libb200mc.so")  # B200MC_SO: A/B-test another build of the same library
CSRC = os.path.join(_PKG, "csrc")

i64, i32, u32, f64, P = C.c_int64, C.c_int32, C.c_uint32, C.c_double, C.c_void_p
PP = C.POINTER(C.c_void_p)


class B200MCError(RuntimeError):
    pass


def build(force: bool = False) -> str:
    """Compile the CUDA extension in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    args = ["make", "-C", CSRC, "-s"]
    if force:
        args.append("-B")
    subprocess.check_call(args)
    return SO_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise B200MCError(
                f"{SO_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C cuda_fortran_mc_simulation_spin_b200/csrc). There is no CPU fallback."
            )
        _lib = C.CDLL(SO_PATH)
        _lib.b200mc_last_error.restype = C.c_char_p
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = lib().b200mc_last_error()
        raise B200MCError(f"b200mc error {rc}: {msg.decode() if msg else ''}")


def fn(name, restype, *argtypes):
    f = getattr(lib(), name)
    f.restype = restype
    f.argtypes = list(argtypes)
    return f


def call(name, *args):
    """call an int-returning entry point whose argtypes were declared with fn()"""
    check(getattr(lib(), name)(*args))
