"""CPU: every handle-taking entry point of include/b200mc.h refuses a NULL handle with an error code (or a sentinel for
the value getters) instead of dereferencing it -- no GPU needed: the check comes before any CUDA call."""
import ctypes as C
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CT = {"void*": C.c_void_p, "int32_t": C.c_int32, "int64_t": C.c_int64, "int": C.c_int, "double": C.c_double,
      "uint32_t": C.c_uint32, "unsigned long long": C.c_ulonglong}


def _prototypes():
    src = open(os.path.join(ROOT, "include", "b200mc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for m in re.finditer(r"\b(int|int32_t|int64_t|double|unsigned long long)\s+(b200mc_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        ret, name, params = m.group(1), m.group(2), [p.strip() for p in m.group(3).split(",")]
        yield ret, name, params


def test_null_handle_is_refused():
    from cuda_fortran_mc_simulation_spin_b200 import _lib
    lib = C.CDLL(_lib.SO_PATH)
    scratch = (C.c_double * 4096)()          # any pointer argument points here (never reached: the handle check is first)
    checked = 0
    for ret, name, params in _prototypes():
        if not params or not re.fullmatch(r"void\s*\*\s*h", params[0]):
            continue                           # creators (void** h), free functions
        if name.endswith("_destroy"):
            continue                           # destroy(NULL) is a no-op by contract (like free)
        args, argtypes = [None], [C.c_void_p]
        for p in params[1:]:
            if "*" in p or "[" in p:
                args.append(C.cast(scratch, C.c_void_p)); argtypes.append(C.c_void_p)
            else:
                t = next(v for k, v in CT.items() if re.match(rf"(const\s+)?{re.escape(k)}\b", p))
                args.append(t(1)); argtypes.append(t)
        f = getattr(lib, name)
        f.argtypes = argtypes
        f.restype = {"int": C.c_int, "int32_t": C.c_int32, "int64_t": C.c_int64, "double": C.c_double,
                     "unsigned long long": C.c_ulonglong}[ret]
        r = f(*args)
        if ret == "int":
            assert r != 0, f"{name}(NULL, ...) returned success"
        elif ret in ("int32_t", "int64_t"):
            assert r == -1, f"{name}(NULL) = {r}"
        elif ret == "double":
            assert r == 0.0, f"{name}(NULL) = {r}"
        checked += 1
    assert checked > 150, checked
    msg = C.c_char_p.in_dll  # noqa: F841  (the message itself is thread-local: fetched through the accessor)
    lib.b200mc_last_error.restype = C.c_char_p
    assert b"handle" in lib.b200mc_last_error().lower() or lib.b200mc_last_error()


def test_destroy_null_is_a_noop():
    from cuda_fortran_mc_simulation_spin_b200 import _lib
    lib = C.CDLL(_lib.SO_PATH)
    for ret, name, params in _prototypes():
        if name.endswith("_destroy"):
            f = getattr(lib, name)
            f.argtypes = [C.c_void_p]
            assert f(None) == 0, name
