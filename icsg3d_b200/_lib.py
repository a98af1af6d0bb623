"""ctypes binding of libicsg3d.so — the C-ABI boundary (include/icsg3d.h).

Prototypes are parsed from the header itself, so the Python side can never drift from the declared
ABI, and `declared_symbols()` lets the CPU test-suite verify that the library exports every one.
There is NO fallback: if the shared library is missing this module raises at first use.
"""
from __future__ import annotations

import ctypes
import re
from pathlib import Path

PKG = Path(__file__).resolve().parent
HEADER = PKG.parent / "include" / "icsg3d.h"
LIBPATH = PKG / "libicsg3d.so"

_CTYPES = {
    "int": ctypes.c_int,
    "int64_t": ctypes.c_int64,
    "uint64_t": ctypes.c_uint64,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "void": None,
}


class Icsg3dError(RuntimeError):
    pass


def _parse_type(t: str):
    t = t.strip()
    if "*" in t:
        if "char" in t:
            return ctypes.c_char_p
        return ctypes.c_void_p
    t = t.replace("const", "").strip()
    if t not in _CTYPES:
        raise ValueError(f"icsg3d.h: unknown C type '{t}'")
    return _CTYPES[t]


def parse_header(path: Path = HEADER):
    """Return {name: (restype, [argtypes], [argnames])} for every prototype in the header."""
    src = path.read_text()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    src = re.sub(r"#[^\n]*", "", src)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][\w\s\*]*?)\b(icsg3d_\w+)\s*\(([^)]*)\)\s*;", src):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        argtypes, argnames = [], []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                mm = re.match(r"(.*?)(\w+)$", a)
                argtypes.append(_parse_type(mm.group(1)))
                argnames.append(mm.group(2))
        protos[name] = (_parse_type(ret), argtypes, argnames)
    return protos


def declared_symbols():
    return sorted(parse_header().keys())


_lib = None
_protos = None


def lib():
    global _lib, _protos
    if _lib is None:
        if not LIBPATH.exists():
            raise Icsg3dError(
                f"{LIBPATH} is missing: build it with `python -m icsg3d_b200.build` "
                "(there is no CPU or PyTorch fallback for the icsg3d hot path)")
        L = ctypes.CDLL(str(LIBPATH))
        _protos = parse_header()
        for name, (ret, argtypes, _) in _protos.items():
            fn = getattr(L, name)  # AttributeError if the .so does not export a declared symbol
            fn.restype = ret
            fn.argtypes = argtypes
        _lib = L
    return _lib


def last_error() -> str:
    return lib().icsg3d_last_error().decode()


# Optional per-call device timing (tools/profile_step.py): when PROFILE is a list, every entry-point call is
# bracketed by CUDA events on the current stream and (name, start, end) is appended.  Off on the product path.
PROFILE = None


def call(name: str, *args):
    """Call an int-returning entry point; raise Icsg3dError with the library's message on failure."""
    fn = getattr(lib(), name)
    if PROFILE is not None:
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROFILE.append((name, e0, e1))
    else:
        rc = fn(*args)
    if rc != 0:
        raise Icsg3dError(f"{name} failed ({rc}): {last_error()}")
    return rc
