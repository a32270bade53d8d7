"""ctypes binding of libsfk.so (the C-ABI in include/sfk.h).

There is deliberately no fallback: if the shared library is missing or the device is not sm_100 the first
call raises.  PyTorch is used only for device memory and streams (tensor.data_ptr(), current_stream()).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsfk.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "sfk.h")

_lib: Optional[ctypes.CDLL] = None
_device_checked = False


class SfkError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load libsfk.so (built in-tree by __graft_entry__.build()). Raises if it is absent."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SfkError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                "There is no CPU / PyTorch fallback for this package."
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in header_prototypes().items():
            fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


_CTYPE = {
    "int": ctypes.c_int,
    "long long": ctypes.c_longlong,
    "float": ctypes.c_float,
    "double": ctypes.c_double,
    "size_t": ctypes.c_size_t,
    "sfk_stream_t": ctypes.c_void_p,
}


def _ctype_of(decl: str):
    decl = decl.strip()
    if "char*" in decl.replace(" ", ""):
        return ctypes.c_char_p
    if "*" in decl:
        return ctypes.c_void_p
    base = re.sub(r"\b(const|unsigned)\b", "", decl)
    base = re.sub(r"\b[a-zA-Z_][a-zA-Z0-9_]*$", "", base.strip()).strip() or decl  # drop the parameter name
    if base not in _CTYPE:
        raise SfkError(f"sfk.h: cannot map C type {decl!r}")
    return _CTYPE[base]


def header_prototypes() -> dict:
    """Parse include/sfk.h into {name: (restype, [argtypes])} so that ctypes marshals 64-bit and float args exactly."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int|size_t|long long)\s+(sfk_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        restype = (ctypes.c_char_p if "char" in ret else ctypes.c_size_t if ret == "size_t"
                   else ctypes.c_longlong if ret == "long long" else ctypes.c_int)
        argtypes = [] if args in ("", "void") else [_ctype_of(a) for a in args.split(",")]
        protos[name] = (restype, argtypes)
    return protos


def declared_symbols() -> list[str]:
    """Every function name declared in include/sfk.h (used by the CPU-side ABI test)."""
    return sorted(header_prototypes())


def last_error() -> str:
    return load().sfk_last_error().decode()


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise SfkError(f"{what or 'libsfk'} failed (rc={rc}): {last_error()}")


def ensure_device() -> None:
    global _device_checked
    if not _device_checked:
        if not torch.cuda.is_available():
            raise SfkError("CUDA is not available; libsfk has no CPU path")
        check(load().sfk_device_check(), "sfk_device_check")
        _device_checked = True


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    """Device address of a tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
