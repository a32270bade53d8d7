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
        _lib = ctypes.CDLL(LIB_PATH)
        _lib.sfk_last_error.restype = ctypes.c_char_p
        _lib.sfk_version.restype = ctypes.c_int
    return _lib


def declared_symbols() -> list[str]:
    """Every function name declared in include/sfk.h (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sfk_[a-z0-9_]+)\s*\(", text)))


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


def stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> ctypes.c_void_p:
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


c_int = ctypes.c_int
c_ll = ctypes.c_longlong
c_float = ctypes.c_float
c_size_t = ctypes.c_size_t
