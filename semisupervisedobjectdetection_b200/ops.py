"""Host-side glue between PyTorch tensors and the libsfk C-ABI.

PyTorch provides device memory, streams and the autograd graph; every arithmetic kernel is in libsfk.so.  Nothing here
falls back to PyTorch math: a missing library or a non-sm_100 device raises (see _lib.ensure_device).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib as L

# ---------------------------------------------------------------------------------------------------------------
# workspaces
# ---------------------------------------------------------------------------------------------------------------
_scratch: dict = {}
WEIGHT_EPOCH = 0  # bumped whenever libsfk writes parameters in place (optimiser step, EMA), invalidating bf16 shadows
LAUNCHES = 0      # number of libsfk entry-point calls (each enqueues >= 1 kernel); bench.py reports it


def bump_weight_epoch() -> None:
    global WEIGHT_EPOCH
    WEIGHT_EPOCH += 1


def lib():
    L.ensure_device()
    return L.load()


def call(fn_name: str, *args) -> None:
    global LAUNCHES
    LAUNCHES += 1
    rc = getattr(lib(), fn_name)(*args)
    if rc != 0:
        raise L.SfkError(f"{fn_name} failed (rc={rc}): {L.last_error()}")


def scratch(nbytes: int, device: torch.device, slot: str = "main") -> torch.Tensor:
    """Reusable byte arena (grows on demand); work enqueued on one stream is ordered, so reuse within a stream is safe."""
    # one arena per (device, slot, stream): work on different streams may run concurrently (teacher || student forward)
    key = (device.index if device.index is not None else torch.cuda.current_device(), slot, stream())
    buf = _scratch.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
        ws = wgrad_stream(device) if slot.startswith("bwd") else None
        if ws is not None:
            buf.record_stream(ws)  # side-stream GEMMs read this arena: the allocator must not recycle it under them
        _scratch[key] = buf
    return buf


# ---- second stream for weight-gradient GEMMs (off the critical path of backward, see csrc/layer.cu) ----
WGRAD_OVERLAP = True
SIDE_STREAM_PRIORITY = 0   # 0 = default; -1 = high (experiment switch, see tools/ab_step.py)
_wgrad_streams: dict = {}
_bwd_parity = 0


def wgrad_stream(device=None):
    """torch.cuda.Stream used for weight-gradient GEMMs on `device` (None when the overlap is disabled)."""
    if not WGRAD_OVERLAP:
        return None
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    st = _wgrad_streams.get(idx)
    if st is None:
        st = torch.cuda.Stream(device=idx, priority=SIDE_STREAM_PRIORITY)
        _wgrad_streams[idx] = st
    return st


_deferred_free: list = []


def defer_free(t: torch.Tensor) -> None:
    """Keep `t` alive until the next join_wgrad(): side-stream GEMMs still read it.  (Tensor.record_stream would also be
    correct, but it makes the caching allocator poll events and, when the next step is enqueued before they fire, grow by
    cudaMalloc inside the steady-state loop — measured: 5 cudaMallocs in 5 steps.  Freed after the join, the blocks are
    ordinary same-stream frees.)"""
    _deferred_free.append(t)


def join_wgrad() -> None:
    """Make the current stream wait for every weight-gradient GEMM enqueued so far (before gradients are consumed)."""
    for st in _wgrad_streams.values():
        torch.cuda.current_stream().wait_stream(st)
    _deferred_free.clear()


def next_bwd_parity() -> int:
    global _bwd_parity
    _bwd_parity ^= 1
    return _bwd_parity


def stream() -> int:
    L.ensure_device()  # raises SfkError (never falls back) when there is no sm_100 device
    # raw handle of torch's current stream (torch.cuda.current_stream() builds a Python object: ~15 us per call)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise L.SfkError(f"{name}: expected a CUDA tensor (libsfk has no CPU path)")
    if t.dtype != dtype:
        raise L.SfkError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if t.device.index != torch.cuda.current_device():
        # libsfk enqueues on the CURRENT device's stream: a tensor of another GPU would be an illegal (or silently peer-mapped)
        # access.  The module / wrapper entry points switch devices themselves (device_of); raw op callers must do it.
        raise L.SfkError(f"{name}: tensor lives on cuda:{t.device.index} but the current device is cuda:"
                         f"{torch.cuda.current_device()}; wrap the call in `with torch.cuda.device(tensor.device)`")
    return t if t.is_contiguous() else t.contiguous()


def device_of(t):
    """Context manager making the device of tensor / module-device string `t` current (no-op for CPU tensors, so that the
    usual 'expected a CUDA tensor' error still surfaces from _chk)."""
    dev = t.device if isinstance(t, torch.Tensor) else torch.device(t)
    if dev.type != "cuda":
        import contextlib
        return contextlib.nullcontext()
    return torch.cuda.device(dev)


def grad_buffer(p: torch.nn.Parameter) -> Optional[torch.Tensor]:
    """fp32 gradient buffer the wgrad kernels accumulate into (None when the parameter is frozen)."""
    if not p.requires_grad:
        return None
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    p._sfk_live = True  # FusedClipAdam only updates parameters whose gradient was produced (Adam skips grad=None)
    return p.grad


# ---------------------------------------------------------------------------------------------------------------
# raw (non-autograd) wrappers used by the modules
# ---------------------------------------------------------------------------------------------------------------
def gemm(a, b, M, N, K, *, lda=None, ldb=None, a_mn=0, b_mn=0, alpha=1.0, bias=None, residual=None, ldr=0, row_scale=None,
         rows_per_scale=1, out=None, ldo=None, out_bf16=True, atomic_add=False, relu=False, splits=0, block_n=0):
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    call("sfk_gemm_bf16", _p(a), lda if lda is not None else a.stride(0), a_mn, _p(b), ldb if ldb is not None else b.stride(0),
         b_mn, M, N, K, alpha, _p(bias), _p(residual), ldr, _p(row_scale), rows_per_scale, _p(out),
         ldo if ldo is not None else out.stride(0), int(out_bf16), int(atomic_add), int(relu), splits, block_n, stream())
    return out


def cast_bf16(src: torch.Tensor, cols_padded: Optional[int] = None) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 [rows, cols_padded] (zero padded)."""
    src2 = src.reshape(src.shape[0], -1)
    rows, cols = src2.shape
    cp = cols_padded or cols
    dst = torch.empty(rows, cp, device=src.device, dtype=torch.bfloat16)
    call("sfk_cast_2d", _p(src2), src2.stride(0), _p(dst), cp, rows, cols, cp, stream())
    return dst


def weight_bf16(p: torch.Tensor) -> torch.Tensor:
    """bf16 [rows, cols] GEMM operand of an fp32 weight.  Parameters owned by FusedClipAdam carry an always-current view of
    its flat bf16 shadow (refreshed by the Adam / EMA kernels); anything else is cast on the fly."""
    w16 = getattr(p, "_sfk_w16", None)
    if w16 is not None and w16.device == p.device:
        if getattr(p, "_sfk_w16_version", -1) != p._version:   # written through PyTorch since the last sync
            src = p.detach().reshape(p.shape[0], -1)
            call("sfk_cast_2d", _p(src), src.stride(0), _p(w16), src.shape[1], src.shape[0], src.shape[1], src.shape[1], stream())
            p._sfk_w16_version = p._version
        return w16.view(p.shape[0], -1)
    return cast_bf16(p.detach())


def layernorm_fwd(x, gamma, beta, out_bf16=True, need_stats=True, eps=1e-5):
    rows, C = x.numel() // x.shape[-1], x.shape[-1]
    y = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32) if need_stats else None
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32) if need_stats else None
    call("sfk_layernorm_fwd", _p(x), int(x.dtype == torch.bfloat16), 0, 0, 0, _p(gamma), _p(beta), _p(y), int(out_bf16), _p(mean),
         _p(rstd), rows, C, eps, stream())
    return y, mean, rstd


def layernorm_bwd(dy, x, mean, rstd, gamma, dgamma, dbeta, dres=None, out_bf16=False):
    rows, C = x.numel() // x.shape[-1], x.shape[-1]
    dx = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16 if out_bf16 else torch.float32)
    call("sfk_layernorm_bwd", _p(dy), int(dy.dtype == torch.bfloat16), _p(x), int(x.dtype == torch.bfloat16), _p(mean), _p(rstd),
         _p(gamma), _p(dres), _p(dx), int(out_bf16), _p(dgamma), _p(dbeta), rows, C, stream())
    return dx


def colsum_into(x: torch.Tensor, out: torch.Tensor, rows: int, C: int, ld: Optional[int] = None):
    call("sfk_colsum", _p(x), int(x.dtype == torch.bfloat16), ld if ld is not None else C, _p(out), rows, C, stream())


def scale_cast(x: torch.Tensor, out_dtype, scale=None, rows_per_scale=1):
    C = x.shape[-1]
    rows = x.numel() // C
    y = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    call("sfk_scale_cast", _p(x), int(x.dtype == torch.bfloat16), _p(y), int(out_dtype == torch.bfloat16), _p(scale), rows_per_scale,
         rows, C, stream())
    return y


# ---------------------------------------------------------------------------------------------------------------
# ctypes mirrors of the structs in include/sfk.h
# ---------------------------------------------------------------------------------------------------------------
class LayerDims(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in ("B", "H", "W", "C", "heads", "sr", "hidden", "ptl")]


_PARAM_FIELDS = ("wq", "wkv", "wo", "wsr", "w1", "w2", "bq", "bkv", "bo", "bsr", "b1", "b2", "ln1_g", "ln1_b", "ln2_g", "ln2_b",
                 "lnsr_g", "lnsr_b", "dw_wt", "dw_b", "drop_scale_attn", "drop_scale_mlp", "dmlp_in")
_GRAD_FIELDS = ("dwq", "dwk", "dwv", "dwo", "dwsr", "dw1", "dw2", "dbq", "dbk", "dbv", "dbo", "dbsr", "db1", "db2", "dln1_g",
                "dln1_b", "dln2_g", "dln2_b", "dlnsr_g", "dlnsr_b", "ddw", "ddb", "next_dmlp", "next_drop_scale", "next_db2")


class LayerParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _PARAM_FIELDS]


class LayerGrads(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in _GRAD_FIELDS]
