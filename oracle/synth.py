"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): deterministic synthetic tensors.

Weights and inputs for the golden fixtures are produced by an integer hash in numpy, so the fixture generator
(which runs the real reference in the build container) and the tests (which run anywhere, including the GPU box
where /root/reference does not exist) build bit-identical tensors without having to store them.
"""
from __future__ import annotations

import zlib

import numpy as np
import torch


def _mix64(x: np.ndarray) -> np.ndarray:
    # splitmix64 finaliser (public-domain constant set), vectorised on uint64
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return x ^ (x >> np.uint64(31))


def uniform01(key: str, shape) -> torch.Tensor:
    """U[0,1) float32 tensor fully determined by (key, shape)."""
    n = int(np.prod(shape)) if len(shape) else 1
    seed = np.uint64(zlib.crc32(key.encode()) * 2654435761 % (1 << 63))
    with np.errstate(over="ignore"):
        idx = np.arange(n, dtype=np.uint64) + seed
        bits = _mix64(idx) >> np.uint64(40)  # 24 random bits
    vals = bits.astype(np.float32) * np.float32(1.0 / (1 << 24))
    return torch.from_numpy(vals.reshape(tuple(shape)).copy())


def normalish(key: str, shape, std: float) -> torch.Tensor:
    """Zero-mean tensor with the requested std (sum of 4 uniforms, Irwin-Hall), float32."""
    acc = torch.zeros(tuple(shape), dtype=torch.float32)
    for i in range(4):
        acc += uniform01(f"{key}#{i}", shape)
    return (acc - 2.0) * (std / 0.5773502691896257)


def synth_state_dict(template: dict, weight_std: float = 0.02, perturb_affine: bool = True) -> dict:
    """Fill a state-dict-shaped template {name: tensor} with deterministic values.

    Linear/Conv weights ~ std 0.02 (the reference init scale, MS:500-514); biases and LayerNorm/BatchNorm affine
    parameters are perturbed away from 0/1 so that parity tests exercise them.
    """
    out = {}
    for name, t in template.items():
        shape = tuple(t.shape)
        if name.endswith("num_batches_tracked"):
            out[name] = torch.zeros(shape, dtype=torch.int64)
        elif name.endswith("running_mean"):
            out[name] = normalish(name, shape, 0.05)
        elif name.endswith("running_var"):
            out[name] = 0.5 + uniform01(name, shape)
        elif name.endswith(".bias"):
            out[name] = normalish(name, shape, 0.02 if perturb_affine else 0.0)
        elif ("layer_norm" in name or "batch_norm" in name) and name.endswith(".weight"):
            out[name] = 1.0 + (normalish(name, shape, 0.05) if perturb_affine else 0.0)
        else:
            out[name] = normalish(name, shape, weight_std)
    return out
