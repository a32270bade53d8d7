"""Flat fp32 parameter / gradient storage and the fused clip + Adam(L2) step (SFM:32-37,158-162).

All parameters handed to the optimiser are re-pointed at views of one flat buffer (in forward-execution order), with
gradient, exp_avg and exp_avg_sq buffers of the same layout.  One kernel launch per contiguous run of *live* parameters
updates them; a parameter is live iff its gradient was written since the last zero_grad(), which reproduces torch.optim.Adam
skipping parameters whose .grad is None (frozen blocks, the unused cls_proj Linears).
"""
from __future__ import annotations

import re
from typing import Iterable, Optional

import torch

from . import ops


_IN_LAYER_RANK = {"attention.self.key.weight": 0, "attention.self.value.weight": 1,   # adjacent -> one [2C, C] K|V GEMM operand
                  "attention.self.key.bias": 2, "attention.self.value.bias": 3}         # adjacent -> one [2C] bias


def _exec_order_key(name: str):
    """Sort key placing parameters in forward-execution order (so backward finalises gradients back-to-front)."""
    m = re.search(r"encoder\.patch_embeddings\.(\d+)\.", name)
    if m:
        return (int(m.group(1)), 0, 0, name)
    m = re.search(r"encoder\.block\.(\d+)\.(\d+)\.", name)
    if m:
        rank = 9
        for suffix, r in _IN_LAYER_RANK.items():
            if name.endswith(suffix):
                rank = r
        return (int(m.group(1)), 1, int(m.group(2)), rank, name)
    m = re.search(r"encoder\.layer_norm\.(\d+)\.", name)
    if m:
        return (int(m.group(1)), 2, 0, name)
    return (99, 0, 0, name)


class FusedClipAdam(torch.optim.Optimizer):
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float, betas=(0.5, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, clip_value: Optional[float] = 1.2, model: Optional[torch.nn.Module] = None):
        params = [p for p in params]
        if model is not None:
            names = {id(p): n for n, p in model.named_parameters()}
            params.sort(key=lambda p: _exec_order_key(names.get(id(p), "~")))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay or 0.0, clip_value=clip_value)
        super().__init__(params, defaults)
        self._flatten()
        self._step = 0

    def _flatten(self):
        ps = [p for g in self.param_groups for p in g["params"]]
        if not ps:
            raise ValueError("optimizer got an empty parameter list")
        dev = ps[0].device
        if dev.type != "cuda":
            raise RuntimeError("FusedClipAdam needs CUDA parameters (libsfk has no CPU path)")
        offs, total = [], 0
        for p in ps:
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8  # every parameter 16-byte aligned in the bf16 shadow too (TMA operand base)
        self.flat_p = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(total, device=dev, dtype=torch.float32)
        self.flat_w16 = torch.zeros(total, device=dev, dtype=torch.bfloat16)  # bf16 GEMM shadow, same element layout
        self._offsets = offs
        self._params = ps
        for p, o in zip(ps, offs):
            n = p.numel()
            view = self.flat_p[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[o:o + n].view(p.shape)
            p._sfk_live = False
            p._sfk_w16 = self.flat_w16[o:o + n].view(p.shape)
        self.sync_shadow()

    def sync_shadow(self):
        """Recast the whole flat bf16 shadow from the fp32 masters (one launch); called after host-side weight writes."""
        n = self.flat_p.numel()
        ops.call("sfk_cast_2d", self.flat_p.data_ptr(), n, self.flat_w16.data_ptr(), n, 1, n, n, ops.stream())
        for p in self._params:
            p._sfk_w16_version = p._version
        ops.bump_weight_epoch()

    # --- gradient bookkeeping ---
    def zero_grad(self, set_to_none: bool = False):
        ops.join_wgrad()
        self.flat_g.zero_()
        for p, o in zip(self._params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
            p._sfk_live = False

    def live_ranges(self):
        """Contiguous [start, end) element ranges of parameters whose gradient was produced this step."""
        ranges, cur = [], None
        for p, o in zip(self._params, self._offsets):
            live = p.requires_grad and getattr(p, "_sfk_live", False)
            end = o + (p.numel() + 7) // 8 * 8
            if live:
                if cur is not None and cur[1] == o:
                    cur[1] = end
                else:
                    cur = [o, end]
                    ranges.append(cur)
            else:
                cur = None
        return [(a, b) for a, b in ranges]

    @torch.no_grad()
    def step(self, closure=None):
        ops.join_wgrad()  # weight-gradient GEMMs run on a side stream
        ranges = self.live_ranges()
        if not ranges:
            return None  # reference: every .grad is None -> Adam touches nothing
        self._step += 1
        g = self.param_groups[0]
        clip = g["clip_value"] if g["clip_value"] is not None else 3.0e38
        for a, b in ranges:
            ops.call("sfk_clip_adam_l2", self.flat_p.data_ptr() + 4 * a, self.flat_g.data_ptr() + 4 * a, self.flat_m.data_ptr() + 4 * a,
                     self.flat_v.data_ptr() + 4 * a, self.flat_w16.data_ptr() + 2 * a, b - a, float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                     float(g["weight_decay"]), float(clip), self._step, ops.stream())
        ops.bump_weight_epoch()
        return None


def mark_live(p: torch.nn.Parameter) -> None:
    p._sfk_live = True


def ema_update(teacher_opt: FusedClipAdam, student_opt: FusedClipAdam, decay: float) -> None:
    """teacher <- decay*teacher + (1-decay)*student over the whole flat buffer (decay = 0 is the hard reset of TS:202-204)."""
    if teacher_opt.flat_p.numel() != student_opt.flat_p.numel():
        raise ValueError("teacher and student have different parameter layouts")
    ops.call("sfk_ema_lerp", teacher_opt.flat_p.data_ptr(), student_opt.flat_p.data_ptr(), teacher_opt.flat_w16.data_ptr(),
             teacher_opt.flat_p.numel(), float(decay), ops.stream())
    ops.bump_weight_epoch()
