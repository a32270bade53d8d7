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
        # torch.optim.Adam keeps one step count per parameter: a parameter that first receives a gradient after K steps
        # (unfroze_encoder, add_cls_token mid-run) starts its bias correction at step 1, not K + 1
        self._steps = [0] * len(ps)
        self._hooks = []
        for p, o in zip(ps, offs):
            n = p.numel()
            view = self.flat_p[o:o + n].view(p.shape)
            view.copy_(p.data)
            p.data = view
            p.grad = self.flat_g[o:o + n].view(p.shape)
            p._sfk_live = False
            p._sfk_w16 = self.flat_w16[o:o + n].view(p.shape)
            # gradients written by autograd itself (host-API ops such as the cls_proj Linear, MS:721) land in the same flat
            # buffer (p.grad is a view of it and AccumulateGrad adds in place): mark such parameters live as well
            self._hooks.append(p.register_post_accumulate_grad_hook(mark_live))
        self.sync_shadow()

    def sync_shadow(self):
        """Recast the whole flat bf16 shadow from the fp32 masters (one launch); called after host-side weight writes."""
        n = self.flat_p.numel()
        with torch.cuda.device(self.flat_p.device):
            ops.call("sfk_cast_2d", self.flat_p.data_ptr(), n, self.flat_w16.data_ptr(), n, 1, n, n, ops.stream())
        for p in self._params:
            p._sfk_w16_version = p._version
        ops.bump_weight_epoch()

    # --- gradient bookkeeping ---
    def zero_grad(self, set_to_none: bool = False):
        with torch.cuda.device(self.flat_p.device):
            ops.join_wgrad()
        self.flat_g.zero_()
        for p, o in zip(self._params, self._offsets):
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * o:
                p.grad = self.flat_g[o:o + p.numel()].view(p.shape)
            p._sfk_live = False

    def live_ranges(self):
        """Contiguous [start, end, step) element ranges of parameters whose gradient was produced this step; a range only
        spans parameters with the same per-parameter step count (they share one bias correction)."""
        ranges, cur = [], None
        for i, (p, o) in enumerate(zip(self._params, self._offsets)):
            live = p.requires_grad and getattr(p, "_sfk_live", False)
            end = o + (p.numel() + 7) // 8 * 8
            if live:
                if cur is not None and cur[1] == o and cur[2] == self._steps[i]:
                    cur[1] = end
                    cur[3].append(i)
                else:
                    cur = [o, end, self._steps[i], [i]]
                    ranges.append(cur)
            else:
                cur = None
        return [(a, b, st, idx) for a, b, st, idx in ranges]

    @torch.no_grad()
    def step(self, closure=None):
        with torch.cuda.device(self.flat_p.device):
            ops.join_wgrad()  # weight-gradient GEMMs run on a side stream
            ranges = self.live_ranges()
            if not ranges:
                return None  # reference: every .grad is None -> Adam touches nothing
            g = self.param_groups[0]
            clip = g["clip_value"] if g["clip_value"] is not None else 3.0e38
            for a, b, st, idx in ranges:
                ops.call("sfk_clip_adam_l2", self.flat_p.data_ptr() + 4 * a, self.flat_g.data_ptr() + 4 * a,
                         self.flat_m.data_ptr() + 4 * a, self.flat_v.data_ptr() + 4 * a, self.flat_w16.data_ptr() + 2 * a, b - a,
                         float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]), float(g["weight_decay"]),
                         float(clip), st + 1, ops.stream())
                for i in idx:
                    self._steps[i] = st + 1
            ops.bump_weight_epoch()
        return None

    # --- checkpointing: the layout of torch.optim.Adam's state dict (state[i] = {step, exp_avg, exp_avg_sq}) ---
    def state_dict(self):
        """torch.optim.Adam-compatible state dict built from the flat moment buffers (parameters that never stepped carry
        no state, as in torch)."""
        state = {}
        for i, (p, o) in enumerate(zip(self._params, self._offsets)):
            if self._steps[i] > 0:
                n = p.numel()
                state[i] = {"step": torch.tensor(float(self._steps[i])),
                            "exp_avg": self.flat_m[o:o + n].view(p.shape).clone(),
                            "exp_avg_sq": self.flat_v[o:o + n].view(p.shape).clone()}
        groups = []
        for gr in self.param_groups:
            packed = {k: v for k, v in gr.items() if k != "params"}
            packed["params"] = list(range(len(self._params)))
            groups.append(packed)
        return {"state": state, "param_groups": groups}

    @torch.no_grad()
    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != len(self.param_groups) or len(groups[0]["params"]) != len(self._params):
            raise ValueError("loaded state dict has a different number of parameters / parameter groups")
        for gr, saved in zip(self.param_groups, groups):
            for k, v in saved.items():
                if k != "params":
                    gr[k] = v
        self.flat_m.zero_()
        self.flat_v.zero_()
        self._steps = [0] * len(self._params)
        for key, st in state_dict["state"].items():
            i = int(key)
            p, o = self._params[i], self._offsets[i]
            n = p.numel()
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} does not match {tuple(p.shape)}")
            self.flat_m[o:o + n].view(p.shape).copy_(st["exp_avg"])
            self.flat_v[o:o + n].view(p.shape).copy_(st["exp_avg_sq"])
            self._steps[i] = int(float(st["step"]))


def mark_live(p: torch.nn.Parameter) -> None:
    p._sfk_live = True


def ema_update(teacher_opt: FusedClipAdam, student_opt: FusedClipAdam, decay: float) -> None:
    """teacher <- decay*teacher + (1-decay)*student over the whole flat buffer (decay = 0 is the hard reset of TS:202-204)."""
    if teacher_opt.flat_p.numel() != student_opt.flat_p.numel():
        raise ValueError("teacher and student have different parameter layouts")
    if teacher_opt.flat_p.device != student_opt.flat_p.device:
        raise ValueError("teacher and student live on different devices")
    with torch.cuda.device(teacher_opt.flat_p.device):
        ops.call("sfk_ema_lerp", teacher_opt.flat_p.data_ptr(), student_opt.flat_p.data_ptr(), teacher_opt.flat_w16.data_ptr(),
                 teacher_opt.flat_p.numel(), float(decay), ops.stream())
        ops.bump_weight_epoch()
