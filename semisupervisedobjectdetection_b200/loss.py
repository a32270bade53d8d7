"""models/Loss.py of the reference (LOSS:16-95) on libsfk kernels: Dice coefficient, MSE and SegmentationLoss.

Same call signatures and return values; the reductions and their gradients run in loss.cu (two-stage deterministic sums).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import ops


def _as4d(t: torch.Tensor) -> torch.Tensor:
    if t.dim() == 3:
        return t.unsqueeze(1)
    if t.dim() == 2:
        return t[None, None]
    return t


class _DiceLossFn(torch.autograd.Function):
    """loss = 1 - mean_n (2|sum g p| + s) / (sum |p| + sum g + s) over channel 0 of pred/gt [N, C, ...]."""

    @staticmethod
    def forward(ctx, pred, gt, smooth, hard_threshold, sample_weight=None):
        pred = ops._chk(pred, torch.float32, "dice prediction")
        if not gt.is_cuda:
            raise RuntimeError("dice target must be a CUDA tensor")
        if gt.dtype not in (torch.int64, torch.float32):
            gt = gt.to(torch.int64) if not gt.is_floating_point() else gt.float()
        gt = gt.contiguous()
        N = pred.shape[0]
        P = pred[0, 0].numel()
        p_stride = pred[0].numel()
        g_stride = gt[0].numel()
        lib = ops.lib()
        ws = ops.scratch(lib.sfk_dice_workspace_bytes(N), pred.device, "loss")
        sums = torch.empty(N, 3, device=pred.device, dtype=torch.float32)
        loss = torch.empty(1, device=pred.device, dtype=torch.float32)
        if sample_weight is not None:
            sample_weight = ops._chk(sample_weight.detach(), torch.float32, "dice sample weights")
            if sample_weight.numel() != N:
                raise ValueError(f"dice: {sample_weight.numel()} sample weights for {N} samples")
        ops.call("sfk_dice_fwd", pred.data_ptr(), p_stride, gt.data_ptr(), int(gt.dtype == torch.int64), g_stride, N, P,
                 int(hard_threshold), float(smooth), ops._p(sample_weight), ws.data_ptr(), sums.data_ptr(), loss.data_ptr(),
                 ops.stream())
        ctx.meta = (N, P, p_stride, g_stride, float(smooth), hard_threshold)
        ctx.weighted = sample_weight is not None
        if ctx.needs_input_grad[0] and not hard_threshold:
            ctx.save_for_backward(pred, gt, sums, *([sample_weight] if sample_weight is not None else []))
        return loss[0]

    @staticmethod
    def backward(ctx, dloss):
        N, P, p_stride, g_stride, smooth, hard = ctx.meta
        if hard:
            return None, None, None, None, None
        pred, gt, sums, *w = ctx.saved_tensors
        dp = torch.zeros_like(pred)  # channels other than 0 get no gradient (LOSS:83 slices channel 0)
        dl = dloss.reshape(1).float().contiguous()
        ops.call("sfk_dice_bwd", pred.data_ptr(), p_stride, gt.data_ptr(), int(gt.dtype == torch.int64), g_stride, sums.data_ptr(),
                 dl.data_ptr(), smooth, ops._p(w[0] if w else None), dp.data_ptr(), p_stride, N, P, 0, ops.stream())
        return dp, None, None, None, None


class _MseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, denom):
        a = ops._chk(a, torch.float32, "mse input")
        b = ops._chk(b, torch.float32, "mse input")
        lib = ops.lib()
        ws = ops.scratch(lib.sfk_loss_workspace_bytes(), a.device, "loss")
        out = torch.empty(2, device=a.device, dtype=torch.float32)
        ops.call("sfk_mse_fwd", a.data_ptr(), b.data_ptr(), a.numel(), float(denom), ws.data_ptr(), out.data_ptr(), ops.stream())
        ctx.denom = float(denom)
        ctx.save_for_backward(a, b)
        return out[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        a, b = ctx.saved_tensors
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        dl = dloss.reshape(1).float().contiguous()
        ops.call("sfk_mse_bwd", a.data_ptr(), b.data_ptr(), dl.data_ptr(), ctx.denom, ops._p(da), ops._p(db), a.numel(), ops.stream())
        return da, db, None


def diceCoeff(pred, gt, smooth_value: float = 1.0, activation="softmax2d"):
    """LOSS:16-41. Only activation None / 'none' (what every reference call site passes, SFM:39-40) is implemented."""
    if activation not in (None, "none"):
        raise NotImplementedError("Activation implemented for 'none' only in the libsfk build (reference call sites use 'none')")
    return 1.0 - _DiceLossFn.apply(_as4d(pred), _as4d(gt), smooth_value, False)


def mse(predicted, gt):
    """LOSS:44-54: mean_n( sum((gt - predicted)^2) / (N*C) )."""
    n, c = gt.shape[0], gt.shape[1]
    return _MseFn.apply(gt.float().contiguous(), predicted.float().contiguous(), float(n * c * n))


def weighted_dice_loss(pred, gt, sample_weight, smooth_value: float = 1.0):
    """(1/N) sum_n w_n * (1 - dice_n): the per-sample form of LOSS:33-41 used by the north-star EMA step, where w_n is the
    keep flag of an unlabeled tile (TS:43 drops such tiles from the batch; here they stay, with zero loss and gradient)."""
    return _DiceLossFn.apply(_as4d(pred).float(), _as4d(gt), smooth_value, False, sample_weight)


class SegmentationLoss(nn.Module):
    __name__ = 'seg_loss'

    def __init__(self, num_classes, activation=None, loss_type='dice'):
        super(SegmentationLoss, self).__init__()
        self.activation = activation
        self.num_classes = num_classes
        self.loss_type = loss_type
        if num_classes != 1:
            raise NotImplementedError("num_classes != 1 is never used by the reference wrapper (SFM:39-41)")

    def forward(self, y_pred, y_true):
        y_pred, y_true = _as4d(y_pred), _as4d(y_true)
        if self.loss_type == 'mse':
            return mse(y_pred, y_true)
        if self.activation not in (None, "none"):
            raise NotImplementedError("only activation='none' is implemented (what the reference wrapper uses)")
        if self.loss_type == 'dice':
            loss = _DiceLossFn.apply(y_pred.float(), y_true, 1.0, False)
        elif self.loss_type == 'dice_argmax':
            loss = _DiceLossFn.apply(y_pred.float(), y_true, 1.0, True)
        else:
            raise ValueError("err lose type")
        if not loss.requires_grad:
            loss.requires_grad_(True)  # LOSS:94: lets a graph-less loss still call backward()
        return loss
