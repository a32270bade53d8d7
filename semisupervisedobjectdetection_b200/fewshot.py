"""Few-shot / domain-prompt losses of main_segformer/segFormer_fewshot_learning.py (FS) on libsfk kernels: the cosine
similarity between the sigmoid cls tokens of two image categories (FS:185-188) and the inter / intra loss mix of
train_autoencoder_iteration (FS:191-237).  The autoencoder ('mse', num_labels = 3) part runs through
SegFormerModel.predict(..., use_loss='mse', output_cls_token=True) as in the reference.
"""
from __future__ import annotations

import torch

from . import ops


class _CosineMeanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        a = ops._chk(a, torch.float32, "cosine input")
        b = ops._chk(b, torch.float32, "cosine input")
        if a.shape != b.shape or a.dim() != 2:
            raise ValueError(f"cosine similarity needs two [n, d] matrices, got {tuple(a.shape)} and {tuple(b.shape)}")
        n, d = a.shape
        stats = torch.empty(n, 3, device=a.device, dtype=torch.float32)
        out = torch.empty(1, device=a.device, dtype=torch.float32)
        ops.call("sfk_cosine_mean_fwd", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), n, d, stats.data_ptr(), out.data_ptr(),
                 ops.stream())
        ctx.save_for_backward(a, b, stats)
        return out[0]

    @staticmethod
    def backward(ctx, dloss):
        a, b, stats = ctx.saved_tensors
        n, d = a.shape
        da = torch.empty_like(a) if ctx.needs_input_grad[0] else None
        db = torch.empty_like(b) if ctx.needs_input_grad[1] else None
        dl = dloss.reshape(1).float().contiguous()
        ops.call("sfk_cosine_mean_bwd", a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), n, d, stats.data_ptr(), dl.data_ptr(),
                 ops._p(da), d, ops._p(db), d, ops.stream())
        return da, db


def compute_similarity(mat1: torch.Tensor, mat2: torch.Tensor) -> torch.Tensor:
    """FS:185-188: mean over the batch of CosineSimilarity(dim=1) between two [B, 1, C] (or [B, C]) cls-token batches."""
    if mat1.dim() == 3:
        mat1 = mat1.squeeze(1)
    if mat2.dim() == 3:
        mat2 = mat2.squeeze(1)
    with ops.device_of(mat1):
        return _CosineMeanFn.apply(mat1.float(), mat2.float())


def fewshot_losses(category_1_loss, category_1_cls_token, category_2_loss, category_2_cls_token):
    """FS:216-236: inter-category loss 0.5 + 0.5 cos(c1, c2), intra-category losses 0.5 - 0.5 cos(first half, last half),
    summed with weight 100 onto each category's reconstruction loss.  Returns
    (summation_loss, intra_loss_1, intra_loss_2, inter_loss)."""
    batch_size = min(category_1_cls_token.shape[0], category_2_cls_token.shape[0])
    half = batch_size // 2
    inter_loss = 0.5 + 0.5 * compute_similarity(category_1_cls_token[:batch_size], category_2_cls_token[:batch_size])
    intra_loss_1 = 0.5 - 0.5 * compute_similarity(category_1_cls_token[:half], category_1_cls_token[-half:])
    intra_loss_2 = 0.5 - 0.5 * compute_similarity(category_2_cls_token[:half], category_2_cls_token[-half:])
    category_1_summation_loss = category_1_loss + 100 * inter_loss + 100 * intra_loss_1
    category_2_summation_loss = category_2_loss + 100 * inter_loss + 100 * intra_loss_2
    summation_loss = (category_1_summation_loss + category_2_summation_loss) / 2
    return summation_loss, intra_loss_1, intra_loss_2, inter_loss


def train_autoencoder_iteration(model, category_1_img, category_2_img):
    """FS:191-237 without the data-loader bookkeeping: reconstruction ('mse') loss of both categories through
    SegFormerModel.predict, the cls-token cosine losses, and their mix.  Same return tuple as the reference."""
    category_1_img = category_1_img.to(device=model.device, dtype=torch.float32)
    category_2_img = category_2_img.to(device=model.device, dtype=torch.float32)
    category_1_loss, category_1_predicted, category_1_cls_token = model.predict(category_1_img, category_1_img, output_cls_token=True,
                                                                                isEval=False, use_loss='mse')
    category_2_loss, category_2_predicted, category_2_cls_token = model.predict(category_2_img, category_2_img, output_cls_token=True,
                                                                                isEval=False, use_loss='mse')
    summation_loss, intra_loss_1, intra_loss_2, inter_loss = fewshot_losses(category_1_loss, category_1_cls_token, category_2_loss,
                                                                            category_2_cls_token)
    return summation_loss, category_1_img, category_1_predicted, category_1_loss, category_2_loss, intra_loss_1, intra_loss_2, inter_loss
