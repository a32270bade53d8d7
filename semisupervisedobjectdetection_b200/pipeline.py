"""Training input pipeline of the reference on the device.

config.py:27-38 builds the albumentations chain RandomCrop(500, 500) -> OneOf([HorizontalFlip, VerticalFlip, RandomRotate90],
p=0.75) -> Normalize(mean 0, std 255, max_pixel_value 1) -> Resize(512, 512); archaeological_georgia_biostyle_dataloader.py
(DL:59-108) applies it to the uint8 PNG tile (and, with INTER_NEAREST, to channel 0 of the mask, which it then min-max scales)
and moves channels first.  Here the host only draws the random decisions (one row of four ints per tile) and ships the raw uint8
tiles; crop, flip / rot90, /255, bilinear resize and the NCHW transpose run in one libsfk kernel, so a step uploads 1 byte per
pixel channel instead of 4.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

OP_NONE, OP_HFLIP, OP_VFLIP, OP_ROT90 = 0, 1, 2, 3   # OP_ROT90 + k == np.rot90(img, k)


class TileAugmenter:
    def __init__(self, crop: int = 500, out_hw=(512, 512), p_oneof: float = 0.75):
        self.crop, self.out_hw, self.p_oneof = crop, tuple(out_hw), p_oneof

    def draw_params(self, batch: int, tile_h: int, tile_w: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """The random decisions of the albumentations chain as an int32 [B, 4] tensor {crop_y0, crop_x0, op, 0}:
        uniform crop corner; with probability p one of the three transforms, uniformly; RandomRotate90 draws k in 0..3."""
        g = generator
        y0 = torch.randint(0, tile_h - self.crop + 1, (batch,), generator=g)
        x0 = torch.randint(0, tile_w - self.crop + 1, (batch,), generator=g)
        apply = torch.rand(batch, generator=g) < self.p_oneof
        which = torch.randint(0, 3, (batch,), generator=g)
        k = torch.randint(0, 4, (batch,), generator=g)
        op = torch.where(which == 0, torch.full_like(which, OP_HFLIP),
                         torch.where(which == 1, torch.full_like(which, OP_VFLIP), OP_ROT90 + k))
        op = torch.where(apply, op, torch.zeros_like(op))
        return torch.stack([y0, x0, op, torch.zeros_like(op)], dim=1).to(torch.int32)

    def __call__(self, tiles_u8: torch.Tensor, params: torch.Tensor, masks_u8: Optional[torch.Tensor] = None, valid_hw=None):
        """tiles_u8: uint8 [B, H, W, C>=3] on the device (channels 0..2 used); params: int32 [B, 4] (draw_params);
        masks_u8: optional uint8 [B, H, W] or [B, H, W, C] (channel 0 used); valid_hw: the (rows, columns) of every tile the
        crop may touch - DL:64 cuts the 23-row Bing banner off the bottom of the decoded PNG, here the raw decode is passed and
        valid_hw = (H - 23, W) keeps the crop out of it without a copy.
        Returns (images f32 [B, 3, OH, OW], masks f32 [B, OH, OW] | None)."""
        with ops.device_of(tiles_u8):
            if tiles_u8.dtype != torch.uint8 or tiles_u8.dim() != 4 or tiles_u8.shape[3] < 3:
                raise ValueError("tiles must be uint8 [B, H, W, C>=3]")
            tiles_u8 = ops._chk(tiles_u8, torch.uint8, "tiles")
            B, H, W, C = tiles_u8.shape
            if params.shape != (B, 4) or params.dtype != torch.int32:
                raise ValueError("params must be int32 [B, 4]")
            host_params = params
            params = ops._chk(params.to(tiles_u8.device, non_blocking=True), torch.int32, "augmentation parameters")
            OH, OW = self.out_hw
            out = torch.empty(B, 3, OH, OW, device=tiles_u8.device, dtype=torch.float32)
            mask_out = minmax = None
            m_args = (None, 0, 0, 0)
            if masks_u8 is not None:
                masks_u8 = ops._chk(masks_u8, torch.uint8, "masks")
                if masks_u8.dim() == 3:
                    masks_u8 = masks_u8.unsqueeze(-1)
                if masks_u8.shape[:3] != (B, H, W):
                    raise ValueError("masks must have the tiles' [B, H, W]")
                mc = masks_u8.shape[3]
                mask_out = torch.empty(B, OH, OW, device=tiles_u8.device, dtype=torch.float32)
                minmax = torch.empty(B, 2, device=tiles_u8.device, dtype=torch.int32)
                m_args = (masks_u8.data_ptr(), H * W * mc, W * mc, mc)
            vh, vw = (H, W) if valid_hw is None else (int(valid_hw[0]), int(valid_hw[1]))
            if vh > H or vw > W:
                raise ValueError("valid_hw exceeds the tile")
            lim = torch.tensor([vh - self.crop, vw - self.crop], dtype=torch.int32)
            pc = host_params.cpu() if host_params.is_cuda else host_params
            if bool((pc[:, :2] < 0).any()) or bool((pc[:, :2] > lim).any()) or bool((pc[:, 2] < 0).any()) or bool((pc[:, 2] > 6).any()):
                raise ValueError("augmentation parameters out of range for this tile size")
            ops.call("sfk_tile_augment_u8", tiles_u8.data_ptr(), H * W * C, W * C, C, vh, vw, params.data_ptr(), self.crop, out.data_ptr(),
                     B, OH, OW, m_args[0], m_args[1], m_args[2], m_args[3], ops._p(mask_out), ops._p(minmax), ops.stream())
            return out, mask_out
