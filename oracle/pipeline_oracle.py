"""TEST INFRASTRUCTURE ONLY (see oracle/README.md): CPU restatement of the reference's training input pipeline.

config.py:27-38 (albumentations==1.3.0, opencv_python==4.7.0.72 pinned in requirements.txt:2,60 - neither albumentations nor its
exact cv2 build is in this image, so the chain is restated from the published algorithms and anchored on the installed
cv2 4.13 by tests/golden/make_golden_extra.py, which stores cv2's outputs as fixtures):
  RandomCrop(500, 500)        albumentations functional.random_crop: img[y0:y0+h, x0:x0+w]
  HorizontalFlip / VerticalFlip   cv2.flip(img, 1) / cv2.flip(img, 0)
  RandomRotate90              np.rot90(img, factor)
  Normalize(mean 0, std 255, max_pixel_value 1)   functional.normalize: img.astype(float32) * np.reciprocal(255, float32)
  Resize(512, 512)            cv2.resize(INTER_LINEAR) for the image, INTER_NEAREST for the mask
and archaeological_georgia_biostyle_dataloader.py:59-108 (DL): channel 0 of the mask, min-max scaling (DL:85-86), channels first
(DL:102).  Pure numpy; loops are vectorised so that a 512 x 512 tile takes milliseconds.
"""
from __future__ import annotations

import numpy as np

OP_NONE, OP_HFLIP, OP_VFLIP, OP_ROT90 = 0, 1, 2, 3


def apply_op(img: np.ndarray, op: int) -> np.ndarray:
    if op == OP_HFLIP:
        return img[:, ::-1]
    if op == OP_VFLIP:
        return img[::-1]
    if op >= OP_ROT90:
        return np.rot90(img, op - OP_ROT90)
    return img


def _linear_coords(dst_size: int, src_size: int):
    """cv2 resize.cpp (INTER_LINEAR): fx = (dx + 0.5) * scale - 0.5 in double, cast to float; sx = floor(fx); fx -= sx."""
    scale = float(src_size) / float(dst_size)
    f = ((np.arange(dst_size, dtype=np.float64) + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int64)
    w = (f - s.astype(np.float32)).astype(np.float32)
    lo = s < 0
    s[lo], w[lo] = 0, 0.0
    hi = s >= src_size - 1
    s[hi], w[hi] = src_size - 1, 0.0
    s1 = np.minimum(s + 1, src_size - 1)
    return s, s1, w


def resize_linear(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """float32 HWC bilinear resize, horizontal pass first then vertical, all in float32 (cv2's HResize / VResize order)."""
    y0, y1, wy = _linear_coords(out_h, img.shape[0])
    x0, x1, wx = _linear_coords(out_w, img.shape[1])
    wx = wx[None, :, None]
    rows = img[:, x0] * (np.float32(1) - wx) + img[:, x1] * wx
    wy = wy[:, None, None]
    return (rows[y0] * (np.float32(1) - wy) + rows[y1] * wy).astype(np.float32)


def resize_nearest(img: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """cv2 INTER_NEAREST: sx = min(floor(dx * scale), src - 1)."""
    ys = np.minimum(np.floor(np.arange(out_h) * (img.shape[0] / out_h)).astype(np.int64), img.shape[0] - 1)
    xs = np.minimum(np.floor(np.arange(out_w) * (img.shape[1] / out_w)).astype(np.int64), img.shape[1] - 1)
    return img[ys][:, xs]


def augment_tile(tile_u8: np.ndarray, params, crop: int = 500, out_hw=(512, 512), mask_u8: np.ndarray = None):
    """One tile through the chain. tile_u8: uint8 [H, W, C>=3]; params = (y0, x0, op, _).  Returns (image f32 [3, OH, OW],
    mask f32 [OH, OW] | None)."""
    y0, x0, op = int(params[0]), int(params[1]), int(params[2])
    img = apply_op(tile_u8[y0:y0 + crop, x0:x0 + crop, 0:3], op)
    img = img.astype(np.float32) * np.reciprocal(np.float32(255.0))
    img = resize_linear(img, out_hw[0], out_hw[1])
    out = np.ascontiguousarray(np.rollaxis(img, 2, 0))
    m = None
    if mask_u8 is not None:
        mm = mask_u8 if mask_u8.ndim == 2 else mask_u8[:, :, 0]
        mm = resize_nearest(apply_op(mm[y0:y0 + crop, x0:x0 + crop], op), out_hw[0], out_hw[1]).astype(np.float32)
        with np.errstate(invalid="ignore", divide="ignore"):
            m = (mm - mm.min()) / (mm.max() - mm.min())   # DL:85-86 (a constant mask is 0/0 there too)
    return out, m
