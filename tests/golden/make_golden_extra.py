"""Fixtures for the rows next to the hot path: few-shot cosine losses and the device input pipeline.

Run in the build container only:   python tests/golden/make_golden_extra.py

* few-shot: `compute_similarity` is extracted (AST) from the UNMODIFIED /root/reference/main_segformer/
  segFormer_fewshot_learning.py and executed with the module-level `similarity_loss = torch.nn.CosineSimilarity(dim=1)` the script
  sets up (FS:353); the inter / intra mix (FS:216-236) is evaluated with that function; autograd gradients are stored too.
* input pipeline: albumentations is not installed here, so the chain of config.py:27-38 is executed with the cv2 / numpy calls
  albumentations 1.3.0 makes (cv2.flip, np.rot90, float32 multiply by 1/255, cv2.resize INTER_LINEAR / INTER_NEAREST), with the
  installed cv2; inputs come from oracle/synth.py so only outputs are stored.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import synth  # noqa: E402

PIPE_CASES = {
    # name: (tile H, W, C, crop, out, [params rows])
    "small": (61, 67, 4, 50, 64, [(3, 5, 0, 0), (11, 0, 1, 0), (0, 17, 2, 0), (9, 9, 3, 0), (1, 2, 4, 0), (10, 16, 5, 0), (8, 1, 6, 0)]),
    "real": (523, 530, 4, 500, 512, [(23, 30, 0, 0), (0, 0, 4, 0), (11, 7, 1, 0)]),
}


def pipe_inputs(name):
    H, W, C, crop, out, rows = PIPE_CASES[name]
    B = len(rows)
    tiles = (synth.uniform01(f"pipe/{name}/tiles", (B, H, W, C)) * 256).to(torch.uint8)
    masks = (synth.uniform01(f"pipe/{name}/masks", (B, H // 8 + 1, W // 8 + 1)) > 0.6).to(torch.uint8) * 255
    masks = masks.repeat_interleave(8, 1).repeat_interleave(8, 2)[:, :H, :W].contiguous()   # blocky site masks
    return tiles, masks, torch.tensor(rows, dtype=torch.int32), crop, out


def load_compute_similarity():
    path = os.path.join(REF, "main_segformer", "segFormer_fewshot_learning.py")
    tree = ast.parse(open(path).read())
    wanted = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "compute_similarity"]
    ns = {"torch": torch, "similarity_loss": torch.nn.CosineSimilarity(dim=1)}
    exec(compile(ast.Module(body=wanted, type_ignores=[]), path, "exec"), ns)
    return ns["compute_similarity"]


def fewshot_goldens():
    cs = load_compute_similarity()
    out = {}
    for tag, (b1, b2) in {"b6": (6, 6), "b5_b4": (5, 4), "b2": (2, 2)}.items():
        c1 = torch.sigmoid((synth.uniform01(f"fs/{tag}/c1", (b1, 1, 512)) - 0.5) * 4).requires_grad_(True)
        c2 = torch.sigmoid((synth.uniform01(f"fs/{tag}/c2", (b2, 1, 512)) - 0.5) * 4).requires_grad_(True)
        l1, l2 = torch.tensor(0.0371), torch.tensor(0.0412)
        bs = min(b1, b2)
        inter = 0.5 + 0.5 * cs(c1[:bs, :, :], c2[:bs, :, :])
        intra1 = 0.5 - 0.5 * cs(c1[:bs // 2, :, :], c1[-(bs // 2):, :, :])
        intra2 = 0.5 - 0.5 * cs(c2[:bs // 2, :, :], c2[-(bs // 2):, :, :])
        total = ((l1 + 100 * inter + 100 * intra1) + (l2 + 100 * inter + 100 * intra2)) / 2
        total.backward()
        out[tag] = dict(inter=inter.detach().clone(), intra1=intra1.detach().clone(), intra2=intra2.detach().clone(),
                        total=total.detach().clone(), g1=c1.grad.clone(), g2=c2.grad.clone(),
                        sim=cs(c1[:bs].detach(), c2[:bs].detach()).clone())
    return out


def pipeline_goldens():
    import cv2
    out = {"cv2_version": cv2.__version__}
    for name in PIPE_CASES:
        tiles, masks, params, crop, osz = pipe_inputs(name)
        imgs, ms = [], []
        for b in range(tiles.shape[0]):
            y0, x0, op = [int(v) for v in params[b, :3]]
            img = tiles[b].numpy()[:, :, 0:3][y0:y0 + crop, x0:x0 + crop]        # DL:64 drops alpha; RandomCrop
            m = masks[b].numpy()[y0:y0 + crop, x0:x0 + crop]
            if op == 1:
                img, m = cv2.flip(np.ascontiguousarray(img), 1), cv2.flip(np.ascontiguousarray(m), 1)
            elif op == 2:
                img, m = cv2.flip(np.ascontiguousarray(img), 0), cv2.flip(np.ascontiguousarray(m), 0)
            elif op >= 3:
                img, m = np.ascontiguousarray(np.rot90(img, op - 3)), np.ascontiguousarray(np.rot90(m, op - 3))
            img = img.astype(np.float32)
            img *= np.reciprocal(np.array([255.0, 255.0, 255.0], dtype=np.float32))   # functional.normalize, mean 0
            img = cv2.resize(img, dsize=(osz, osz), interpolation=cv2.INTER_LINEAR)
            m = cv2.resize(np.ascontiguousarray(m), dsize=(osz, osz), interpolation=cv2.INTER_NEAREST).astype(np.float32)
            m = (m - np.min(m)) / (np.max(m) - np.min(m))                        # DL:85-86
            imgs.append(torch.from_numpy(np.ascontiguousarray(np.rollaxis(img, 2, 0))))
            ms.append(torch.from_numpy(m))
        img_t, m_t = torch.stack(imgs), torch.stack(ms)
        if name == "real":   # keep the fixture small: a strided sample plus sums
            out[name] = dict(img_sample=img_t[:, :, ::5, ::7].clone(), mask_sample=m_t[:, ::5, ::7].clone(),
                             img_sum=img_t.double().sum(dim=(1, 2, 3)), mask_sum=m_t.double().sum(dim=(1, 2)))
        else:
            out[name] = dict(img=img_t, mask=m_t)
    return out


def main():
    res = {"fewshot": fewshot_goldens(), "pipeline": pipeline_goldens()}
    path = os.path.join(HERE, "extra.pt")
    torch.save(res, path)
    print("extra.pt", os.path.getsize(path) // 1024, "KiB;", {k: float(v["total"]) for k, v in res["fewshot"].items()})


if __name__ == "__main__":
    main()
