"""Decode-head kernel timings at the MiT-b5 512x512 / 16-tile shapes (CUDA-graph replay, device time per launch).
Usage: python tools/decode_bench.py   (env SFK_UPSAMPLE_SEP=0|1)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gemm_bench import time_graph
from semisupervisedobjectdetection_b200 import _lib as L

dev = "cuda"
lib = L.load()
L.ensure_device()
for opt in ("upsample_sep",):
    if os.environ.get("SFK_" + opt.upper()) is not None:
        L.check(lib.sfk_set_option(opt.encode(), int(os.environ["SFK_" + opt.upper()])))
        print(opt, "=", os.environ["SFK_" + opt.upper()], flush=True)
torch.manual_seed(0)
B, H, W, C = 16, 128, 128, 768
dims = ((64, 64), (32, 32), (16, 16))
z0 = torch.randn(B, H, W, C, device=dev).bfloat16()
zs = [torch.randn(B, h, w, C, device=dev).bfloat16() for (h, w) in dims]
bias = torch.randn(C, device=dev)
out = torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16)


def fwd():
    L.check(lib.sfk_upsample_add4(z0.data_ptr(), zs[0].data_ptr(), dims[0][0], dims[0][1], zs[1].data_ptr(), dims[1][0], dims[1][1],
                                  zs[2].data_ptr(), dims[2][0], dims[2][1], bias.data_ptr(), out.data_ptr(), B, H, W, C, L.stream_ptr()))


us = time_graph(fwd, reps=5)
print(f"DECODE upsample_add4 {B}x{H}x{W}x{C}: {us:7.1f} us  {2 * z0.numel() * 2 / us / 1e3:6.0f} GB/s algorithmic", flush=True)
dz = [torch.empty(B, h, w, C, device=dev, dtype=torch.bfloat16) for (h, w) in dims]
tot = 0.0
for i, (h, w) in enumerate(dims):
    def bwd(i=i, h=h, w=w):
        L.check(lib.sfk_upsample_concat_bwd(z0.data_ptr(), dz[i].data_ptr(), B, h, w, C, H, W, C, 0, L.stream_ptr()))
    us = time_graph(bwd, reps=5)
    tot += us
    print(f"DECODE upsample_concat_bwd -> {h}x{w}: {us:7.1f} us  {z0.numel() * 2 / us / 1e3:6.0f} GB/s algorithmic", flush=True)
print(f"DECODE transposed upsampling, three levels: {tot:7.1f} us", flush=True)
