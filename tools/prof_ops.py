"""Stage-3-shaped (MiT-b5, 16 tiles of 512x512) launches of the row-wise / attention kernels, for one ncu --set full capture.
Usage: python tools/prof_ops.py"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import check_ops as CO
from semisupervisedobjectdetection_b200 import _lib as L

dev = "cuda"
lib = L.load()
L.ensure_device()
CO.lib = lib
torch.manual_seed(0)
rows, C = 16384, 320
x = torch.randn(rows, C, device=dev)
g = torch.randn(C, device=dev)
b = torch.randn(C, device=dev)
y = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
mean = torch.empty(rows, device=dev)
rstd = torch.empty(rows, device=dev)
dy = torch.randn(rows, C, device=dev).bfloat16()
dres = torch.randn(rows, C, device=dev)
dx = torch.empty(rows, C, device=dev)
dg = torch.zeros(C, device=dev)
db = torch.zeros(C, device=dev)
scale = torch.rand(16, device=dev)
y16 = torch.empty(rows, C, device=dev, dtype=torch.bfloat16)
B, heads, T, Tk, D = 16, 5, 1024, 256, 64
q = torch.randn(B, T, C, device=dev).bfloat16()
kv = torch.randn(B, Tk, 2 * C, device=dev).bfloat16()
dO = torch.randn(B, T, C, device=dev).bfloat16()
delta = torch.empty(B, heads, T, device=dev)
dq = torch.zeros(B, T, C, device=dev)
dkv = torch.zeros(B, Tk, 2 * C, device=dev)
sc = 1.0 / math.sqrt(D)
for it in range(3):
    L.check(lib.sfk_layernorm_fwd(x.data_ptr(), 0, 0, 0, 0, g.data_ptr(), b.data_ptr(), y.data_ptr(), 1, mean.data_ptr(),
                                  rstd.data_ptr(), rows, C, 1e-5, L.stream_ptr()))
    L.check(lib.sfk_layernorm_bwd(dy.data_ptr(), 1, x.data_ptr(), 0, mean.data_ptr(), rstd.data_ptr(), g.data_ptr(),
                                  dres.data_ptr(), dx.data_ptr(), 0, dg.data_ptr(), db.data_ptr(), rows, C, L.stream_ptr()))
    L.check(lib.sfk_scale_cast(x.data_ptr(), 0, y16.data_ptr(), 1, scale.data_ptr(), 1024, rows, C, L.stream_ptr()))
    o, lse = CO.run_attn_fwd(q, kv, heads, D, sc)
    L.check(lib.sfk_attention_bwd(q.data_ptr(), C, T * C, kv.data_ptr(), kv.data_ptr() + 2 * C, 2 * C, Tk * 2 * C, o.data_ptr(),
                                  dO.data_ptr(), C, T * C, lse.data_ptr(), delta.data_ptr(), dq.data_ptr(), C, T * C,
                                  dkv.data_ptr(), dkv.data_ptr() + 4 * C, 2 * C, Tk * 2 * C, B, heads, T, Tk, D, sc,
                                  L.stream_ptr()))
torch.cuda.synchronize()
print("done")
