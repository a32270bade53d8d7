"""Timing-only experiments on the GEMM main loop (wrong results by design): per k-block cost of the MMA side and of the operand
stream for different tile widths.  Usage: python tools/gemm_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from check_gemm import gemm
from gemm_bench import time_graph
from semisupervisedobjectdetection_b200 import _lib as L

lib = L.load()
L.ensure_device()
if os.environ.get("SFK_GEMM_CLUSTER2") is not None:
    L.check(lib.sfk_set_option(b"gemm_cluster2", int(os.environ["SFK_GEMM_CLUSTER2"])))
M, K = 18944, 2560          # 148 row tiles x 40 k-blocks: one tile per CTA
a = torch.randn(M, K, device="cuda").bfloat16()
for bn in (64, 96, 128, 160, 192, 224, 256):
    b = torch.randn(bn, K, device="cuda").bfloat16()
    out = torch.empty(M, bn, device="cuda", dtype=torch.bfloat16)
    res = {}
    for dbg in (0, 4, 8, 12):
        L.check(lib.sfk_set_option(b"gemm_dbg", dbg))
        res[dbg] = time_graph(lambda: gemm(a, b, M, bn, K, out_bf16=True, out=out, block_n=bn))
    L.check(lib.sfk_set_option(b"gemm_dbg", 0))
    kb = K // 64
    print(f"PROBE bn={bn:3d}: full {res[0]:6.2f} us | no-load {res[4]:6.2f} | no-mma {res[8]:6.2f} | neither {res[12]:6.2f} || per k-block: "
          f"full {(res[0]-res[12])/kb*1000:5.0f} ns, mma-only {(res[4]-res[12])/kb*1000:5.0f} ns, load-only {(res[8]-res[12])/kb*1000:5.0f} ns "
          f"(operands {16 + bn // 8} KB)", flush=True)
