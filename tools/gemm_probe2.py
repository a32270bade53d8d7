"""Timing-only experiment: GEMM main-loop cost per k-block against the operand ring depth.  Usage: python tools/gemm_probe2.py"""
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
M, K = 18944, 2560          # 148 row tiles x 40 k-blocks: one tile per CTA
a = torch.randn(M, K, device="cuda").bfloat16()
for bn in (64, 160, 256):
    b = torch.randn(bn, K, device="cuda").bfloat16()
    out = torch.empty(M, bn, device="cuda", dtype=torch.bfloat16)
    for st in (2, 3, 4, 5, 6, 8):
        L.check(lib.sfk_set_option(b"gemm_max_stages", st))
        t = time_graph(lambda: gemm(a, b, M, bn, K, out_bf16=True, out=out, block_n=bn))
        print(f"PROBE2 bn={bn:3d} max_stages={st}: {t:6.2f} us", flush=True)
L.check(lib.sfk_set_option(b"gemm_max_stages", 0))
