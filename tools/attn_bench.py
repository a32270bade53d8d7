"""Attention forward / backward timings on the four MiT-b5 stage shapes of the 16-tile 512x512 step (CUDA-graph replay of 20
launches, device time per launch).  Usage: python tools/attn_bench.py [--bwd]   (env SFK_ATTN_FWD256=0|1)"""
import argparse
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import check_ops as CO
from gemm_bench import time_graph
from semisupervisedobjectdetection_b200 import _lib as L

dev = "cuda"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bwd", action="store_true")
    args = ap.parse_args()
    lib = L.load()
    L.ensure_device()
    CO.lib = lib
    for opt in ("attn_fwd256", "attn_bwd_persist", "attn_dbg"):
        if os.environ.get("SFK_" + opt.upper()) is not None:
            L.check(lib.sfk_set_option(opt.encode(), int(os.environ["SFK_" + opt.upper()])))
            print(opt, "=", os.environ["SFK_" + opt.upper()], flush=True)
    torch.manual_seed(0)
    B = 16
    tot = 0.0
    for name, heads, T, depth in [("s1", 1, 16384, 3), ("s2", 2, 4096, 6), ("s3", 5, 1024, 40), ("s4", 8, 256, 3)]:
        D, Tk = 64, 256
        C = heads * D
        q = torch.randn(B, T, C, device=dev).bfloat16()
        kv = torch.randn(B, Tk, 2 * C, device=dev).bfloat16()
        sc = 1.0 / math.sqrt(D)
        o = torch.empty(B, T, C, device=dev, dtype=torch.bfloat16)
        lse = torch.empty(B, heads, T, device=dev, dtype=torch.float32)

        def fwd():
            L.check(lib.sfk_attention_fwd(q.data_ptr(), C, T * C, kv.data_ptr(), kv.data_ptr() + 2 * C, 2 * C, Tk * 2 * C, o.data_ptr(),
                                          C, T * C, lse.data_ptr(), B, heads, T, Tk, D, sc, L.stream_ptr()))

        if os.environ.get("SFK_ATTN_TRACE"):
            import ctypes
            L.check(lib.sfk_set_option(b"attn_dbg", 2))
            fwd()
            torch.cuda.synchronize()
            buf = (ctypes.c_longlong * 16)()
            L.check(lib.sfk_debug_attn_trace(buf))
            L.check(lib.sfk_set_option(b"attn_dbg", 0))
            t = list(buf)
            names = ["start", "setup", "qk_landed", "s_ready", "max_done", "p_written", "o_ready", "end"]
            print("TRACE " + name + ": " + " ".join(f"{n}={(t[i] - t[0]) / 1965.0:.2f}us" for i, n in enumerate(names)), flush=True)
        us = time_graph(fwd)
        fl = 4.0 * B * heads * T * Tk * D
        byt = 2.0 * (2 * q.numel() + kv.numel())
        print(f"ATTN fwd {name} B{B} h{heads} T{T} Tk{Tk}: {us:7.2f} us  {fl/us/1e6:6.1f} TF/s  {byt/us/1e3:6.0f} GB/s algorithmic", flush=True)
        tot += us * 2 * depth
        if args.bwd:
            dO = torch.randn(B, T, C, device=dev).bfloat16()
            delta = torch.empty(B, heads, T, device=dev)
            dq = torch.zeros(B, T, C, device=dev)
            dkv = torch.zeros(B, Tk, 2 * C, device=dev)

            def bwd():
                L.check(lib.sfk_attention_bwd(q.data_ptr(), C, T * C, kv.data_ptr(), kv.data_ptr() + 2 * C, 2 * C, Tk * 2 * C,
                                              o.data_ptr(), dO.data_ptr(), C, T * C, lse.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                              C, T * C, dkv.data_ptr(), dkv.data_ptr() + 4 * C, 2 * C, Tk * 2 * C, B, heads, T, Tk,
                                              D, sc, L.stream_ptr()))

            usb = time_graph(bwd)
            print(f"ATTN bwd {name}: {usb:7.2f} us  {2.5*fl/usb/1e6:6.1f} TF/s", flush=True)
    print(f"TOTAL fwd per step (teacher + student): {tot/1000:.2f} ms", flush=True)


if __name__ == "__main__":
    main()
