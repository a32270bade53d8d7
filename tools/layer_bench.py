"""Device time of one SegformerLayer forward / backward per MiT-b5 stage (16 tiles of 512x512), host overhead excluded:
each call sequence is captured in a CUDA graph (single stream, 4 layers per graph) and replayed.

Usage: python tools/layer_bench.py [--stages 1,2,3,4] [--ptl 0]
The per-step estimate printed at the end is sum_stage depth * (2 * fwd + bwd) (teacher + student forward, student backward).
"""
import argparse
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semisupervisedobjectdetection_b200 import _lib as L
from semisupervisedobjectdetection_b200 import ops

DEV = "cuda"
B = 16
STAGES = {1: (128, 64, 1, 8, 3), 2: (64, 128, 2, 4, 6), 3: (32, 320, 5, 2, 40), 4: (16, 512, 8, 1, 3)}  # H, C, heads, sr, depth


def rnd(*shape, dtype=torch.float32, scale=0.05):
    return (torch.randn(*shape, device=DEV) * scale).to(dtype)


def build(stage, ptl):
    H, C, heads, sr, depth = STAGES[stage]
    hid = 4 * C
    dims = ops.LayerDims(B, H, H, C, heads, sr, hid, ptl)
    keep = {}
    p = ops.LayerParams()
    g = ops.LayerGrads()

    def setp(name, t):
        keep["p_" + name] = t
        setattr(p, name, t.data_ptr())

    def setg(name, shape):
        t = torch.zeros(*shape, device=DEV)
        keep["g_" + name] = t
        setattr(g, name, t.data_ptr())

    bf = torch.bfloat16
    setp("wq", rnd(C, C, dtype=bf)); setp("wkv", rnd(2 * C, C, dtype=bf)); setp("wo", rnd(C, C, dtype=bf))
    if sr > 1:
        setp("wsr", rnd(C, sr * sr * C, dtype=bf)); setp("bsr", rnd(C)); setp("lnsr_g", 1 + rnd(C)); setp("lnsr_b", rnd(C))
    setp("w1", rnd(hid, C, dtype=bf)); setp("w2", rnd(C, hid, dtype=bf))
    setp("bq", rnd(C)); setp("bkv", rnd(2 * C)); setp("bo", rnd(C)); setp("b1", rnd(hid)); setp("b2", rnd(C))
    setp("ln1_g", 1 + rnd(C)); setp("ln1_b", rnd(C)); setp("ln2_g", 1 + rnd(C)); setp("ln2_b", rnd(C))
    setp("dw_wt", rnd(9, hid, scale=0.2)); setp("dw_b", rnd(hid))
    setp("drop_scale_attn", torch.ones(B, device=DEV)); setp("drop_scale_mlp", torch.ones(B, device=DEV))
    for n, shp in (("dwq", (C, C)), ("dwk", (C, C)), ("dwv", (C, C)), ("dwo", (C, C)), ("dw1", (hid, C)), ("dw2", (C, hid)),
                   ("dbq", (C,)), ("dbk", (C,)), ("dbv", (C,)), ("dbo", (C,)), ("db1", (hid,)), ("db2", (C,)),
                   ("dln1_g", (C,)), ("dln1_b", (C,)), ("dln2_g", (C,)), ("dln2_b", (C,)), ("ddw", (hid, 1, 3, 3)), ("ddb", (hid,))):
        setg(n, shp)
    if sr > 1:
        setg("dwsr", (C, C, sr, sr)); setg("dbsr", (C,)); setg("dlnsr_g", (C,)); setg("dlnsr_b", (C,))
    return dims, p, g, keep


def time_graph(fn, reps=4, replays=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(gr, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    gr.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(replays):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000.0 / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--stages", default="1,2,3,4")
    ap.add_argument("--ptl", type=int, default=0)
    ap.add_argument("--plain", action="store_true", help="no graphs: one forward and three backward calls (for an ncu launch list)")
    args = ap.parse_args()
    L.ensure_device()
    lib = L.load()
    for opt in ("gemm_occ2", "gemm_cluster2", "gemm_epi_tma", "gemm_bstat", "attn_bwd_persist"):
        if os.environ.get("SFK_" + opt.upper()) is not None:
            L.check(lib.sfk_set_option(opt.encode(), int(os.environ["SFK_" + opt.upper()])))
            print(opt, "=", os.environ["SFK_" + opt.upper()], flush=True)
    torch.manual_seed(0)
    total = 0.0
    for stage in [int(s) for s in args.stages.split(",")]:
        H, C, heads, sr, depth = STAGES[stage]
        ptl = args.ptl
        dims, p, g, keep = build(stage, ptl)
        N = H * H
        x = rnd(B, N, C, scale=1.0)
        front = rnd(B, ptl, C, scale=1.0) if ptl else None
        out = torch.empty(B, N, C, device=DEV)
        front_out = torch.empty(B, ptl, C, device=DEV) if ptl else None
        dout = rnd(B, N, C, scale=1.0)
        dfront_out = rnd(B, ptl, C, scale=1.0) if ptl else None
        dx = torch.empty(B, N, C, device=DEV)
        dfront = torch.empty(B, ptl, C, device=DEV) if ptl else None
        sb = lib.sfk_layer_saved_bytes(ctypes.byref(dims))
        cb = lib.sfk_layer_scratch_bytes(ctypes.byref(dims))
        saved = torch.empty(sb, dtype=torch.uint8, device=DEV)
        scratch = torch.empty(cb + 4096, dtype=torch.uint8, device=DEV)

        def fwd():
            L.check(lib.sfk_layer_fwd(ctypes.byref(dims), ctypes.byref(p), x.data_ptr(), ops._p(front), out.data_ptr(),
                                      ops._p(front_out), saved.data_ptr(), sb, scratch.data_ptr(), scratch.numel(), L.stream_ptr()),
                    "layer_fwd")

        def bwd():
            L.check(lib.sfk_layer_bwd(ctypes.byref(dims), ctypes.byref(p), ctypes.byref(g), x.data_ptr(), dout.data_ptr(),
                                      ops._p(dfront_out), dx.data_ptr(), ops._p(dfront), saved.data_ptr(), sb, scratch.data_ptr(),
                                      scratch.numel(), L.stream_ptr(), None, 0), "layer_bwd")

        if args.plain:
            fwd()
            for _ in range(3):
                bwd()
            torch.cuda.synchronize()
            continue
        tf = time_graph(fwd)
        tb = time_graph(bwd)
        est = depth * (2 * tf + tb)
        total += est
        print(f"LAYER stage {stage} (H={H} C={C} sr={sr} ptl={ptl}): fwd {tf:8.1f} us | bwd {tb:8.1f} us | x{depth} layers -> "
              f"{est/1000:.2f} ms per step", flush=True)
    print(f"LAYER TOTAL (single stream, encoder layers only): {total/1000:.2f} ms per step", flush=True)


if __name__ == "__main__":
    main()
