"""Per-shape GEMM timings of one MiT-b5 teacher-student step (16 tiles of 512x512 per GPU), host overhead excluded.

Every (shape, layout, epilogue) class of the step is launched 20 times inside one CUDA graph and the graph is replayed, so
the number is the device time of back-to-back launches (what the kernel costs inside the step), not the ~17 us ctypes call.
cuBLAS (torch.matmul, same graph treatment) is printed for context only; it is not used by the product.

Usage: python tools/gemm_bench.py [--quick] [--json out.json] [--filter substr]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from check_gemm import gemm
from semisupervisedobjectdetection_b200 import _lib as L

DEV = "cuda"
B = 16
STAGES = [  # tokens per tile, C, sr, depth
    (16384, 64, 8, 3), (4096, 128, 4, 6), (1024, 320, 2, 40), (256, 512, 1, 3)]


def step_shapes():
    """[(name, kind, M, N, K, count_per_step)] kind: fwd_bf16 | fwd_res | dgrad | dgrad_res | wgrad"""
    out = []
    for si, (n, C, sr, depth) in enumerate(STAGES, 1):
        T, KV, H = B * n, B * 256, 4 * C
        f, b = 2 * depth, depth  # teacher + student forward, student backward
        out += [(f"s{si}.q", "fwd_bf16", T, C, C, f), (f"s{si}.kv", "fwd_bf16", KV, 2 * C, C, f),
                (f"s{si}.o", "fwd_res", T, C, C, f), (f"s{si}.ffn1", "fwd_bf16", T, H, C, f),
                (f"s{si}.ffn2", "fwd_res", T, C, H, f)]
        if sr > 1:
            out += [(f"s{si}.sr", "fwd_f32", KV, C, sr * sr * C, f)]
        out += [(f"s{si}.ffn2.dgrad", "dgrad", T, H, C, b), (f"s{si}.ffn1.dgrad", "dgrad", T, C, H, b),
                (f"s{si}.o.dgrad", "dgrad", T, C, C, b), (f"s{si}.q.dgrad", "dgrad_res", T, C, C, b)]
        if sr > 1:
            out += [(f"s{si}.kv.dgrad", "dgrad", KV, C, 2 * C, b), (f"s{si}.sr.dgrad", "dgrad", KV, sr * sr * C, C, b)]
        else:
            out += [(f"s{si}.kv.dgrad", "dgrad_f32", KV, C, 2 * C, b)]
        out += [(f"s{si}.ffn2.wgrad", "wgrad", C, H, T, b), (f"s{si}.ffn1.wgrad", "wgrad", H, C, T, b),
                (f"s{si}.o.wgrad", "wgrad", C, C, T, 2 * b), (f"s{si}.kv.wgrad", "wgrad", C, C, KV, 2 * b)]
        if sr > 1:
            out += [(f"s{si}.sr.wgrad", "wgrad", C, sr * sr * C, KV, b)]
    return out


def make_case(kind, M, N, K):
    """returns (launch closure for sfk, launch closure for cuBLAS)"""
    if kind.startswith("fwd"):
        a = torch.randn(M, K, device=DEV).bfloat16()
        w = torch.randn(N, K, device=DEV).bfloat16()
        bias = torch.randn(N, device=DEV)
        if kind == "fwd_bf16":
            out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
            return (lambda: gemm(a, w, M, N, K, bias=bias, out_bf16=True, out=out)), (lambda: torch.matmul(a, w.t()))
        if kind == "fwd_red":   # residual already in the output buffer: scaled accumulator reduce-added in place (split-K style epilogue)
            out = torch.zeros(M, N, device=DEV, dtype=torch.float32)
            rs = torch.ones(B, device=DEV)
            return (lambda: gemm(a, w, M, N, K, bias=bias, row_scale=rs, rows_per_scale=max(1, M // B), atomic=True, splits=1, out=out)), \
                   (lambda: torch.matmul(a, w.t()))
        res = torch.randn(M, N, device=DEV) if kind == "fwd_res" else None
        rs = torch.ones(B, device=DEV) if kind == "fwd_res" else None
        out = torch.empty(M, N, device=DEV, dtype=torch.float32)
        return (lambda: gemm(a, w, M, N, K, bias=bias, residual=res, row_scale=rs, rows_per_scale=max(1, M // B), out=out)), \
               (lambda: torch.matmul(a, w.t()))
    if kind.startswith("dgrad"):
        dy = torch.randn(M, K, device=DEV).bfloat16()
        w = torch.randn(K, N, device=DEV).bfloat16()   # [n_out = K, k_in = N]
        if kind == "dgrad":
            out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
            return (lambda: gemm(dy, w, M, N, K, b_mn=1, out_bf16=True, out=out)), (lambda: torch.matmul(dy, w))
        if kind == "dgrad_f32":
            out = torch.empty(M, N, device=DEV, dtype=torch.float32)
            return (lambda: gemm(dy, w, M, N, K, b_mn=1, out=out)), (lambda: torch.matmul(dy, w))
        res = torch.randn(M, N, device=DEV)
        out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
        return (lambda: gemm(dy, w, M, N, K, b_mn=1, residual=res, out_bf16=True, out=out)), (lambda: torch.matmul(dy, w))
    if kind == "wgrad":
        dy = torch.randn(K, M, device=DEV).bfloat16()
        x = torch.randn(K, N, device=DEV).bfloat16()
        out = torch.zeros(M, N, device=DEV)
        return (lambda: gemm(dy, x, M, N, K, a_mn=1, b_mn=1, atomic=True, out=out)), (lambda: torch.matmul(dy.t(), x))
    raise ValueError(kind)


def time_graph(fn, reps=20, replays=5):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps):
                fn()
    torch.cuda.synchronize()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(replays):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000.0 / reps)
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true", help="stage-3 shapes only")
    ap.add_argument("--json", default=None)
    ap.add_argument("--filter", default=None)
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--as-red", action="store_true", help="time the fwd_res shapes with the in-place reduce-add epilogue instead")
    args = ap.parse_args()
    L.ensure_device()
    for opt in ("gemm_occ2", "gemm_cluster2", "gemm_epi_tma", "gemm_bstat", "gemm_dbg", "gemm_epi_bufs", "gemm_krot"):
        if os.environ.get("SFK_" + opt.upper()) is not None:
            L.check(L.load().sfk_set_option(opt.encode(), int(os.environ["SFK_" + opt.upper()])))
            print(opt, "=", os.environ["SFK_" + opt.upper()], flush=True)
    torch.manual_seed(0)
    rows, tot_sfk, tot_cub, tot_ideal = [], 0.0, 0.0, 0.0
    for name, kind, M, N, K, count in step_shapes():
        if args.quick and not name.startswith("s3."):
            continue
        if args.filter and args.filter not in name:
            continue
        if args.as_red and kind == "fwd_res":
            kind = "fwd_red"
        f_sfk, f_cub = make_case(kind, M, N, K)
        us = time_graph(f_sfk)
        us_c = float("nan") if args.no_cublas else time_graph(f_cub)
        fl = 2.0 * M * N * K
        ideal = fl / 1389.9e12 * 1e6
        rows.append(dict(name=name, kind=kind, M=M, N=N, K=K, count=count, sfk_us=us, cublas_us=us_c, tflops=fl / us / 1e6))
        tot_sfk += us * count
        tot_cub += us_c * count
        tot_ideal += ideal * count
        print(f"GEMM {name:16s} {kind:9s} {M:7d}x{N:5d}x{K:6d} x{count:3d}: sfk {us:7.2f} us {fl/us/1e6:7.1f} TF/s | cublas {us_c:7.2f} us"
              f" | tensor floor {ideal:6.2f} us", flush=True)
    print(f"TOTAL per step: sfk {tot_sfk/1000:.2f} ms | cublas {tot_cub/1000:.2f} ms | tensor floor {tot_ideal/1000:.2f} ms", flush=True)
    if args.json:
        with open(args.json, "w") as f:
            json.dump(dict(rows=rows, total_sfk_ms=tot_sfk / 1000, total_cublas_ms=tot_cub / 1000, tensor_floor_ms=tot_ideal / 1000), f, indent=1)


if __name__ == "__main__":
    main()
