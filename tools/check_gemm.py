"""GPU bring-up check for sfk_gemm_bf16 (run under gpurun). Usage: python tools/check_gemm.py <group>
groups: nt | nn | tn | epi | perf
Each group runs in its own process so a trap in one layout does not poison the others."""
import ctypes
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semisupervisedobjectdetection_b200 import _lib as L


def gemm(a, b, M, N, K, a_mn=0, b_mn=0, alpha=1.0, bias=None, residual=None, row_scale=None, rows_per_scale=1,
         out_bf16=False, atomic=False, relu=False, splits=0, block_n=0, out=None):
    lib = L.load()
    L.ensure_device()
    if out is None:
        out = torch.zeros(M, N, device="cuda", dtype=torch.bfloat16 if out_bf16 else torch.float32)
    rc = lib.sfk_gemm_bf16(L.ptr(a), L.c_ll(a.stride(0)), a_mn, L.ptr(b), L.c_ll(b.stride(0)), b_mn, M, N, K,
                           L.c_float(alpha), L.ptr(bias), L.ptr(residual),
                           L.c_ll(residual.stride(0) if residual is not None else 0), L.ptr(row_scale), rows_per_scale,
                           L.ptr(out), L.c_ll(out.stride(0)), int(out_bf16), int(atomic), int(relu), splits, block_n,
                           L.stream_ptr())
    L.check(rc, "sfk_gemm_bf16")
    return out


def report(name, got, ref):
    torch.cuda.synchronize()
    err = (got.float() - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-9
    ok = err / scale < 2e-2
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e}", flush=True)
    return ok


def main():
    group = sys.argv[1]
    if os.environ.get("SFK_GEMM_CLUSTER2") is not None:
        L.check(L.load().sfk_set_option(b"gemm_cluster2", 2 * int(os.environ["SFK_GEMM_CLUSTER2"])))  # 1 -> "whenever possible"
        print("gemm_cluster2 =", os.environ["SFK_GEMM_CLUSTER2"], flush=True)
    for opt in ("gemm_occ2", "gemm_epi_tma", "gemm_bstat", "gemm_dbg", "gemm_epi_bufs", "gemm_krot"):
        if os.environ.get("SFK_" + opt.upper()) is not None:
            L.check(L.load().sfk_set_option(opt.encode(), int(os.environ["SFK_" + opt.upper()])))
            print(opt, "=", os.environ["SFK_" + opt.upper()], flush=True)
    torch.manual_seed(0)
    dev = "cuda"
    allok = True
    if group == "nt":
        for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 128, 0), (256, 256, 512, 0), (300, 320, 152, 0),
                              (1024, 1280, 320, 0), (4096, 64, 64, 0), (512, 768, 3072, 0), (130, 40, 72, 0)]:
            a = torch.randn(M, K, device=dev).bfloat16()
            b = torch.randn(N, K, device=dev).bfloat16()
            ref = a.float() @ b.float().t()
            out = gemm(a, b, M, N, K, block_n=bn)
            allok &= report(f"NT M{M} N{N} K{K} bn{bn}", out, ref)
            outb = gemm(a, b, M, N, K, block_n=bn, out_bf16=True)
            allok &= report(f"NT-bf16out M{M} N{N} K{K}", outb, ref)
    elif group == "nn":
        # dgrad: dX[M,K'] = dY[M,N'] @ W[N',K'] ; B stored [k=N', n=K'] n-contiguous -> b_mn
        for (M, Nout, Kc) in [(128, 64, 64), (256, 128, 256), (1000, 320, 1280), (512, 3072, 768), (384, 192, 72)]:
            dy = torch.randn(M, Kc, device=dev).bfloat16()
            w = torch.randn(Kc, Nout, device=dev).bfloat16()
            ref = dy.float() @ w.float()
            out = gemm(dy, w, M, Nout, Kc, b_mn=1)
            allok &= report(f"NN M{M} N{Nout} K{Kc}", out, ref)
    elif group == "tn":
        # wgrad: dW[Nout,Kin] = dY[T,Nout]^T @ X[T,Kin]
        for (T, Nout, Kin, splits) in [(128, 128, 64, 1), (512, 128, 128, 1), (4096, 64, 64, 0), (2048, 320, 1280, 0),
                                       (1000, 256, 320, 0), (16384, 768, 3072, 0)]:
            dy = torch.randn(T, Nout, device=dev).bfloat16()
            x = torch.randn(T, Kin, device=dev).bfloat16()
            ref = dy.float().t() @ x.float()
            out = torch.zeros(Nout, Kin, device=dev)
            gemm(dy, x, Nout, Kin, T, a_mn=1, b_mn=1, atomic=True, splits=splits, out=out)
            allok &= report(f"TN T{T} N{Nout} K{Kin} splits{splits}", out, ref)
    elif group == "epi":
        M, N, K = 512, 320, 320
        a = torch.randn(M, K, device=dev).bfloat16()
        b = torch.randn(N, K, device=dev).bfloat16()
        bias = torch.randn(N, device=dev)
        res = torch.randn(M, N, device=dev)
        rs = torch.rand(4, device=dev)
        ref = (a.float() @ b.float().t() + bias) * 0.5 * rs.repeat_interleave(128)[:, None] + res
        out = gemm(a, b, M, N, K, alpha=0.5, bias=bias, residual=res, row_scale=rs, rows_per_scale=128)
        allok &= report("epilogue bias+scale+residual", out, ref)
        out = gemm(a, b, M, N, K, bias=bias, relu=True, out_bf16=True)
        allok &= report("epilogue bias+relu bf16", out, torch.relu(a.float() @ b.float().t() + bias))
    elif group == "v2":
        # TMA-store epilogue + B-stationary main loop: every epilogue combination on full, ragged and multi-column-tile shapes
        shapes = [(512, 320, 320, 0), (16384, 1280, 320, 0), (16384, 320, 320, 0), (1000, 200, 152, 0), (4096, 640, 320, 0),
                  (2048, 64, 64, 0), (2048, 256, 64, 0), (130, 40, 72, 0), (4096, 320, 1280, 0), (4096, 2048, 512, 0),
                  (16384, 1280, 320, 128), (16384, 1280, 320, 224), (3000, 96, 512, 32)]
        for (M, N, K, bn) in shapes:
            a = torch.randn(M, K, device=dev).bfloat16()
            b = torch.randn(N, K, device=dev).bfloat16()
            bias = torch.randn(N, device=dev)
            res = torch.randn(M, N, device=dev)
            rs = torch.rand((M + 255) // 256, device=dev) + 0.5
            base = a.float() @ b.float().t()
            rsv = rs.repeat_interleave(256)[:M, None]
            out = gemm(a, b, M, N, K, bias=bias, out_bf16=True, block_n=bn)
            allok &= report(f"v2 bf16+bias M{M} N{N} K{K} bn{bn}", out, base + bias)
            out = gemm(a, b, M, N, K, bias=bias, residual=res, row_scale=rs, rows_per_scale=256, block_n=bn)
            allok &= report(f"v2 f32+bias+scale+res M{M} N{N} K{K} bn{bn}", out, (base + bias) * rsv + res)
            out = gemm(a, b, M, N, K, residual=res, out_bf16=True, relu=True, block_n=bn)
            allok &= report(f"v2 bf16+res+relu M{M} N{N} K{K} bn{bn}", out, torch.relu(base + res))
            out = gemm(a, b, M, N, K, alpha=0.25, block_n=bn)
            allok &= report(f"v2 f32 alpha M{M} N{N} K{K} bn{bn}", out, 0.25 * base)
            # dgrad layout (MN-major B) with residual, bf16 out
            if N % 8 == 0:
                w = torch.randn(K, N, device=dev).bfloat16()
                out = gemm(a, w, M, N, K, b_mn=1, residual=res, out_bf16=True)
                allok &= report(f"v2 dgrad+res M{M} N{N} K{K}", out, a.float() @ w.float() + res)
        # split-K reduce-add epilogue accumulates into existing contents
        T, No, Ki = 16384, 320, 1280
        dy = torch.randn(T, No, device=dev).bfloat16()
        x = torch.randn(T, Ki, device=dev).bfloat16()
        out = torch.ones(No, Ki, device=dev)
        gemm(dy, x, No, Ki, T, a_mn=1, b_mn=1, atomic=True, out=out)
        gemm(dy, x, No, Ki, T, a_mn=1, b_mn=1, atomic=True, out=out)
        allok &= report("v2 wgrad reduce-add twice", out, 2 * (dy.float().t() @ x.float()) + 1)
    elif group == "trace":
        lib = L.load()
        buf = torch.zeros(256, device=dev, dtype=torch.int64)
        cases = [("q_s3 16384x320x320", 16384, 320, 320, 0), ("ffn1_s3 16384x1280x320", 16384, 1280, 320, 0),
                 ("ffn1_s3 bn256", 16384, 1280, 320, 256), ("ffn2_s3 16384x320x1280", 16384, 320, 1280, 0),
                 ("ffn1_s1 262144x256x64", 262144, 256, 64, 0), ("q_s1 262144x64x64", 262144, 64, 64, 0)]
        for name, M, N, K, bn in cases:
            a = torch.randn(M, K, device=dev).bfloat16()
            b = torch.randn(N, K, device=dev).bfloat16()
            bias = torch.randn(N, device=dev)
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                gemm(a, b, M, N, K, out_bf16=True, out=out, bias=bias, block_n=bn)
            torch.cuda.synchronize()
            buf.zero_()
            lib.sfk_debug_gemm_trace(buf.data_ptr())
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            gemm(a, b, M, N, K, out_bf16=True, out=out, bias=bias, block_n=bn)
            e1.record()
            torch.cuda.synchronize()
            lib.sfk_debug_gemm_trace(None)
            t = buf.cpu().tolist()
            rel = [(x - t[0]) / 1965.0 if x else float("nan") for x in t]   # SM cycles at the 1965 MHz boost clock -> us
            print(f"TRACE {name}: event {e0.elapsed_time(e1)*1000:.1f} us | us since CTA0 start: setup {rel[1]:.2f} tma0 {rel[2]:.2f} "
                  f"landed {rel[3]:.2f} last-mma {rel[4]:.2f} acc-ready {rel[5]:.2f} tile-stored {rel[6]:.2f} cta-done {rel[7]:.2f}", flush=True)
            tiles = [(rel[8 + 2 * i], rel[9 + 2 * i]) for i in range(28) if t[8 + 2 * i]]
            print("      per tile (acc-ready, stored): " + " ".join(f"({a0:.2f},{a1:.2f})" for a0, a1 in tiles), flush=True)
            for label, base in (("producer starts tile", 64), ("accumulator free", 96), ("last k-block issued", 160)):
                print(f"      {label:22s}: " + " ".join(f"{rel[base + i]:.2f}" for i in range(24) if t[base + i]), flush=True)
    elif group == "perf":
        shapes = [("q_s3 16384x320x320", 16384, 320, 320), ("kv_s3 4096x640x320", 4096, 640, 320),
                  ("fuse 262144x768x3072", 262144, 768, 3072), ("ffn1_s3 16384x1280x320", 16384, 1280, 320),
                  ("ffn2_s3 16384x320x1280", 16384, 320, 1280), ("ffn1_s1 262144x256x64", 262144, 256, 64),
                  ("q_s1 262144x64x64", 262144, 64, 64), ("ffn1_s4 4096x2048x512", 4096, 2048, 512),
                  ("square 8192", 8192, 8192, 8192)]
        for name, M, N, K in shapes:
            a = torch.randn(M, K, device=dev).bfloat16()
            b = torch.randn(N, K, device=dev).bfloat16()
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            for _ in range(3):
                gemm(a, b, M, N, K, out_bf16=True, out=out)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 10
            e0.record()
            for _ in range(iters):
                gemm(a, b, M, N, K, out_bf16=True, out=out)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = 2.0 * M * N * K / ms / 1e9
            # cuBLAS for context only
            for _ in range(3):
                torch.matmul(a, b.t())
            e0.record()
            for _ in range(iters):
                torch.matmul(a, b.t())
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1) / iters
            print(f"PERF {name}: sfk {ms:.4f} ms {tf:.1f} TF/s | cublas {ms2:.4f} ms {2.0*M*N*K/ms2/1e9:.1f} TF/s", flush=True)
    print("GROUP", group, "ALLOK" if allok else "FAILED", flush=True)
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
