"""GPU bring-up check of the non-GEMM ops against plain PyTorch fp32. Usage: python tools/check_ops.py <group>
groups: ln | dwconv | im2col | attn_fwd | attn_bwd | attn_perf"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from semisupervisedobjectdetection_b200 import _lib as L

dev = "cuda"
lib = None


def report(name, got, ref, tol=2e-2):
    torch.cuda.synchronize()
    got = got.float()
    err = (got - ref).abs().max().item()
    scale = ref.abs().max().item() + 1e-12
    ok = (err / scale < tol) and bool(torch.isfinite(got).all())
    print(f"{'OK ' if ok else 'BAD'} {name}: max_abs_err={err:.4e} ref_max={scale:.3e} rel={err/scale:.3e}", flush=True)
    return ok


def attn_ref(q, k, v, heads, scale):
    B, T, C = q.shape
    d = C // heads
    qh = q.float().reshape(B, T, heads, d).transpose(1, 2)
    kh = k.float().reshape(B, -1, heads, d).transpose(1, 2)
    vh = v.float().reshape(B, -1, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    p = s.softmax(-1)
    return (p @ vh).transpose(1, 2).reshape(B, T, C)


def run_attn_fwd(q, kv, heads, D, scale):
    B, T, C = q.shape
    Tk = kv.shape[1]
    o = torch.empty(B, T, C, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, heads, T, device=dev, dtype=torch.float32)
    rc = lib.sfk_attention_fwd(q.data_ptr(), C, T * C, kv.data_ptr(), kv.data_ptr() + 2 * C, 2 * C, Tk * 2 * C, o.data_ptr(), C,
                               T * C, lse.data_ptr(), B, heads, T, Tk, D, scale, L.stream_ptr())
    L.check(rc, "attention_fwd")
    return o, lse


def main():
    global lib
    lib = L.load()
    L.ensure_device()
    group = sys.argv[1]
    torch.manual_seed(0)
    allok = True
    if group == "ln":
        for rows, C in [(1000, 64), (4096, 320), (513, 512), (77, 32), (300, 160), (64, 1024)]:
            x = torch.randn(rows, C, device=dev) * 2 + 0.5
            g = torch.randn(C, device=dev)
            b = torch.randn(C, device=dev)
            ref = F.layer_norm(x, (C,), g, b, 1e-5)
            for ybf in (0, 1):
                y = torch.empty(rows, C, device=dev, dtype=torch.bfloat16 if ybf else torch.float32)
                mean = torch.empty(rows, device=dev)
                rstd = torch.empty(rows, device=dev)
                L.check(lib.sfk_layernorm_fwd(x.data_ptr(), 0, 0, 0, 0, g.data_ptr(), b.data_ptr(), y.data_ptr(), ybf,
                                              mean.data_ptr(), rstd.data_ptr(), rows, C, 1e-5, L.stream_ptr()))
                allok &= report(f"ln_fwd rows{rows} C{C} bf16out{ybf}", y, ref, 1e-2 if ybf else 1e-5)
            # backward
            xr = x.clone().requires_grad_(True)
            gr = g.clone().requires_grad_(True)
            br = b.clone().requires_grad_(True)
            dy = torch.randn(rows, C, device=dev)
            dres = torch.randn(rows, C, device=dev)
            F.layer_norm(xr, (C,), gr, br, 1e-5).backward(dy)
            dx = torch.empty(rows, C, device=dev)
            dg = torch.zeros(C, device=dev)
            db = torch.zeros(C, device=dev)
            L.check(lib.sfk_layernorm_bwd(dy.data_ptr(), 0, x.data_ptr(), 0, mean.data_ptr(), rstd.data_ptr(), g.data_ptr(),
                                          dres.data_ptr(), dx.data_ptr(), 0, dg.data_ptr(), db.data_ptr(), rows, C,
                                          L.stream_ptr()))
            allok &= report(f"ln_bwd dx rows{rows} C{C}", dx, xr.grad + dres, 1e-4)
            allok &= report(f"ln_bwd dgamma", dg, gr.grad, 1e-4)
            allok &= report(f"ln_bwd dbeta", db, br.grad, 1e-4)
    elif group == "dwconv":
        for B, H, W, C in [(2, 16, 16, 128), (1, 7, 5, 64), (2, 32, 32, 256), (3, 2, 2, 1024), (2, 9, 13, 640)]:
            x = torch.randn(B, H, W, C, device=dev).bfloat16()
            w = torch.randn(C, 1, 3, 3, device=dev) * 0.3
            bias = torch.randn(C, device=dev) * 0.1
            wt = torch.empty(9, C, device=dev)
            L.check(lib.sfk_dw_weight_to_tap(w.data_ptr(), wt.data_ptr(), C, L.stream_ptr()))
            y = torch.empty_like(x)
            L.check(lib.sfk_dwconv3x3(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), y.data_ptr(), B, H, W, C, 0, L.stream_ptr()))
            xr = x.float().permute(0, 3, 1, 2).clone().requires_grad_(True)
            wr = w.clone().requires_grad_(True)
            br = bias.clone().requires_grad_(True)
            pre = F.conv2d(xr, wr, br, padding=1, groups=C)
            ref = F.gelu(pre)
            allok &= report(f"dwconv_gelu fwd B{B} H{H} W{W} C{C}", y, ref.permute(0, 2, 3, 1), 1e-2)
            dg = torch.randn(B, H, W, C, device=dev).bfloat16()
            ref.backward(dg.float().permute(0, 3, 1, 2))
            dpre = torch.empty_like(x)
            dwt = torch.zeros(9, C, device=dev)
            dbias = torch.zeros(C, device=dev)
            L.check(lib.sfk_dwconv3x3_gelu_bwd(x.data_ptr(), wt.data_ptr(), bias.data_ptr(), dg.data_ptr(), dpre.data_ptr(),
                                               dwt.data_ptr(), dbias.data_ptr(), B, H, W, C, L.stream_ptr()))
            dx = torch.empty_like(x)
            L.check(lib.sfk_dwconv3x3(dpre.data_ptr(), wt.data_ptr(), None, dx.data_ptr(), B, H, W, C, 1, L.stream_ptr()))
            dw = torch.zeros(C, 1, 3, 3, device=dev)
            L.check(lib.sfk_dw_wgrad_from_tap(dwt.data_ptr(), dw.data_ptr(), C, L.stream_ptr()))
            allok &= report("dwconv bwd dx", dx, xr.grad.permute(0, 2, 3, 1), 2e-2)
            allok &= report("dwconv bwd dw", dw, wr.grad, 2e-2)
            allok &= report("dwconv bwd db", dbias, br.grad, 2e-2)
    elif group == "im2col":
        for (B, H, W, C, K, S, P) in [(2, 16, 16, 64, 3, 2, 1), (2, 8, 8, 32, 8, 8, 0), (1, 12, 20, 128, 2, 2, 0), (2, 4, 4, 160, 3, 2, 1)]:
            x = torch.randn(B, H, W, C, device=dev).bfloat16()
            Ho, Wo = (H + 2 * P - K) // S + 1, (W + 2 * P - K) // S + 1
            A = torch.empty(B * Ho * Wo, K * K * C, device=dev, dtype=torch.bfloat16)
            L.check(lib.sfk_im2col_nhwc(x.data_ptr(), A.data_ptr(), B, H, W, C, K, S, P, H * W, 0, L.stream_ptr()))
            w = torch.randn(48, C, K, K, device=dev)
            ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, stride=S, padding=P)  # [B,48,Ho,Wo]
            wk = w.permute(0, 2, 3, 1).reshape(48, -1)
            got = (A.float() @ wk.t()).reshape(B, Ho, Wo, 48).permute(0, 3, 1, 2)
            allok &= report(f"im2col_nhwc K{K} S{S} P{P} C{C}", got, ref, 2e-3)  # cudnn conv runs TF32
            # col2im == transpose of im2col: <A(x), dA> == <x, col2im(dA)>
            dA = torch.randn_like(A)
            dx = torch.zeros(B, H, W, C, device=dev)
            L.check(lib.sfk_col2im_nhwc(dA.data_ptr(), dx.data_ptr(), B, H, W, C, K, S, P, H * W, 0, 0, L.stream_ptr()))
            xr = x.float().clone().requires_grad_(True)
            cols = F.unfold(xr.permute(0, 3, 1, 2), K, padding=P, stride=S)  # [B, C*K*K, L] with (c,ky,kx) order
            cols = cols.reshape(B, C, K * K, Ho * Wo).permute(0, 3, 2, 1).reshape(B * Ho * Wo, K * K * C)
            (cols * dA.float()).sum().backward()
            allok &= report("col2im_nhwc", dx, xr.grad, 1e-3)
        x = torch.rand(2, 3, 32, 32, device=dev)
        Kp = 152
        A = torch.empty(2 * 8 * 8, Kp, device=dev, dtype=torch.bfloat16)
        L.check(lib.sfk_im2col_nchw_f32(x.data_ptr(), A.data_ptr(), 2, 3, 32, 32, 7, 4, 3, Kp, L.stream_ptr()))
        w = torch.randn(16, 3, 7, 7, device=dev)
        ref = F.conv2d(x, w, stride=4, padding=3)
        got = (A.float()[:, :147] @ w.reshape(16, -1).t()).reshape(2, 8, 8, 16).permute(0, 3, 1, 2)
        allok &= report("im2col_nchw_f32 7x7 s4", got, ref, 1e-2)
        allok &= report("im2col_nchw_f32 pad zero", A[:, 147:], torch.zeros(128, 5, device=dev), 1.0) if False else allok
    elif group in ("attn_fwd", "attn_bwd"):
        cases = [(2, 1, 256, 128, 64), (2, 2, 300, 256, 64), (1, 5, 1024, 266, 64), (2, 8, 64, 64, 64), (2, 1, 16, 4, 32),
                 (2, 2, 67, 19, 32), (1, 8, 256, 259, 32), (1, 1, 16384, 256, 64), (2, 2, 130, 700, 64)]
        for (B, heads, T, Tk, D) in cases:
            C = heads * D
            q = (torch.randn(B, T, C, device=dev) * 1.5).bfloat16()
            kv = (torch.randn(B, Tk, 2 * C, device=dev) * 1.5).bfloat16()
            scale = 1.0 / math.sqrt(D)
            o, lse = run_attn_fwd(q, kv, heads, D, scale)
            if group == "attn_fwd":
                ref = attn_ref(q, kv[..., :C], kv[..., C:], heads, scale)
                allok &= report(f"attn_fwd B{B} h{heads} T{T} Tk{Tk} D{D}", o, ref, 2e-2)
                # lse check
                qh = q.float().reshape(B, T, heads, D).transpose(1, 2)
                kh = kv[..., :C].float().reshape(B, Tk, heads, D).transpose(1, 2)
                lse_ref = torch.logsumexp(qh @ kh.transpose(-1, -2) * scale, dim=-1) * 1.4426950408889634
                allok &= report("   lse(log2)", lse, lse_ref, 1e-3)
            else:
                qr = q.float().clone().requires_grad_(True)
                kvr = kv.float().clone().requires_grad_(True)
                ref = attn_ref(qr, kvr[..., :C], kvr[..., C:], heads, scale)
                dO = torch.randn(B, T, C, device=dev).bfloat16()
                ref.backward(dO.float())
                delta = torch.empty(B, heads, T, device=dev)
                dq = torch.zeros(B, T, C, device=dev)
                dkv = torch.zeros(B, Tk, 2 * C, device=dev)
                rc = lib.sfk_attention_bwd(q.data_ptr(), C, T * C, kv.data_ptr(), kv.data_ptr() + 2 * C, 2 * C, Tk * 2 * C,
                                           o.data_ptr(), dO.data_ptr(), C, T * C, lse.data_ptr(), delta.data_ptr(), dq.data_ptr(),
                                           C, T * C, dkv.data_ptr(), dkv.data_ptr() + 4 * C, 2 * C, Tk * 2 * C, B, heads, T, Tk,
                                           D, scale, L.stream_ptr())
                L.check(rc, "attention_bwd")
                allok &= report(f"attn_bwd dq B{B} h{heads} T{T} Tk{Tk} D{D}", dq, qr.grad, 3e-2)
                allok &= report("   dk", dkv[..., :C], kvr.grad[..., :C], 3e-2)
                allok &= report("   dv", dkv[..., C:], kvr.grad[..., C:], 3e-2)
    elif group == "attn_perf":
        for (name, B, heads, T, Tk, D) in [("stage1", 16, 1, 16384, 256, 64), ("stage2", 16, 2, 4096, 256, 64),
                                          ("stage3", 16, 5, 1024, 256, 64), ("stage4", 16, 8, 256, 256, 64)]:
            C = heads * D
            q = torch.randn(B, T, C, device=dev).bfloat16()
            kv = torch.randn(B, Tk, 2 * C, device=dev).bfloat16()
            scale = 1.0 / math.sqrt(D)
            for _ in range(3):
                o, lse = run_attn_fwd(q, kv, heads, D, scale)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                o, lse = run_attn_fwd(q, kv, heads, D, scale)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            fl = 4.0 * B * heads * T * Tk * D
            dO = torch.randn(B, T, C, device=dev).bfloat16()
            delta = torch.empty(B, heads, T, device=dev)
            dq = torch.zeros(B, T, C, device=dev)
            dkv = torch.zeros(B, Tk, 2 * C, device=dev)

            def bwd():
                L.check(lib.sfk_attention_bwd(q.data_ptr(), C, T * C, kv.data_ptr(), kv.data_ptr() + 2 * C, 2 * C, Tk * 2 * C,
                                              o.data_ptr(), dO.data_ptr(), C, T * C, lse.data_ptr(), delta.data_ptr(),
                                              dq.data_ptr(), C, T * C, dkv.data_ptr(), dkv.data_ptr() + 4 * C, 2 * C, Tk * 2 * C,
                                              B, heads, T, Tk, D, scale, L.stream_ptr()))
            for _ in range(3):
                bwd()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                bwd()
            e1.record()
            torch.cuda.synchronize()
            msb = e0.elapsed_time(e1) / 10
            print(f"PERF attn {name}: fwd {ms:.4f} ms {fl/ms/1e9:.1f} TF/s | bwd {msb:.4f} ms {2.5*fl/msb/1e9:.1f} TF/s", flush=True)
    print("GROUP", group, "ALLOK" if allok else "FAILED", flush=True)
    sys.exit(0 if allok else 1)


if __name__ == "__main__":
    main()
