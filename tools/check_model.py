"""GPU bring-up: whole-model parity against the golden fixtures (reference outputs) + quick b5 timing.
Usage: python tools/check_model.py golden | b5 [batch]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from semisupervisedobjectdetection_b200 import ops
from semisupervisedobjectdetection_b200.configuration import SegformerConfig
from semisupervisedobjectdetection_b200.loss import SegmentationLoss
from semisupervisedobjectdetection_b200.modeling import SegformerForSemanticSegmentation, upsample_bilinear
from tests import helpers
from tests.golden import cases


def rel(got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    return float(((got - ref).abs().max() / (ref.abs().max() + 1e-12)).detach())


def run_golden():
    ok_all = True
    for name in cases.MODEL_CASES:
        gold = helpers.load_golden(name)
        case, ocfg, sd, x, prompts, cls, labels, mask = helpers.case_inputs(name)
        cfg = SegformerConfig(**case["config"])
        model = SegformerForSemanticSegmentation(cfg)
        model.load_state_dict(sd, strict=True)
        model.cuda()
        train = case.get("train_mode", False)
        model.train() if train else model.eval()
        enc = model.segformer.encoder
        pts = [None] * 4
        cts = [None] * 4
        for s in range(4):
            if prompts[s] is not None:
                pts[s] = prompts[s].clone().cuda().requires_grad_(True)
                enc.prompt_tokens[s] = pts[s]
            if cls[s] is not None:
                cts[s] = cls[s].clone().cuda().requires_grad_(True)
                enc.cls_token[s] = cts[s]
        xg = x.cuda()
        lab = labels.cuda() if labels is not None else None
        out, cls_token = model(pixel_values=xg, labels=lab, output_hidden_states=True)
        res = {"logits": rel(out.logits, gold["logits"])}
        for i, (h, g) in enumerate(zip(out.hidden_states, gold["hidden_states"])):
            res[f"hs{i}"] = rel(h, g)
        hw = case["hw"]
        probs = torch.squeeze(upsample_bilinear(out.logits, (hw, hw), sigmoid=True), 1)
        res["probs"] = rel(probs, gold["probs"])
        loss = None
        if lab is not None:
            res["hf_loss"] = abs(float(out.loss) - float(gold["hf_loss"])) / abs(float(gold["hf_loss"]))
            loss = out.loss
        if cfg.num_labels == 1:
            dice = SegmentationLoss(1, loss_type="dice", activation="none")(probs, mask.cuda())
            res["dice"] = abs(float(dice) - float(gold["dice_loss"])) / abs(float(gold["dice_loss"]))
            loss = dice if loss is None else dice + loss
        if "cls_token" in gold:
            res["cls_token"] = rel(cls_token, gold["cls_token"])
        if train:
            res["bn_rm"] = rel(model.decode_head.batch_norm.running_mean, gold["bn_running_mean"])
            res["bn_rv"] = rel(model.decode_head.batch_norm.running_var, gold["bn_running_var"])
        model.zero_grad()
        loss.backward()
        torch.cuda.synchronize()
        named = dict(model.named_parameters())
        worst_g, worst_k = 0.0, None
        for k, want in gold["grads"].items():
            g = named[k].grad
            if g is None:
                res[f"MISSING grad {k}"] = 1.0
                continue
            r = rel(cases.sample_flat(g.cpu()), want)
            if r > worst_g:
                worst_g, worst_k = r, k
        res["grad_worst"] = worst_g
        worst_n, worst_nk = 0.0, None
        for k, want in gold["grad_norms"].items():
            g = named[k].grad
            if g is None:
                worst_n, worst_nk = 9.9, k + " (missing)"
                continue
            r = abs(float(g.norm()) - want) / (want + 1e-12)
            if r > worst_n and want > 1e-7:
                worst_n, worst_nk = r, k
        res["gradnorm_worst"] = worst_n
        for s, want in gold.get("prompt_grads", {}).items():
            res[f"prompt_grad{s}"] = rel(pts[s].grad, want)
        for s, want in gold.get("cls_grads", {}).items():
            res[f"cls_grad{s}"] = rel(cts[s].grad, want)
        # bf16 tolerance: 5e-2 of the tensor max; train-mode BatchNorm backward (batch statistics recomputed from the
        # bf16 fuse output, batch of 3 tiny tiles) amplifies rounding on a few near-cancelling LayerNorm-weight gradients
        tol = {k: 5e-2 for k in res}
        if train:
            tol["grad_worst"] = 2e-1
            tol["gradnorm_worst"] = 1e-1
        bad = {k: v for k, v in res.items() if v > tol[k]}
        print(f"{'OK ' if not bad else 'BAD'} {name}: " + " ".join(f"{k}={v:.2e}" for k, v in res.items()), flush=True)
        print(f"      worst grad {worst_k} | worst norm {worst_nk}", flush=True)
        ok_all &= not bad
    print("GOLDEN", "ALLOK" if ok_all else "FAILED", flush=True)
    return ok_all


def run_b5(batch):
    torch.manual_seed(0)
    cfg = SegformerConfig.from_variant("b5", num_labels=1)
    model = SegformerForSemanticSegmentation(cfg).cuda().train()
    x = torch.rand(batch, 3, 512, 512, device="cuda")
    mask = (torch.rand(batch, 512, 512, device="cuda") > 0.9).to(torch.int64)
    lossf = SegmentationLoss(1, loss_type="dice", activation="none")

    def step():
        model.zero_grad(set_to_none=False)
        out, _ = model(pixel_values=x)
        probs = torch.squeeze(upsample_bilinear(out.logits, (512, 512), sigmoid=True), 1)
        loss = lossf(probs, mask)
        loss.backward()
        return loss

    for _ in range(2):
        loss = step()
    torch.cuda.synchronize()
    print("b5 loss", float(loss), "max mem GB", torch.cuda.max_memory_allocated() / 2 ** 30, flush=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record()
    n = 5
    l0 = ops.LAUNCHES
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"b5 batch {batch} fwd+bwd: {ms:.2f} ms/step (wall {1000*(time.time()-t0)/n:.2f} ms) -> {batch/ms*1000:.1f} img/s, "
          f"{657.8*batch/ms:.1f} TFLOP/s, {(ops.LAUNCHES-l0)/n:.0f} libsfk calls/step", flush=True)
    # forward only
    with torch.no_grad():
        for _ in range(2):
            model(pixel_values=x)
        e0.record()
        for _ in range(n):
            model(pixel_values=x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"b5 batch {batch} fwd only: {ms:.2f} ms -> {batch/ms*1000:.1f} img/s, {219.27*batch/ms:.1f} TFLOP/s", flush=True)


def run_patch_parity():
    """Every fused path against the plainer kernel sequence it replaced, on the same inputs (MiT-b0 at 512x512, where every
    stage tiles: Wo = 16), one switch flipped at a time:
      sr_tma_patch   SR conv operands through the rank-5 TMA patch view vs the gathered (im2col) copy
      sr_patch_store SR conv input gradient scattered by the dgrad GEMM's TMA stores (+ bf16 residual in the Q dgrad) vs col2im
      layer_link     LayerNorm-1 backward emitting the preceding layer's bf16 MixFFN gradient + dense2 bias gradient vs its own pass
      attn_fwd256    single-pass attention forward (Tk <= 256) vs the chunked online-softmax kernel
      fwd_kv_overlap K/V branch of the forward pass on an auxiliary stream next to the Q projection vs one stream
    Backward-only switches must leave the logits unchanged and the gradients equal to bf16-rounding noise."""
    from semisupervisedobjectdetection_b200 import _lib as L
    from semisupervisedobjectdetection_b200 import modeling as M
    lib = L.load()
    torch.manual_seed(0)
    cfg = SegformerConfig.from_variant("b0", num_labels=1)
    model = SegformerForSemanticSegmentation(cfg).cuda().eval()
    x = torch.rand(2, 3, 512, 512, device="cuda")
    mask = (torch.rand(2, 512, 512, device="cuda") > 0.9).to(torch.int64)
    lossf = SegmentationLoss(1, loss_type="dice", activation="none")

    def run():
        model.zero_grad(set_to_none=False)
        out, _ = model(pixel_values=x)
        probs = torch.squeeze(upsample_bilinear(out.logits, (512, 512), sigmoid=True), 1)
        lossf(probs, mask).backward()
        torch.cuda.synchronize()
        return out.logits.detach().clone(), {n: p.grad.detach().clone() for n, p in model.named_parameters() if p.grad is not None}

    base = run()
    ok = True
    for name, tol_logits, tol_grad in [("sr_tma_patch", 1e-5, 1e-3), ("sr_patch_store", 1e-6, 4e-3), ("layer_link", 1e-6, 1e-3),
                                       ("fwd_kv_overlap", 0.0, 1e-3),   # same kernels on two streams: identical logits

                                       # two kernels that round the softmax numerators to bf16 against different maxima
                                       # (per-chunk running maximum vs whole-row maximum): bf16-level differences
                                       ("attn_fwd256", 1e-2, 2e-2)]:
        if name == "layer_link":
            M.LAYER_LINK = False
        else:
            L.check(lib.sfk_set_option(name.encode(), 0))
        alt = run()
        if name == "layer_link":
            M.LAYER_LINK = True
        else:
            L.check(lib.sfk_set_option(name.encode(), 1))
        dl = (base[0] - alt[0]).abs().max().item() / (alt[0].abs().max().item() + 1e-12)
        print(f"{'OK ' if dl <= tol_logits else 'BAD'} {name}: logits fused vs plain: rel {dl:.2e}", flush=True)
        ok &= dl <= tol_logits
        worst, wn = 0.0, ""
        for n, g0 in alt[1].items():
            if ".key." in n:
                continue   # softmax is invariant to a key shift: these gradients are mostly summation-order noise around zero
            e = (base[1][n] - g0).abs().max().item() / (g0.abs().max().item() + 1e-12)
            if e > worst:
                worst, wn = e, n
        print(f"{'OK ' if worst < tol_grad else 'BAD'} {name}: gradients fused vs plain: worst rel {worst:.2e} ({wn})", flush=True)
        ok &= worst < tol_grad
    print("PATCH", "ALLOK" if ok else "FAILED", flush=True)
    return ok


if __name__ == "__main__":
    if sys.argv[1] == "patch":
        sys.exit(0 if run_patch_parity() else 1)
    if sys.argv[1] == "golden":
        sys.exit(0 if run_golden() else 1)
    run_b5(int(sys.argv[2]) if len(sys.argv) > 2 else 2)
