"""Generate the golden fixtures in tests/golden/*.pt by executing the UNMODIFIED reference from /root/reference.

Run in the build container only (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py

Inputs and weights come from oracle/synth.py (integer-hash generator), so the fixtures only store OUTPUTS of the
reference; the tests rebuild identical inputs/weights on any machine.
"""
from __future__ import annotations

import ast
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import synth  # noqa: E402
from tests.golden import cases  # noqa: E402


def import_reference():
    sys.path.insert(0, REF)
    stub = types.ModuleType("segformer_source.dependency_versions_check")
    stub.dep_version_check = lambda *a, **k: None
    sys.modules["segformer_source.dependency_versions_check"] = stub
    from segformer_source.models.segformer import SegformerConfig, SegformerForSemanticSegmentation  # type: ignore
    import models.Loss as ref_loss  # type: ignore
    return SegformerConfig, SegformerForSemanticSegmentation, ref_loss


def load_threshold_fn(ref_loss):
    """Extract threshold_pseudo_masks (TS:24-56) from the script without running its __main__ / visdom imports.
    The function hard-codes `.cuda()` (TS:41); on this GPU-less box Tensor.cuda is patched to identity for the call."""
    path = os.path.join(REF, "main_segformer", "segFormer_semi_teacherstudent_main.py")
    tree = ast.parse(open(path).read())
    wanted = [n for n in tree.body if (isinstance(n, ast.FunctionDef) and n.name == "threshold_pseudo_masks")
              or (isinstance(n, ast.Assign) and any(getattr(t, "id", "") in ("PESUDO_MAKS_THRESHOLD", "CONFIDENT_THRESHOLD")
                                                    for t in n.targets))]
    mod = ast.Module(body=wanted, type_ignores=[])
    ns = {"torch": torch, "loss_func": ref_loss.SegmentationLoss(1, loss_type="dice", activation="none")}
    exec(compile(mod, path, "exec"), ns)
    fn = ns["threshold_pseudo_masks"]

    def call(*a, **k):
        orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *aa, **kk: self.clone()  # non-leaf copy, like a real H2D copy
        try:
            return fn(*a, **k)
        finally:
            torch.Tensor.cuda = orig

    return call


def build_model(SegformerConfig, Model, case):
    cfg = SegformerConfig(**case["config"])
    model = Model(cfg)
    template = {k: v for k, v in model.state_dict().items()}
    sd = synth.synth_state_dict(template)
    model.load_state_dict(sd, strict=True)
    return model, cfg


def run_case(name, case, SegformerConfig, Model, ref_loss):
    model, cfg = build_model(SegformerConfig, Model, case)
    train_bn = case.get("train_mode", False)
    model.train() if train_bn else model.eval()
    b, hw = case["batch"], case["hw"]
    x = synth.uniform01(f"{name}/x", (b, 3, hw, hw))
    enc = model.segformer.encoder
    for s, spec in enumerate(case.get("prompts", [None] * 4)):
        if spec is not None:
            enc.prompt_tokens[s] = synth.uniform01(f"{name}/prompt{s}", spec).requires_grad_(True)
    for s, n in enumerate(case.get("cls", [0] * 4)):
        if n:
            enc.cls_token[s] = synth.uniform01(f"{name}/cls{s}", (n, cfg.hidden_sizes[s])).requires_grad_(True)
    out = {}
    labels = None
    if case.get("labels") == "binary":
        labels = (synth.uniform01(f"{name}/labels", (b, hw, hw)) > 0.6).to(torch.int64)
        ign = synth.uniform01(f"{name}/ignore", (b, hw, hw)) > 0.95
        labels = torch.where(ign, torch.full_like(labels, 255), labels)
    elif case.get("labels") == "multi":
        labels = (synth.uniform01(f"{name}/labels", (b, hw, hw)) * cfg.num_labels).to(torch.int64).clamp_(max=cfg.num_labels - 1)
        ign = synth.uniform01(f"{name}/ignore", (b, hw, hw)) > 0.95
        labels = torch.where(ign, torch.full_like(labels, 255), labels)
    outputs, cls_token = model(pixel_values=x, labels=labels, output_hidden_states=True)
    out["logits"] = outputs.logits.detach().clone()
    out["hidden_states"] = [h.detach().clone() for h in outputs.hidden_states]
    if labels is not None:
        out["hf_loss"] = outputs.loss.detach().clone()
    if cls_token is not None:
        out["cls_token"] = cls_token.detach().clone()  # sigmoid(stage-4 cls row), MS:848-850
    if train_bn:
        out["bn_running_mean"] = model.decode_head.batch_norm.running_mean.detach().clone()
        out["bn_running_var"] = model.decode_head.batch_norm.running_var.detach().clone()
    # wrapper tail (SFM:117-129): upsample, sigmoid, squeeze, dice against an int64 mask
    up = torch.nn.functional.interpolate(outputs.logits, size=(hw, hw), mode="bilinear", align_corners=False)
    probs = torch.squeeze(torch.sigmoid(up), 1)
    out["probs"] = probs.detach().clone()
    if cfg.num_labels == 1:
        mask = (synth.uniform01(f"{name}/mask", (b, hw, hw)) > 0.7).to(torch.int64)
        dice = ref_loss.SegmentationLoss(1, loss_type="dice", activation="none")(probs, mask)
        out["dice_loss"] = dice.detach().clone()
        loss = dice if labels is None else dice + outputs.loss
    else:
        loss = outputs.loss
    model.zero_grad()
    loss.backward()
    grads = {}
    for k, p in model.named_parameters():
        if p.grad is not None and any(tag in k for tag in cases.GRAD_KEYS):
            grads[k] = cases.sample_flat(p.grad.detach())
    out["grads"] = grads
    out["grad_norms"] = {k: float(p.grad.norm()) for k, p in model.named_parameters() if p.grad is not None}
    for s in range(4):
        if enc.prompt_tokens[s] is not None and enc.prompt_tokens[s].grad is not None:
            out.setdefault("prompt_grads", {})[s] = enc.prompt_tokens[s].grad.detach().clone()
        if enc.cls_token[s] is not None and enc.cls_token[s].grad is not None:
            out.setdefault("cls_grads", {})[s] = enc.cls_token[s].grad.detach().clone()
    return out


def loss_and_threshold_goldens(ref_loss):
    out = {}
    p = synth.uniform01("loss/p", (3, 64, 64))
    g = (synth.uniform01("loss/g", (3, 64, 64)) > 0.5)
    for kind in ("dice", "dice_argmax"):
        f = ref_loss.SegmentationLoss(1, loss_type=kind, activation="none")
        out[f"{kind}/int64"] = f(p, g.to(torch.int64)).detach().clone()
        out[f"{kind}/float"] = f(p, g.float()).detach().clone()
    img = synth.uniform01("loss/img", (3, 3, 64, 64))
    out["mse"] = ref_loss.SegmentationLoss(1, loss_type="mse")(img, synth.uniform01("loss/rec", (3, 3, 64, 64))).detach().clone()
    # dice gradient w.r.t. pred
    pr = p.clone().requires_grad_(True)
    ref_loss.SegmentationLoss(1, loss_type="dice", activation="none")(pr, g.to(torch.int64)).backward()
    out["dice/grad"] = pr.grad.detach().clone()

    thr = load_threshold_fn(ref_loss)
    for tag, (n, hw, sharp) in {"mixed": (6, 64, 6.0), "all_conf": (2, 48, 40.0), "none_conf": (2, 48, 0.3)}.items():
        logits = (synth.uniform01(f"thr/{tag}", (n, hw, hw)) - 0.5) * sharp
        if tag == "mixed":
            logits[1] = logits[1] * 0.05          # unconfident sample -> dropped
            logits[4] = -20.0                     # confident but (almost) empty -> pixel_num <= 1000 -> dropped
        probs = torch.sigmoid(logits)
        img = synth.uniform01(f"thr/{tag}/img", (n, 3, hw, hw))
        for allow in (True, False):
            ci, cm, cp, conf, loss = thr(img, probs, allow_throw_sample=allow)
            key = f"threshold/{tag}/allow{int(allow)}"
            out[key] = {
                "kept_imgs": None if ci is None else ci.detach().clone(),
                "kept_masks": None if cm is None else cm.detach().clone(),
                "kept_probs": cp.detach().clone() if torch.is_tensor(cp) else None,
                "confidence": conf.detach().clone(),
                "loss": loss.detach().clone(),
            }
    return out


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    SegformerConfig, Model, ref_loss = import_reference()
    for name, case in cases.MODEL_CASES.items():
        res = run_case(name, case, SegformerConfig, Model, ref_loss)
        path = os.path.join(HERE, f"{name}.pt")
        torch.save(res, path)
        print(f"{name}: logits {tuple(res['logits'].shape)} mean {res['logits'].mean():.6f} -> {os.path.getsize(path)/1024:.0f} KiB")
    lt = loss_and_threshold_goldens(ref_loss)
    torch.save(lt, os.path.join(HERE, "loss_threshold.pt"))
    print("loss_threshold:", {k: (float(v) if torch.is_tensor(v) and v.numel() == 1 else "...") for k, v in lt.items()})


if __name__ == "__main__":
    main()
