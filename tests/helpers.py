"""Shared test helpers: rebuild the synthetic weights / inputs of a golden case and run the oracle on them."""
from __future__ import annotations

import os

import torch

from oracle import segformer_oracle as O
from oracle import synth
from tests.golden import cases

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN_DIR, f"{name}.pt"), weights_only=False)


def oracle_config(case) -> O.OracleConfig:
    cfg = O.OracleConfig()
    for k, v in case["config"].items():
        setattr(cfg, k, v)
    return cfg


def case_inputs(name):
    """(cfg, state_dict, x, prompts, cls, labels, mask) for a golden case, all deterministic."""
    case = cases.MODEL_CASES[name]
    cfg = oracle_config(case)
    sd = synth.synth_state_dict(O.state_dict_template(cfg))
    b, hw = case["batch"], case["hw"]
    x = synth.uniform01(f"{name}/x", (b, 3, hw, hw))
    prompts = [None] * 4
    for s, spec in enumerate(case.get("prompts", [None] * 4)):
        if spec is not None:
            prompts[s] = synth.uniform01(f"{name}/prompt{s}", spec)
    cls = [None] * 4
    for s, n in enumerate(case.get("cls", [0] * 4)):
        if n:
            cls[s] = synth.uniform01(f"{name}/cls{s}", (n, cfg.hidden_sizes[s]))
    labels = None
    if case.get("labels") == "binary":
        labels = (synth.uniform01(f"{name}/labels", (b, hw, hw)) > 0.6).to(torch.int64)
        ign = synth.uniform01(f"{name}/ignore", (b, hw, hw)) > 0.95
        labels = torch.where(ign, torch.full_like(labels, 255), labels)
    elif case.get("labels") == "multi":
        labels = (synth.uniform01(f"{name}/labels", (b, hw, hw)) * cfg.num_labels).to(torch.int64).clamp_(max=cfg.num_labels - 1)
        ign = synth.uniform01(f"{name}/ignore", (b, hw, hw)) > 0.95
        labels = torch.where(ign, torch.full_like(labels, 255), labels)
    mask = (synth.uniform01(f"{name}/mask", (b, hw, hw)) > 0.7).to(torch.int64) if cfg.num_labels == 1 else None
    return case, cfg, sd, x, prompts, cls, labels, mask
