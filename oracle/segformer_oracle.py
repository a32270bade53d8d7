"""TEST INFRASTRUCTURE ONLY — CPU restatement (plain PyTorch fp32, functional) of the reference SegFormer path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product path (semisupervisedobjectdetection_b200/) never does.

Parity status: PINNED against the real reference executed in the build container — tests/golden/make_golden.py
imports /root/reference (unmodified) and stores its outputs in tests/golden/*.pt; tests/test_oracle_golden.py checks
this restatement against those fixtures (fp32, atol 1e-5).  The reference ships no tests/golden vectors of its own
(SURVEY.md §4), so "pinned" here means pinned by executing the reference, not by upstream fixtures.

Citations: MS = segformer_source/models/segformer/modeling_segformer.py, SFM = models/SegFormerModel.py,
LOSS = models/Loss.py, TS = main_segformer/segFormer_semi_teacherstudent_main.py (all under the reference root).

Everything is written as pure functions over a state dict (keys exactly as in the reference checkpoint layout,
`segformer.encoder...` / `decode_head...`), so it shares no code structure with the nn.Module-based reference.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
LN_EPS = 1e-5  # every nn.LayerNorm in MS uses the torch default, config.layer_norm_eps is ignored (MS:133,174,325,333,418)
BN_EPS = 1e-5
BN_MOMENTUM = 0.1


@dataclass
class OracleConfig:
    """Field names follow SegformerConfig (configuration_segformer.py:105-154)."""
    num_channels: int = 3
    num_encoder_blocks: int = 4
    depths: Sequence[int] = (2, 2, 2, 2)
    sr_ratios: Sequence[int] = (8, 4, 2, 1)
    hidden_sizes: Sequence[int] = (32, 64, 160, 256)
    patch_sizes: Sequence[int] = (7, 3, 3, 3)
    strides: Sequence[int] = (4, 2, 2, 2)
    num_attention_heads: Sequence[int] = (1, 2, 5, 8)
    mlp_ratios: Sequence[int] = (4, 4, 4, 4)
    decoder_hidden_size: int = 256
    num_labels: int = 1
    semantic_loss_ignore_index: int = 255
    drop_path_rate: float = 0.1
    classifier_dropout_prob: float = 0.1

    @staticmethod
    def mit_b5(num_labels: int = 1) -> "OracleConfig":
        return OracleConfig(depths=(3, 6, 40, 3), hidden_sizes=(64, 128, 320, 512), decoder_hidden_size=768,
                            num_labels=num_labels)

    @staticmethod
    def mit_b0(num_labels: int = 1) -> "OracleConfig":
        return OracleConfig(num_labels=num_labels)


def state_dict_template(cfg: OracleConfig) -> Dict[str, Tensor]:
    """Names and shapes of the reference SegformerForSemanticSegmentation state dict (MS:123-133,163-174,236,286,
    301-307,325-335,417-419,713-715,738-748). Values are zeros; fill with oracle.synth.synth_state_dict."""
    sd: Dict[str, Tensor] = {}

    def lin(prefix, out_f, in_f):
        sd[prefix + ".weight"] = torch.zeros(out_f, in_f)
        sd[prefix + ".bias"] = torch.zeros(out_f)

    def ln(prefix, c):
        sd[prefix + ".weight"] = torch.zeros(c)
        sd[prefix + ".bias"] = torch.zeros(c)

    enc = "segformer.encoder."
    for i in range(cfg.num_encoder_blocks):
        cin = cfg.num_channels if i == 0 else cfg.hidden_sizes[i - 1]
        c = cfg.hidden_sizes[i]
        k = cfg.patch_sizes[i]
        sd[f"{enc}patch_embeddings.{i}.proj.weight"] = torch.zeros(c, cin, k, k)
        sd[f"{enc}patch_embeddings.{i}.proj.bias"] = torch.zeros(c)
        ln(f"{enc}patch_embeddings.{i}.layer_norm", c)
        for j in range(cfg.depths[i]):
            p = f"{enc}block.{i}.{j}."
            ln(p + "layer_norm_1", c)
            lin(p + "attention.self.query", c, c)
            lin(p + "attention.self.key", c, c)
            lin(p + "attention.self.value", c, c)
            sr = cfg.sr_ratios[i]
            if sr > 1:
                sd[p + "attention.self.sr.weight"] = torch.zeros(c, c, sr, sr)
                sd[p + "attention.self.sr.bias"] = torch.zeros(c)
                ln(p + "attention.self.layer_norm", c)
            lin(p + "attention.output.dense", c, c)
            ln(p + "layer_norm_2", c)
            hid = int(c * cfg.mlp_ratios[i])
            lin(p + "mlp.dense1", hid, c)
            sd[p + "mlp.dwconv.dwconv.weight"] = torch.zeros(hid, 1, 3, 3)
            sd[p + "mlp.dwconv.dwconv.bias"] = torch.zeros(hid)
            lin(p + "mlp.dense2", c, hid)
        ln(f"{enc}layer_norm.{i}", c)
    dec = cfg.decoder_hidden_size
    for i in range(cfg.num_encoder_blocks):
        lin(f"decode_head.linear_c.{i}.proj", dec, cfg.hidden_sizes[i])
        lin(f"decode_head.linear_c.{i}.cls_proj", dec, 512)  # hard-coded 512 (MS:733)
    sd["decode_head.linear_fuse.weight"] = torch.zeros(dec, dec * cfg.num_encoder_blocks, 1, 1)
    sd["decode_head.batch_norm.weight"] = torch.zeros(dec)
    sd["decode_head.batch_norm.bias"] = torch.zeros(dec)
    sd["decode_head.batch_norm.running_mean"] = torch.zeros(dec)
    sd["decode_head.batch_norm.running_var"] = torch.zeros(dec)
    sd["decode_head.batch_norm.num_batches_tracked"] = torch.zeros((), dtype=torch.int64)
    sd["decode_head.classifier.weight"] = torch.zeros(cfg.num_labels, dec, 1, 1)
    sd["decode_head.classifier.bias"] = torch.zeros(cfg.num_labels)
    return sd


# --------------------------------------------------------------------------------------------------------------
# building blocks
# --------------------------------------------------------------------------------------------------------------
def _ln(x: Tensor, sd, prefix: str) -> Tensor:
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], LN_EPS)


def _linear(x: Tensor, sd, prefix: str) -> Tensor:
    return F.linear(x, sd[prefix + ".weight"], sd[prefix + ".bias"])


def patch_embed(x_nchw: Tensor, sd, prefix: str, patch: int, stride: int) -> Tuple[Tensor, int, int]:
    """MS:135-142: conv(k=patch, s=stride, p=patch//2) -> tokens [B, H*W, C] -> LayerNorm."""
    y = F.conv2d(x_nchw, sd[prefix + ".proj.weight"], sd[prefix + ".proj.bias"], stride=stride, padding=patch // 2)
    b, c, h, w = y.shape
    tokens = y.reshape(b, c, h * w).transpose(1, 2)
    return _ln(tokens, sd, prefix + ".layer_norm"), h, w


def efficient_self_attention(x: Tensor, h: int, w: int, sd, prefix: str, heads: int, sr: int, ptl: int) -> Tensor:
    """MS:181-230. x: [B, ptl+N, C] (already layer-normed). Prompt/cls rows bypass the SR conv but pass through the
    SR LayerNorm (MS:194-205)."""
    b, t, c = x.shape
    d = c // heads
    q = _linear(x, sd, prefix + ".query").reshape(b, t, heads, d).transpose(1, 2)
    kv_src = x
    if sr > 1:
        front, spatial = x[:, :ptl], x[:, ptl:]
        fmap = spatial.transpose(1, 2).reshape(b, c, h, w)
        red = F.conv2d(fmap, sd[prefix + ".sr.weight"], sd[prefix + ".sr.bias"], stride=sr)
        red = red.reshape(b, c, -1).transpose(1, 2)
        if ptl:
            red = torch.cat([front, red], dim=1)
        kv_src = _ln(red, sd, prefix + ".layer_norm")
    tk = kv_src.shape[1]
    k = _linear(kv_src, sd, prefix + ".key").reshape(b, tk, heads, d).transpose(1, 2)
    v = _linear(kv_src, sd, prefix + ".value").reshape(b, tk, heads, d).transpose(1, 2)
    scores = (q @ k.transpose(-1, -2)) / math.sqrt(d)
    probs = scores.softmax(dim=-1)
    ctx = (probs @ v).transpose(1, 2).reshape(b, t, c)
    return ctx


def mix_ffn(x: Tensor, h: int, w: int, sd, prefix: str) -> Tensor:
    """MS:310-317 + MS:288-294: dense1 -> depthwise 3x3 (+bias) -> exact-erf GELU -> dense2."""
    b, n, _ = x.shape
    y = _linear(x, sd, prefix + ".dense1")
    hid = y.shape[-1]
    fmap = y.transpose(1, 2).reshape(b, hid, h, w)
    fmap = F.conv2d(fmap, sd[prefix + ".dwconv.dwconv.weight"], sd[prefix + ".dwconv.dwconv.bias"], padding=1, groups=hid)
    y = F.gelu(fmap.reshape(b, hid, n).transpose(1, 2))
    return _linear(y, sd, prefix + ".dense2")


def encoder_layer(x: Tensor, h: int, w: int, sd, prefix: str, heads: int, sr: int, ptl: int,
                  keep_scale: Optional[Tuple[Tensor, Tensor]] = None) -> Tuple[Tensor, Tensor]:
    """MS:337-363. Returns (spatial tokens [B,N,C], front rows [B,ptl,C]).
    keep_scale = (s_attn[B], s_mlp[B]) are the per-sample drop-path multipliers mask/keep_prob (MS:95-102)."""
    a = efficient_self_attention(_ln(x, sd, prefix + "layer_norm_1"), h, w, sd, prefix + "attention.self", heads, sr, ptl)
    a = _linear(a, sd, prefix + "attention.output.dense")
    if keep_scale is not None:
        a = a * keep_scale[0].view(-1, 1, 1)
    x = a + x
    front, x = x[:, :ptl], x[:, ptl:]
    m = mix_ffn(_ln(x, sd, prefix + "layer_norm_2"), h, w, sd, prefix + "mlp")
    if keep_scale is not None:
        m = m * keep_scale[1].view(-1, 1, 1)
    return m + x, front


def encoder(pixel_values: Tensor, sd, cfg: OracleConfig, prompt_tokens: Optional[List[Optional[Tensor]]] = None,
            cls_tokens: Optional[List[Optional[Tensor]]] = None,
            keep_scales: Optional[List[Tuple[Tensor, Tensor]]] = None):
    """MS:421-487. Returns (list of 4 NCHW maps, per-stage cls rows or None).
    keep_scales: optional list (one entry per global layer index) of drop-path multipliers."""
    prompt_tokens = prompt_tokens or [None] * cfg.num_encoder_blocks
    cls_tokens = cls_tokens or [None] * cfg.num_encoder_blocks
    x = pixel_values
    bsz = x.shape[0]
    maps, cls_out = [], []
    gidx = 0
    for s in range(cfg.num_encoder_blocks):
        pfx = f"segformer.encoder.patch_embeddings.{s}"
        x, h, w = patch_embed(x, sd, pfx, cfg.patch_sizes[s], cfg.strides[s])
        stage_cls = None
        running_cls = None
        for j in range(cfg.depths[s]):
            ptl = 0
            pt = prompt_tokens[s]
            if pt is not None:
                rows = pt[j] if pt.dim() == 3 else pt            # MS:447-450
                x = torch.cat([rows.unsqueeze(0).expand(bsz, -1, -1), x], dim=1)
                ptl = rows.shape[0]
            if cls_tokens[s] is not None:
                if j == 0:
                    running_cls = cls_tokens[s].unsqueeze(0).expand(bsz, -1, -1)   # MS:456-458
                x = torch.cat([running_cls, x], dim=1)
                ptl += running_cls.shape[1]
            ks = keep_scales[gidx] if keep_scales is not None else None
            x, front = encoder_layer(x, h, w, sd, f"segformer.encoder.block.{s}.{j}.", cfg.num_attention_heads[s],
                                     cfg.sr_ratios[s], ptl, ks)
            if cls_tokens[s] is not None:
                running_cls = front[:, 0:1, :]                    # MS:463-465 (first front row only)
                stage_cls = running_cls
            gidx += 1
        x = _ln(x, sd, f"segformer.encoder.layer_norm.{s}")
        x = x.reshape(bsz, h, w, -1).permute(0, 3, 1, 2).contiguous()
        maps.append(x)
        cls_out.append(stage_cls)
    if any(c is None for c in cls_out):
        cls_out = None
    return maps, cls_out


def decode_head(maps: List[Tensor], sd, cfg: OracleConfig, cls_token: Optional[Tensor] = None, bn_training: bool = False,
                dropout_mask: Optional[Tensor] = None, bn_buffers: Optional[Dict[str, Tensor]] = None) -> Tensor:
    """MS:752-782. maps: 4 NCHW stage outputs. cls_token: sigmoid(stage-4 cls row) [B,1,512] or None.
    dropout_mask: optional {0,1/(1-p)} multiplier [B,dec,H,W] applied before the classifier (train mode)."""
    bsz = maps[-1].shape[0]
    size = maps[0].shape[2:]
    ups = []
    for i, fmap in enumerate(maps):
        _, c, h, w = fmap.shape
        t = _linear(fmap.flatten(2).transpose(1, 2), sd, f"decode_head.linear_c.{i}.proj")
        if cls_token is not None:
            t = t + _linear(cls_token, sd, f"decode_head.linear_c.{i}.cls_proj")   # broadcast over tokens (MS:720-723)
        t = t.transpose(1, 2).reshape(bsz, -1, h, w)
        ups.append(F.interpolate(t, size=size, mode="bilinear", align_corners=False))
    fused = F.conv2d(torch.cat(ups[::-1], dim=1), sd["decode_head.linear_fuse.weight"])   # order c4,c3,c2,c1 (MS:774)
    rm, rv = sd["decode_head.batch_norm.running_mean"], sd["decode_head.batch_norm.running_var"]
    if bn_training:
        rm = rm.clone() if bn_buffers is None else bn_buffers["running_mean"]
        rv = rv.clone() if bn_buffers is None else bn_buffers["running_var"]
    y = F.batch_norm(fused, rm, rv, sd["decode_head.batch_norm.weight"], sd["decode_head.batch_norm.bias"],
                     training=bn_training, momentum=BN_MOMENTUM, eps=BN_EPS)
    y = F.relu(y)
    if dropout_mask is not None:
        y = y * dropout_mask
    return F.conv2d(y, sd["decode_head.classifier.weight"], sd["decode_head.classifier.bias"])


def forward_logits(pixel_values: Tensor, sd, cfg: OracleConfig, prompt_tokens=None, cls_tokens=None,
                   bn_training: bool = False, keep_scales=None, dropout_mask=None, bn_buffers=None):
    """MS:800-851: encoder -> sigmoid(last-stage cls) -> decode head. Returns (logits [B,nl,H/4,W/4], maps, cls list)."""
    maps, cls_list = encoder(pixel_values, sd, cfg, prompt_tokens, cls_tokens, keep_scales)
    cls_in = torch.sigmoid(cls_list[-1]) if cls_list is not None else None
    logits = decode_head(maps, sd, cfg, cls_in, bn_training, dropout_mask, bn_buffers)
    return logits, maps, cls_list


def hf_loss(logits: Tensor, labels: Tensor, cfg: OracleConfig) -> Tensor:
    """MS:853-868: in-model loss (use_loss='bce' path of SFM:109-110)."""
    up = F.interpolate(logits, size=labels.shape[-2:], mode="bilinear", align_corners=False)
    if cfg.num_labels > 1:
        return F.cross_entropy(up, labels, ignore_index=cfg.semantic_loss_ignore_index)
    valid = ((labels >= 0) & (labels != cfg.semantic_loss_ignore_index)).float()
    per_px = F.binary_cross_entropy_with_logits(up.squeeze(1), labels.float(), reduction="none")
    return (per_px * valid).mean()


def predict_probs(logits: Tensor, out_hw: Tuple[int, int]) -> Tensor:
    """SFM:117-125: bilinear upsample to the image size, sigmoid, squeeze the channel dim."""
    up = F.interpolate(logits, size=out_hw, mode="bilinear", align_corners=False)
    return torch.squeeze(torch.sigmoid(up), 1)


# --------------------------------------------------------------------------------------------------------------
# losses (LOSS:16-95)
# --------------------------------------------------------------------------------------------------------------
def _as4d(t: Tensor) -> Tensor:
    if t.dim() == 3:
        return t.unsqueeze(1)
    if t.dim() == 2:
        return t[None, None]
    return t


def dice_score(pred: Tensor, gt: Tensor, smooth: float = 1.0) -> Tensor:
    """LOSS:33-41 with activation 'none': mean over samples of (2|Σ g·p| + s) / (Σ|p| + Σ g + s)."""
    n = gt.shape[0]
    p = pred.reshape(n, -1)
    g = gt.reshape(n, -1)
    tp = (g * p).sum(1).abs()
    fp = p.abs().sum(1)
    fn = g.sum(1)
    return ((2 * tp + smooth) / (fp + fn + smooth)).mean()


def segmentation_loss(pred: Tensor, target: Tensor, loss_type: str = "dice") -> Tensor:
    """LOSS:66-95 for num_classes == 1 (the only value the wrapper uses, SFM:39-41)."""
    pred, target = _as4d(pred), _as4d(target)
    if loss_type == "mse":
        n = target.shape[0]
        err = ((target.reshape(n, -1) - pred.reshape(n, -1)) ** 2).sum(1)
        return (err / (target.shape[0] * target.shape[1])).mean()      # LOSS:49-54
    if loss_type == "dice_argmax":
        pred = torch.where(pred >= 0.5, 1, 0)
    elif loss_type != "dice":
        raise ValueError(loss_type)
    return 1 - dice_score(pred[:, 0:1], target[:, 0:1])


# --------------------------------------------------------------------------------------------------------------
# teacher-student glue (TS:18-56, TS:110-126)
# --------------------------------------------------------------------------------------------------------------
PSEUDO_MASK_THRESHOLD = 0.7
CONFIDENT_THRESHOLD = 0.7


def threshold_pseudo_masks(img: Tensor, probs: Tensor, allow_throw_sample: bool = True,
                           thr: float = PSEUDO_MASK_THRESHOLD, conf_thr: float = CONFIDENT_THRESHOLD):
    """TS:24-56. Returns (imgs_kept | None, int64 masks_kept | None, probs_kept | [], confidence[N], mean dice loss)."""
    n = probs.shape[0]
    flat = probs.reshape(n, -1)
    pixel_num = flat.abs().sum(1)
    confident = torch.where((flat >= thr) | (flat <= 1 - thr), 1, 0)
    confidence = confident.sum(1) / probs[0].numel()
    pseudo = torch.where(probs >= thr, 1, 0)
    keep = [i for i in range(n)
            if (not allow_throw_sample) or (pixel_num[i] > 1000 and confidence[i] >= conf_thr)]
    total = torch.zeros((), dtype=torch.float32)
    for i in keep:
        total = total + segmentation_loss(probs[i], pseudo[i], "dice")
    if keep:
        idx = torch.tensor(keep)
        return img[idx], pseudo[idx], probs[idx], confidence, total / len(keep)
    return None, None, [], confidence, torch.tensor(float("nan"))   # reference: 0-tensor / len([]) -> nan


def denoise_teacher_probs(teacher_probs: Tensor, gt: Tensor) -> Tensor:
    """TS:111-114: clamp(t + 0.2*gt - 0.1, 0, 1) written with the same op order."""
    t = teacher_probs + 0.2 * gt
    t = t - 0.1
    t = torch.where(t > 1, torch.ones_like(t), t)
    return torch.where(t < 0, torch.zeros_like(t), t)


# --------------------------------------------------------------------------------------------------------------
# optimiser (SFM:32-37,158-162) and teacher refresh (TS:202-204; EMA with decay d, d=0 == hard copy)
# --------------------------------------------------------------------------------------------------------------
def clip_adam_l2_step(param: Tensor, grad: Tensor, exp_avg: Tensor, exp_avg_sq: Tensor, step: int, lr: float,
                      weight_decay: float, clip: float = 1.2, beta1: float = 0.5, beta2: float = 0.999, eps: float = 1e-8):
    """clip_grad_value_(1.2) then torch.optim.Adam (L2-in-gradient, not AdamW). `step` is 1-based."""
    g = grad.clamp(-clip, clip)
    if weight_decay != 0:
        g = g + weight_decay * param
    m = beta1 * exp_avg + (1 - beta1) * g
    v = beta2 * exp_avg_sq + (1 - beta2) * g * g
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = v.sqrt() / math.sqrt(bc2) + eps
    new_p = param - (lr / bc1) * (m / denom)
    return new_p, m, v


def ema_update(teacher: Tensor, student: Tensor, decay: float) -> Tensor:
    return decay * teacher + (1 - decay) * student


# ---------------------------------------------------------------------------------------------------------------
# few-shot / domain-prompt cls-token losses (main_segformer/segFormer_fewshot_learning.py, "FS")
# ---------------------------------------------------------------------------------------------------------------
def compute_similarity(mat1: Tensor, mat2: Tensor) -> Tensor:
    """FS:185-188 with similarity_loss = torch.nn.CosineSimilarity(dim=1) (FS:353): mean cosine over the batch."""
    mat1 = mat1.squeeze(1) if mat1.dim() == 3 else mat1
    mat2 = mat2.squeeze(1) if mat2.dim() == 3 else mat2
    return torch.mean(F.cosine_similarity(mat1, mat2, dim=1, eps=1e-8))


def fewshot_losses(category_1_loss, category_1_cls_token, category_2_loss, category_2_cls_token):
    """FS:216-236 -> (summation_loss, intra_loss_1, intra_loss_2, inter_loss)."""
    batch_size = min(category_1_cls_token.shape[0], category_2_cls_token.shape[0])
    inter_loss = 0.5 + 0.5 * compute_similarity(category_1_cls_token[:batch_size, :, :], category_2_cls_token[:batch_size, :, :])
    intra_loss_1 = 0.5 - 0.5 * compute_similarity(category_1_cls_token[:batch_size // 2, :, :],
                                                  category_1_cls_token[-(batch_size // 2):, :, :])
    intra_loss_2 = 0.5 - 0.5 * compute_similarity(category_2_cls_token[:batch_size // 2, :, :],
                                                  category_2_cls_token[-(batch_size // 2):, :, :])
    category_1_summation_loss = category_1_loss + 100 * inter_loss + 100 * intra_loss_1
    category_2_summation_loss = category_2_loss + 100 * inter_loss + 100 * intra_loss_2
    return (category_1_summation_loss + category_2_summation_loss) / 2, intra_loss_1, intra_loss_2, inter_loss
