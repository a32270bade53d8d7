"""B200-native SegFormer (MiT-b0..b5 encoder + all-MLP decode head) behind the reference's module API.

Class names, constructor / forward signatures, return tuples, attribute paths and state-dict keys follow
segformer_source/models/segformer/modeling_segformer.py (MS) of the reference so that models/SegFormerModel.py and the
three training scripts can use this file unchanged; every forward/backward computation runs in libsfk.so (sm_100a CUDA).
nn.Linear / nn.Conv2d / nn.LayerNorm / nn.BatchNorm2d instances below are parameter containers only (identical names,
shapes and init) — their PyTorch forward is never called.
"""
from __future__ import annotations

import ctypes
import os
import math
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch
from torch import nn

from . import _lib as L
from . import ops
from .configuration import SegformerConfig

LN_EPS = 1e-5  # MS:133,174,325,333,418 construct nn.LayerNorm with the default eps; config.layer_norm_eps is unused


# ---------------------------------------------------------------------------------------------------------------
# output containers (ModelOutput look-alikes: attribute + index access)
# ---------------------------------------------------------------------------------------------------------------
class _Output:
    _fields: Tuple[str, ...] = ()

    def to_tuple(self):
        return tuple(getattr(self, f) for f in self._fields if getattr(self, f) is not None)

    def __getitem__(self, i):
        if isinstance(i, str):
            return getattr(self, i)
        return self.to_tuple()[i]

    def __iter__(self):
        return iter(self.to_tuple())


@dataclass
class BaseModelOutput(_Output):
    last_hidden_state: Optional[torch.Tensor] = None
    hidden_states: Optional[Tuple[torch.Tensor, ...]] = None
    attentions: Optional[Tuple[torch.Tensor, ...]] = None
    _fields = ("last_hidden_state", "hidden_states", "attentions")


class SemanticSegmenterOutput(_Output):
    """loss / logits / hidden_states / attentions.  `hidden_states` may be given as a zero-argument callable: the NCHW fp32
    stage maps are then built on first access.  The reference wrapper switches config.output_hidden_states on (SFM:17) but no
    project script ever reads the maps, and building them eagerly cost 16 layout-copy launches and ~0.9 GB of
    HBM traffic per teacher-student step."""
    _fields = ("loss", "logits", "hidden_states", "attentions")

    def __init__(self, loss=None, logits=None, hidden_states=None, attentions=None):
        self.loss, self.logits, self.attentions = loss, logits, attentions
        self._hidden, self._hidden_grad_mode = hidden_states, torch.is_grad_enabled()

    @property
    def hidden_states(self):
        if callable(self._hidden):
            with torch.set_grad_enabled(self._hidden_grad_mode):   # as if built inside the forward call
                self._hidden = self._hidden()
        return self._hidden

    @hidden_states.setter
    def hidden_states(self, value):
        self._hidden = value

    def __repr__(self):
        return f"SemanticSegmenterOutput(loss={self.loss!r}, logits={self.logits!r}, hidden_states=<{'lazy' if callable(self._hidden) else type(self._hidden).__name__}>)"


def _versions(params) -> tuple:
    return tuple(p._version for p in params) + (ops.WEIGHT_EPOCH, params[0].data_ptr())


def sh_ptr(p) -> int:
    w16 = getattr(p, "_sfk_w16", None)
    return -1 if w16 is None else w16.data_ptr()


def _any_requires_grad(*tensors) -> bool:
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors)


class _GradAnchor(nn.Module):
    """Zero-size tensor with requires_grad=True handed to the autograd Functions so that their backward runs (and
    writes parameter gradients in place) even when the activation input itself does not require grad."""

    def __init__(self):
        super().__init__()
        self._anchor = None

    def get(self, params, device):
        if not (torch.is_grad_enabled() and any(p.requires_grad for p in params)):
            return None
        if self._anchor is None or self._anchor.device != device:
            self._anchor = torch.zeros(0, device=device, requires_grad=True)
        return self._anchor


# ---------------------------------------------------------------------------------------------------------------
# drop path (MS:85-117)
# ---------------------------------------------------------------------------------------------------------------
def drop_path_scale(batch: int, drop_prob: float, training: bool, device) -> Optional[torch.Tensor]:
    """Per-sample multiplier floor(keep + U[0,1)) / keep (MS:95-102) or None when inactive."""
    if drop_prob == 0.0 or not training:
        return None
    keep = 1.0 - drop_prob
    return torch.floor(keep + torch.rand(batch, device=device, dtype=torch.float32)) / keep


_KEEP_CACHE: dict = {}


def drop_path_scales_batched(batch: int, drop_probs, training: bool, device):
    """All per-sample drop-path multipliers of one forward pass in three tensor ops instead of four per layer:
    returns a list (one entry per layer) of None or a [2, batch] tensor (attention branch, MLP branch; MS:350,358)."""
    if not training or all(p == 0.0 for p in drop_probs):
        return [None] * len(drop_probs)
    ck = (tuple(drop_probs), str(device))
    keep = _KEEP_CACHE.get(ck)
    if keep is None:  # built once: a host->device copy per forward would serialise the stream
        keep = torch.tensor([1.0 - p for p in drop_probs], dtype=torch.float32).view(-1, 1, 1).to(device)
        _KEEP_CACHE[ck] = keep
    scales = torch.floor(keep + torch.rand(len(drop_probs), 2, batch, device=device, dtype=torch.float32)) / keep
    return [None if p == 0.0 else scales[i] for i, p in enumerate(drop_probs)]


class SegformerDropPath(nn.Module):
    def __init__(self, drop_prob: Optional[float] = None) -> None:
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, hidden_states: torch.Tensor) -> torch.Tensor:
        s = drop_path_scale(hidden_states.shape[0], self.drop_prob or 0.0, self.training, hidden_states.device)
        if s is None:
            return hidden_states
        flat = hidden_states.reshape(hidden_states.shape[0], -1).float().contiguous()
        rows_per = flat.shape[1] // 4
        out = ops.scale_cast(flat.reshape(-1, 4), torch.float32, s, rows_per)
        return out.reshape(hidden_states.shape).to(hidden_states.dtype)

    def extra_repr(self) -> str:
        return "p={}".format(self.drop_prob)


# ---------------------------------------------------------------------------------------------------------------
# overlapping patch embedding (MS:120-142)
# ---------------------------------------------------------------------------------------------------------------
class _PatchEmbedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x_in, anchor, mod, is_image, Hin, Win):
        K, S = mod.patch_size, mod.stride
        P = K // 2
        C, Cin = mod.proj.out_channels, mod.proj.in_channels
        B = x_in.shape[0]
        Ho, Wo = (Hin + 2 * P - K) // S + 1, (Win + 2 * P - K) // S + 1
        M = B * Ho * Wo
        sh = mod._shadow()
        Kp = sh["w"].shape[1]
        A = torch.empty(M, Kp, device=x_in.device, dtype=torch.bfloat16)
        if is_image:
            x_in = ops._chk(x_in, torch.float32, "pixel_values")
            ops.call("sfk_im2col_nchw_f32", x_in.data_ptr(), A.data_ptr(), B, Cin, Hin, Win, K, S, P, Kp, ops.stream())
        else:
            x_in = ops._chk(x_in, torch.bfloat16, "patch-embed input tokens")
            ops.call("sfk_im2col_nhwc", x_in.data_ptr(), A.data_ptr(), B, Hin, Win, Cin, K, S, P, Hin * Win, 0, ops.stream())
        y = ops.gemm(A, sh["w"], M, C, Kp, bias=mod.proj.bias, out_bf16=False)
        out, mean, rstd = ops.layernorm_fwd(y, mod.layer_norm.weight, mod.layer_norm.bias, out_bf16=False,
                                            need_stats=any(ctx.needs_input_grad))
        ctx.mod, ctx.meta, ctx.sh = mod, (is_image, B, Hin, Win, Ho, Wo, K, S, P, C, Cin, Kp, M), sh
        if any(ctx.needs_input_grad):
            ctx.save_for_backward(A, y, mean, rstd)
        return out.view(B, Ho * Wo, C)

    @staticmethod
    def backward(ctx, dout):
        mod = ctx.mod
        is_image, B, Hin, Win, Ho, Wo, K, S, P, C, Cin, Kp, M = ctx.meta
        A, y, mean, rstd = ctx.saved_tensors
        dout = ops._chk(dout, torch.float32, "patch-embed grad")
        ln = mod.layer_norm
        dy = ops.layernorm_bwd(dout.view(M, C), y, mean, rstd, ln.weight, ops.grad_buffer(ln.weight), ops.grad_buffer(ln.bias),
                               out_bf16=True)
        gw = ops.grad_buffer(mod.proj.weight)
        if gw is not None:
            stage = torch.zeros(C, Kp, device=dout.device, dtype=torch.float32)
            ops.gemm(dy, A, C, Kp, M, lda=C, ldb=Kp, a_mn=1, b_mn=1, out=stage, out_bf16=False, atomic_add=True)
            if is_image:  # rows already in OIHW-flatten order, only the K padding has to go
                ops.call("sfk_conv_wgrad_ohwi_to_oihw", stage.data_ptr(), gw.data_ptr(), C, Cin * K * K, 1, 1, Kp, ops.stream())
            else:
                ops.call("sfk_conv_wgrad_ohwi_to_oihw", stage.data_ptr(), gw.data_ptr(), C, Cin, K, K, Kp, ops.stream())
        gb = ops.grad_buffer(mod.proj.bias)
        if gb is not None:
            ops.colsum_into(dy, gb, M, C)
        if is_image:
            ops.join_wgrad()  # the image patch-embed is the last node of backward: gradients are complete after this point
        dx = None
        if (not is_image) and ctx.needs_input_grad[0]:
            dA = ops.gemm(dy, ctx.sh["w"], M, Kp, C, lda=C, ldb=Kp, b_mn=1, out_bf16=True)
            dx32 = torch.empty(B, Hin * Win, Cin, device=dout.device, dtype=torch.float32)
            ops.call("sfk_col2im_nhwc", dA.data_ptr(), dx32.data_ptr(), B, Hin, Win, Cin, K, S, P, Hin * Win, 0, 0, ops.stream())
            dx = ops.scale_cast(dx32, torch.bfloat16)
        return dx, None, None, None, None, None


class SegformerOverlapPatchEmbeddings(nn.Module):
    """Construct the overlapping patch embeddings (MS:120-142): strided conv as a tensor-core GEMM over gathered patches,
    bias in the GEMM epilogue, LayerNorm on the fp32 result."""

    def __init__(self, patch_size, stride, num_channels, hidden_size):
        super().__init__()
        self.patch_size, self.stride = patch_size, stride
        self.proj = nn.Conv2d(num_channels, hidden_size, kernel_size=patch_size, stride=stride, padding=patch_size // 2)
        self.layer_norm = nn.LayerNorm(hidden_size)
        self._anchor = _GradAnchor()
        self._sh, self._sh_key = None, None

    def _shadow(self):
        params = [self.proj.weight]
        key = _versions(params)
        if self._sh_key != key:
            w = self.proj.weight.detach()
            O, I, KH, KW = w.shape
            if I % 8 == 0:   # channels-last patches: [O, (ky,kx,c)]
                dst = torch.empty(O, KH * KW * I, device=w.device, dtype=torch.bfloat16)
                ops.call("sfk_conv_weight_ohwi", w.data_ptr(), dst.data_ptr(), O, I, KH, KW, KH * KW * I, ops.stream())
            else:            # image input: OIHW-flatten order, K padded to a multiple of 8 for the 16-byte TMA pitch
                kp = (I * KH * KW + 7) // 8 * 8
                dst = ops.cast_bf16(w.reshape(O, -1), kp)
            self._sh, self._sh_key = {"w": dst}, key
        return self._sh

    def _forward_tokens(self, x_in, is_image, Hin, Win):
        K, S = self.patch_size, self.stride
        P = K // 2
        Ho, Wo = (Hin + 2 * P - K) // S + 1, (Win + 2 * P - K) // S + 1
        anchor = self._anchor.get([self.proj.weight, self.proj.bias, self.layer_norm.weight, self.layer_norm.bias], x_in.device)
        return _PatchEmbedFn.apply(x_in, anchor, self, is_image, Hin, Win), Ho, Wo

    def forward(self, pixel_values):
        """pixel_values: NCHW float tensor -> (embeddings [B, H*W, C] fp32, H, W) as in MS:135-142."""
        B, Cin, Hin, Win = pixel_values.shape
        if Cin % 8 == 0:
            tokens = pixel_values.permute(0, 2, 3, 1).reshape(B, Hin * Win, Cin).to(torch.bfloat16).contiguous()
            return self._forward_tokens(tokens, False, Hin, Win)
        return self._forward_tokens(pixel_values.float().contiguous(), True, Hin, Win)


# ---------------------------------------------------------------------------------------------------------------
# attention / MixFFN sub-modules: parameter containers with the reference constructor signatures.
# SegformerLayer runs the fused per-layer kernel sequence; these classes' own forward methods drive the same
# libsfk ops for callers that use them stand-alone.
# ---------------------------------------------------------------------------------------------------------------
class SegformerEfficientSelfAttention(nn.Module):
    def __init__(self, config, hidden_size, num_attention_heads, sequence_reduction_ratio):
        super().__init__()
        self.hidden_size = hidden_size
        self.num_attention_heads = num_attention_heads
        if self.hidden_size % self.num_attention_heads != 0:
            raise ValueError(f"The hidden size ({self.hidden_size}) is not a multiple of the number of attention "
                             f"heads ({self.num_attention_heads})")
        self.attention_head_size = int(self.hidden_size / self.num_attention_heads)
        self.all_head_size = self.num_attention_heads * self.attention_head_size
        self.query = nn.Linear(self.hidden_size, self.all_head_size)
        self.key = nn.Linear(self.hidden_size, self.all_head_size)
        self.value = nn.Linear(self.hidden_size, self.all_head_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)
        self.sr_ratio = sequence_reduction_ratio
        if sequence_reduction_ratio > 1:
            self.sr = nn.Conv2d(hidden_size, hidden_size, kernel_size=sequence_reduction_ratio, stride=sequence_reduction_ratio)
            self.layer_norm = nn.LayerNorm(hidden_size)

    def transpose_for_scores(self, hidden_states):
        new_shape = hidden_states.size()[:-1] + (self.num_attention_heads, self.attention_head_size)
        return hidden_states.view(new_shape).permute(0, 2, 1, 3)

    def forward(self, hidden_states, height, width, prompt_token_len=0, output_attentions=False):
        raise NotImplementedError(
            "stand-alone SegformerEfficientSelfAttention.forward is not wired in this build; call SegformerLayer / "
            "SegformerEncoder (the fused libsfk path). There is no PyTorch fallback.")


class SegformerSelfOutput(nn.Module):
    def __init__(self, config, hidden_size):
        super().__init__()
        self.dense = nn.Linear(hidden_size, hidden_size)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        raise NotImplementedError("use SegformerLayer (fused libsfk path); there is no PyTorch fallback")


class SegformerAttention(nn.Module):
    def __init__(self, config, hidden_size, num_attention_heads, sequence_reduction_ratio):
        super().__init__()
        self.self = SegformerEfficientSelfAttention(config=config, hidden_size=hidden_size,
                                                    num_attention_heads=num_attention_heads,
                                                    sequence_reduction_ratio=sequence_reduction_ratio)
        self.output = SegformerSelfOutput(config, hidden_size=hidden_size)
        self.pruned_heads = set()

    def prune_heads(self, heads):
        raise NotImplementedError("head pruning (MS:257-273) is not used by any reference script and is not supported")

    def forward(self, hidden_states, height, width, prompt_token_len=0, output_attentions=False):
        raise NotImplementedError("use SegformerLayer (fused libsfk path); there is no PyTorch fallback")


class SegformerDWConv(nn.Module):
    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)

    def forward(self, hidden_states, height, width):
        raise NotImplementedError("use SegformerLayer (fused libsfk path); there is no PyTorch fallback")


class SegformerMixFFN(nn.Module):
    def __init__(self, config, in_features, hidden_features=None, out_features=None):
        super().__init__()
        out_features = out_features or in_features
        if config.hidden_act != "gelu":
            raise NotImplementedError("only the exact-erf 'gelu' activation of the reference configs is implemented")
        self.dense1 = nn.Linear(in_features, hidden_features)
        self.dwconv = SegformerDWConv(hidden_features)
        self.dense2 = nn.Linear(hidden_features, out_features)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, height, width):
        raise NotImplementedError("use SegformerLayer (fused libsfk path); there is no PyTorch fallback")


# ---------------------------------------------------------------------------------------------------------------
# one encoder layer (MS:320-363) -> one libsfk call per direction
# ---------------------------------------------------------------------------------------------------------------
class _BwdLink:
    """Hand-over between the backward passes of two consecutive layers of a stage (no prompt / cls rows).  The LayerNorm-1
    backward of layer i+1 writes dx = the output gradient of layer i; in the same pass it can emit what layer i's backward
    would otherwise start with: the drop-path-scaled bf16 copy of that gradient and its column sums (dense2 bias gradient).
    Layer i+1 fills `dmlp` / `dx_ptr`; layer i takes the copy only if autograd hands it exactly that dx tensor."""
    __slots__ = ("layer", "drop_mlp", "dmlp", "dx_ptr")

    def __init__(self, layer, drop_mlp):
        self.layer, self.drop_mlp, self.dmlp, self.dx_ptr = layer, drop_mlp, None, None


LAYER_LINK = os.environ.get("SFK_LAYER_LINK", "1") != "0"


class _LayerFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, front, anchor, layer, H, W, drop_attn, drop_mlp, link_in=None, link_out=None):
        x = ops._chk(x, torch.float32, "hidden_states")
        B, N, C = x.shape
        ptl = 0 if front is None else front.shape[1]
        if front is not None:
            front = ops._chk(front, torch.float32, "prompt/cls rows")
        att = layer.attention.self
        dims = ops.LayerDims(B, H, W, C, att.num_attention_heads, att.sr_ratio, layer.mlp.dense1.out_features, ptl)
        lib = ops.lib()
        need_grad = any(ctx.needs_input_grad)
        saved_bytes = lib.sfk_layer_saved_bytes(ctypes.byref(dims))
        scratch_bytes = lib.sfk_layer_scratch_bytes(ctypes.byref(dims))
        saved = (torch.empty(saved_bytes, dtype=torch.uint8, device=x.device) if need_grad
                 else ops.scratch(saved_bytes, x.device, "saved_nograd"))
        scratch = ops.scratch(scratch_bytes, x.device)
        out = torch.empty(B, N, C, device=x.device, dtype=torch.float32)
        front_out = torch.empty(B, ptl, C, device=x.device, dtype=torch.float32) if ptl else None
        params, keep = layer._param_struct(drop_attn, drop_mlp)
        ops.call("sfk_layer_fwd", ctypes.byref(dims), ctypes.byref(params), x.data_ptr(), ops._p(front), out.data_ptr(),
                 ops._p(front_out), saved.data_ptr(), saved_bytes, scratch.data_ptr(), scratch.numel(), ops.stream())
        if need_grad:
            ctx.layer, ctx.dims, ctx.saved, ctx.keep, ctx.ptl = layer, dims, saved, (params, keep, drop_attn, drop_mlp), ptl
            ctx.link_in, ctx.link_out = link_in, link_out
            ctx.save_for_backward(x)
        if front_out is None:
            return out, None
        return out, front_out

    @staticmethod
    def backward(ctx, dout, dfront_out):
        (x,) = ctx.saved_tensors
        layer, dims, ptl = ctx.layer, ctx.dims, ctx.ptl
        params = ctx.keep[0]
        B, N, C = x.shape
        if dout is None:
            dout = torch.zeros_like(x)
        link_in, link_out = ctx.link_in, ctx.link_out
        dmlp_in = None
        if link_out is not None and link_out.dmlp is not None:
            # the following layer already produced this layer's MixFFN input gradient (and its dense2 bias gradient)
            if link_out.dx_ptr != dout.data_ptr():
                raise RuntimeError("SegformerLayer backward: the output gradient is not the tensor the following layer produced "
                                   "(a second consumer of an intermediate layer output?); set SFK_LAYER_LINK=0 to disable the "
                                   "cross-layer backward fusion")
            dmlp_in, link_out.dmlp = link_out.dmlp, None
        dout = ops._chk(dout, torch.float32, "layer grad")
        if dfront_out is not None:
            dfront_out = ops._chk(dfront_out, torch.float32, "front grad")
        lib = ops.lib()
        scratch_bytes = lib.sfk_layer_scratch_bytes(ctypes.byref(dims))
        # weight-gradient GEMMs go to a side stream; consecutive calls alternate between two scratch arenas (parity)
        ws = ops.wgrad_stream(x.device)
        parity = ops.next_bwd_parity()
        scratch = ops.scratch(scratch_bytes, x.device, f"bwd{parity}")
        grads, gkeep = layer._grad_struct()
        dx = torch.empty_like(x)
        dfront = torch.empty(B, ptl, C, device=x.device, dtype=torch.float32) if ptl else None
        if dmlp_in is not None:
            params.dmlp_in = dmlp_in.data_ptr()
        dmlp_next = None
        if link_in is not None and ptl == 0 and grads.dln1_g and grads.dln1_b:
            dmlp_next = torch.empty(B * N, C, device=x.device, dtype=torch.bfloat16)
            db2_prev = ops.grad_buffer(link_in.layer.mlp.dense2.bias)
            gkeep.append(db2_prev)
            grads.next_dmlp, grads.next_drop_scale = dmlp_next.data_ptr(), ops._p(link_in.drop_mlp)
            grads.next_db2 = None if db2_prev is None else db2_prev.data_ptr()
        ops.call("sfk_layer_bwd", ctypes.byref(dims), ctypes.byref(params), ctypes.byref(grads), x.data_ptr(), dout.data_ptr(),
                 ops._p(dfront_out), dx.data_ptr(), ops._p(dfront), ctx.saved.data_ptr(), ctx.saved.numel(), scratch.data_ptr(),
                 scratch.numel(), ops.stream(), None if ws is None else ws.cuda_stream, parity)
        params.dmlp_in = None
        if dmlp_next is not None:
            link_in.dmlp, link_in.dx_ptr = dmlp_next, dx.data_ptr()
        if ws is not None:
            ops.defer_free(ctx.saved)  # the saved activations are still being read by the side-stream wgrad GEMMs
            if dmlp_in is not None:
                ops.defer_free(dmlp_in)  # ... and so is the handed-over gradient copy (dense2 wgrad operand)
        ctx.saved = None
        return dx, dfront, None, None, None, None, None, None, None, None


class SegformerLayer(nn.Module):
    """This corresponds to the Block class in the original implementation (MS:320-363)."""

    def __init__(self, config, hidden_size, num_attention_heads, drop_path, sequence_reduction_ratio, mlp_ratio):
        super().__init__()
        if config.hidden_dropout_prob != 0.0 or config.attention_probs_dropout_prob != 0.0:
            raise NotImplementedError("hidden/attention dropout > 0 is not implemented (every reference config uses 0.0)")
        self.layer_norm_1 = nn.LayerNorm(hidden_size)
        self.attention = SegformerAttention(config, hidden_size=hidden_size, num_attention_heads=num_attention_heads,
                                            sequence_reduction_ratio=sequence_reduction_ratio)
        self.drop_path = SegformerDropPath(drop_path) if drop_path > 0.0 else nn.Identity()
        self.layer_norm_2 = nn.LayerNorm(hidden_size)
        mlp_hidden_size = int(hidden_size * mlp_ratio)
        self.mlp = SegformerMixFFN(config, in_features=hidden_size, hidden_features=mlp_hidden_size)
        self._anchor = _GradAnchor()
        self._sh, self._sh_key, self._plist = None, None, None
        self._link_out = None

    # -- bf16 shadow weights in GEMM layouts (refreshed when the fp32 masters change) --
    def _params(self):
        if self._plist is None:
            self._plist = [p for p in self.parameters()]
        return self._plist

    def _shadow(self):
        key = _versions(self._params())
        if self._sh_key != key:
            a, m = self.attention.self, self.mlp
            C = a.hidden_size
            sh = self._sh
            flat_ok = sh is not None and sh.get("flat", False) and self._sh_key[-1] == key[-1] and self._sh_key[:-2] == key[:-2]
            if flat_ok:
                # only libsfk's weight epoch moved (optimiser / EMA step): the Linear operands are live views of the flat bf16
                # shadow; just re-run the two layout permutations into their existing buffers
                if a.sr_ratio > 1:
                    sr = a.sr_ratio
                    ops.call("sfk_conv_weight_ohwi", a.sr.weight.data_ptr(), sh["wsr"].data_ptr(), C, C, sr, sr, sr * sr * C, ops.stream())
                ops.call("sfk_dw_weight_to_tap", m.dwconv.dwconv.weight.data_ptr(), sh["dw_wt"].data_ptr(), m.dense1.out_features,
                         ops.stream())
                self._sh_key = key
                return sh
            sh = {}
            sh["wq"] = ops.weight_bf16(a.query.weight)
            wk, wv = ops.weight_bf16(a.key.weight), ops.weight_bf16(a.value.weight)
            same = wk.untyped_storage().data_ptr() == wv.untyped_storage().data_ptr()
            if same and wv.data_ptr() == wk.data_ptr() + 2 * C * C:   # adjacent in FusedClipAdam's flat shadow: a free [2C, C] view
                wkv = torch.as_strided(wk, (2 * C, C), (C, 1))
            else:
                wkv = torch.empty(2 * C, C, device=a.key.weight.device, dtype=torch.bfloat16)
                ops.call("sfk_cast_2d", a.key.weight.data_ptr(), C, wkv.data_ptr(), C, C, C, C, ops.stream())
                ops.call("sfk_cast_2d", a.value.weight.data_ptr(), C, wkv[C:].data_ptr(), C, C, C, C, ops.stream())
            sh["wkv"] = wkv
            bk, bv = a.key.bias.detach(), a.value.bias.detach()
            if bk.untyped_storage().data_ptr() == bv.untyped_storage().data_ptr() and bv.data_ptr() == bk.data_ptr() + 4 * C:
                sh["bkv"] = torch.as_strided(bk, (2 * C,), (1,))
            else:
                sh["bkv"] = torch.cat([bk, bv]).contiguous()
            sh["wo"] = ops.weight_bf16(self.attention.output.dense.weight)
            if a.sr_ratio > 1:
                sr = a.sr_ratio
                wsr = torch.empty(C, sr * sr * C, device=wkv.device, dtype=torch.bfloat16)
                ops.call("sfk_conv_weight_ohwi", a.sr.weight.data_ptr(), wsr.data_ptr(), C, C, sr, sr, sr * sr * C, ops.stream())
                sh["wsr"] = wsr
            sh["w1"] = ops.weight_bf16(m.dense1.weight)
            sh["w2"] = ops.weight_bf16(m.dense2.weight)
            hid = m.dense1.out_features
            wt = torch.empty(9, hid, device=wkv.device, dtype=torch.float32)
            ops.call("sfk_dw_weight_to_tap", m.dwconv.dwconv.weight.data_ptr(), wt.data_ptr(), hid, ops.stream())
            sh["dw_wt"] = wt
            sh["flat"] = all(getattr(p, "_sfk_w16", None) is not None for p in
                             (a.query.weight, a.key.weight, a.value.weight, self.attention.output.dense.weight, m.dense1.weight,
                              m.dense2.weight)) and sh["wkv"].data_ptr() == sh_ptr(a.key.weight) and sh["bkv"].data_ptr() == a.key.bias.data_ptr()
            self._sh, self._sh_key = sh, key
        return self._sh

    def _param_struct(self, drop_attn, drop_mlp):
        sh = self._shadow()
        a, m = self.attention.self, self.mlp
        p = ops.LayerParams()
        p.wq, p.wkv, p.wo = sh["wq"].data_ptr(), sh["wkv"].data_ptr(), sh["wo"].data_ptr()
        p.wsr = sh["wsr"].data_ptr() if "wsr" in sh else None
        p.w1, p.w2 = sh["w1"].data_ptr(), sh["w2"].data_ptr()
        p.bq, p.bkv, p.bo = a.query.bias.data_ptr(), sh["bkv"].data_ptr(), self.attention.output.dense.bias.data_ptr()
        p.bsr = a.sr.bias.data_ptr() if a.sr_ratio > 1 else None
        p.b1, p.b2 = m.dense1.bias.data_ptr(), m.dense2.bias.data_ptr()
        p.ln1_g, p.ln1_b = self.layer_norm_1.weight.data_ptr(), self.layer_norm_1.bias.data_ptr()
        p.ln2_g, p.ln2_b = self.layer_norm_2.weight.data_ptr(), self.layer_norm_2.bias.data_ptr()
        if a.sr_ratio > 1:
            p.lnsr_g, p.lnsr_b = a.layer_norm.weight.data_ptr(), a.layer_norm.bias.data_ptr()
        p.dw_wt, p.dw_b = sh["dw_wt"].data_ptr(), m.dwconv.dwconv.bias.data_ptr()
        p.drop_scale_attn = ops._p(drop_attn)
        p.drop_scale_mlp = ops._p(drop_mlp)
        return p, sh

    def _grad_struct(self):
        a, m, o = self.attention.self, self.mlp, self.attention.output.dense
        g = ops.LayerGrads()
        bufs = []

        def gp(param):
            t = ops.grad_buffer(param)
            bufs.append(t)
            return None if t is None else t.data_ptr()

        g.dwq, g.dwk, g.dwv, g.dwo = gp(a.query.weight), gp(a.key.weight), gp(a.value.weight), gp(o.weight)
        g.dbq, g.dbk, g.dbv, g.dbo = gp(a.query.bias), gp(a.key.bias), gp(a.value.bias), gp(o.bias)
        if a.sr_ratio > 1:
            g.dwsr, g.dbsr = gp(a.sr.weight), gp(a.sr.bias)
            g.dlnsr_g, g.dlnsr_b = gp(a.layer_norm.weight), gp(a.layer_norm.bias)
        g.dw1, g.dw2, g.db1, g.db2 = gp(m.dense1.weight), gp(m.dense2.weight), gp(m.dense1.bias), gp(m.dense2.bias)
        g.dln1_g, g.dln1_b = gp(self.layer_norm_1.weight), gp(self.layer_norm_1.bias)
        g.dln2_g, g.dln2_b = gp(self.layer_norm_2.weight), gp(self.layer_norm_2.bias)
        g.ddw, g.ddb = gp(m.dwconv.dwconv.weight), gp(m.dwconv.dwconv.bias)
        return g, bufs

    def forward(self, hidden_states, height, width, prompt_token_len=0, output_attentions=False):
        """hidden_states: [B, prompt_token_len + H*W, C]. Returns ((layer_output [B, H*W, C],), cls_token [B, ptl, C])."""
        if output_attentions:
            raise NotImplementedError("output_attentions=True would materialise the score tensor the fused kernel avoids")
        front = hidden_states[:, :prompt_token_len, :] if prompt_token_len else None
        spatial = hidden_states[:, prompt_token_len:, :] if prompt_token_len else hidden_states
        out, front_out = self._forward_split(spatial, front, height, width)
        if front_out is None:
            front_out = out[:, :0, :]
        return (out,), front_out

    def drop_prob(self) -> float:
        return self.drop_path.drop_prob if isinstance(self.drop_path, SegformerDropPath) else 0.0

    def _forward_split(self, spatial, front, height, width, scales=None, link_in=None):
        B = spatial.shape[0]
        if scales is None and self.training and self.drop_prob() > 0.0:
            # the same DropPath module is applied twice with independent draws (MS:350,358)
            scales = drop_path_scales_batched(B, [self.drop_prob()], True, spatial.device)[0]
        s_attn = scales[0] if scales is not None else None
        s_mlp = scales[1] if scales is not None else None
        anchor = self._anchor.get(self._params(), spatial.device)
        link_out = _BwdLink(self, s_mlp) if (LAYER_LINK and front is None and torch.is_grad_enabled()) else None
        self._link_out = link_out   # picked up by the encoder loop right after this call
        return _LayerFn.apply(spatial.float(), None if front is None else front.float(), anchor, self, height, width, s_attn, s_mlp,
                              link_in if front is None else None, link_out)


# ---------------------------------------------------------------------------------------------------------------
# stage-final LayerNorm (MS:470): fp32 residual stream -> bf16 tokens consumed by the next patch-embed and the head
# ---------------------------------------------------------------------------------------------------------------
class _StageNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, anchor, ln):
        x = ops._chk(x, torch.float32, "stage output")
        need = any(ctx.needs_input_grad)
        y, mean, rstd = ops.layernorm_fwd(x, ln.weight, ln.bias, out_bf16=True, need_stats=need)
        if need:
            ctx.ln = ln
            ctx.save_for_backward(x, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, mean, rstd = ctx.saved_tensors
        ln = ctx.ln
        dy = ops._chk(dy, torch.bfloat16, "stage-norm grad")
        dx = ops.layernorm_bwd(dy, x, mean, rstd, ln.weight, ops.grad_buffer(ln.weight), ops.grad_buffer(ln.bias), out_bf16=False)
        return dx, None, None


def _tokens_to_nchw(tokens: torch.Tensor, H: int, W: int) -> torch.Tensor:
    """[B, H*W, C] bf16 tokens -> fp32 NCHW map (MS:475). Pure layout/dtype plumbing for API consumers."""
    B, _, C = tokens.shape
    return tokens.float().reshape(B, H, W, C).permute(0, 3, 1, 2).contiguous()


class SegformerEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        total = sum(config.depths)
        drop_path_decays = [x.item() for x in torch.linspace(0, config.drop_path_rate, total)]
        self.patch_embeddings = nn.ModuleList([
            SegformerOverlapPatchEmbeddings(patch_size=config.patch_sizes[i], stride=config.strides[i],
                                            num_channels=config.num_channels if i == 0 else config.hidden_sizes[i - 1],
                                            hidden_size=config.hidden_sizes[i])
            for i in range(config.num_encoder_blocks)])
        blocks = []
        cur = 0
        # Python lists assigned from outside by the wrapper (SFM:83-99), exactly as in MS:391-395
        self.prompt_tokens = []
        self.cls_token = []
        for i in range(config.num_encoder_blocks):
            self.prompt_tokens.append(None)
            self.cls_token.append(None)
            if i != 0:
                cur += config.depths[i - 1]
            blocks.append(nn.ModuleList([
                SegformerLayer(config, hidden_size=config.hidden_sizes[i], num_attention_heads=config.num_attention_heads[i],
                               drop_path=drop_path_decays[cur + j], sequence_reduction_ratio=config.sr_ratios[i],
                               mlp_ratio=config.mlp_ratios[i])
                for j in range(config.depths[i])]))
        self.block = nn.ModuleList(blocks)
        self.layer_norm = nn.ModuleList([nn.LayerNorm(config.hidden_sizes[i]) for i in range(config.num_encoder_blocks)])
        self._anchor = _GradAnchor()
        self._ddp_bucketer = None        # set by ddp.attach(): overlaps the gradient all-reduce with backward
        self._ddp_stage_offsets = None

    def _refresh_layouts(self):
        """After an optimiser / EMA step only libsfk's weight epoch has moved: the Linear operands are live views of the flat bf16
        shadow, and what is left to redo are the layout permutations (SR-conv weights OIHW -> [O,(ky,kx,c)] bf16, depthwise
        weights -> tap-major).  One batched launch for the whole encoder instead of two tiny launches per layer."""
        layers = [layer for blk in self.block for layer in blk]
        pending = []
        for layer in layers:
            key = _versions(layer._params())
            if layer._sh_key == key:
                continue
            sh = layer._sh
            if sh is None or not sh.get("flat", False) or layer._sh_key[-1] != key[-1] or layer._sh_key[:-2] != key[:-2]:
                return   # first use or weights written through PyTorch: the per-layer path rebuilds everything
            pending.append((layer, key))
        if not pending:
            return
        entries = []
        for layer in layers:
            a, m, sh = layer.attention.self, layer.mlp, layer._sh
            C = a.hidden_size
            if a.sr_ratio > 1:
                sr = a.sr_ratio
                entries.append((a.sr.weight.data_ptr(), sh["wsr"].data_ptr(), 0, C, C, sr, sr, 0, sr * sr * C))
            entries.append((m.dwconv.dwconv.weight.data_ptr(), sh["dw_wt"].data_ptr(), 1, m.dense1.out_features, 0, 0, 0, 0, 0))
        assert all(e[3] * max(e[4], 1) * max(e[5], 1) * max(e[6], 1) < 2 ** 31 for e in entries)   # the kernel indexes in 32 bits
        sig = tuple(entries)
        cached = getattr(self, "_layout_table", None)
        if cached is None or cached[0] != sig:
            import numpy as np
            dt = np.dtype([("src", "<u8"), ("dst", "<u8"), ("kind", "<i4"), ("O", "<i4"), ("I", "<i4"), ("KH", "<i4"), ("KW", "<i4"),
                           ("pad", "<i4"), ("ld", "<i8")])   # sfk_layout_entry (include/sfk.h)
            arr = np.array(entries, dtype=dt)
            table = torch.from_numpy(arr.view(np.uint8).copy()).to(layers[0].layer_norm_1.weight.device)
            cached = (sig, table)
            self._layout_table = cached
        ops.call("sfk_weight_layouts_batched", cached[1].data_ptr(), len(entries), ops.stream())
        for layer, key in pending:
            layer._sh_key = key

    def _forward_tokens(self, pixel_values):
        """Fast path: returns ([(tokens bf16 [B,N_i,C_i], H_i, W_i)] per stage, cls list | None)."""
        if pixel_values.dim() != 4:
            raise ValueError("pixel_values must be [B, C, H, W]")
        self._refresh_layouts()
        B = pixel_values.shape[0]
        x_in, is_image = pixel_values.float().contiguous(), True
        Hin, Win = pixel_values.shape[2], pixel_values.shape[3]
        stages, block_cls = [], []
        all_layers = [layer for blk in self.block for layer in blk]
        dp_scales = drop_path_scales_batched(B, [l.drop_prob() for l in all_layers], self.training, pixel_values.device)
        layer_no = 0
        for idx, (emb, blk, norm) in enumerate(zip(self.patch_embeddings, self.block, self.layer_norm)):
            hidden, H, W = emb._forward_tokens(x_in, is_image, Hin, Win)
            if self._ddp_bucketer is not None and hidden.requires_grad:
                # when this gradient exists, every gradient of this stage's blocks and of everything after it is final
                bucketer, off = self._ddp_bucketer, self._ddp_stage_offsets[idx]
                hidden.register_hook(lambda g, b=bucketer, o=off: b.notify(o))
            block_cls.append(None)
            running_cls = None
            link = None   # cross-layer backward hand-over (see _BwdLink); restarts at every stage
            pt = self.prompt_tokens[idx]
            ct = self.cls_token[idx]
            for i, layer in enumerate(blk):
                front_parts = []
                if ct is not None:  # token order is [cls | prompt | spatial] (MS:452,459)
                    if i == 0:
                        running_cls = ct.to(hidden.device).float().unsqueeze(0).expand(B, -1, -1)
                    front_parts.append(running_cls)
                if pt is not None:
                    rows = pt[i] if pt.dim() == 3 else pt
                    front_parts.append(rows.to(hidden.device).float().unsqueeze(0).expand(B, -1, -1))
                front = torch.cat(front_parts, dim=1).contiguous() if front_parts else None
                hidden, front_out = layer._forward_split(hidden, front, H, W, dp_scales[layer_no], link)
                link = layer._link_out
                layer_no += 1
                if ct is not None:
                    running_cls = front_out[:, 0:1, :]  # MS:463-465
                    block_cls[-1] = running_cls
            anchor = self._anchor.get([norm.weight, norm.bias], hidden.device)
            tokens = _StageNormFn.apply(hidden, anchor, norm)
            stages.append((tokens, H, W))
            x_in, is_image, Hin, Win = tokens, False, H, W
        cls_out = None if any(c is None for c in block_cls) else block_cls
        return stages, cls_out

    def forward(self, pixel_values, output_attentions=False, output_hidden_states=False, return_dict=True):
        if output_attentions:
            raise NotImplementedError("output_attentions=True is not supported by the fused attention kernel")
        with ops.device_of(pixel_values):
            stages, cls_out = self._forward_tokens(pixel_values)
        maps = tuple(_tokens_to_nchw(t, H, W) for (t, H, W) in stages)
        last = maps[-1]
        if not self.config.reshape_last_stage:
            t, H, W = stages[-1]
            last = t.float()
            maps = maps[:-1] + (last,)
        all_hidden = maps if output_hidden_states else None
        if not return_dict:
            return tuple(v for v in [last, all_hidden, None] if v is not None)
        out = BaseModelOutput(last_hidden_state=last, hidden_states=all_hidden, attentions=None)
        out._token_stages = stages
        return out, cls_out


def _init_weights(module, config):
    """MS:500-514."""
    if isinstance(module, (nn.Linear, nn.Conv2d)):
        module.weight.data.normal_(mean=0.0, std=config.initializer_range)
        if module.bias is not None:
            module.bias.data.zero_()
    elif isinstance(module, nn.LayerNorm):
        module.bias.data.zero_()
        module.weight.data.fill_(1.0)


class SegformerPreTrainedModel(nn.Module):
    config_class = SegformerConfig
    base_model_prefix = "segformer"
    main_input_name = "pixel_values"

    def __init__(self, config):
        super().__init__()
        self.config = config

    def post_init(self):
        self.apply(lambda m: _init_weights(m, self.config))

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, *model_args, ignore_mismatched_sizes=False, **kwargs):
        """Offline stand-in for PreTrainedModel.from_pretrained (called at SFM:14-16).  The architecture comes from the MiT
        variant named in the argument.  Weights are taken, in this order, from
          * the argument itself when it is a state-dict file or a Hugging Face snapshot directory
            (pytorch_model.bin / model.safetensors),
          * $SFK_PRETRAINED_DIR/<name> or $SFK_PRETRAINED_DIR/<basename of name> (same formats),
        loaded with strict=False (mismatched shapes skipped when ignore_mismatched_sizes=True, like the reference's call).
        Without either the network keeps the reference's random init (MS:500-514) and a warning says so: there is no
        network access to fetch nvidia/mit-b5."""
        import warnings
        cfg_kwargs = {k: v for k, v in kwargs.items() if k in ("num_labels", "reshape_last_stage", "drop_path_rate",
                                                               "classifier_dropout_prob", "decoder_hidden_size")}
        name = str(pretrained_model_name_or_path)
        source = _find_checkpoint(name)
        variant = kwargs.get("variant")
        if variant is None:
            try:
                SegformerConfig.from_variant(name)
                variant = name
            except ValueError:
                if source is None:
                    raise
                variant = "b5"   # the reference only ever loads MiT-b5 (SFM:14)
        config = SegformerConfig.from_variant(variant, **cfg_kwargs)
        model = cls(config, *model_args)
        if source is None:
            warnings.warn(f"from_pretrained({name!r}): no local checkpoint found (set SFK_PRETRAINED_DIR or pass a state-dict / "
                          "snapshot path); the model keeps its RANDOM initialisation, unlike the reference which starts from "
                          "the ImageNet-pretrained encoder", RuntimeWarning, stacklevel=2)
            return model
        state = _normalise_checkpoint_keys(_read_checkpoint(source), model)
        own = model.state_dict()
        if ignore_mismatched_sizes:
            state = {k: v for k, v in state.items() if k not in own or own[k].shape == v.shape}
        missing, unexpected = model.load_state_dict(state, strict=False)
        loaded = len([k for k in state if k in own])
        if loaded == 0:
            raise RuntimeError(f"from_pretrained: {source} shares no parameter name with {cls.__name__}")
        print(f"from_pretrained: loaded {loaded} tensors from {source} ({len(missing)} missing, {len(unexpected)} unexpected)")
        return model


def _find_checkpoint(name: str):
    """Path of a local checkpoint for `name`, or None."""
    cands = [name]
    root = os.environ.get("SFK_PRETRAINED_DIR")
    if root:
        cands += [os.path.join(root, name), os.path.join(root, os.path.basename(name.rstrip("/")))]
    for c in cands:
        if os.path.isfile(c):
            return c
        if os.path.isdir(c):
            for fn in ("pytorch_model.bin", "model.safetensors", "model.pth", "model.pt"):
                if os.path.isfile(os.path.join(c, fn)):
                    return os.path.join(c, fn)
    return None


def _read_checkpoint(path: str) -> dict:
    if path.endswith(".safetensors"):
        try:
            from safetensors.torch import load_file
        except ImportError as e:   # no silent fallback: say what is missing
            raise RuntimeError(f"{path}: reading .safetensors needs the `safetensors` package") from e
        return load_file(path)
    state = torch.load(path, map_location="cpu")
    if isinstance(state, dict) and "state_dict" in state and isinstance(state["state_dict"], dict):
        state = state["state_dict"]
    return state


def _normalise_checkpoint_keys(state: dict, model: nn.Module) -> dict:
    """Map Hugging Face / reference-wrapper key spellings onto this module's: HF SegFormer checkpoints use
    'segformer.encoder.*' / 'decode_head.*' (identical to ours); encoder-only dumps use 'encoder.*'; the reference wrapper
    saves 'model.segformer.*' (SFM:22-29)."""
    own = set(model.state_dict().keys())
    out = {}
    for k, v in state.items():
        for cand in (k, k[len("model."):] if k.startswith("model.") else None, "segformer." + k if k.startswith("encoder.") else None):
            if cand is not None and cand in own:
                out[cand] = v
                break
        else:
            out[k] = v
    return out


class SegformerModel(SegformerPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.encoder = SegformerEncoder(config)
        self.post_init()

    def forward(self, pixel_values, output_attentions=None, output_hidden_states=None, return_dict=None):
        output_attentions = output_attentions if output_attentions is not None else self.config.output_attentions
        output_hidden_states = output_hidden_states if output_hidden_states is not None else self.config.output_hidden_states
        return_dict = return_dict if return_dict is not None else self.config.use_return_dict
        res = self.encoder(pixel_values, output_attentions=output_attentions, output_hidden_states=output_hidden_states,
                           return_dict=return_dict)
        if not return_dict:
            return res
        return res


class SegformerForImageClassification(SegformerPreTrainedModel):
    """Exported by the reference package (MS:615-703) but used by none of its scripts; not part of the hot path."""

    def __init__(self, config):
        super().__init__(config)
        raise NotImplementedError("SegformerForImageClassification is outside the teacher-student segmentation path")


# ---------------------------------------------------------------------------------------------------------------
# decode head (MS:706-782)
# ---------------------------------------------------------------------------------------------------------------
class SegformerMLP(nn.Module):
    """Linear Embedding (MS:706-724)."""

    def __init__(self, config: SegformerConfig, input_dim, cls_dim=None):
        super().__init__()
        self.proj = nn.Linear(input_dim, config.decoder_hidden_size)
        if cls_dim is not None:
            self.cls_proj = nn.Linear(cls_dim, config.decoder_hidden_size)

    def forward(self, hidden_states: torch.Tensor, cls_token=None):
        raise NotImplementedError("use SegformerDecodeHead (fused libsfk path); there is no PyTorch fallback")


# SURVEY 7 algebraic shortcut for the decode head (default on; SFK_DECODE_FOLD=0 runs the reference's literal sequence)
DECODE_FOLD = os.environ.get("SFK_DECODE_FOLD", "1") != "0"


class _DecodeHeadFn(torch.autograd.Function):
    """linear_c x4 -> bilinear upsample + concat (c4,c3,c2,c1) -> 1x1 fuse GEMM -> BN -> ReLU -> Dropout -> classifier."""

    @staticmethod
    def forward(ctx, t0, t1, t2, t3, cls_vecs, anchor, head, dims_hw, training):
        toks = [ops._chk(t, torch.bfloat16, "encoder tokens") for t in (t0, t1, t2, t3)]
        B = toks[0].shape[0]
        dec = head.config.decoder_hidden_size
        NL = head.config.num_labels
        (H0, W0) = dims_hw[0]
        M = B * H0 * W0
        sh = head._shadow()
        dev = toks[0].device
        fold = DECODE_FOLD and cls_vecs is None
        folded = None
        if fold:
            # SURVEY 7 algebraic shortcut: fuse(concat(up(linear_c_i(x_i)))) == sum_i up((W_f,i W_c,i) x_i) + sum_i W_f,i b_c,i
            # (bilinear weights sum to one, so biases pass through the upsampling).  Executes 25 GF instead of 1 237 GF per
            # 16-tile batch and never materialises the [M, 4*dec] concat buffer.
            wf16 = sh["wf"]
            wf32 = head.linear_fuse.weight.detach().view(dec, 4 * dec)
            folded, z = [], []
            bias_f = torch.zeros(dec, device=dev, dtype=torch.float32)
            for i, t in enumerate(toks):
                Hi, Wi = dims_hw[i]
                Ci = t.shape[2]
                blk = slice((3 - i) * dec, (4 - i) * dec)
                folded.append(ops.gemm(wf16[:, blk], sh["wc"][i], dec, Ci, dec, lda=4 * dec, ldb=Ci, b_mn=1, out_bf16=True))
                bias_f += torch.mv(wf32[:, blk], head.linear_c[i].proj.bias.detach())   # [dec]-sized host-API op
                z.append(ops.gemm(t.view(B * Hi * Wi, Ci), folded[i], B * Hi * Wi, dec, Ci, out_bf16=True))
            fused = torch.empty(M, dec, device=dev, dtype=torch.bfloat16)
            ops.call("sfk_upsample_add4", z[0].data_ptr(), z[1].data_ptr(), dims_hw[1][0], dims_hw[1][1], z[2].data_ptr(),
                     dims_hw[2][0], dims_hw[2][1], z[3].data_ptr(), dims_hw[3][0], dims_hw[3][1], bias_f.data_ptr(),
                     fused.data_ptr(), B, H0, W0, dec, ops.stream())
            cat = None
        else:
            cat = torch.empty(M, 4 * dec, device=dev, dtype=torch.bfloat16)
            for i, t in enumerate(toks):
                Hi, Wi = dims_hw[i]
                Ci = t.shape[2]
                proj = ops.gemm(t.view(B * Hi * Wi, Ci), sh["wc"][i], B * Hi * Wi, dec, Ci, bias=head.linear_c[i].proj.bias, out_bf16=True)
                if cls_vecs is not None:
                    ops.call("sfk_add_sample_bias", proj.data_ptr(), cls_vecs[i].contiguous().data_ptr(), B * Hi * Wi, Hi * Wi, dec, ops.stream())
                ops.call("sfk_upsample_concat", proj.data_ptr(), cat.data_ptr(), B, Hi, Wi, dec, H0, W0, 4 * dec, (3 - i) * dec, ops.stream())
            fused = ops.gemm(cat, sh["wf"], M, dec, 4 * dec, out_bf16=True)
        bn = head.batch_norm
        mean = torch.empty(dec, device=dev, dtype=torch.float32)
        rstd = torch.empty(dec, device=dev, dtype=torch.float32)
        if training:
            lib = ops.lib()
            wsb = lib.sfk_bn_stats_workspace_bytes(M, dec)
            ws = ops.scratch(wsb, dev, "bn")
            ops.call("sfk_bn_batch_stats", fused.data_ptr(), M, dec, bn.eps, bn.momentum, mean.data_ptr(), rstd.data_ptr(),
                     bn.running_mean.data_ptr(), bn.running_var.data_ptr(), ws.data_ptr(), ws.numel(), ops.stream())
            bn.num_batches_tracked += 1
        else:
            ops.call("sfk_bn_eval_stats", bn.running_mean.data_ptr(), bn.running_var.data_ptr(), bn.eps, mean.data_ptr(),
                     rstd.data_ptr(), dec, ops.stream())
        drop_p = float(head.dropout.p) if training else 0.0
        seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if drop_p > 0 else 0
        logits = torch.empty(B, NL, H0, W0, device=dev, dtype=torch.float32)
        wc = head.classifier.weight
        ops.call("sfk_bn_relu_cls_fwd", fused.data_ptr(), mean.data_ptr(), rstd.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(),
                 wc.data_ptr(), head.classifier.bias.data_ptr(), logits.data_ptr(), M, dec, NL, H0 * W0, drop_p, seed, ops.stream())
        if any(ctx.needs_input_grad):
            ctx.head, ctx.sh, ctx.meta = head, sh, (B, dec, NL, dims_hw, M, drop_p, seed, training, cls_vecs is not None)
            ctx.folded = folded
            ctx.save_for_backward(cat if cat is not None else fused.new_empty(0), fused, mean, rstd, *toks)
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        head, sh = ctx.head, ctx.sh
        B, dec, NL, dims_hw, M, drop_p, seed, training, has_cls = ctx.meta
        cat, fused, mean, rstd, *toks = ctx.saved_tensors
        dev = cat.device
        dlogits = ops._chk(dlogits, torch.float32, "logits grad")
        bn, cls = head.batch_norm, head.classifier
        (H0, W0) = dims_hw[0]
        sums = torch.empty(2, dec, device=dev, dtype=torch.float32)
        gwc, gbc = ops.grad_buffer(cls.weight), ops.grad_buffer(cls.bias)
        if gwc is None:
            gwc = torch.zeros_like(cls.weight)
        if gbc is None:
            gbc = torch.zeros_like(cls.bias)
        dfused = torch.empty(M, dec, device=dev, dtype=torch.bfloat16)
        ops.call("sfk_bn_relu_cls_bwd", fused.data_ptr(), mean.data_ptr(), rstd.data_ptr(), bn.weight.data_ptr(), bn.bias.data_ptr(),
                 cls.weight.data_ptr(), dlogits.data_ptr(), sums.data_ptr(), gwc.data_ptr(), gbc.data_ptr(), dfused.data_ptr(), M, dec,
                 NL, H0 * W0, drop_p, seed, int(training), ops.stream())
        gb, gg = ops.grad_buffer(bn.bias), ops.grad_buffer(bn.weight)
        if gb is not None:
            ops.colsum_into(sums[0], gb, 1, dec)
        if gg is not None:
            ops.colsum_into(sums[1], gg, 1, dec)
        gwf = ops.grad_buffer(head.linear_fuse.weight)
        if ctx.folded is not None:
            return _DecodeHeadFn._backward_folded(ctx, dfused, gwf, toks)
        if gwf is not None:
            ops.gemm(dfused, cat, dec, 4 * dec, M, lda=dec, ldb=4 * dec, a_mn=1, b_mn=1, out=gwf.view(dec, 4 * dec), out_bf16=False,
                     atomic_add=True)
        dcat = ops.gemm(dfused, sh["wf"], M, 4 * dec, dec, lda=dec, ldb=4 * dec, b_mn=1, out_bf16=True)
        dtoks = []
        dcls = torch.zeros(4, B, dec, device=dev, dtype=torch.float32) if has_cls else None
        for i, t in enumerate(toks):
            Hi, Wi = dims_hw[i]
            Ci = t.shape[2]
            rows = B * Hi * Wi
            dproj = torch.empty(rows, dec, device=dev, dtype=torch.bfloat16)
            ops.call("sfk_upsample_concat_bwd", dcat.data_ptr(), dproj.data_ptr(), B, Hi, Wi, dec, H0, W0, 4 * dec, (3 - i) * dec,
                     ops.stream())
            lin = head.linear_c[i].proj
            gw, gbi = ops.grad_buffer(lin.weight), ops.grad_buffer(lin.bias)
            if gw is not None:
                ops.gemm(dproj, t.view(rows, Ci), dec, Ci, rows, lda=dec, ldb=Ci, a_mn=1, b_mn=1, out=gw, out_bf16=False, atomic_add=True)
            if gbi is not None:
                ops.colsum_into(dproj, gbi, rows, dec)
            if has_cls:
                for b in range(B):
                    ops.colsum_into(dproj[b * Hi * Wi:(b + 1) * Hi * Wi], dcls[i, b], Hi * Wi, dec)
            if ctx.needs_input_grad[i]:
                dtoks.append(ops.gemm(dproj, sh["wc"][i], rows, Ci, dec, lda=dec, ldb=Ci, b_mn=1, out_bf16=True).view(B, Hi * Wi, Ci))
            else:
                dtoks.append(None)
        return dtoks[0], dtoks[1], dtoks[2], dtoks[3], dcls, None, None, None, None

    @staticmethod
    def _backward_folded(ctx, dfused, gwf, toks):
        """Backward of the folded head: dz_0 = dfused, dz_i = up^T(dfused); per stage dx_i = dz_i W'_i, dW'_i = dz_i^T x_i,
        then the chain rule through W'_i = W_f,i W_c,i and b' = sum_i W_f,i b_c,i onto the reference's parameters."""
        head, sh = ctx.head, ctx.sh
        B, dec, NL, dims_hw, M, drop_p, seed, training, has_cls = ctx.meta
        dev = dfused.device
        (H0, W0) = dims_hw[0]
        wf16 = sh["wf"]
        wf32 = head.linear_fuse.weight.detach().view(dec, 4 * dec)
        dbias = torch.zeros(dec, device=dev, dtype=torch.float32)    # gradient of the folded bias b'
        ops.colsum_into(dfused, dbias, M, dec)
        gwf2 = gwf.view(dec, 4 * dec) if gwf is not None else None
        dtoks = []
        for i, t in enumerate(toks):
            Hi, Wi = dims_hw[i]
            Ci = t.shape[2]
            rows = B * Hi * Wi
            blk = slice((3 - i) * dec, (4 - i) * dec)
            if i == 0:
                dz = dfused
            else:
                dz = torch.empty(rows, dec, device=dev, dtype=torch.bfloat16)
                ops.call("sfk_upsample_concat_bwd", dfused.data_ptr(), dz.data_ptr(), B, Hi, Wi, dec, H0, W0, dec, 0, ops.stream())
            lin = head.linear_c[i].proj
            gw, gbi = ops.grad_buffer(lin.weight), ops.grad_buffer(lin.bias)
            if gw is not None or gwf2 is not None:
                dwp = torch.zeros(dec, Ci, device=dev, dtype=torch.float32)      # dW'_i
                ops.gemm(dz, t.view(rows, Ci), dec, Ci, rows, lda=dec, ldb=Ci, a_mn=1, b_mn=1, out=dwp, out_bf16=False, atomic_add=True)
                dwp16 = ops.cast_bf16(dwp)
                if gwf2 is not None:
                    # dW_f,i += dW'_i W_c,i^T + db' (x) b_c,i
                    ops.gemm(dwp16, sh["wc"][i], dec, dec, Ci, out=gwf2[:, blk], ldo=4 * dec, out_bf16=False, atomic_add=True)
                    gwf2[:, blk].addr_(dbias, lin.bias.detach())
                if gw is not None:
                    # dW_c,i += W_f,i^T dW'_i
                    ops.gemm(wf16[:, blk], dwp16, dec, Ci, dec, lda=4 * dec, ldb=Ci, a_mn=1, b_mn=1, out=gw, out_bf16=False, atomic_add=True)
            if gbi is not None:
                gbi += torch.mv(wf32[:, blk].t(), dbias)                            # db_c,i += W_f,i^T db'
            if ctx.needs_input_grad[i]:
                dtoks.append(ops.gemm(dz, ctx.folded[i], rows, Ci, dec, lda=dec, ldb=Ci, b_mn=1, out_bf16=True).view(B, Hi * Wi, Ci))
            else:
                dtoks.append(None)
        return dtoks[0], dtoks[1], dtoks[2], dtoks[3], None, None, None, None, None


class SegformerDecodeHead(SegformerPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.linear_c = nn.ModuleList([SegformerMLP(config, input_dim=config.hidden_sizes[i], cls_dim=512)
                                       for i in range(config.num_encoder_blocks)])
        self.linear_fuse = nn.Conv2d(in_channels=config.decoder_hidden_size * config.num_encoder_blocks,
                                     out_channels=config.decoder_hidden_size, kernel_size=1, bias=False)
        self.batch_norm = nn.BatchNorm2d(config.decoder_hidden_size)
        self.activation = nn.ReLU()
        self.dropout = nn.Dropout(config.classifier_dropout_prob)
        self.classifier = nn.Conv2d(config.decoder_hidden_size, config.num_labels, kernel_size=1)
        self.config = config
        self._anchor = _GradAnchor()
        self._sh, self._sh_key = None, None
        if config.num_encoder_blocks != 4:
            raise NotImplementedError("the fused decode head is written for the 4-stage MiT encoders")

    def _shadow(self):
        params = [m.proj.weight for m in self.linear_c] + [self.linear_fuse.weight]
        key = _versions(params)
        if self._sh_key != key:
            dec = self.config.decoder_hidden_size
            self._sh = {"wc": [ops.weight_bf16(m.proj.weight) for m in self.linear_c],
                        "wf": ops.weight_bf16(self.linear_fuse.weight)}
            self._sh_key = key
        return self._sh

    def _forward_tokens(self, stages, cls_token=None):
        toks = [s[0] for s in stages]
        dims_hw = tuple((s[1], s[2]) for s in stages)
        cls_vecs = None
        if cls_token is not None:
            # cls_proj (MS:721): a [B,1,512] x [512,dec] product per stage — 4 tiny matrix-vector products on the host API
            cls_vecs = torch.stack([torch.nn.functional.linear(cls_token.float(), m.cls_proj.weight, m.cls_proj.bias)[:, 0, :]
                                    for m in self.linear_c])
        own = [p for p in self.parameters()]
        anchor = self._anchor.get(own, toks[0].device)
        return _DecodeHeadFn.apply(toks[0], toks[1], toks[2], toks[3], cls_vecs, anchor, self, dims_hw, self.training)

    def forward(self, encoder_hidden_states, cls_token=None) -> torch.Tensor:
        """encoder_hidden_states: 4 NCHW maps (or [B,N,C] for the last one when reshape_last_stage is False)."""
        batch_size = encoder_hidden_states[-1].shape[0]
        stages = []
        for h in encoder_hidden_states:
            if h.dim() == 3:  # MS:757-761
                side = int(math.sqrt(h.shape[1]))
                h = h.reshape(batch_size, side, side, -1).permute(0, 3, 1, 2)
            B, C, H, W = h.shape
            stages.append((h.permute(0, 2, 3, 1).reshape(B, H * W, C).to(torch.bfloat16).contiguous(), H, W))
        with ops.device_of(stages[0][0]):
            return self._forward_tokens(stages, cls_token)


# ---------------------------------------------------------------------------------------------------------------
# in-model losses (MS:853-868)
# ---------------------------------------------------------------------------------------------------------------
class _UpsampleFn(torch.autograd.Function):
    """F.interpolate(mode='bilinear', align_corners=False) on NCHW fp32, optional fused sigmoid (SFM:117-124)."""

    @staticmethod
    def forward(ctx, x, H, W, act):
        x = ops._chk(x, torch.float32, "logits")
        B, C, h, w = x.shape
        y = torch.empty(B, C, H, W, device=x.device, dtype=torch.float32)
        ops.call("sfk_upsample_nchw_fwd", x.data_ptr(), y.data_ptr(), B * C, h, w, H, W, act, ops.stream())
        ctx.meta = (B, C, h, w, H, W, act)
        if act and ctx.needs_input_grad[0]:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, dy):
        B, C, h, w, H, W, act = ctx.meta
        dy = ops._chk(dy, torch.float32, "probs grad")
        yact = ctx.saved_tensors[0] if act else None
        dx = torch.empty(B, C, h, w, device=dy.device, dtype=torch.float32)
        ops.call("sfk_upsample_nchw_bwd", dy.data_ptr(), ops._p(yact), dx.data_ptr(), B * C, h, w, H, W, 0, ops.stream())
        return dx, None, None, None


def upsample_bilinear(x: torch.Tensor, size, sigmoid: bool = False) -> torch.Tensor:
    return _UpsampleFn.apply(x, int(size[0]), int(size[1]), 1 if sigmoid else 0)


class _BceValidFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labels, ignore_index):
        z = ops._chk(z, torch.float32, "upsampled logits")
        labels = ops._chk(labels, torch.int64, "labels")
        n = z.numel()
        lib = ops.lib()
        ws = ops.scratch(lib.sfk_loss_workspace_bytes(), z.device, "loss")
        out = torch.empty(2, device=z.device, dtype=torch.float32)
        ops.call("sfk_bce_valid_fwd", z.data_ptr(), labels.data_ptr(), ignore_index, n, ws.data_ptr(), out.data_ptr(), ops.stream())
        ctx.ignore = ignore_index
        ctx.save_for_backward(z, labels)
        return out[0]

    @staticmethod
    def backward(ctx, dloss):
        z, labels = ctx.saved_tensors
        dz = torch.empty_like(z)
        dl = dloss.reshape(1).float().contiguous()
        ops.call("sfk_bce_valid_bwd", z.data_ptr(), labels.data_ptr(), ctx.ignore, dl.data_ptr(), dz.data_ptr(), z.numel(), ops.stream())
        return dz, None, None


class _CeIgnoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, labels, ignore_index):
        z = ops._chk(z, torch.float32, "upsampled logits")
        labels = ops._chk(labels, torch.int64, "labels")
        N, Lc = z.shape[0], z.shape[1]
        P = z.numel() // (N * Lc)
        lib = ops.lib()
        ws = ops.scratch(lib.sfk_loss_workspace_bytes(), z.device, "loss")
        out = torch.empty(2, device=z.device, dtype=torch.float32)
        ops.call("sfk_ce_ignore_fwd", z.data_ptr(), labels.data_ptr(), ignore_index, N, Lc, P, ws.data_ptr(), out.data_ptr(), ops.stream())
        ctx.meta = (ignore_index, N, Lc, P)
        ctx.save_for_backward(z, labels, out)
        return out[0].clone()

    @staticmethod
    def backward(ctx, dloss):
        z, labels, out = ctx.saved_tensors
        ignore_index, N, Lc, P = ctx.meta
        dz = torch.empty_like(z)
        dl = dloss.reshape(1).float().contiguous()
        ops.call("sfk_ce_ignore_bwd", z.data_ptr(), labels.data_ptr(), ignore_index, out.data_ptr(), dl.data_ptr(), dz.data_ptr(), N, Lc,
                 P, ops.stream())
        return dz, None, None


class SegformerForSemanticSegmentation(SegformerPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.segformer = SegformerModel(config)
        self.decode_head = SegformerDecodeHead(config)
        self.post_init()

    def _forward_impl(self, pixel_values, labels):
        enc = self.segformer.encoder
        stages, cls_list = enc._forward_tokens(pixel_values)
        cls_token = None
        if cls_list is not None:
            cls_token = torch.sigmoid(cls_list[-1])  # MS:848-850, [B,1,C4]
        logits = self.decode_head._forward_tokens(stages, cls_token)
        loss = None
        if labels is not None:
            up = upsample_bilinear(logits, labels.shape[-2:])
            if self.config.num_labels > 1:
                loss = _CeIgnoreFn.apply(up, labels, int(self.config.semantic_loss_ignore_index))
            elif self.config.num_labels == 1:
                loss = _BceValidFn.apply(up.squeeze(1), labels, int(self.config.semantic_loss_ignore_index))
            else:
                raise ValueError(f"Number of labels should be >=0: {self.config.num_labels}")
        return logits, loss, stages, cls_token

    def forward(self, pixel_values, labels=None, output_attentions=None, output_hidden_states=None, return_dict=None):
        """Returns (SemanticSegmenterOutput, cls_token) exactly like the reference's modified forward (MS:800-882)."""
        return_dict = return_dict if return_dict is not None else self.config.use_return_dict
        output_hidden_states = output_hidden_states if output_hidden_states is not None else self.config.output_hidden_states
        if output_attentions:
            raise NotImplementedError("output_attentions=True is not supported by the fused attention kernel")
        if not return_dict:
            raise NotImplementedError("return_dict=False is not used by any reference script and is not supported")
        with ops.device_of(pixel_values):
            logits, loss, stages, cls_token = self._forward_impl(pixel_values, labels)
        hidden = (lambda: tuple(_tokens_to_nchw(t, H, W) for (t, H, W) in stages)) if output_hidden_states else None
        return SemanticSegmenterOutput(loss=loss, logits=logits, hidden_states=hidden, attentions=None), cls_token
