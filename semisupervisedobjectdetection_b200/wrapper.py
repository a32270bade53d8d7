"""models/SegFormerModel.py of the reference (SFM:9-219) on the libsfk path.

Same class name, constructor arguments, methods, return arities and attribute names (.model .optimizer .scheduler .device
.loss_function .loss_argmax_function .mse_loss_function .activation_fn .num_labels), so the reference's
segFormer_main.py / segFormer_transfer_learning_main.py / segFormer_semi_teacherstudent_main.py run against it unchanged.
The implementation is organised around three private helpers (`_run_model`, `_probabilities`, `_loss_for`) instead of the
reference's inline branches; behavioural quirks of the reference are kept on purpose and cited (SURVEY.md §0 M-list).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import loss as myLoss
from . import ops
from .modeling import SegformerForSemanticSegmentation, upsample_bilinear
from .optim import FusedClipAdam

_CHECKPOINT_DIR = "../checkpoints"                      # SFM:21: paths are relative to main_segformer/
_CLASSIFIER_KEYS = ("model.decode_head.classifier.weight", "model.decode_head.classifier.bias")


class SegFormerModel(nn.Module):
    def __init__(self, pretrain_weight=None, lr=None, weight_decay=None, scheduler=None, device="cuda:0", use_dice_loss=True,
                 num_labels=1, *args, variant="nvidia/mit-b5", **kwargs):
        super().__init__(*args, **kwargs)
        self.device, self.num_labels, self.use_dice_loss = device, num_labels, use_dice_loss
        # SFM:14-16 downloads nvidia/mit-b5; offline this builds the same architecture with the reference's random init
        net = SegformerForSemanticSegmentation.from_pretrained(variant, ignore_mismatched_sizes=True, num_labels=num_labels,
                                                               reshape_last_stage=True)
        net.config.output_hidden_states = True
        self.model = net.to(device)
        if pretrain_weight:
            self._load_checkpoint(pretrain_weight)
        # Adam(betas=(0.5, 0.999), L2 weight decay) and clip_grad_value_(1.2) run as one fused kernel (SFM:32-35,161-162);
        # only parameters trainable at construction time are optimised, like the reference's filter().
        trainable = [p for p in self.model.parameters() if p.requires_grad is not False]
        self.optimizer = FusedClipAdam(trainable, lr=lr, weight_decay=weight_decay, betas=(0.5, 0.999), clip_value=1.2,
                                       model=self.model)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma=scheduler)
        # SFM:38-42: all three losses look at channel 0 only (num_classes = 1)
        for attr, kind, extra in (("loss_argmax_function", "dice_argmax", {"activation": "none"}),
                                  ("loss_function", "dice", {"activation": "none"}),
                                  ("mse_loss_function", "mse", {})):
            setattr(self, attr, myLoss.SegmentationLoss(1, loss_type=kind, **extra))
        self.activation_fn = nn.Sigmoid()

    def _load_checkpoint(self, name):
        state = torch.load(os.path.join(_CHECKPOINT_DIR, name), map_location=torch.device(self.device))
        if self.num_labels == 1:
            # SFM:24-28: a multi-class checkpoint keeps only classifier row 0
            for key in _CLASSIFIER_KEYS:
                state[key] = state[key][0].unsqueeze(dim=0)
        self.load_state_dict(state, strict=False)
        print("Pretrained model loaded")

    # ------------------------------------------------------------------------------------------------ SFM:46-101
    def _encoder_blocks(self):
        return self.model.segformer.encoder.block

    @staticmethod
    def _set_trainable(modules, flag):
        for module in modules:
            for p in module.parameters():
                p.requires_grad = flag

    def frozen_encoder(self, layers_num=None, layers=None):
        """Freeze whole stages: explicit indices in `layers`, else the first `layers_num` (default: all four). Only
        encoder.block is touched — patch embeddings and stage norms stay trainable (SFM:56-59)."""
        blocks = self._encoder_blocks()
        if layers is not None:
            chosen = [blocks[idx] for idx in layers]
        else:
            count = self.model.config.num_encoder_blocks if layers_num is None else layers_num
            chosen = list(blocks[:count])
        self._set_trainable(chosen, False)
        print('encoder is frozen')

    def unfroze_encoder(self):
        self._set_trainable(self._encoder_blocks(), True)

    def _install_tokens(self, slot_name, counts, shape_of):
        """torch.rand(..., requires_grad=True) moved to the device: a non-leaf copy of a CPU leaf, exactly the reference's
        `torch.rand(...).cuda()` (SFM:83-97, SURVEY M6: such tokens receive gradients that no optimiser applies)."""
        slots = getattr(self.model.segformer.encoder, slot_name)
        for stage, n in enumerate(counts):
            if n:
                slots[stage] = torch.rand(*shape_of(stage, n), requires_grad=True).to(self.device)

    def add_prompt_token(self, token_num_per_block=[10, 10, 10, 10], isSamePerLayer=True):
        cfg = self.model.config

        def shape_of(stage, n):
            width = cfg.hidden_sizes[stage]
            return (n, width) if isSamePerLayer else (cfg.depths[stage], n, width)

        self._install_tokens("prompt_tokens", token_num_per_block, shape_of)
        print('prompt tokens are added')

    def add_cls_token(self, token_num_per_block=[1, 1, 1, 1]):
        widths = self.model.config.hidden_sizes
        self._install_tokens("cls_token", token_num_per_block, lambda stage, n: (n, widths[stage]))
        print('CLS tokens are added')

    # ------------------------------------------------------------------------------------------------ SFM:103-139
    def _run_model(self, images, labels=None):
        """-> (SemanticSegmenterOutput, cls_token); labels only when the in-model BCE / CE loss is wanted."""
        if labels is None:
            return self.model(pixel_values=images)
        return self.model(pixel_values=images, labels=labels)

    @staticmethod
    def _probabilities(logits, image_hw, sigmoid=True):
        """Bilinear upsampling of the quarter-resolution logits to the image size fused with the sigmoid (SFM:117-124)."""
        return upsample_bilinear(logits, list(image_hw), sigmoid=sigmoid)

    def _loss_for(self, use_loss, outputs, images, probs, target):
        if use_loss == 'dice':
            return self.loss_function(probs, target)
        if use_loss == 'bce':
            return outputs.loss                                            # computed inside the model (MS:853-868)
        if use_loss == 'mse':
            raw = self._probabilities(outputs.logits, images.shape[2:], sigmoid=False)
            return self.mse_loss_function(images, raw)                     # SFM:133 compares with the PRE-sigmoid map
        return self.loss_argmax_function(probs, target)                    # any other string: hard-thresholded Dice

    def predict(self, img, mask=None, isEval=True, use_loss='dice', output_cls_token=False):
        # the reference accepts any device string (SFM:11); libsfk launches on the current device, so make ours current
        with ops.device_of(self.device):
            return self._predict(img, mask, isEval, use_loss, output_cls_token)

    def _predict(self, img, mask, isEval, use_loss, output_cls_token):
        if not isEval:
            self.model.eval()   # SFM:104-105: the flag is inverted in the reference; kept
        img = img.to(self.device)
        target = None if mask is None else mask.to(self.device, dtype=torch.int64)   # SFM:108: float masks truncate
        outputs, cls_token = self._run_model(img, target if use_loss == 'bce' else None)
        predict_masks = torch.squeeze(self._probabilities(outputs.logits, img.shape[2:]), 1)
        if target is None:
            return predict_masks
        loss = self._loss_for(use_loss, outputs, img, predict_masks, target)
        return (loss, predict_masks, cls_token) if output_cls_token else (loss, predict_masks)

    def eval_one_epoch(self, imgs, masks):
        self.model.eval()
        with torch.no_grad():
            return self.predict(imgs, masks, use_loss='argmax')

    def train_one_epoch(self, imgs, masks, use_loss='dice', output_cls_token=False):
        self.model.train()
        result = self.predict(imgs, masks, isEval=False, use_loss=use_loss, output_cls_token=output_cls_token)
        self.train_from_loss(result[0])
        return result

    def train_from_loss(self, loss):
        """SFM:158-162. zero_grad -> backward unless NaN -> clip_grad_value_(1.2) -> Adam, the last two fused in
        FusedClipAdam.step(); a step with no gradient anywhere touches nothing (the reference under torch 2.0.1)."""
        with ops.device_of(self.device):
            self.optimizer.zero_grad()
            if not torch.isnan(loss):
                loss.backward()
            self.optimizer.step()

    def scheduler_step(self):
        self.scheduler.step()

    def show_mask(self, vis, img, mask, title=""):
        """visdom helper (SFM:166-175): paint the mask into channel 0 of the image and show it."""
        canvas = img.cpu().numpy()
        if mask is not None:
            canvas[0, :, :] = mask.detach().cpu().numpy()
        vis.image(canvas, opts=dict(title=title))

    # ------------------------------------------------------------------------------------------------ SFM:177-219
    def _autoencode(self, imgs):
        """Autoencoder helpers: post-sigmoid reconstruction vs the image (SFM:192-194, 214-216)."""
        with ops.device_of(self.device):
            imgs = imgs.to(self.device)
            outputs, _ = self._run_model(imgs)
            recon = self._probabilities(outputs.logits, imgs.shape[2:])
            return self.mse_loss_function(imgs, recon), recon

    def eval_one_epoch_without_mask(self, imgs):
        self.model.eval()
        with torch.no_grad():
            return self._autoencode(imgs)

    def train_one_epoch_without_mask(self, imgs):
        self.model.train()
        loss, recon = self._autoencode(imgs)
        self.train_from_loss(loss)
        return loss, recon
