"""models/SegFormerModel.py of the reference (SFM:9-219) on the libsfk path.

Same class name, constructor arguments, methods, return arities and attribute names (.model .optimizer .scheduler .device
.loss_function .loss_argmax_function .mse_loss_function .activation_fn .num_labels), so the reference's
segFormer_main.py / segFormer_transfer_learning_main.py / segFormer_semi_teacherstudent_main.py run against it unchanged.
Behavioural quirks of the reference are kept on purpose and cited inline (SURVEY.md §0 M-list).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from . import loss as myLoss
from . import ops
from .modeling import SegformerForSemanticSegmentation, upsample_bilinear
from .optim import FusedClipAdam


class SegFormerModel(nn.Module):
    def __init__(self, pretrain_weight=None, lr=None, weight_decay=None, scheduler=None, device="cuda:0", use_dice_loss=True,
                 num_labels=1, *args, variant="nvidia/mit-b5", **kwargs):
        super().__init__(*args, **kwargs)
        # SFM:14-16 downloads nvidia/mit-b5; offline this builds the same architecture with the reference's random init
        self.model = SegformerForSemanticSegmentation.from_pretrained(variant, ignore_mismatched_sizes=True,
                                                                      num_labels=num_labels, reshape_last_stage=True)
        self.model.config.output_hidden_states = True
        self.model.to(device)
        self.device = device
        if pretrain_weight:
            ctx = torch.load(os.path.join('../checkpoints', pretrain_weight), map_location=torch.device(device))
            if num_labels == 1:  # SFM:24-28: keep row 0 of a multi-class classifier
                ctx['model.decode_head.classifier.weight'] = ctx['model.decode_head.classifier.weight'][0].unsqueeze(dim=0)
                ctx['model.decode_head.classifier.bias'] = ctx['model.decode_head.classifier.bias'][0].unsqueeze(dim=0)
            self.load_state_dict(ctx, strict=False)
            print("Pretrained model loaded")
        # Adam(betas=(0.5,0.999), L2 weight decay) + clip_grad_value_(1.2) fused into one kernel (SFM:32-35,161-162)
        self.optimizer = FusedClipAdam(filter(lambda x: x.requires_grad is not False, self.model.parameters()), lr=lr,
                                       weight_decay=weight_decay, betas=(0.5, 0.999), clip_value=1.2, model=self.model)
        self.scheduler = torch.optim.lr_scheduler.ExponentialLR(self.optimizer, gamma=scheduler)
        self.loss_argmax_function = myLoss.SegmentationLoss(1, loss_type='dice_argmax', activation='none')
        self.loss_function = myLoss.SegmentationLoss(1, loss_type='dice', activation='none')
        self.mse_loss_function = myLoss.SegmentationLoss(1, loss_type='mse')
        self.activation_fn = torch.nn.Sigmoid()
        self.use_dice_loss = use_dice_loss
        self.num_labels = num_labels

    # ---- SFM:46-101 ----
    def frozen_encoder(self, layers_num=None, layers=None):
        if layers_num is None:
            layers_num = self.model.config.num_encoder_blocks
        if layers is not None:
            for layer_id in layers:
                for param in self.model.segformer.encoder.block[layer_id].parameters():
                    param.requires_grad = False
        else:
            for param in self.model.segformer.encoder.block[:layers_num].parameters():
                param.requires_grad = False
        print('encoder is frozen')

    def unfroze_encoder(self):
        for param in self.model.segformer.encoder.block.parameters():
            param.requires_grad = True

    def add_prompt_token(self, token_num_per_block=[10, 10, 10, 10], isSamePerLayer=True):
        cfg = self.model.config
        enc = self.model.segformer.encoder
        for i, _ in enumerate(token_num_per_block):
            if token_num_per_block[i] == 0:
                continue
            shape = ((token_num_per_block[i], cfg.hidden_sizes[i]) if isSamePerLayer
                     else (cfg.depths[i], token_num_per_block[i], cfg.hidden_sizes[i]))
            # SFM:83-90: torch.rand(..., requires_grad=True).cuda() -> a non-leaf CUDA copy of a CPU leaf
            enc.prompt_tokens[i] = torch.rand(*shape, requires_grad=True).to(self.device)
        print('prompt tokens are added')

    def add_cls_token(self, token_num_per_block=[1, 1, 1, 1]):
        cfg = self.model.config
        enc = self.model.segformer.encoder
        for i, _ in enumerate(token_num_per_block):
            if token_num_per_block[i] == 0:
                continue
            enc.cls_token[i] = torch.rand(token_num_per_block[i], cfg.hidden_sizes[i], requires_grad=True).to(self.device)
        print('CLS tokens are added')

    # ---- SFM:103-139 ----
    def predict(self, img, mask=None, isEval=True, use_loss='dice', output_cls_token=False):
        if not isEval:
            self.model.eval()  # SFM:104-105: the flag is inverted in the reference; kept
        img = img.to(self.device)
        if mask is not None:
            mask = mask.to(self.device, dtype=torch.int64)
        if use_loss == 'bce':
            outputs, cls_token = self.model(pixel_values=img, labels=mask)
        else:
            outputs, cls_token = self.model(pixel_values=img)
        logits = outputs.logits
        size = list(img.shape)
        sig = upsample_bilinear(logits, size[2:], sigmoid=True)  # SFM:117-124 fused: bilinear x4 + sigmoid
        predict_masks = torch.squeeze(sig, 1)
        if mask is None:
            return predict_masks
        if use_loss == 'dice':
            loss = self.loss_function(predict_masks, mask)
        elif use_loss == 'bce':
            loss = outputs.loss
        elif use_loss == 'mse':
            loss = self.mse_loss_function(img, upsample_bilinear(logits, size[2:], sigmoid=False))  # SFM:133 pre-sigmoid
        else:
            loss = self.loss_argmax_function(predict_masks, mask)
        if output_cls_token:
            return loss, predict_masks, cls_token
        return loss, predict_masks

    def eval_one_epoch(self, imgs, masks):
        self.model.eval()
        with torch.no_grad():
            return self.predict(imgs, masks, use_loss='argmax')

    def train_one_epoch(self, imgs, masks, use_loss='dice', output_cls_token=False):
        self.model.train()
        if output_cls_token:
            loss, predict_masks, cls_tokens = self.predict(imgs, masks, isEval=False, use_loss=use_loss, output_cls_token=True)
            self.train_from_loss(loss)
            return loss, predict_masks, cls_tokens
        loss, predict_masks = self.predict(imgs, masks, isEval=False, use_loss=use_loss, output_cls_token=False)
        self.train_from_loss(loss)
        return loss, predict_masks

    def train_from_loss(self, loss):
        """SFM:158-162. zero_grad -> backward unless NaN -> clip_grad_value_(1.2) -> Adam, the last two fused in
        FusedClipAdam.step(); a step with no gradient anywhere touches nothing (the reference under torch 2.0.1)."""
        self.optimizer.zero_grad()
        if not torch.isnan(loss):
            loss.backward()
        self.optimizer.step()

    def scheduler_step(self):
        self.scheduler.step()

    def show_mask(self, vis, img, mask, title=""):
        mask_img = img.cpu().numpy()
        if mask is not None:
            mask = mask.detach().cpu().numpy()
            mask_img[0, :, :] = mask
        vis.image(mask_img, opts=dict(title=title))

    # ---- SFM:177-219 (autoencoder helpers) ----
    def eval_one_epoch_without_mask(self, imgs):
        self.model.eval()
        with torch.no_grad():
            imgs = imgs.to(self.device)
            outputs, _ = self.model(pixel_values=imgs)
            up = upsample_bilinear(outputs.logits, list(imgs.shape)[2:], sigmoid=True)
            return self.mse_loss_function(imgs, up), up

    def train_one_epoch_without_mask(self, imgs):
        self.model.train()
        imgs = imgs.to(self.device)
        outputs, _ = self.model(pixel_values=imgs)
        up = upsample_bilinear(outputs.logits, list(imgs.shape)[2:], sigmoid=True)
        loss = self.mse_loss_function(imgs, up)
        self.train_from_loss(loss)
        return loss, up
