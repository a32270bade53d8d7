"""Teacher-student glue of main_segformer/segFormer_semi_teacherstudent_main.py (TS) on libsfk kernels:
threshold_pseudo_masks (TS:24-56), the denoise-label trick (TS:110-118), the loss mix (TS:125-126), the reference-faithful
labeled / unlabeled iterations (TS:80-127) and the north-star EMA-teacher step (teacher <- d*teacher + (1-d)*student;
d = 0 reproduces the hard reset of TS:202-204).
"""
from __future__ import annotations

from typing import Optional

import torch

from . import loss as myLoss
from . import ops
from .optim import ema_update

PESUDO_MAKS_THRESHOLD = 0.7  # (sic) TS:18
CONFIDENT_THRESHOLD = 0.7    # TS:19

loss_func = myLoss.SegmentationLoss(1, loss_type='dice', activation='none')  # TS:253


def pseudo_label_stats(masks: torch.Tensor, allow_throw_sample: bool = True, threshold: Optional[float] = None,
                       confident_threshold: Optional[float] = None):
    """Device part of threshold_pseudo_masks: returns (pseudo int64 [N,...], stats f32 [N,5]) with
    stats[n] = {sum|p|, confidence, #(p >= thr), per-sample Dice loss against its own pseudo mask, keep flag}."""
    thr = PESUDO_MAKS_THRESHOLD if threshold is None else threshold
    cthr = CONFIDENT_THRESHOLD if confident_threshold is None else confident_threshold
    masks = ops._chk(masks.detach(), torch.float32, "probabilities")
    N = masks.shape[0]
    P = masks[0].numel()
    pseudo = torch.empty(masks.shape, device=masks.device, dtype=torch.int64)
    stats = torch.empty(N, 5, device=masks.device, dtype=torch.float32)
    lib = ops.lib()
    ws = ops.scratch(lib.sfk_pseudo_workspace_bytes(N), masks.device, "loss")
    # comparisons run in fp32 against float32(thr) and float32(1 - thr) (1 - thr evaluated in double first), like torch
    ops.call("sfk_pseudo_label_threshold", masks.data_ptr(), pseudo.data_ptr(), stats.data_ptr(), N, P, float(thr), float(1 - thr),
             float(cthr), 1000.0, int(allow_throw_sample), ws.data_ptr(), ops.stream())
    return pseudo, stats


def threshold_pseudo_masks(img, masks, allow_throw_sample=True):
    """TS:24-56. Returns (confident_img | None, confident_mask (int64) | None, confident_predicted | [], confidence [N],
    mean per-sample Dice loss).  One host read of the N keep flags replaces the reference's 2N per-sample syncs."""
    pseudo, stats = pseudo_label_stats(masks, allow_throw_sample)
    confidence = stats[:, 1].clone()
    keep = stats[:, 4] > 0.5
    idx = torch.nonzero(keep).flatten()
    if idx.numel() == 0:
        return None, None, [], confidence, torch.tensor(float("nan"), device=masks.device)
    mean_loss = stats[idx, 3].sum() / idx.numel()
    mean_loss.requires_grad_(True)  # the reference's accumulator is a requires_grad leaf copy (TS:41); grads reach nothing (M10)
    return img[idx], pseudo[idx], masks[idx], confidence, mean_loss


def denoise_teacher_probs(teacher_probs: torch.Tensor, ground_truth: torch.Tensor) -> torch.Tensor:
    """TS:111-114: clamp(t + 0.2*gt - 0.1, 0, 1)."""
    t = ops._chk(teacher_probs.detach(), torch.float32, "teacher probabilities")
    g = ops._chk(ground_truth, torch.float32, "ground truth")
    out = torch.empty_like(t)
    ops.call("sfk_denoise_labels", t.data_ptr(), g.data_ptr(), out.data_ptr(), t.numel(), ops.stream())
    return out


def unlabeled_iteration(teacher, img, train_teacher: bool = False):
    """TS:80-95. Teacher forward under no_grad, thresholding, optional (weight-neutral, M10) teacher step."""
    with torch.no_grad():
        img = img.to(device=teacher.device, dtype=torch.float32)
        predicted = teacher.predict(img)
    out = threshold_pseudo_masks(img, predicted)
    if out[0] is not None and train_teacher:
        teacher.train_from_loss(out[4])
    return out


def labeled_iteration(teacher, student, img, ground_truth, supervise_weight: float = 0.8, denoise_label: bool = True):
    """TS:99-127: teacher fwd+bwd+Adam on its denoised pseudo labels, then student fwd+bwd+Adam on
    w * Dice(student, gt) + (1 - w) * Dice(student, teacher mask)."""
    img = img.to(device=teacher.device, dtype=torch.float32)
    ground_truth = ground_truth.to(device=teacher.device, dtype=torch.float32)
    if not denoise_label:
        teacher_loss_gt, _ = teacher.train_one_epoch(img, ground_truth)
        with torch.no_grad():
            teacher_masks = teacher.predict(img)
    else:
        t0 = teacher.predict(img)
        t = denoise_teacher_probs(t0, ground_truth)
        _, conf_mask, _, _, _ = threshold_pseudo_masks(img, t, allow_throw_sample=False)
        teacher_loss_gt = loss_func(t0, conf_mask)
        teacher_masks = conf_mask
        teacher.train_from_loss(teacher_loss_gt)
    student_loss, student_pred = student.predict(img, ground_truth)
    self_sup = loss_func(student_pred, teacher_masks)
    loss = supervise_weight * student_loss + (1 - supervise_weight) * self_sup
    student.train_from_loss(loss)
    return loss, teacher_loss_gt, student_loss, self_sup


_SIDE_STREAMS: dict = {}
_MAIN_STREAMS: dict = {}
# Experiment switch: run the step's dependent chain (student forward / backward, optimiser) on a high-priority stream so that the
# weight-gradient GEMMs and the teacher forward (default-priority side streams) only back-fill.  Measured on B200 (A/B/A/B in
# one process group, tools/ab_step.py): 56.9 ms with the priority stream against 52.5 ms without - it is OFF.
HIGH_PRIORITY_MAIN = False


def _main_stream(device) -> "torch.cuda.Stream":
    key = str(device)
    if key not in _MAIN_STREAMS:
        _MAIN_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1)
    return _MAIN_STREAMS[key]



def _side_stream(device) -> "torch.cuda.Stream":
    key = str(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device)
    return _SIDE_STREAMS[key]


def ema_teacher_student_step(teacher, student, img_labeled, ground_truth, img_unlabeled, supervise_weight: float = 0.8,
                             ema_decay: float = 0.99, grad_sync=None, overlap_teacher: bool = True):
    """North-star variant: EMA teacher forward on labeled + unlabeled tiles -> confidence-thresholded pseudo labels ->
    student forward/backward on both -> fused clip+Adam -> rank-local EMA teacher update.
    `grad_sync`: optional callable run between backward and the optimiser step (data-parallel all-reduce)."""
    with ops.device_of(student.device):
        if not HIGH_PRIORITY_MAIN:
            return _ema_teacher_student_step(teacher, student, img_labeled, ground_truth, img_unlabeled, supervise_weight,
                                             ema_decay, grad_sync, overlap_teacher)
        cur, hp = torch.cuda.current_stream(), _main_stream(student.device)
        hp.wait_stream(cur)
        with torch.cuda.stream(hp):
            loss = _ema_teacher_student_step(teacher, student, img_labeled, ground_truth, img_unlabeled, supervise_weight,
                                             ema_decay, grad_sync, overlap_teacher)
        cur.wait_stream(hp)
        loss.record_stream(cur)
        return loss


def _ema_teacher_student_step(teacher, student, img_labeled, ground_truth, img_unlabeled, supervise_weight, ema_decay,
                              grad_sync, overlap_teacher):
    dev = student.device
    img_l = img_labeled.to(device=dev, dtype=torch.float32)
    img_u = img_unlabeled.to(device=dev, dtype=torch.float32)
    gt = ground_truth.to(device=dev, dtype=torch.float32)
    imgs = torch.cat([img_l, img_u], dim=0)
    nl = img_l.shape[0]
    if overlap_teacher:
        # the (gradient-free) teacher forward and the student forward are independent until the losses: run the teacher on a
        # side stream so the two chains of small kernels fill each other's bubbles
        main, side = torch.cuda.current_stream(), _side_stream(dev)
        side.wait_stream(main)
        with torch.cuda.stream(side), torch.no_grad():
            t_probs = teacher.predict(imgs)
        s_probs = student.predict(imgs)
        main.wait_stream(side)
        t_probs.record_stream(main)
    else:
        with torch.no_grad():
            t_probs = teacher.predict(imgs)
        s_probs = student.predict(imgs)
    t_l = denoise_teacher_probs(t_probs[:nl], gt)
    pseudo_l, _ = pseudo_label_stats(t_l, allow_throw_sample=False)
    pseudo_u, stats_u = pseudo_label_stats(t_probs[nl:].contiguous(), allow_throw_sample=True)
    keep_u = (stats_u[:, 4] > 0.5).float()
    # static shapes: rejected unlabeled tiles (TS:43 filters them out) stay in the batch with weight 0, i.e. they add neither
    # loss nor gradient; every kept tile weighs 1/N_u, like a labeled tile's 1/N_l
    sup = loss_func(s_probs[:nl], gt.to(torch.int64))
    self_l = loss_func(s_probs[:nl], pseudo_l)
    self_u = myLoss.weighted_dice_loss(s_probs[nl:], pseudo_u, keep_u)
    loss = supervise_weight * sup + (1 - supervise_weight) * (self_l + self_u) * 0.5
    student.optimizer.zero_grad()
    loss.backward()
    if grad_sync is not None:
        grad_sync()
    student.optimizer.step()
    ema_update(teacher.optimizer, student.optimizer, ema_decay)
    return loss


class PipelinedEmaLoop:
    """Training loop driver around ema_teacher_student_step for batches that live in (pinned) host memory.

    The reference's loop (TS:80-127, 230-260) does `.to(device)` and reads losses with `.item()` inline, which serialises the
    host-to-device copy, the step and the read-back.  Here batch i+1 is copied on a copy stream while step i computes, and the
    loss of step i is read from a pinned buffer after step i+1 has been enqueued, so neither the copy nor the read-back stalls
    the GPU.  Every batch is still copied from host memory and every loss still comes back to the host - one step later.

        loop = PipelinedEmaLoop(teacher, student)
        for loss in loop.run(batches):      # batches: iterable of (img_labeled, ground_truth, img_unlabeled) host tensors
            ...                              # python floats, in batch order
    """

    def __init__(self, teacher, student, supervise_weight: float = 0.8, ema_decay: float = 0.99, grad_sync=None,
                 overlap_teacher: bool = True, before_step=None):
        self.teacher, self.student = teacher, student
        self.kw = dict(supervise_weight=supervise_weight, ema_decay=ema_decay, grad_sync=grad_sync, overlap_teacher=overlap_teacher)
        self.before_step = before_step
        self.dev = torch.device(student.device)
        self._copy = torch.cuda.Stream(device=self.dev)
        self._loss_host = torch.empty(2, dtype=torch.float32).pin_memory()
        self._events = [torch.cuda.Event(), torch.cuda.Event()]
        self._bufs = [None, None]      # double-buffered device copies of a batch
        self._used = [None, None]      # event recorded on the compute stream after the step that read slot k

    def _upload(self, batch, slot):
        """Enqueue the host-to-device copies of one batch on the copy stream into the double-buffered device inputs of `slot`
        (no allocation in steady state); returns (device tensors, ready event)."""
        bufs = self._bufs[slot]
        if bufs is None or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(bufs, batch)):
            bufs = tuple(torch.empty(t.shape, dtype=t.dtype, device=self.dev) for t in batch)
            self._bufs[slot] = bufs
        if self._used[slot] is not None:
            self._copy.wait_event(self._used[slot])   # the step that last read these buffers must be done with them
        with torch.cuda.stream(self._copy):
            for b, t in zip(bufs, batch):
                b.copy_(t, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy)
        return bufs, ev

    def run(self, batches):
        with ops.device_of(self.dev):
            it = iter(batches)
            try:
                nxt = self._upload(next(it), 0)
            except StopIteration:
                return
            pending = None   # (slot, event) of the loss not yet handed to the caller
            i = 0
            main = torch.cuda.current_stream()
            while nxt is not None:
                (img_l, gt, img_u), ready = nxt
                try:
                    nxt = self._upload(next(it), (i + 1) & 1)       # overlaps with the step enqueued below
                except StopIteration:
                    nxt = None
                main.wait_event(ready)
                if self.before_step is not None:
                    self.before_step()
                loss = ema_teacher_student_step(self.teacher, self.student, img_l, gt, img_u, **self.kw)
                slot = i & 1
                self._loss_host[slot:slot + 1].copy_(loss.detach().reshape(1), non_blocking=True)
                self._events[slot].record(main)
                self._used[slot] = self._events[slot]
                if pending is not None:
                    pslot, pev = pending
                    pev.synchronize()
                    yield float(self._loss_host[pslot])
                pending = (slot, self._events[slot])
                i += 1
            if pending is not None:
                pslot, pev = pending
                pev.synchronize()
                yield float(self._loss_host[pslot])
