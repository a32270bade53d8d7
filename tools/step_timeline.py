"""Phase timeline of the MiT-b5 512x512 EMA teacher-student step (what bench.py times): CUDA events at the phase boundaries
of the main stream and of the teacher side stream, averaged over a few steps.  Shows where the overlapped step spends its time
(the sum of isolated kernel times says little once three streams share the SMs).

Usage: python tools/step_timeline.py [--steps 5] [--no-overlap]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import warnings

import torch

from semisupervisedobjectdetection_b200 import loss as myLoss
from semisupervisedobjectdetection_b200 import ops
from semisupervisedobjectdetection_b200 import teacher_student as TS
from semisupervisedobjectdetection_b200.optim import ema_update
from semisupervisedobjectdetection_b200.wrapper import SegFormerModel


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--no-overlap", action="store_true")
    ap.add_argument("--tile", type=int, default=512)
    args = ap.parse_args()
    dev = "cuda:0"
    torch.manual_seed(1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        student = SegFormerModel(None, lr=3e-5, weight_decay=5e-5, scheduler=0.97, device=dev)
        teacher = SegFormerModel(None, lr=5e-7, weight_decay=5e-5, scheduler=0.97, device=dev)
    teacher.load_state_dict(student.state_dict())
    teacher.optimizer.sync_shadow()
    teacher.train()
    student.train()
    for p in teacher.model.parameters():
        p.requires_grad_(False)
    if args.no_overlap:
        ops.WGRAD_OVERLAP = False
    t = args.tile
    img_l, img_u = torch.rand(8, 3, t, t, device=dev), torch.rand(8, 3, t, t, device=dev)
    gt = (torch.rand(8, t, t, device=dev) > 0.9).float()
    names = ["start", "teacher_fwd_done(side)", "student_fwd_done", "losses_done", "backward_done", "optimizer_done", "ema_done"]
    acc = {n: 0.0 for n in names}

    def step(record):
        ev = {n: torch.cuda.Event(enable_timing=True) for n in names}
        main_s, side = torch.cuda.current_stream(), TS._side_stream(dev)
        imgs = torch.cat([img_l, img_u], dim=0)
        ev["start"].record()
        if not args.no_overlap:
            side.wait_stream(main_s)
            with torch.cuda.stream(side), torch.no_grad():
                t_probs = teacher.predict(imgs)
                ev["teacher_fwd_done(side)"].record()
            s_probs = student.predict(imgs)
            ev["student_fwd_done"].record()
            main_s.wait_stream(side)
            t_probs.record_stream(main_s)
        else:
            with torch.no_grad():
                t_probs = teacher.predict(imgs)
            ev["teacher_fwd_done(side)"].record()
            s_probs = student.predict(imgs)
            ev["student_fwd_done"].record()
        t_l = TS.denoise_teacher_probs(t_probs[:8], gt)
        pseudo_l, _ = TS.pseudo_label_stats(t_l, allow_throw_sample=False)
        pseudo_u, stats_u = TS.pseudo_label_stats(t_probs[8:].contiguous(), allow_throw_sample=True)
        keep_u = (stats_u[:, 4] > 0.5).float()
        sup = TS.loss_func(s_probs[:8], gt.to(torch.int64))
        self_l = TS.loss_func(s_probs[:8], pseudo_l)
        self_u = myLoss.weighted_dice_loss(s_probs[8:], pseudo_u, keep_u)
        loss = 0.8 * sup + 0.2 * (self_l + self_u) * 0.5
        ev["losses_done"].record()
        student.optimizer.zero_grad()
        loss.backward()
        ev["backward_done"].record()
        student.optimizer.step()
        ev["optimizer_done"].record()
        ema_update(teacher.optimizer, student.optimizer, 0.99)
        ev["ema_done"].record()
        torch.cuda.synchronize()
        if record:
            for n in names:
                acc[n] += ev["start"].elapsed_time(ev[n])

    for _ in range(3):
        step(False)
    for _ in range(args.steps):
        step(True)
    print(f"TIMELINE (ms since step start, mean of {args.steps} steps, overlap={'off' if args.no_overlap else 'on'}):")
    for n in names:
        print(f"  {n:24s} {acc[n] / args.steps:8.2f}")


if __name__ == "__main__":
    main()
