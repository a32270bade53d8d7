"""BASELINE.json config 5: MiT-b5 inference / pseudo-label throughput sweep (teacher forward + x4 bilinear upsample + sigmoid +
confidence thresholding) over tile sizes 512 / 1024 / 2048 and batch sizes 1..64.  CUDA events, 3 warm-ups, inputs resident.
Also times the supervised training step of config 2 (b5, batch 16, Dice) when --train is given.

Usage: python tools/infer_sweep.py [--tiles 512,1024,2048] [--batches 1,2,4,8,16,32,64] [--train] [--ts] [--prompt] [--json out.json]
"""
import argparse
import json
import os
import sys
import warnings

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from semisupervisedobjectdetection_b200 import teacher_student as TS
from semisupervisedobjectdetection_b200.wrapper import SegFormerModel

FWD_GFLOP = {512: 219.27, 1024: 1120.29, 2048: 8372.38}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", default="512,1024,2048")
    ap.add_argument("--batches", default="1,2,4,8,16,32,64")
    ap.add_argument("--max-pixels", type=float, default=64 * 512 * 512, help="skip (tile, batch) pairs with more input pixels")
    ap.add_argument("--train", action="store_true")
    ap.add_argument("--ts", action="store_true", help="also time the reference-faithful teacher-student iteration (config 3-i)")
    ap.add_argument("--prompt", action="store_true", help="also time the prompt-tuning step (config 4)")
    ap.add_argument("--json", default=None)
    args = ap.parse_args()
    dev = "cuda:0"
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        w = SegFormerModel(None, lr=5e-5, weight_decay=5e-5, scheduler=0.97, device=dev)
    rows = []
    w.model.eval()
    for tile in [int(t) for t in args.tiles.split(",")]:
        for bs in [int(b) for b in args.batches.split(",")]:
            if bs * tile * tile > args.max_pixels:
                continue
            x = torch.rand(bs, 3, tile, tile, device=dev)

            def run():
                with torch.no_grad():
                    p = w.predict(x)
                    return TS.pseudo_label_stats(p, allow_throw_sample=True)

            for _ in range(3):
                run()
            torch.cuda.synchronize()
            iters = 5 if tile >= 2048 else 10
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            tf = FWD_GFLOP.get(tile, 0.0) * bs / ms
            rows.append(dict(mode="inference", tile=tile, batch=bs, ms=ms, images_per_s=bs / ms * 1e3, tflops=tf))
            print(f"INFER tile {tile:4d} batch {bs:2d}: {ms:8.2f} ms  {bs / ms * 1e3:8.1f} images/s  {tf:6.1f} TF/s (reference-algorithmic flops)",
                  flush=True)
            del x
            torch.cuda.empty_cache()
    if args.train:
        w.model.train()
        x = torch.rand(16, 3, 512, 512, device=dev)
        m = (torch.rand(16, 512, 512, device=dev) > 0.9).float()
        for _ in range(3):
            w.train_one_epoch(x, m)
        w.model.train()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            w.model.train()   # train_one_epoch -> predict(isEval=False) flips to eval; the benchmark keeps train mode
            w.train_one_epoch(x, m)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rows.append(dict(mode="supervised_train", tile=512, batch=16, ms=ms, images_per_s=16 / ms * 1e3, tflops=657.8 * 16 / ms))
        print(f"TRAIN supervised b5 512 batch 16 (config 2): {ms:8.2f} ms/step  {16 / ms * 1e3:8.1f} images/s", flush=True)
    if args.ts:
        # config 3-i, the reference-faithful iteration (TS:80-127): teacher forward on 8 unlabeled tiles + thresholding (host
        # read of the keep flags, as the reference does per sample); teacher fwd+bwd+Adam and student fwd+bwd+Adam on 8 labeled
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            teacher = SegFormerModel(None, lr=5e-7, weight_decay=5e-5, scheduler=0.97, device=dev)
            student = SegFormerModel(None, lr=3e-5, weight_decay=5e-5, scheduler=0.97, device=dev)
        xl, xu = torch.rand(8, 3, 512, 512, device=dev), torch.rand(8, 3, 512, 512, device=dev)
        gt = (torch.rand(8, 512, 512, device=dev) > 0.9).float()

        def it():
            teacher.model.train(); student.model.train()
            TS.unlabeled_iteration(teacher, xu, train_teacher=False)
            TS.labeled_iteration(teacher, student, xl, gt)

        for _ in range(3):
            it()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            it()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        rows.append(dict(mode="reference_faithful_ts_iteration", tile=512, batch=16, ms=ms, images_per_s=16 / ms * 1e3,
                         tflops=(219.27 * 8 + 2 * 657.8 * 8) / ms))
        print(f"TS reference-faithful iteration (config 3-i, 8 unlabeled + 8 labeled): {ms:8.2f} ms  {16 / ms * 1e3:8.1f} images/s",
              flush=True)
        del teacher, student
        torch.cuda.empty_cache()
    if args.prompt:
        # config 4: frozen encoder blocks 0,1 + P learnable prompt tokens per stage (segFormer_transfer_learning_main.py:181-184)
        for P in (1, 10, 100):
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                wp = SegFormerModel(None, lr=5e-5, weight_decay=5e-5, scheduler=0.97, device=dev)
            wp.frozen_encoder(layers=[0, 1])
            wp.add_prompt_token([P] * 4)
            x = torch.rand(16, 3, 512, 512, device=dev)
            m = (torch.rand(16, 512, 512, device=dev) > 0.9).float()
            for _ in range(3):
                wp.model.train()
                wp.train_one_epoch(x, m)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                wp.model.train()
                wp.train_one_epoch(x, m)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            rows.append(dict(mode=f"prompt_tuning_P{P}", tile=512, batch=16, ms=ms, images_per_s=16 / ms * 1e3))
            print(f"PROMPT tuning step (config 4, blocks 0,1 frozen, {P} prompt tokens per stage, batch 16): {ms:8.2f} ms  "
                  f"{16 / ms * 1e3:8.1f} images/s", flush=True)
            del wp
            torch.cuda.empty_cache()
    if args.json:
        with open(args.json, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
