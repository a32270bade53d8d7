"""Host-side profile of one teacher-student step (where does the CPU enqueue time go?)."""
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from semisupervisedobjectdetection_b200 import teacher_student as TS
from semisupervisedobjectdetection_b200.wrapper import SegFormerModel

dev = "cuda:0"
torch.manual_seed(0)
student = SegFormerModel(None, lr=3e-5, weight_decay=5e-5, scheduler=0.97, device=dev)
teacher = SegFormerModel(None, lr=5e-7, weight_decay=5e-5, scheduler=0.97, device=dev)
teacher.load_state_dict(student.state_dict())
teacher.optimizer.sync_shadow()
teacher.train(); student.train()
for p in teacher.model.parameters():
    p.requires_grad_(False)
l = torch.rand(8, 3, 512, 512, device=dev); u = torch.rand(8, 3, 512, 512, device=dev)
gt = (torch.rand(8, 512, 512, device=dev) > 0.9).float()
for _ in range(3):
    TS.ema_teacher_student_step(teacher, student, l, gt, u)
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
for _ in range(3):
    TS.ema_teacher_student_step(teacher, student, l, gt, u)
pr.disable()
t1 = time.perf_counter()
torch.cuda.synchronize()
print(f"cpu enqueue per step {1000*(t1-t0)/3:.1f} ms (under cProfile)")
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
st.sort_stats("cumulative").print_stats(30)
