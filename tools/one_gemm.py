"""Launch one GEMM shape a few times (for an ncu capture). Usage: python tools/one_gemm.py M N K [kind]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from check_gemm import gemm
from semisupervisedobjectdetection_b200 import _lib as L

M, N, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
kind = sys.argv[4] if len(sys.argv) > 4 else "nt"
for opt in ("gemm_epi8", "gemm_occ2"):
    if os.environ.get("SFK_" + opt.upper()) is not None:
        L.check(L.load().sfk_set_option(opt.encode(), int(os.environ["SFK_" + opt.upper()])))
dev = "cuda"
a = torch.randn(M, K, device=dev).bfloat16()
b = torch.randn(N, K, device=dev).bfloat16()
bias = torch.randn(N, device=dev)
if kind == "res":
    out = torch.empty(M, N, device=dev)
    kw = dict(bias=bias, residual=torch.randn(M, N, device=dev))
else:
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    kw = dict(out_bf16=True, bias=bias)
for _ in range(6):
    gemm(a, b, M, N, K, out=out, **kw)
torch.cuda.synchronize()
print("done")
