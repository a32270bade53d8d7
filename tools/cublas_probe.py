"""Launch torch.matmul (cuBLAS) on the stage-3 GEMM shapes, for an ncu capture of the library's kernel choice (context only;
cuBLAS is not used by the product).  Usage: python tools/cublas_probe.py"""
import torch

for (M, N, K) in [(16384, 320, 1280), (16384, 320, 320), (16384, 1280, 320), (4096, 320, 1280)]:
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, K, device="cuda").bfloat16()
    for _ in range(3):
        c = torch.matmul(a, b.t())
torch.cuda.synchronize()
print("done")
