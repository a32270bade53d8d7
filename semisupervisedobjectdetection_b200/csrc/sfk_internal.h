// Internal (C++) declarations shared between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfk {

struct GemmDesc {
  // D[M,N] = alpha * A·B^T (+bias[N]) (*row_scale[row / rows_per_scale]) (+residual fp32) -> out (fp32 | bf16 | fp32 +=)
  int M = 0, N = 0, K = 0;
  const void* a = nullptr;   // bf16. a_mn==0: [M, lda] (K contiguous). a_mn==1: [K, lda] (M contiguous)
  long long lda = 0;
  int a_mn = 0;
  const void* b = nullptr;   // bf16. b_mn==0: [N, ldb] (K contiguous). b_mn==1: [K, ldb] (N contiguous)
  long long ldb = 0;
  int b_mn = 0;
  float alpha = 1.f;
  const float* bias = nullptr;
  const float* residual = nullptr;
  long long ldr = 0;
  const float* row_scale = nullptr;
  int rows_per_scale = 1;
  void* out = nullptr;
  long long ldo = 0;
  int out_bf16 = 0;
  int atomic_add = 0;
  int relu = 0;
  int splits = 0;   // 0 = auto
  int block_n = 0;  // 0 = auto
};

int gemm_bf16(const GemmDesc& g, cudaStream_t stream);

}  // namespace sfk
