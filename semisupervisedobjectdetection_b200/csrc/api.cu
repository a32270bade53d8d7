// extern "C" entry points of libsfk.so (declared in include/sfk.h). Thin argument marshalling only.
#include "common.cuh"
#include "sfk_internal.h"
#include "../../include/sfk.h"

using namespace sfk;

extern "C" {

int sfk_gemm_bf16(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, int M, int N, int K,
                  float alpha, const float* bias, const float* residual, long long ldr, const float* row_scale,
                  int rows_per_scale, void* out, long long ldo, int out_bf16, int atomic_add, int relu, int splits,
                  int block_n, sfk_stream_t stream) {
  GemmDesc g;
  g.M = M; g.N = N; g.K = K;
  g.a = a; g.lda = lda; g.a_mn = a_mn;
  g.b = b; g.ldb = ldb; g.b_mn = b_mn;
  g.alpha = alpha; g.bias = bias; g.residual = residual; g.ldr = ldr;
  g.row_scale = row_scale; g.rows_per_scale = rows_per_scale;
  g.out = out; g.ldo = ldo; g.out_bf16 = out_bf16; g.atomic_add = atomic_add; g.relu = relu;
  g.splits = splits; g.block_n = block_n;
  return gemm_bf16(g, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
