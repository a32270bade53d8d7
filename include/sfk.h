/* sfk.h — C ABI of libsfk.so: B200 (sm_100a) kernels for the SegFormer (MiT-b0..b5) teacher-student hot path.
 *
 * Conventions:
 *   - every entry point returns 0 on success and a negative code on failure; the message is available through
 *     sfk_last_error() (thread-local).  No exception crosses this boundary.
 *   - the caller owns every buffer (PyTorch's caching allocator in the shipped host code); the library never
 *     allocates or frees device memory and keeps no pointer after returning.
 *   - all work is enqueued on the stream passed in (a cudaStream_t); there are no hidden synchronisations and
 *     no use of the default stream, so calls are capturable in CUDA graphs.
 *   - unsupported shape / dtype / architecture is an error, never a fallback.  There is no CPU path.
 *   - "bf16" buffers are raw __nv_bfloat16, "f32" raw float, "i64" int64_t.  Activations are token-major
 *     ([B, tokens, C] == NHWC) unless stated otherwise.
 *
 * Reference citations are relative to the upstream repository root:
 *   MS   = segformer_source/models/segformer/modeling_segformer.py
 *   SFM  = models/SegFormerModel.py
 *   LOSS = models/Loss.py
 *   TS   = main_segformer/segFormer_semi_teacherstudent_main.py
 */
#ifndef SFK_H_
#define SFK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SFK_VERSION 1

typedef void* sfk_stream_t; /* cudaStream_t */

/* ---- library -------------------------------------------------------------------------------------------- */
const char* sfk_last_error(void);
int sfk_version(void);
/* 0 iff the current CUDA device is compute capability 10.x (sm_100). */
int sfk_device_check(void);

/* ---- dense contractions (tcgen05 + TMA + TMEM) ------------------------------------------------------------
 * D[M,N] = alpha * (A · B^T) (+ bias[N]) (* row_scale[row / rows_per_scale]) (+ residual[M,N] f32) (relu)
 *   a: bf16.  a_mn == 0: stored [M, lda], K contiguous.   a_mn == 1: stored [K, lda], M contiguous.
 *   b: bf16.  b_mn == 0: stored [N, ldb], K contiguous.   b_mn == 1: stored [K, ldb], N contiguous.
 *   out: f32 or bf16 [M, ldo]; atomic_add != 0 accumulates into an f32 out with red.global.add (split-K).
 * Replaces every nn.Linear / 1x1 Conv2d / (kernel == stride) Conv2d on the path and their autograd gradients:
 *   MS:163-165 (query/key/value), MS:171-173 (sr), MS:236 (SelfOutput.dense), MS:301,307 (MixFFN dense1/2),
 *   MS:713 (SegformerMLP.proj), MS:738 (linear_fuse). */
int sfk_gemm_bf16(const void* a, long long lda, int a_mn, const void* b, long long ldb, int b_mn, int M, int N, int K,
                  float alpha, const float* bias, const float* residual, long long ldr, const float* row_scale,
                  int rows_per_scale, void* out, long long ldo, int out_bf16, int atomic_add, int relu, int splits,
                  int block_n, sfk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SFK_H_ */
