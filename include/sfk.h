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

/* ---- fused attention (tcgen05 QK^T -> online softmax -> PV; scores never touch HBM) -------------------------
 * Replaces MS:207-226 (matmul / div sqrt(d) / softmax / dropout(p=0) / matmul / head merge).
 *   q: bf16, token t of sample b at q + b*q_batch_stride + t*ldq, head h owns columns [h*D, (h+1)*D) (MS:176-179)
 *   k, v: bf16 addressed likewise with ldkv / kv_batch_stride (Tk rows incl. prompt / cls rows, MS:203-208)
 *   o: bf16 context [B, T, heads*D] in the merged-head layout of MS:224-226;  lse: f32 [B, heads, T] (log2 domain)
 * D must be 32 (MiT-b0) or 64 (MiT-b1..b5). */
int sfk_attention_fwd(const void* q, long long ldq, long long q_batch_stride, const void* k, const void* v,
                      long long ldkv, long long kv_batch_stride, void* o, long long ldo, long long o_batch_stride,
                      float* lse, int B, int heads, int T, int Tk, int D, float scale, sfk_stream_t stream);
/* Backward of the above. dq, dk, dv are f32 and accumulated with red.global.add: the caller zero-fills them.
 * delta is an f32 workspace of B*heads*T elements. */
int sfk_attention_bwd(const void* q, long long ldq, long long q_batch_stride, const void* k, const void* v,
                      long long ldkv, long long kv_batch_stride, const void* o, const void* dO, long long ldo,
                      long long o_batch_stride, const float* lse, float* delta, float* dq, long long ld_dq,
                      long long dq_batch_stride, float* dk, float* dv, long long ld_dkv, long long dkv_batch_stride,
                      int B, int heads, int T, int Tk, int D, float scale, sfk_stream_t stream);

/* ---- LayerNorm (MS:133,174,325,333,418; eps inside the sqrt, biased variance, affine) --------------------------
 * Rows may be remapped: src_row = row + (period > 0 ? (row / period) * extra : 0) + offset  (prompt rows in front). */
int sfk_layernorm_fwd(const void* x, int x_bf16, long long map_period, long long map_extra, long long map_offset,
                      const float* gamma, const float* beta, void* y, int y_bf16, float* mean, float* rstd,
                      long long rows, int C, float eps, sfk_stream_t stream);
/* dx = LN'(dy) (+ dres), dgamma += , dbeta += (f32 atomics; pass NULL to skip the parameter gradients). */
int sfk_layernorm_bwd(const void* dy, int dy_bf16, const void* x, int x_bf16, const float* mean, const float* rstd,
                      const float* gamma, const float* dres, void* dx, int dx_bf16, float* dgamma, float* dbeta,
                      long long rows, int C, sfk_stream_t stream);

/* ---- MixFFN depthwise 3x3 (+bias) + exact-erf GELU, channels-last (MS:283-294, MS:311-313) --------------------
 * x, y: bf16 [B,H,W,C]; wt: f32 tap-major [9][C] copy of dwconv.weight[C,1,3,3] (sfk_dw_weight_to_tap).
 * mode 0: y = gelu(conv(x)+bias).  mode 1: y = conv with flipped taps (input gradient), bias ignored. */
int sfk_dwconv3x3(const void* x, const float* wt, const float* bias, void* y, int B, int H, int W, int C, int mode,
                  sfk_stream_t stream);
/* dpre = dg * gelu'(conv(x)+bias) (bf16), dwt[9][C] += , dbias[C] += */
int sfk_dwconv3x3_gelu_bwd(const void* x, const float* wt, const float* bias, const void* dg, void* dpre, float* dwt,
                           float* dbias, int B, int H, int W, int C, sfk_stream_t stream);
int sfk_dw_weight_to_tap(const float* w, float* wt, int C, sfk_stream_t stream);
int sfk_dw_wgrad_from_tap(const float* dwt, float* dw, int C, sfk_stream_t stream);

/* ---- patch gather / scatter for the overlapping patch-embed conv (MS:123-131) and the SR conv (MS:171-173) ---- */
int sfk_im2col_nhwc(const void* x, void* A, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                    long long tok_offset, sfk_stream_t stream);
int sfk_im2col_nchw_f32(const float* x, void* A, int B, int Cin, int H, int W, int K, int S, int P, int Kp,
                        sfk_stream_t stream);
int sfk_col2im_nhwc(const void* dA, float* dx, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                    long long tok_offset, int accumulate, sfk_stream_t stream);
int sfk_conv_weight_ohwi(const float* w, void* out, int O, int I, int KH, int KW, long long ldo, sfk_stream_t stream);
int sfk_conv_wgrad_ohwi_to_oihw(const float* g, float* dw, int O, int I, int KH, int KW, long long ldg,
                                sfk_stream_t stream);
int sfk_cast_2d(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols, int cols_padded,
                sfk_stream_t stream);

/* ---- small row-wise helpers ------------------------------------------------------------------------------- */
/* y[row,:] = x[row,:] * scale[row / rows_per_scale] with dtype conversion (drop-path backward, casts) */
int sfk_scale_cast(const void* x, int x_bf16, void* y, int y_bf16, const float* scale, int rows_per_scale,
                   long long rows, int C, sfk_stream_t stream);
/* out[c] += sum_rows x[row, c] (bias gradients) */
int sfk_colsum(const void* x, int x_bf16, long long ld, float* out, long long rows, int C, sfk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SFK_H_ */
