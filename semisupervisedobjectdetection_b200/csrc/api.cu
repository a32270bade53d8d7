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


int sfk_attention_fwd(const void* q, long long ldq, long long q_batch_stride, const void* k, const void* v,
                      long long ldkv, long long kv_batch_stride, void* o, long long ldo, long long o_batch_stride,
                      float* lse, int B, int heads, int T, int Tk, int D, float scale, sfk_stream_t stream) {
  AttentionDesc a;
  a.B = B; a.heads = heads; a.T = T; a.Tk = Tk; a.D = D; a.scale = scale;
  a.q = q; a.ldq = ldq; a.q_batch_stride = q_batch_stride;
  a.k = k; a.v = v; a.ldkv = ldkv; a.kv_batch_stride = kv_batch_stride;
  a.o = o; a.ldo = ldo; a.o_batch_stride = o_batch_stride; a.lse = lse;
  return attention_fwd(a, static_cast<cudaStream_t>(stream));
}

int sfk_attention_bwd(const void* q, long long ldq, long long q_batch_stride, const void* k, const void* v,
                      long long ldkv, long long kv_batch_stride, const void* o, const void* dO, long long ldo,
                      long long o_batch_stride, const float* lse, float* delta, float* dq, long long ld_dq,
                      long long dq_batch_stride, float* dk, float* dv, long long ld_dkv, long long dkv_batch_stride,
                      int B, int heads, int T, int Tk, int D, float scale, sfk_stream_t stream) {
  AttentionBwdDesc a;
  a.B = B; a.heads = heads; a.T = T; a.Tk = Tk; a.D = D; a.scale = scale;
  a.q = q; a.ldq = ldq; a.q_batch_stride = q_batch_stride;
  a.k = k; a.v = v; a.ldkv = ldkv; a.kv_batch_stride = kv_batch_stride;
  a.o = const_cast<void*>(o); a.ldo = ldo; a.o_batch_stride = o_batch_stride; a.lse = const_cast<float*>(lse);
  a.dO = dO; a.delta = delta;
  a.dq = dq; a.ld_dq = ld_dq; a.dq_batch_stride = dq_batch_stride;
  a.dk = dk; a.dv = dv; a.ld_dkv = ld_dkv; a.dkv_batch_stride = dkv_batch_stride;
  return attention_bwd(a, static_cast<cudaStream_t>(stream));
}

int sfk_layernorm_fwd(const void* x, int x_bf16, long long map_period, long long map_extra, long long map_offset,
                      const float* gamma, const float* beta, void* y, int y_bf16, float* mean, float* rstd,
                      long long rows, int C, float eps, sfk_stream_t stream) {
  RowMap m; m.period = map_period; m.extra = map_extra; m.offset = map_offset;
  return layernorm_fwd(x, x_bf16, m, gamma, beta, y, y_bf16, mean, rstd, rows, C, eps, static_cast<cudaStream_t>(stream));
}

int sfk_layernorm_bwd(const void* dy, int dy_bf16, const void* x, int x_bf16, const float* mean, const float* rstd,
                      const float* gamma, const float* dres, void* dx, int dx_bf16, float* dgamma, float* dbeta,
                      long long rows, int C, sfk_stream_t stream) {
  RowMap id;
  return layernorm_bwd(dy, dy_bf16, x, x_bf16, id, mean, rstd, gamma, dres, id, dx, dx_bf16, id, dgamma, dbeta, rows, C,
                       static_cast<cudaStream_t>(stream));
}

int sfk_dwconv3x3(const void* x, const float* wt, const float* bias, void* y, int B, int H, int W, int C, int mode,
                  sfk_stream_t stream) {
  return dwconv3x3(x, wt, bias, y, B, H, W, C, mode, static_cast<cudaStream_t>(stream));
}
int sfk_dwconv3x3_gelu_bwd(const void* x, const float* wt, const float* bias, const void* dg, void* dpre, float* dwt,
                           float* dbias, int B, int H, int W, int C, sfk_stream_t stream) {
  return dwconv3x3_gelu_bwd(x, wt, bias, dg, dpre, dwt, dbias, B, H, W, C, static_cast<cudaStream_t>(stream));
}
int sfk_dw_weight_to_tap(const float* w, float* wt, int C, sfk_stream_t stream) {
  return dw_weight_to_tap(w, wt, C, static_cast<cudaStream_t>(stream));
}
int sfk_dw_wgrad_from_tap(const float* dwt, float* dw, int C, sfk_stream_t stream) {
  return dw_wgrad_from_tap(dwt, dw, C, static_cast<cudaStream_t>(stream));
}
int sfk_im2col_nhwc(const void* x, void* A, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                    long long tok_offset, sfk_stream_t stream) {
  return im2col_nhwc(x, A, B, H, W, C, K, S, P, tok_stride, tok_offset, static_cast<cudaStream_t>(stream));
}
int sfk_im2col_nchw_f32(const float* x, void* A, int B, int Cin, int H, int W, int K, int S, int P, int Kp,
                        sfk_stream_t stream) {
  return im2col_nchw_f32(x, A, B, Cin, H, W, K, S, P, Kp, static_cast<cudaStream_t>(stream));
}
int sfk_col2im_nhwc(const void* dA, float* dx, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                    long long tok_offset, int accumulate, sfk_stream_t stream) {
  return col2im_nhwc(dA, dx, B, H, W, C, K, S, P, tok_stride, tok_offset, accumulate, static_cast<cudaStream_t>(stream));
}
int sfk_conv_weight_ohwi(const float* w, void* out, int O, int I, int KH, int KW, long long ldo, sfk_stream_t stream) {
  return conv_weight_ohwi(w, out, O, I, KH, KW, ldo, static_cast<cudaStream_t>(stream));
}
int sfk_conv_wgrad_ohwi_to_oihw(const float* g, float* dw, int O, int I, int KH, int KW, long long ldg,
                                sfk_stream_t stream) {
  return conv_wgrad_ohwi_to_oihw(g, dw, O, I, KH, KW, ldg, static_cast<cudaStream_t>(stream));
}
int sfk_cast_2d(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols, int cols_padded,
                sfk_stream_t stream) {
  return cast_2d(src, lds, dst, ldd, rows, cols, cols_padded, static_cast<cudaStream_t>(stream));
}
int sfk_scale_cast(const void* x, int x_bf16, void* y, int y_bf16, const float* scale, int rows_per_scale,
                   long long rows, int C, sfk_stream_t stream) {
  RowMap id;
  return scale_cast(x, x_bf16, id, y, y_bf16, id, scale, rows_per_scale, rows, C, static_cast<cudaStream_t>(stream));
}
int sfk_colsum(const void* x, int x_bf16, long long ld, float* out, long long rows, int C, sfk_stream_t stream) {
  return colsum(x, x_bf16, ld, out, rows, C, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
