// extern "C" entry points of libsfk.so (declared in include/sfk.h). Thin argument marshalling only.
#include <string>
#include "common.cuh"
#include "sfk_internal.h"
#include "../../include/sfk.h"

using namespace sfk;

namespace sfk { extern int g_gemm_cluster2; extern int g_gemm_occ2; extern int g_gemm_epi_tma; extern int g_gemm_bstat; extern int g_gemm_epi_bufs; extern int g_gemm_max_stages; extern int g_gemm_krot; extern int g_gemm_dbg; extern int g_ln_bwd_ring; extern unsigned long long* g_gemm_trace; void set_dw_bwd_vec(int v); extern int g_sr_tma_patch; extern int g_sr_patch_store; extern int g_fwd_kv_overlap; extern int g_upsample_sep; extern int g_attn_bwd_persist; extern int g_attn_fwd256; int set_attn_dbg(int v); int get_attn_trace(long long* out16); }

extern "C" {

int sfk_set_option(const char* name, int value) {
  if (name != nullptr && std::string(name) == "gemm_cluster2") { sfk::g_gemm_cluster2 = value < 0 ? 0 : (value > 2 ? 2 : value); return 0; }
  if (name != nullptr && std::string(name) == "gemm_occ2") { sfk::g_gemm_occ2 = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "gemm_epi_tma") { sfk::g_gemm_epi_tma = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "gemm_bstat") { sfk::g_gemm_bstat = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "upsample_sep") { sfk::g_upsample_sep = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "fwd_kv_overlap") { sfk::g_fwd_kv_overlap = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "sr_patch_store") { sfk::g_sr_patch_store = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "gemm_krot") { sfk::g_gemm_krot = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "gemm_max_stages") { sfk::g_gemm_max_stages = value; return 0; }
  if (name != nullptr && std::string(name) == "attn_dbg") return sfk::set_attn_dbg(value);
  if (name != nullptr && std::string(name) == "attn_fwd256") { sfk::g_attn_fwd256 = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "gemm_epi_bufs") { sfk::g_gemm_epi_bufs = value >= 2 ? 2 : 1; return 0; }
  if (name != nullptr && std::string(name) == "gemm_dbg") { sfk::g_gemm_dbg = value; return 0; }
  if (name != nullptr && std::string(name) == "ln_bwd_ring") { sfk::g_ln_bwd_ring = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "attn_bwd_persist") { sfk::g_attn_bwd_persist = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "sr_tma_patch") { sfk::g_sr_tma_patch = value ? 1 : 0; return 0; }
  if (name != nullptr && std::string(name) == "dw_bwd_vec") { sfk::set_dw_bwd_vec(value); return 0; }
  sfk::set_error("sfk_set_option: unknown option");
  return -1;
}

/* debug: device buffer of >= 8 u64 receiving CTA 0's globaltimer at phase boundaries of every following GEMM (NULL = off) */
int sfk_debug_attn_trace(long long* host_i64x16) { return sfk::get_attn_trace(host_i64x16); }
int sfk_debug_gemm_trace(void* device_u64x8) {
  sfk::g_gemm_trace = static_cast<unsigned long long*>(device_u64x8);
  return 0;
}

int sfk_gemm_plan(int M, int N, int K, int a_mn, int b_mn, int out_bf16, int atomic_add, int splits, int block_n, int* out8) {
  if (out8 == nullptr) { set_error("sfk_gemm_plan: null output"); return -1; }
  GemmDesc g;
  g.M = M; g.N = N; g.K = K; g.a_mn = a_mn; g.b_mn = b_mn; g.out_bf16 = out_bf16; g.atomic_add = atomic_add;
  g.splits = splits; g.block_n = block_n;
  return gemm_plan(g, out8);
}

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
  return dwconv3x3(x, wt, bias, y, nullptr, B, H, W, C, mode, static_cast<cudaStream_t>(stream));
}
int sfk_dwconv3x3_gelu_bwd(const void* x, const float* wt, const float* bias, const void* dg, void* dpre, float* dwt,
                           float* dbias, int B, int H, int W, int C, sfk_stream_t stream) {
  return dwconv3x3_gelu_bwd(x, wt, bias, dg, dpre, dwt, dbias, B, H, W, C, static_cast<cudaStream_t>(stream));
}
int sfk_dw_weight_to_tap(const float* w, float* wt, int C, sfk_stream_t stream) {
  return dw_weight_to_tap(w, wt, C, static_cast<cudaStream_t>(stream));
}
int sfk_weight_layouts_batched(const sfk_layout_entry* table, int n, sfk_stream_t stream) {
  return weight_layouts_batched(table, n, static_cast<cudaStream_t>(stream));
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


#define ST static_cast<cudaStream_t>(stream)
int sfk_upsample_concat(const void* in, void* out, int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off,
                        sfk_stream_t stream) { return upsample_concat(in, out, B, h, w, C, H, W, ld_out, ch_off, ST); }
int sfk_upsample_concat_bwd(const void* dout, void* din, int B, int h, int w, int C, int H, int W, long long ld_out,
                            int ch_off, sfk_stream_t stream) {
  return upsample_concat_bwd(dout, din, B, h, w, C, H, W, ld_out, ch_off, ST);
}
int sfk_upsample_add4(const void* z0, const void* z1, int h1, int w1, const void* z2, int h2, int w2, const void* z3, int h3,
                      int w3, const float* bias, void* out, int B, int H, int W, int C, sfk_stream_t stream) {
  return upsample_add4(z0, z1, h1, w1, z2, h2, w2, z3, h3, w3, bias, out, B, H, W, C, ST);
}
int sfk_add_sample_bias(void* out, const float* vec, long long rows, int rows_per_sample, int C, sfk_stream_t stream) {
  return add_sample_bias(out, vec, rows, rows_per_sample, C, ST);
}
size_t sfk_bn_stats_workspace_bytes(long long M, int C) { return bn_stats_workspace_bytes(M, C); }
int sfk_bn_batch_stats(const void* x, long long M, int C, float eps, float momentum, float* mean, float* rstd,
                       float* running_mean, float* running_var, float* workspace, size_t workspace_bytes,
                       sfk_stream_t stream) {
  return bn_batch_stats(x, M, C, eps, momentum, mean, rstd, running_mean, running_var, workspace, workspace_bytes, ST);
}
int sfk_bn_eval_stats(const float* running_mean, const float* running_var, float eps, float* mean, float* rstd, int C,
                      sfk_stream_t stream) { return bn_eval_stats(running_mean, running_var, eps, mean, rstd, C, ST); }
int sfk_bn_relu_cls_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const float* wc, const float* bc, float* logits, long long M, int C, int NL, int pix_per_sample,
                        float drop_p, long long seed, sfk_stream_t stream) {
  return bn_relu_cls_fwd(x, mean, rstd, gamma, beta, wc, bc, logits, M, C, NL, pix_per_sample, drop_p,
                         (unsigned long long)seed, ST);
}
int sfk_bn_relu_cls_bwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const float* wc, const float* dlogits, float* sums, float* dwc, float* dbc, void* dx, long long M,
                        int C, int NL, int pix_per_sample, float drop_p, long long seed, int train_stats,
                        sfk_stream_t stream) {
  return bn_relu_cls_bwd(x, mean, rstd, gamma, beta, wc, dlogits, sums, dwc, dbc, dx, M, C, NL, pix_per_sample, drop_p,
                         (unsigned long long)seed, train_stats, ST);
}
int sfk_upsample_nchw_fwd(const float* x, float* y, long long planes, int h, int w, int H, int W, int act,
                          sfk_stream_t stream) { return upsample_nchw_fwd(x, y, planes, h, w, H, W, act, ST); }
int sfk_upsample_nchw_bwd(const float* dy, const float* yact, float* dx, long long planes, int h, int w, int H, int W,
                          int accumulate, sfk_stream_t stream) {
  return upsample_nchw_bwd(dy, yact, dx, planes, h, w, H, W, accumulate, ST);
}
size_t sfk_dice_workspace_bytes(int N) { return dice_workspace_bytes(N); }
int sfk_dice_fwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, int N, long long P,
                 int hard_threshold, float smooth, const float* sample_weight, float* workspace, float* sums, float* loss,
                 sfk_stream_t stream) {
  return dice_fwd(p, p_stride, g, g_is_i64, g_stride, N, P, hard_threshold, smooth, sample_weight, workspace, sums, loss, ST);
}
int sfk_dice_bwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, const float* sums,
                 const float* dloss, float smooth, const float* sample_weight, float* dp, long long dp_stride, int N,
                 long long P, int accumulate, sfk_stream_t stream) {
  return dice_bwd(p, p_stride, g, g_is_i64, g_stride, sums, dloss, smooth, sample_weight, dp, dp_stride, N, P, accumulate, ST);
}
size_t sfk_loss_workspace_bytes(void) { return loss_workspace_bytes(); }
int sfk_bce_valid_fwd(const float* z, const void* labels_i64, long long ignore_index, long long n, float* workspace,
                      float* out, sfk_stream_t stream) {
  return bce_valid_fwd(z, (const long long*)labels_i64, ignore_index, n, workspace, out, ST);
}
int sfk_bce_valid_bwd(const float* z, const void* labels_i64, long long ignore_index, const float* dloss, float* dz,
                      long long n, sfk_stream_t stream) {
  return bce_valid_bwd(z, (const long long*)labels_i64, ignore_index, dloss, dz, n, ST);
}
int sfk_ce_ignore_fwd(const float* z, const void* labels_i64, long long ignore_index, int N, int L, long long P,
                      float* workspace, float* out, sfk_stream_t stream) {
  return ce_ignore_fwd(z, (const long long*)labels_i64, ignore_index, N, L, P, workspace, out, ST);
}
int sfk_ce_ignore_bwd(const float* z, const void* labels_i64, long long ignore_index, const float* out, const float* dloss,
                      float* dz, int N, int L, long long P, sfk_stream_t stream) {
  return ce_ignore_bwd(z, (const long long*)labels_i64, ignore_index, out, dloss, dz, N, L, P, ST);
}
int sfk_mse_fwd(const float* a, const float* b, long long n, float denom, float* workspace, float* out, sfk_stream_t stream) {
  return mse_fwd(a, b, n, denom, workspace, out, ST);
}
int sfk_mse_bwd(const float* a, const float* b, const float* dloss, float denom, float* da, float* db, long long n,
                sfk_stream_t stream) { return mse_bwd(a, b, dloss, denom, da, db, n, ST); }
size_t sfk_pseudo_workspace_bytes(int N) { return pseudo_workspace_bytes(N); }
int sfk_pseudo_label_threshold(const float* p, void* pseudo_i64, float* stats, int N, long long P, float thr_hi,
                               float thr_lo, float conf_thr, float min_pixels, int allow_throw, float* workspace,
                               sfk_stream_t stream) {
  return pseudo_label_threshold(p, (long long*)pseudo_i64, stats, N, P, thr_hi, thr_lo, conf_thr, min_pixels, allow_throw,
                                workspace, ST);
}
int sfk_denoise_labels(const float* t0, const float* gt, float* out, long long n, sfk_stream_t stream) {
  return denoise_labels(t0, gt, out, n, ST);
}
int sfk_cosine_mean_fwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, float* stats, float* out,
                        sfk_stream_t stream) {
  return cosine_mean_fwd(a, lda, b, ldb, n, d, stats, out, ST);
}
int sfk_cosine_mean_bwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, const float* stats,
                        const float* dloss, float* da, long long ldda, float* db, long long lddb, sfk_stream_t stream) {
  return cosine_mean_bwd(a, lda, b, ldb, n, d, stats, dloss, da, ldda, db, lddb, ST);
}
int sfk_tile_augment_u8(const void* src, long long batch_stride, long long row_stride, int pix_stride, int Hs, int Ws,
                        const int* params, int crop, float* out, int B, int OH, int OW, const void* mask,
                        long long m_batch_stride, long long m_row_stride, int m_pix_stride, float* mask_out, int* minmax,
                        sfk_stream_t stream) {
  return tile_augment_u8(src, batch_stride, row_stride, pix_stride, Hs, Ws, params, crop, out, B, OH, OW, mask, m_batch_stride,
                         m_row_stride, m_pix_stride, mask_out, minmax, ST);
}
int sfk_clip_adam_l2(float* p, const float* g, float* m, float* v, void* w_bf16, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float clip, int step, sfk_stream_t stream) {
  return clip_adam_l2(p, g, m, v, w_bf16, n, lr, beta1, beta2, eps, weight_decay, clip, step, ST);
}
int sfk_ema_lerp(float* t, const float* s, void* t_bf16, long long n, float decay, sfk_stream_t stream) {
  return ema_lerp(t, s, t_bf16, n, decay, ST);
}
#undef ST

}  // extern "C"
