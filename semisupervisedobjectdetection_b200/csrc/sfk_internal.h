// Internal (C++) declarations shared between the kernel translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sfk {

// Row remap used where a [B, N, C] tensor lives inside a [B, P+N, C] one (prompt / cls rows in front):
//   mapped = row + (period > 0 ? (row / period) * extra : 0) + offset
struct RowMap {
  long long period = 0;
  long long extra = 0;
  long long offset = 0;
};

// k == stride patch view of a channels-last bf16 map x[b][y][x][c] (the SR conv's implicit im2col, MS:171-173):
//   P[m, (ky*sr + kx)*C + c] = x[b, oy*sr + ky, ox*sr + kx, c],  m = (b*Ho + oy)*Wo + ox
// For fixed ky the (kx, c) run is contiguous in memory, so the matrix is a rank-5 strided tensor TMA can tile directly:
// no gather kernel, no materialised [M, sr*sr*C] copy.  The operand pointer is x + first spatial token.
struct PatchView {
  int enabled = 0;
  int B = 0, Ho = 0, Wo = 0, sr = 0, C = 0;
  long long batch_stride = 0;   // elements between samples (tokens per sample * C, prompt rows included)
};
// rows of the patch matrix one TMA box can cover: `rows` consecutive m must be whole ox-rows inside one sample
inline bool patch_view_ok(const PatchView& v, int rows) {
  return v.enabled && v.Wo >= 1 && v.Wo <= rows && rows % v.Wo == 0 && v.Ho % (rows / v.Wo) == 0 && (v.sr * v.C) % 64 == 0;
}

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
  RowMap res_map;            // residual row = map(row)
  RowMap out_map;            // output row = map(row)
  const float* row_scale = nullptr;
  int rows_per_scale = 1;
  void* out = nullptr;
  long long ldo = 0;
  int out_bf16 = 0;
  int atomic_add = 0;
  int relu = 0;
  int splits = 0;   // 0 = auto
  int block_n = 0;  // 0 = auto
  PatchView a_patch;   // A (K-major, M = B*Ho*Wo rows, K = sr*sr*C) read through the patch view of `a`
  PatchView b_patch;   // B (MN-major: [K = B*Ho*Wo rows, N = sr*sr*C]) read through the patch view of `b`  (wgrad)
  PatchView out_patch; // bf16 output [M = B*Ho*Wo patches, N = sr*sr*C] scattered through the patch view of `out` (col2im of a
                       // k == stride conv's input gradient as part of the store: `out` is the [B, H, W, C] activation gradient)
  int res_bf16 = 0;    // `residual` points to bf16 [M, ldr] instead of fp32 (bf16 output only)
  int max_ctas = 0;    // > 0: cap on the persistent grid (leave SMs to a GEMM running concurrently on another stream)
};

// column-sum slot bank of a stream (colsum_slots.cuh): 1 for the layer's auxiliary streams (registered by layer.cu), else 0
void register_aux_stream(cudaStream_t s);
int colsum_bank(cudaStream_t s);
int gemm_bf16(const GemmDesc& g, cudaStream_t stream);
int weight_layouts_batched(const void* table_dev /* sfk_layout_entry[n], include/sfk.h */, int n, cudaStream_t stream);
// launch plan only (no device work): out = {block_n, splits, stages, grid, cta_pair, tmem_cols, smem_bytes, tiles}
int gemm_plan(const GemmDesc& g, int out[8]);

// ---- attention.cu ----
struct AttentionDesc {
  int B = 0, heads = 0, T = 0, Tk = 0, D = 0;
  float scale = 1.f;
  const void* q = nullptr;       // bf16, token t of sample b at q + b*q_batch_stride + t*ldq (+ head*D)
  long long ldq = 0, q_batch_stride = 0;
  const void* k = nullptr;       // bf16, same addressing with ldkv / kv_batch_stride
  const void* v = nullptr;
  long long ldkv = 0, kv_batch_stride = 0;
  void* o = nullptr;             // bf16 out
  long long ldo = 0, o_batch_stride = 0;
  float* lse = nullptr;          // [B, heads, T], log2 domain
};
struct AttentionBwdDesc : AttentionDesc {
  const void* dO = nullptr;      // bf16, addressed like o
  float* delta = nullptr;        // workspace [B, heads, T]
  int zero_fill = 0;             // 1: the delta pre-pass zero-fills dq and dk|dv itself (dq: heads*D columns of every row; dk|dv: one
                                 // contiguous [B, Tk, ld_dkv] block starting at dk) - the caller then skips its two memsets
  float* dq = nullptr;           // fp32, += (caller zero-fills unless zero_fill)
  long long ld_dq = 0, dq_batch_stride = 0;
  float* dk = nullptr;           // fp32, += (caller zero-fills)
  float* dv = nullptr;
  long long ld_dkv = 0, dkv_batch_stride = 0;
};
int attention_fwd(const AttentionDesc& a, cudaStream_t stream);
int attention_bwd(const AttentionBwdDesc& a, cudaStream_t stream);

// ---- rowwise.cu ----
int layernorm_fwd(const void* x, int x_bf16, RowMap in_map, const float* gamma, const float* beta, void* y, int y_bf16,
                  float* mean, float* rstd, long long rows, int C, float eps, cudaStream_t stream);
int layernorm_bwd_aux(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const float* dres,
                      void* dx, float* dgamma, float* dbeta, long long rows, int C, void* aux, const float* aux_scale, int aux_rps,
                      float* aux_colsum, cudaStream_t stream);
int layernorm_bwd_colsum(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                         int dx_bf16, float* dgamma, float* dbeta, long long rows, int C, float* dx_colsum, cudaStream_t stream);
int layernorm_bwd(const void* dy, int dy_bf16, const void* x, int x_bf16, RowMap x_map, const float* mean,
                  const float* rstd, const float* gamma, const float* dres, RowMap dres_map, void* dx, int dx_bf16,
                  RowMap dx_map, float* dgamma, float* dbeta, long long rows, int C, cudaStream_t stream);
int scale_cast(const void* x, int x_bf16, RowMap in_map, void* y, int y_bf16, RowMap out_map, const float* scale,
               int rows_per_scale, long long rows, int C, cudaStream_t stream);
int scale_cast_sum(const void* x, int x_bf16, RowMap in_map, void* y, int y_bf16, RowMap out_map, const float* scale,
                   int rows_per_scale, long long rows, int C, float* colsum0, float* colsum1, int split, cudaStream_t stream);
int colsum(const void* x, int x_bf16, long long ld, float* out, long long rows, int C, cudaStream_t stream);
int concat_front(const float* front, long long front_batch_stride, const float* x, float* out, int B, int P, int N, int C,
                 cudaStream_t stream);
int front_grad_sum(const float* g, float* dfront, int B, int P, int T, int C, cudaStream_t stream);

// ---- conv.cu ----
int dwconv3x3(const void* x, const float* wt, const float* bias, void* y, float* colsum, int B, int H, int W, int C, int mode,
              cudaStream_t stream);
int dwconv3x3_gelu_bwd(const void* x, const float* wt, const float* bias, const void* dg, void* dpre, float* dwt,
                       float* dbias, int B, int H, int W, int C, cudaStream_t stream, int dw_oihw = 0);
int im2col_nhwc(const void* x, void* A, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                long long tok_offset, cudaStream_t stream);
int im2col_nchw_f32(const float* x, void* A, int B, int Cin, int H, int W, int K, int S, int P, int Kp, cudaStream_t stream);
int col2im_nhwc(const void* dA, float* dx, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                long long tok_offset, int accumulate, cudaStream_t stream);
int conv_weight_ohwi(const float* w, void* out, int O, int I, int KH, int KW, long long ldo, cudaStream_t stream);
int conv_wgrad_ohwi_to_oihw(const float* g, float* dw, int O, int I, int KH, int KW, long long ldg, cudaStream_t stream);
int cast_2d(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols, int cols_padded,
            cudaStream_t stream);
int dw_weight_to_tap(const float* w, float* wt, int C, cudaStream_t stream);
int dw_wgrad_from_tap(const float* dwt, float* dw, int C, cudaStream_t stream);


// ---- decode.cu ----
int upsample_concat(const void* in, void* out, int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off,
                    cudaStream_t stream);
int upsample_add4(const void* z0, const void* z1, int h1, int w1, const void* z2, int h2, int w2, const void* z3, int h3, int w3,
                  const float* bias, void* out, int B, int H, int W, int C, cudaStream_t stream);
int upsample_concat_bwd(const void* dout, void* din, int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off,
                        cudaStream_t stream);
int add_sample_bias(void* out, const float* vec, long long rows, int rows_per_sample, int C, cudaStream_t stream);
size_t bn_stats_workspace_bytes(long long M, int C);
int bn_batch_stats(const void* x, long long M, int C, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                   float* running_var, float* workspace, size_t workspace_bytes, cudaStream_t stream);
int bn_eval_stats(const float* rm, const float* rv, float eps, float* mean, float* rstd, int C, cudaStream_t stream);
int bn_relu_cls_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* wc,
                    const float* bc, float* logits, long long M, int C, int NL, int pix_per_sample, float drop_p,
                    unsigned long long seed, cudaStream_t stream);
int bn_relu_cls_bwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* wc,
                    const float* dlogits, float* sums, float* dwc, float* dbc, void* dx, long long M, int C, int NL,
                    int pix_per_sample, float drop_p, unsigned long long seed, int train_stats, cudaStream_t stream);

// ---- loss.cu ----
int upsample_nchw_fwd(const float* x, float* y, long long planes, int h, int w, int H, int W, int act, cudaStream_t stream);
int upsample_nchw_bwd(const float* dy, const float* yact, float* dx, long long planes, int h, int w, int H, int W,
                      int accumulate, cudaStream_t stream);
size_t dice_workspace_bytes(int N);
int dice_fwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, int N, long long P,
             int hard_threshold, float smooth, const float* sample_weight, float* workspace, float* sums, float* loss,
             cudaStream_t stream);
int dice_bwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, const float* sums,
             const float* dloss, float smooth, const float* sample_weight, float* dp, long long dp_stride, int N, long long P,
             int accumulate, cudaStream_t stream);
size_t loss_workspace_bytes();
int bce_valid_fwd(const float* z, const long long* labels, long long ignore_index, long long n, float* workspace, float* out,
                  cudaStream_t stream);
int bce_valid_bwd(const float* z, const long long* labels, long long ignore_index, const float* dloss, float* dz, long long n,
                  cudaStream_t stream);
int ce_ignore_fwd(const float* z, const long long* labels, long long ignore_index, int N, int L, long long P, float* workspace,
                  float* out, cudaStream_t stream);
int ce_ignore_bwd(const float* z, const long long* labels, long long ignore_index, const float* out, const float* dloss,
                  float* dz, int N, int L, long long P, cudaStream_t stream);
int mse_fwd(const float* a, const float* b, long long n, float denom, float* workspace, float* out, cudaStream_t stream);
int mse_bwd(const float* a, const float* b, const float* dloss, float denom, float* da, float* db, long long n,
            cudaStream_t stream);
size_t pseudo_workspace_bytes(int N);
int pseudo_label_threshold(const float* p, long long* pseudo, float* stats, int N, long long P, float thr_hi, float thr_lo,
                           float conf_thr, float min_pixels, int allow_throw, float* workspace, cudaStream_t stream);
int denoise_labels(const float* t0, const float* gt, float* out, long long n, cudaStream_t stream);

// ---- pipeline.cu ----
int cosine_mean_fwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, float* stats, float* out,
                    cudaStream_t stream);
int cosine_mean_bwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, const float* stats,
                    const float* dloss, float* da, long long ldda, float* db, long long lddb, cudaStream_t stream);
int tile_augment_u8(const void* src, long long batch_stride, long long row_stride, int pix_stride, int Hs, int Ws, const int* params,
                    int crop, float* out, int B, int OH, int OW, const void* mask, long long m_batch_stride, long long m_row_stride,
                    int m_pix_stride, float* mask_out, int* minmax, cudaStream_t stream);

// ---- optim.cu ----
int clip_adam_l2(float* p, const float* g, float* m, float* v, void* w16, long long n, float lr, float beta1, float beta2,
                 float eps, float wd, float clip, int step, cudaStream_t stream);
int ema_lerp(float* t, const float* s, void* w16, long long n, float decay, cudaStream_t stream);

}  // namespace sfk
