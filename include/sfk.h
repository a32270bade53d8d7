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

/* Number of kernels libsfk has launched in this process (bench.py's gpu_launches). */
long long sfk_launch_count(void);
/* Live timing of the dominant kernel: CUDA events around every GEMM launch between begin and end (end synchronises). */
int sfk_profile_gemm_begin(void);
int sfk_profile_gemm_end(double* total_ms, double* total_flops, long long* launches);
/* After sfk_profile_gemm_end: one CSV line per profiled GEMM launch (M,N,K,flags,microseconds). */
int sfk_profile_gemm_dump(const char* path);

/* Tuning switches (name: default = meaning):
 *  "gemm_epi_tma": 1 = staged TMA-store / reduce-add epilogue;  "gemm_epi_bufs": 2 = staging buffers per epilogue warp;
 *  "gemm_bstat": 1 = weight tile resident in shared memory (automatic for K <= 128, on request up to K = 512);
 *  "gemm_cluster2": 1 = CTA-pair tiles (tcgen05 cta_group::2, 256-row MMA, B split across the pair): 0 never, 1 for large
 *      problems (M*N*K >= 1e11, K >= 1024), 2 whenever the shape allows;
 *  "gemm_occ2": 0 = light CTAs, two per SM, for K <= 512;  "gemm_krot": 0 = per-CTA rotated k-block order;
 *  "gemm_max_stages": 0 = cap on the operand ring depth (experiments);  "gemm_dbg": 0 = timing-only experiment bits
 *      (1 no stores, 2 no proxy fence, 4 no operand loads, 8 no MMAs, 16 no tiles; results are wrong by construction);
 *  "dw_bwd_vec": 4 = channels per thread of the depthwise-conv backward kernel (2 | 4);
 *  "attn_bwd_persist": 1 = persistent evenly split attention backward;  "attn_fwd256": 1 = single-pass forward for Tk <= 256;
 *  "attn_dbg": 0 = experiment bits (1 skip the row-maximum pass, 2 record sfk_debug_attn_trace);
 *  "sr_tma_patch": 1 = the SR conv's GEMMs read their k == stride patches through a rank-5 TMA view (implicit im2col);
 *  "sr_patch_store": 1 = the SR conv's input-gradient GEMM stores through the same view (no col2im pass);
 *  "upsample_sep": 1 = separable shared-memory up/down-sampling kernels of the folded decode head;
 *  "fwd_kv_overlap": 1 = sfk_layer_fwd runs the K/V branch (SR conv -> LayerNorm -> K|V projection) on an auxiliary stream next
 *      to the Q projection, whose grid is capped to the SMs the SR conv's GEMM leaves free (joined before attention). */
int sfk_set_option(const char* name, int value);

/* Debug aid: device buffer (>= 256 x uint64; slots 8.. hold per-tile timelines, see tools/check_gemm.py trace) receiving CTA 0's globaltimer (ns) at the phase boundaries of each following
 * GEMM launch: 0 start, 1 setup done, 2 first TMA issued, 3 first operands landed, 4 last MMA issued, 5 accumulator ready,
 * 6 first tile stored, 7 CTA done.  NULL switches it off. */
int sfk_debug_gemm_trace(void* device_u64x8);
/* Debug: after a forward attention launch with sfk_set_option("attn_dbg", 2): SM cycle counts of CTA (0,0,0) at its phase
 * boundaries (start, setup done, Q/K landed, S ready, row maxima exchanged, P written, O ready, end), copied to 16 host int64. */
int sfk_debug_attn_trace(long long* host_i64x16);

/* Launch plan sfk_gemm_bf16 would use for this shape, without touching the device (host logic, testable on CPU):
 * out8 = {block_n, splits, pipeline stages, grid (CTAs), mode bits (1 = CTA pair, 2 = B-stationary main loop,
 *         4 = TMA-store epilogue), TMEM columns per CTA, dynamic shared memory bytes, output tiles}.
 * splits / block_n: 0 = automatic, as in sfk_gemm_bf16. */
int sfk_gemm_plan(int M, int N, int K, int a_mn, int b_mn, int out_bf16, int atomic_add, int splits, int block_n, int* out8);

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
/* All weight-layout permutations of a model in one launch (after an optimiser / EMA step): `table` is a DEVICE array of n
 * entries; kind 0 = sfk_conv_weight_ohwi (src f32 OIHW [O,I,KH,KW] -> dst bf16 [O, ld], (ky,kx,c) order),
 * kind 1 = sfk_dw_weight_to_tap (src f32 [O,1,3,3] -> dst f32 [9][O]; I, KH, KW, ld unused). */
typedef struct sfk_layout_entry { const void* src; void* dst; int kind, O, I, KH, KW, pad; long long ld; } sfk_layout_entry;
int sfk_weight_layouts_batched(const sfk_layout_entry* table, int n, sfk_stream_t stream);

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

/* ---- one SegformerLayer (MS:320-363) per call ------------------------------------------------------------------
 * x: f32 [B, N, C] spatial tokens; front: f32 [B, ptl, C] prompt / cls rows placed in front (MS:446-460) or NULL.
 * out: f32 [B, N, C]; front_out: f32 [B, ptl, C] (the `cls_token` slice returned by MS:352,363).
 * Parameter gradients are accumulated (+=) directly into the f32 buffers of sfk_layer_grads (NULL = frozen, MS a14). */
typedef struct { int B, H, W, C, heads, sr, hidden, ptl; } sfk_layer_dims;
typedef struct {
  const void *wq, *wkv, *wo, *wsr, *w1, *w2;   /* bf16 shadows: [C,C], [2C,C] (key rows, value rows), [C,C], [C,(ky,kx,c)], [4C,C], [C,4C] */
  const float *bq, *bkv, *bo, *bsr, *b1, *b2;
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b, *lnsr_g, *lnsr_b;
  const float *dw_wt, *dw_b;                   /* depthwise weight tap-major [9][4C], bias [4C] */
  const float *drop_scale_attn, *drop_scale_mlp; /* per-sample drop-path multipliers mask/keep (MS:95-102) or NULL */
  const void *dmlp_in;                         /* backward only, optional: bf16 [B*H*W, C] = dout * drop_scale_mlp already produced (with db2
                                                  accumulated) by the following layer's backward through next_dmlp below */
} sfk_layer_params;
typedef struct {
  float *dwq, *dwk, *dwv, *dwo, *dwsr, *dw1, *dw2;   /* same shapes as the nn.Parameters (sr: OIHW) */
  float *dbq, *dbk, *dbv, *dbo, *dbsr, *db1, *db2;
  float *dln1_g, *dln1_b, *dln2_g, *dln2_b, *dlnsr_g, *dlnsr_b;
  float *ddw, *ddb;                                  /* depthwise [4C,1,3,3], [4C] */
  /* optional (ptl == 0): the LayerNorm-1 backward that writes dx also writes the PRECEDING layer's MixFFN input gradient
   * next_dmlp = bf16(dx * next_drop_scale[sample]) and accumulates its column sums into next_db2 (that layer's dense2 bias
   * gradient), so the preceding layer's backward can take it through sfk_layer_params.dmlp_in instead of re-reading dx */
  void *next_dmlp; const float *next_drop_scale; float *next_db2;
} sfk_layer_grads;
size_t sfk_layer_saved_bytes(const sfk_layer_dims* dims);
size_t sfk_layer_scratch_bytes(const sfk_layer_dims* dims);
int sfk_layer_fwd(const sfk_layer_dims* dims, const sfk_layer_params* params, const float* x, const float* front, float* out,
                  float* front_out, void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, sfk_stream_t stream);
/* wgrad_stream (optional, may be NULL): second stream for the weight-gradient GEMMs, which are off the critical path.
 * The caller must alternate `parity` (0/1) AND the scratch arena between consecutive calls, keep `saved` alive until the
 * side stream has drained, and make its main stream wait for wgrad_stream before consuming the parameter gradients. */
int sfk_layer_bwd(const sfk_layer_dims* dims, const sfk_layer_params* params, const sfk_layer_grads* grads, const float* x,
                  const float* dout, const float* dfront_out, float* dx, float* dfront, void* saved, size_t saved_bytes,
                  void* scratch, size_t scratch_bytes, sfk_stream_t stream, sfk_stream_t wgrad_stream, int parity);

/* ---- decode head (MS:752-782) ------------------------------------------------------------------------------ */
/* out[b,y,x,ch_off:ch_off+C] = bilinear(in[b,:,:,0:C]) (align_corners=False), bf16 channels-last; and its transpose */
int sfk_upsample_concat(const void* in, void* out, int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off,
                        sfk_stream_t stream);
int sfk_upsample_concat_bwd(const void* dout, void* din, int B, int h, int w, int C, int H, int W, long long ld_out,
                            int ch_off, sfk_stream_t stream);
/* Folded decode head (bilinear upsampling commutes with the 1x1 convs of MS:769-775):
 * out[b,y,x,:] = z0[b,y,x,:] + sum_i bilinear(z_i)[b,y,x,:] + bias[:]; z0, out: bf16 [B,H,W,C]; z_i: bf16 [B,h_i,w_i,C]; bias f32 [C] */
int sfk_upsample_add4(const void* z0, const void* z1, int h1, int w1, const void* z2, int h2, int w2, const void* z3, int h3,
                      int w3, const float* bias, void* out, int B, int H, int W, int C, sfk_stream_t stream);
/* out[row,:] (bf16) += vec[row / rows_per_sample, :] (f32): cls_proj broadcast add (MS:720-723) */
int sfk_add_sample_bias(void* out, const float* vec, long long rows, int rows_per_sample, int C, sfk_stream_t stream);
/* BatchNorm2d statistics over x bf16 [M, C]: batch mean / rstd (+ running-stat update) or eval-mode stats */
size_t sfk_bn_stats_workspace_bytes(long long M, int C);
int sfk_bn_batch_stats(const void* x, long long M, int C, float eps, float momentum, float* mean, float* rstd,
                       float* running_mean, float* running_var, float* workspace, size_t workspace_bytes,
                       sfk_stream_t stream);
int sfk_bn_eval_stats(const float* running_mean, const float* running_var, float eps, float* mean, float* rstd, int C,
                      sfk_stream_t stream);
/* logits[b,l,pix] = classifier(dropout(relu(bn(x)))) fused (MS:775-780); x bf16 [B*pix, C]; NL <= 8 */
int sfk_bn_relu_cls_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const float* wc, const float* bc, float* logits, long long M, int C, int NL, int pix_per_sample,
                        float drop_p, long long seed, sfk_stream_t stream);
/* sums: f32 [2][C] (out: dbeta, dgamma); dwc [NL,C] += ; dbc [NL] += ; dx bf16 [M, C] */
int sfk_bn_relu_cls_bwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                        const float* wc, const float* dlogits, float* sums, float* dwc, float* dbc, void* dx, long long M,
                        int C, int NL, int pix_per_sample, float drop_p, long long seed, int train_stats,
                        sfk_stream_t stream);

/* ---- wrapper tail + losses (SFM:117-125, LOSS:16-95, MS:853-868) ---------------------------------------------- */
/* y[p,:,:] = act(bilinear(x[p,:,:])) over `planes` f32 planes; act: 0 none, 1 sigmoid */
int sfk_upsample_nchw_fwd(const float* x, float* y, long long planes, int h, int w, int H, int W, int act,
                          sfk_stream_t stream);
/* dx = bilinear^T(dy * (yact ? yact*(1-yact) : 1)) */
int sfk_upsample_nchw_bwd(const float* dy, const float* yact, float* dx, long long planes, int h, int w, int H, int W,
                          int accumulate, sfk_stream_t stream);
size_t sfk_dice_workspace_bytes(int N);
/* sums f32 [N,3] = {tp, fp, fn}; target g is i64 or f32.  sample_weight == NULL: loss[0] = 1 - mean_n score_n with
 * score_n = (2|tp|+s)/(fp+fn+s) (LOSS:33-41).  sample_weight f32 [N] (north-star EMA step: keep flags of the unlabeled tiles):
 * loss[0] = (1/N) sum_n w_n (1 - score_n), so a rejected tile (w = 0) contributes neither loss nor gradient. */
int sfk_dice_fwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, int N, long long P,
                 int hard_threshold, float smooth, const float* sample_weight, float* workspace, float* sums, float* loss,
                 sfk_stream_t stream);
int sfk_dice_bwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, const float* sums,
                 const float* dloss, float smooth, const float* sample_weight, float* dp, long long dp_stride, int N,
                 long long P, int accumulate, sfk_stream_t stream);
size_t sfk_loss_workspace_bytes(void);
/* out[0] = mean over ALL pixels of BCEWithLogits(z, y) * [y >= 0 && y != ignore] (MS:863-866); out[1] = #valid */
int sfk_bce_valid_fwd(const float* z, const void* labels_i64, long long ignore_index, long long n, float* workspace,
                      float* out, sfk_stream_t stream);
int sfk_bce_valid_bwd(const float* z, const void* labels_i64, long long ignore_index, const float* dloss, float* dz,
                      long long n, sfk_stream_t stream);
/* CrossEntropyLoss(ignore_index) over z f32 [N,L,P], labels i64 [N,P]; out[0] = loss, out[1] = #non-ignored */
int sfk_ce_ignore_fwd(const float* z, const void* labels_i64, long long ignore_index, int N, int L, long long P,
                      float* workspace, float* out, sfk_stream_t stream);
int sfk_ce_ignore_bwd(const float* z, const void* labels_i64, long long ignore_index, const float* out, const float* dloss,
                      float* dz, int N, int L, long long P, sfk_stream_t stream);
/* out[0] = sum((a-b)^2) / denom  (LOSS:44-54 with denom = N*C*N) */
int sfk_mse_fwd(const float* a, const float* b, long long n, float denom, float* workspace, float* out, sfk_stream_t stream);
int sfk_mse_bwd(const float* a, const float* b, const float* dloss, float denom, float* da, float* db, long long n,
                sfk_stream_t stream);

/* ---- teacher-student glue (TS:24-56, TS:111-114, TS:202-204) ------------------------------------------------- */
size_t sfk_pseudo_workspace_bytes(int N);
/* pseudo i64 [N,P] = p >= thr_hi; stats f32 [N,5] = {sum|p|, confidence, #positive, dice_loss_n, keep} */
int sfk_pseudo_label_threshold(const float* p, void* pseudo_i64, float* stats, int N, long long P, float thr_hi,
                               float thr_lo, float conf_thr, float min_pixels, int allow_throw, float* workspace,
                               sfk_stream_t stream);
int sfk_denoise_labels(const float* t0, const float* gt, float* out, long long n, sfk_stream_t stream);

/* ---- few-shot / domain-prompt cls-token losses (main_segformer/segFormer_fewshot_learning.py:185-188, 216-226) ---------
 * out[0] = mean_i cos(a_i, b_i) over n rows of d floats (torch.nn.CosineSimilarity(dim=1), eps 1e-8);
 * stats f32 [n,3] = {<a,b>, |a|, |b|} is kept for the backward pass; da / db may be NULL. */
int sfk_cosine_mean_fwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, float* stats, float* out,
                        sfk_stream_t stream);
int sfk_cosine_mean_bwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, const float* stats,
                        const float* dloss, float* da, long long ldda, float* db, long long lddb, sfk_stream_t stream);

/* ---- training input pipeline on the device (config.py:27-38, archaeological_georgia_biostyle_dataloader.py:59-108) --------
 * src: uint8 tiles, pixel (b, y, x, c) at src + b*batch_stride + y*row_stride + x*pix_stride + c (bytes; c = 0..2 used, so an
 * RGBA PNG decode is passed as is); params i32 [B,4] = {crop_y0, crop_x0, op, 0}, op: 0 none, 1 HorizontalFlip,
 * 2 VerticalFlip, 3+k np.rot90(img, k) - the host draws them (albumentations' RandomCrop / OneOf(p=0.75) / RandomRotate90).
 * out f32 [B,3,OH,OW] = Resize(INTER_LINEAR)(Normalize(mean 0, std 255, max_pixel_value 1)(transformed crop x crop tile)).
 * mask (optional, may be NULL): uint8, channel 0 used; mask_out f32 [B,OH,OW] = (m - min m) / (max m - min m) of the
 * INTER_NEAREST-resized transformed crop; minmax: i32 [B,2] workspace. */
int sfk_tile_augment_u8(const void* src, long long batch_stride, long long row_stride, int pix_stride, int Hs, int Ws,
                        const int* params, int crop, float* out, int B, int OH, int OW, const void* mask,
                        long long m_batch_stride, long long m_row_stride, int m_pix_stride, float* mask_out, int* minmax,
                        sfk_stream_t stream);

/* ---- optimiser (SFM:32-37,158-162) and teacher refresh --------------------------------------------------------- */
/* clip_grad_value_(clip) then Adam with L2 weight decay over flat f32 buffers; step is 1-based.
 * w_bf16 (optional, same element layout): bf16 shadow of the updated weights written in the same pass. */
int sfk_clip_adam_l2(float* p, const float* g, float* m, float* v, void* w_bf16, long long n, float lr, float beta1,
                     float beta2, float eps, float weight_decay, float clip, int step, sfk_stream_t stream);
/* t = decay*t + (1-decay)*s ; t_bf16 (optional): bf16 shadow of the updated t */
int sfk_ema_lerp(float* t, const float* s, void* t_bf16, long long n, float decay, sfk_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* SFK_H_ */
