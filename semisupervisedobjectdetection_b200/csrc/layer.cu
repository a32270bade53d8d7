// One SegformerLayer (MS:320-363) forward / backward as a single host call that enqueues the whole kernel
// sequence on the caller's stream:
//   LN1 -> Q | (SR conv -> LN) -> KV -> fused attention -> O-proj (+drop-path, +residual) -> split prompt rows ->
//   LN2 -> dense1 -> depthwise3x3+GELU -> dense2 (+drop-path, +residual)
// Prompt / cls rows ("front" rows, MS:446-465) are carried in front of the spatial tokens: T = P + N.
// The residual stream is fp32, GEMM operands bf16, accumulation fp32.  Activations needed by backward live in a
// caller-provided `saved` arena; everything else in a reusable `scratch` arena.  Parameter gradients are accumulated
// straight into the caller's fp32 gradient buffers (split-K red.global.add), there is no separate "+=" pass.
#include <map>
#include <mutex>

#include "common.cuh"
#include "sfk_internal.h"
#include "../../include/sfk.h"

namespace sfk {

// sfk_set_option("sr_tma_patch", 0|1): read the SR conv's patches through a rank-5 TMA view (default) or gather them first
int g_sr_tma_patch = 1;
// sfk_set_option("sr_patch_store", 0|1): the SR conv's input-gradient GEMM stores through the patch view (default) or writes the
// patch-matrix gradient and runs col2im
int g_sr_patch_store = 1;
// sfk_set_option("fwd_kv_overlap", 0|1): forward runs the K/V branch (SR conv -> LayerNorm -> K|V projection) on an auxiliary
// stream next to the Q projection.  The SR conv's GEMM has only M / 128 x 2 tiles (64 CTAs at 16 tiles of 512 x 512): alone it
// leaves 84 of the 148 SMs idle for its whole 20-k-block main loop.  The Q projection's persistent grid is capped to exactly
// those SMs (GEMM CTAs own an SM's shared memory, so the two kernels are co-resident iff their grids add up to <= 148).
int g_fwd_kv_overlap = 1;
// (The same trick in backward - the K/V chain next to the dq cast - was built and measured: 47.5 vs 47.3 ms/step, no gain; it
// needed a second bank of the column-sum slot area, which colsum_slots.cuh keeps for kernels launched on an auxiliary stream.)

namespace {

struct Arena {
  uint8_t* base;
  size_t off = 0, cap;
  Arena(void* b, size_t c) : base(static_cast<uint8_t*>(b)), cap(c) {}
  template <typename T> T* take(size_t n) {
    off = (off + 255) & ~size_t(255);
    T* p = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return p;
  }
};

struct Dims {
  long long B, H, W, C, heads, sr, hidden, P;
  long long N, T, Nkv, Tk, D;
};

Dims make_dims(const sfk_layer_dims& d) {
  Dims x;
  x.B = d.B; x.H = d.H; x.W = d.W; x.C = d.C; x.heads = d.heads; x.sr = d.sr; x.hidden = d.hidden; x.P = d.ptl;
  x.N = x.H * x.W;
  x.T = x.P + x.N;
  x.Nkv = x.sr > 1 ? (x.H / x.sr) * (x.W / x.sr) : x.N;
  x.Tk = x.P + x.Nkv;
  x.D = x.C / x.heads;
  return x;
}

struct Saved {
  float* xcat;      // [B,T,C] fp32 (only when P > 0; otherwise the caller's x is used)
  float *mean1, *rstd1, *mean2, *rstd2, *mean_sr, *rstd_sr;
  __nv_bfloat16* xn1;   // [B*T, C]
  __nv_bfloat16* q;     // [B*T, C]
  float* sr_out;        // [B*Tk, C] fp32 (sr > 1)
  __nv_bfloat16* kvn;   // [B*Tk, C] (sr > 1; else aliases xn1)
  __nv_bfloat16* kv;    // [B*Tk, 2C]
  __nv_bfloat16* ctx;   // [B*T, C]
  float* lse;           // [B*heads*T]
  float* x1;            // [B,T,C] fp32
  __nv_bfloat16* xn2;   // [B*N, C]
  __nv_bfloat16* h1;    // [B*N, hidden]
  __nv_bfloat16* g;     // [B*N, hidden]
};

Saved carve_saved(Arena& a, const Dims& d) {
  Saved s;
  s.xcat = d.P > 0 ? a.take<float>(d.B * d.T * d.C) : nullptr;
  s.mean1 = a.take<float>(d.B * d.T);
  s.rstd1 = a.take<float>(d.B * d.T);
  s.mean2 = a.take<float>(d.B * d.N);
  s.rstd2 = a.take<float>(d.B * d.N);
  s.xn1 = a.take<__nv_bfloat16>(d.B * d.T * d.C);
  s.q = a.take<__nv_bfloat16>(d.B * d.T * d.C);
  if (d.sr > 1) {
    s.mean_sr = a.take<float>(d.B * d.Tk);
    s.rstd_sr = a.take<float>(d.B * d.Tk);
    s.sr_out = a.take<float>(d.B * d.Tk * d.C);
    s.kvn = a.take<__nv_bfloat16>(d.B * d.Tk * d.C);
  } else {
    s.mean_sr = s.rstd_sr = nullptr;
    s.sr_out = nullptr;
    s.kvn = s.xn1;
  }
  s.kv = a.take<__nv_bfloat16>(d.B * d.Tk * 2 * d.C);
  s.ctx = a.take<__nv_bfloat16>(d.B * d.T * d.C);
  s.lse = a.take<float>(d.B * d.heads * d.T);
  s.x1 = a.take<float>(d.B * d.T * d.C);
  s.xn2 = a.take<__nv_bfloat16>(d.B * d.N * d.C);
  s.h1 = a.take<__nv_bfloat16>(d.B * d.N * d.hidden);
  s.g = a.take<__nv_bfloat16>(d.B * d.N * d.hidden);
  return s;
}

int check_dims(const Dims& d) {
  SFK_REQUIRE(d.B > 0 && d.H > 0 && d.W > 0 && d.C > 0 && d.heads > 0 && d.hidden > 0 && d.sr >= 1 && d.P >= 0,
              "layer: bad dims");
  SFK_REQUIRE(d.C % d.heads == 0 && (d.D == 32 || d.D == 64), "layer: head dim %lld unsupported (32 or 64)", d.D);
  SFK_REQUIRE(d.C % 8 == 0 && d.hidden % 8 == 0, "layer: C and hidden must be multiples of 8");
  SFK_REQUIRE(d.sr == 1 || (d.H % d.sr == 0 && d.W % d.sr == 0), "layer: H,W must be divisible by the reduction ratio");
  return 0;
}

RowMap spatial_map(const Dims& d) {  // compact spatial row (b*N + t) -> row in a [B, P+N] sequence
  RowMap m;
  if (d.P > 0) { m.period = d.N; m.extra = d.P; m.offset = d.P; }
  return m;
}
RowMap front_map(long long P, long long rest) {  // compact front row (b*P + t) -> row in a [B, P+rest] sequence
  RowMap m;
  m.period = P; m.extra = rest; m.offset = 0;
  return m;
}
RowMap kv_spatial_map(const Dims& d) {  // compact reduced row (b*Nkv + t) -> row in [B, P+Nkv]
  RowMap m;
  if (d.P > 0) { m.period = d.Nkv; m.extra = d.P; m.offset = d.P; }
  return m;
}

// Weight-gradient GEMMs are off the critical path of backward (nothing downstream in the layer consumes them), so they
// can run on a second stream and fill the bubbles of the latency-bound dgrad chain.  fork(): side stream waits for what
// the main stream has produced so far.  The caller alternates between two scratch arenas (parity) on consecutive calls;
// a call first makes the main stream wait for the side work that last used its arena.
struct SideCtx {
  cudaStream_t main = nullptr, side = nullptr;
  cudaEvent_t fork_ev[8] = {nullptr};
  cudaEvent_t done_ev[2] = {nullptr, nullptr};
  bool done_valid[2] = {false, false};
  int next_fork = 0;
};
static SideCtx g_side;

static int side_begin(cudaStream_t main, cudaStream_t side, int parity) {
  if (side == nullptr) return 0;
  if (g_side.done_ev[0] == nullptr) {
    for (auto& e : g_side.fork_ev) SFK_REQUIRE(cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess, "event create failed");
    for (auto& e : g_side.done_ev) SFK_REQUIRE(cudaEventCreateWithFlags(&e, cudaEventDisableTiming) == cudaSuccess, "event create failed");
  }
  g_side.main = main;
  g_side.side = side;
  if (g_side.done_valid[parity])  // side work of the call that last used this scratch arena must be finished
    SFK_REQUIRE(cudaStreamWaitEvent(main, g_side.done_ev[parity], 0) == cudaSuccess, "stream wait failed");
  return 0;
}
// returns the stream wgrad work should be enqueued on (side stream after a fork, else the main stream)
static cudaStream_t side_fork(cudaStream_t main, cudaStream_t side) {
  if (side == nullptr) return main;
  cudaEvent_t e = g_side.fork_ev[g_side.next_fork];
  g_side.next_fork = (g_side.next_fork + 1) & 7;
  if (cudaEventRecord(e, main) != cudaSuccess || cudaStreamWaitEvent(side, e, 0) != cudaSuccess) return main;
  return side;
}
static int side_end(cudaStream_t side, int parity) {
  if (side == nullptr) return 0;
  SFK_REQUIRE(cudaEventRecord(g_side.done_ev[parity], side) == cudaSuccess, "event record failed");
  g_side.done_valid[parity] = true;
  return 0;
}

// auxiliary stream + fork / join events per caller stream (created on first use; the library allocates no device memory)
struct FwdAux {
  cudaStream_t s = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
static std::mutex g_aux_mu;
static std::map<cudaStream_t, FwdAux> g_aux;
static FwdAux* fwd_aux(cudaStream_t main) {
  std::lock_guard<std::mutex> lock(g_aux_mu);
  auto it = g_aux.find(main);
  if (it != g_aux.end()) return &it->second;
  FwdAux a;
  if (cudaStreamCreateWithFlags(&a.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
  if (cudaEventCreateWithFlags(&a.fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&a.join, cudaEventDisableTiming) != cudaSuccess)
    return nullptr;
  register_aux_stream(a.s);   // its slot-using kernels take the second column-sum bank (colsum_slots.cuh)
  return &(g_aux[main] = a);
}

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

PatchView sr_patch_view(const Dims& d) {
  PatchView v;
  v.enabled = g_sr_tma_patch;
  v.B = (int)d.B; v.Ho = (int)(d.H / d.sr); v.Wo = (int)(d.W / d.sr); v.sr = (int)d.sr; v.C = (int)d.C;
  v.batch_stride = d.T * d.C;
  return v;
}
GemmDesc linear_fwd(const void* a, long long lda, const void* w, long long ldw, long long M, long long N, long long K,
                    const float* bias, void* out, long long ldo, int out_bf16) {
  GemmDesc g;
  g.M = (int)M; g.N = (int)N; g.K = (int)K;
  g.a = a; g.lda = lda; g.b = w; g.ldb = ldw;
  g.bias = bias; g.out = out; g.ldo = ldo; g.out_bf16 = out_bf16;
  return g;
}
// dX[M, K_in] = dY[M, N_out] · W[N_out, K_in]
GemmDesc linear_dgrad(const void* dy, long long lddy, const void* w, long long ldw, long long M, long long n_out, long long k_in,
                      void* out, long long ldo, int out_bf16) {
  GemmDesc g;
  g.M = (int)M; g.N = (int)k_in; g.K = (int)n_out;
  g.a = dy; g.lda = lddy; g.b = w; g.ldb = ldw; g.b_mn = 1;
  g.out = out; g.ldo = ldo; g.out_bf16 = out_bf16;
  return g;
}
// dW[N_out, K_in] += dY[tokens, N_out]^T · X[tokens, K_in]
GemmDesc linear_wgrad(const void* dy, long long lddy, const void* x, long long ldx, long long tokens, long long n_out,
                      long long k_in, float* dw, long long lddw) {
  GemmDesc g;
  g.M = (int)n_out; g.N = (int)k_in; g.K = (int)tokens;
  g.a = dy; g.lda = lddy; g.a_mn = 1; g.b = x; g.ldb = ldx; g.b_mn = 1;
  g.out = dw; g.ldo = lddw; g.atomic_add = 1;
  return g;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
size_t layer_saved_bytes(const sfk_layer_dims& dd) {
  Dims d = make_dims(dd);
  Arena a(nullptr, 0);
  carve_saved(a, d);
  return a.off + 256;
}

size_t layer_scratch_bytes(const sfk_layer_dims& dd) {
  Dims d = make_dims(dd);
  // generous upper bound covering forward (im2col) and backward temporaries
  size_t tok = (size_t)d.B * d.T, sp = (size_t)d.B * d.N, kv = (size_t)d.B * d.Tk;
  size_t b = 0;
  b += sp * d.hidden * 2 * 3;                 // dg, dpre, dh1
  b += sp * d.C * 2 * 2;                      // dmlp, dxn2
  b += tok * d.C * (4 + 2 + 2 + 4 + 2 + 4 + 2 + 4 + 2);  // dx1, da, dctx, dq(f32), dq(bf16), dxn1(f32), dxn1 total(bf16), dxcat, dxn1 SR part (bf16)
  b += kv * 2 * d.C * (4 + 2);                // dkv f32 + bf16
  b += kv * d.C * (4 + 2 + 2 + 2);            // dkvn, dsr, dsr compact
  b += (size_t)d.B * d.heads * d.T * 4;       // delta
  if (d.sr > 1) b += (size_t)d.B * d.Nkv * d.sr * d.sr * d.C * 2 * 2 + (size_t)d.C * d.sr * d.sr * d.C * 4;  // A_sr, dA_sr, dWsr
  b += (size_t)10 * d.hidden * 4;             // depthwise weight-grad staging
  return b + 64 * 256;
}

// ------------------------------------------------------------------------------------------------
int layer_forward(const sfk_layer_dims& dd, const sfk_layer_params& p, const float* x, const float* front, float* out,
                  float* front_out, void* saved_buf, size_t saved_bytes, void* scratch_buf, size_t scratch_bytes,
                  cudaStream_t st) {
  const Dims d = make_dims(dd);
  TRY(check_dims(d));
  SFK_REQUIRE((d.P > 0) == (front != nullptr), "layer_forward: front rows must be given iff ptl > 0");
  Arena sa(saved_buf, saved_bytes);
  Saved s = carve_saved(sa, d);
  SFK_REQUIRE(sa.off <= saved_bytes, "layer_forward: saved arena too small (%zu < %zu)", saved_bytes, sa.off);
  Arena sc(scratch_buf, scratch_bytes);
  const long long C = d.C, rows_all = d.B * d.T, rows_sp = d.B * d.N, rows_kv = d.B * d.Tk;
  const RowMap id;

  const float* xcat = x;
  if (d.P > 0) {
    TRY(concat_front(front, d.P * C, x, s.xcat, (int)d.B, (int)d.P, (int)d.N, (int)C, st));
    xcat = s.xcat;
  }
  // LN1 over all rows
  TRY(layernorm_fwd(xcat, 0, id, p.ln1_g, p.ln1_b, s.xn1, 1, s.mean1, s.rstd1, rows_all, (int)C, 1e-5f, st));
  // The K/V branch (SR conv -> LayerNorm -> K|V projection) depends only on xn1: with a spatial reduction it runs on an
  // auxiliary stream next to the Q projection (see g_fwd_kv_overlap) and is joined before the attention kernel.
  GemmDesc gq = linear_fwd(s.xn1, C, p.wq, C, rows_all, C, C, p.bq, s.q, C, 1);
  cudaStream_t kvs = st;
  FwdAux* aux = nullptr;
  GemmDesc gsr;
  __nv_bfloat16* A = nullptr;
  bool sr_gather = false;
  if (d.sr > 1) {
    const long long Ksr = d.sr * d.sr * C;
    A = sc.take<__nv_bfloat16>(d.B * d.Nkv * Ksr);
    SFK_REQUIRE(sc.off <= scratch_bytes, "layer_forward: scratch arena too small");
    // k == stride conv as a GEMM over the patch view of xn1: TMA tiles the rank-5 strided tensor directly (implicit
    // im2col); the gather kernel is only the fallback for grids whose ox-rows do not tile a 128-row block
    const PatchView pv = sr_patch_view(d);
    if (patch_view_ok(pv, 128)) {
      gsr = linear_fwd(s.xn1 + d.P * C, Ksr, p.wsr, Ksr, d.B * d.Nkv, C, Ksr, p.bsr, s.sr_out, C, 0);
      gsr.a_patch = pv;
    } else {
      sr_gather = true;
      gsr = linear_fwd(A, Ksr, p.wsr, Ksr, d.B * d.Nkv, C, Ksr, p.bsr, s.sr_out, C, 0);
    }
    gsr.out_map = kv_spatial_map(d);
    if (g_fwd_kv_overlap) {
      int plan[8];
      const int free_sms = gemm_plan(gsr, plan) == 0 ? num_sms() - plan[3] : 0;   // plan[3] = the SR GEMM's grid
      if (free_sms >= num_sms() / 4 && (aux = fwd_aux(st)) != nullptr) {
        SFK_REQUIRE(cudaEventRecord(aux->fork, st) == cudaSuccess && cudaStreamWaitEvent(aux->s, aux->fork, 0) == cudaSuccess,
                    "layer_forward: stream fork failed");
        kvs = aux->s;
        gq.max_ctas = free_sms;
      }
    }
  }
  // Q
  TRY(gemm_bf16(gq, st));
  // spatial reduction
  if (d.sr > 1) {
    if (sr_gather) TRY(im2col_nhwc(s.xn1, A, (int)d.B, (int)d.H, (int)d.W, (int)C, (int)d.sr, (int)d.sr, 0, d.T, d.P, kvs));
    TRY(gemm_bf16(gsr, kvs));
    if (d.P > 0)  // prompt rows bypass the conv but do go through the SR LayerNorm (MS:194-205)
      TRY(scale_cast(s.xn1, 1, front_map(d.P, d.N), s.sr_out, 0, front_map(d.P, d.Nkv), nullptr, 1, d.B * d.P, (int)C, kvs));
    TRY(layernorm_fwd(s.sr_out, 0, id, p.lnsr_g, p.lnsr_b, s.kvn, 1, s.mean_sr, s.rstd_sr, rows_kv, (int)C, 1e-5f, kvs));
  }
  // K | V in one GEMM
  TRY(gemm_bf16(linear_fwd(s.kvn, C, p.wkv, C, rows_kv, 2 * C, C, p.bkv, s.kv, 2 * C, 1), kvs));
  if (kvs != st)
    SFK_REQUIRE(cudaEventRecord(aux->join, kvs) == cudaSuccess && cudaStreamWaitEvent(st, aux->join, 0) == cudaSuccess,
                "layer_forward: stream join failed");
  // fused attention
  {
    AttentionDesc a;
    a.B = (int)d.B; a.heads = (int)d.heads; a.T = (int)d.T; a.Tk = (int)d.Tk; a.D = (int)d.D;
    a.scale = 1.0f / sqrtf((float)d.D);
    a.q = s.q; a.ldq = C; a.q_batch_stride = d.T * C;
    a.k = s.kv; a.v = s.kv + C; a.ldkv = 2 * C; a.kv_batch_stride = d.Tk * 2 * C;
    a.o = s.ctx; a.ldo = C; a.o_batch_stride = d.T * C; a.lse = s.lse;
    TRY(attention_fwd(a, st));
  }
  // output projection + drop-path + residual (MS:346-351)
  {
    GemmDesc g = linear_fwd(s.ctx, C, p.wo, C, rows_all, C, C, p.bo, s.x1, C, 0);
    g.residual = xcat; g.ldr = C;
    g.row_scale = p.drop_scale_attn; g.rows_per_scale = (int)d.T;
    TRY(gemm_bf16(g, st));
  }
  // LN2 on the spatial rows only (MS:352-355)
  TRY(layernorm_fwd(s.x1, 0, spatial_map(d), p.ln2_g, p.ln2_b, s.xn2, 1, s.mean2, s.rstd2, rows_sp, (int)C, 1e-5f, st));
  // MixFFN
  TRY(gemm_bf16(linear_fwd(s.xn2, C, p.w1, C, rows_sp, d.hidden, C, p.b1, s.h1, d.hidden, 1), st));
  TRY(dwconv3x3(s.h1, p.dw_wt, p.dw_b, s.g, nullptr, (int)d.B, (int)d.H, (int)d.W, (int)d.hidden, 0, st));
  {
    GemmDesc g = linear_fwd(s.g, d.hidden, p.w2, d.hidden, rows_sp, C, d.hidden, p.b2, out, C, 0);
    g.residual = s.x1; g.ldr = C; g.res_map = spatial_map(d);
    g.row_scale = p.drop_scale_mlp; g.rows_per_scale = (int)d.N;
    TRY(gemm_bf16(g, st));
  }
  if (d.P > 0 && front_out != nullptr)
    TRY(scale_cast(s.x1, 0, front_map(d.P, d.N), front_out, 0, id, nullptr, 1, d.B * d.P, (int)C, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------
int layer_backward(const sfk_layer_dims& dd, const sfk_layer_params& p, const sfk_layer_grads& gr, const float* x,
                   const float* dout, const float* dfront_out, float* dx, float* dfront, void* saved_buf, size_t saved_bytes,
                   void* scratch_buf, size_t scratch_bytes, cudaStream_t st, cudaStream_t side, int parity) {
  const Dims d = make_dims(dd);
  TRY(check_dims(d));
  TRY(side_begin(st, side, parity & 1));
  Arena sa(saved_buf, saved_bytes);
  Saved s = carve_saved(sa, d);
  SFK_REQUIRE(sa.off <= saved_bytes, "layer_backward: saved arena too small");
  Arena sc(scratch_buf, scratch_bytes);
  const long long C = d.C, Hd = d.hidden, rows_all = d.B * d.T, rows_sp = d.B * d.N, rows_kv = d.B * d.Tk;
  const RowMap id;
  const float* xcat = d.P > 0 ? s.xcat : x;

  __nv_bfloat16* dmlp = sc.take<__nv_bfloat16>(rows_sp * C);
  __nv_bfloat16* dg = sc.take<__nv_bfloat16>(rows_sp * Hd);
  __nv_bfloat16* dpre = sc.take<__nv_bfloat16>(rows_sp * Hd);
  __nv_bfloat16* dh1 = sc.take<__nv_bfloat16>(rows_sp * Hd);
  __nv_bfloat16* dxn2 = sc.take<__nv_bfloat16>(rows_sp * C);
  float* dx1 = sc.take<float>(rows_all * C);
  __nv_bfloat16* da = sc.take<__nv_bfloat16>(rows_all * C);
  __nv_bfloat16* dctx = sc.take<__nv_bfloat16>(rows_all * C);
  float* dq32 = sc.take<float>(rows_all * C);
  __nv_bfloat16* dq16 = sc.take<__nv_bfloat16>(rows_all * C);
  float* dkv32 = sc.take<float>(rows_kv * 2 * C);
  __nv_bfloat16* dkv16 = sc.take<__nv_bfloat16>(rows_kv * 2 * C);
  float* delta = sc.take<float>(d.B * d.heads * d.T);
  float* dxn1 = sc.take<float>(rows_all * C);
  __nv_bfloat16* dxn1t = sc.take<__nv_bfloat16>(rows_all * C);
  SFK_REQUIRE(sc.off <= scratch_bytes, "layer_backward: scratch arena too small (%zu < %zu)", scratch_bytes, sc.off);

  // ---- MixFFN branch ----
  // drop-path backward + bf16 cast, fused with the dense2 bias gradient (column sums of dmlp)
  // (the following layer's LayerNorm-1 backward may already have produced both: dmlp_in)
  if (p.dmlp_in != nullptr) dmlp = const_cast<__nv_bfloat16*>(static_cast<const __nv_bfloat16*>(p.dmlp_in));
  else TRY(scale_cast_sum(dout, 0, id, dmlp, 1, id, p.drop_scale_mlp, (int)d.N, rows_sp, (int)C, gr.db2, nullptr, 0, st));
  TRY(gemm_bf16(linear_dgrad(dmlp, C, p.w2, Hd, rows_sp, C, Hd, dg, Hd, 1), st));
  if (gr.dw2) TRY(gemm_bf16(linear_wgrad(dmlp, C, s.g, Hd, rows_sp, C, Hd, gr.dw2, Hd), side_fork(st, side)));
  {
    // depthwise weight / bias gradients accumulate straight into the parameter gradients ([4C, 1, 3, 3] layout and [4C]):
    // no tap-major staging buffer, memset or layout pass
    const bool want = gr.ddw != nullptr;
    TRY(dwconv3x3_gelu_bwd(s.h1, p.dw_wt, p.dw_b, dg, dpre, want ? gr.ddw : nullptr, want ? gr.ddb : nullptr, (int)d.B,
                           (int)d.H, (int)d.W, (int)Hd, st, 1));
  }
  // input gradient of the depthwise conv; the dense1 bias gradient (column sums of dh1) is produced by the same pass
  TRY(dwconv3x3(dpre, p.dw_wt, nullptr, dh1, gr.db1, (int)d.B, (int)d.H, (int)d.W, (int)Hd, 1, st));
  TRY(gemm_bf16(linear_dgrad(dh1, Hd, p.w1, C, rows_sp, Hd, C, dxn2, C, 1), st));
  if (gr.dw1) TRY(gemm_bf16(linear_wgrad(dh1, Hd, s.xn2, C, rows_sp, Hd, C, gr.dw1, C), side_fork(st, side)));
  // LN2 backward + residual gradient -> dx1 (spatial rows); front rows of dx1 come from the returned cls/prompt rows
  // without prompt rows LN2's dx IS the whole dx1, so the same pass also emits the drop-path-scaled bf16 copy (+ its column
  // sums = o-proj bias gradient) the attention branch consumes
  const bool fuse_da = d.P == 0 && gr.dln2_g != nullptr && gr.dln2_b != nullptr;
  if (fuse_da) {
    TRY(layernorm_bwd_aux(dxn2, s.x1, s.mean2, s.rstd2, p.ln2_g, dout, dx1, gr.dln2_g, gr.dln2_b, rows_sp, (int)C, da,
                          p.drop_scale_attn, (int)d.T, gr.dbo, st));
  } else {
    TRY(layernorm_bwd(dxn2, 1, s.x1, 0, spatial_map(d), s.mean2, s.rstd2, p.ln2_g, dout, id, dx1, 0, spatial_map(d), gr.dln2_g,
                      gr.dln2_b, rows_sp, (int)C, st));
  }
  if (d.P > 0) {
    if (dfront_out != nullptr) {
      TRY(scale_cast(dfront_out, 0, id, dx1, 0, front_map(d.P, d.N), nullptr, 1, d.B * d.P, (int)C, st));
    } else {
      for (long long b = 0; b < d.B; ++b) {
        cudaError_t e = cudaMemsetAsync(dx1 + b * d.T * C, 0, (size_t)d.P * C * sizeof(float), st);
        SFK_REQUIRE(e == cudaSuccess, "layer_backward: memset failed");
      }
    }
  }
  // ---- attention branch ----
  if (!fuse_da) TRY(scale_cast_sum(dx1, 0, id, da, 1, id, p.drop_scale_attn, (int)d.T, rows_all, (int)C, gr.dbo, nullptr, 0, st));
  TRY(gemm_bf16(linear_dgrad(da, C, p.wo, C, rows_all, C, C, dctx, C, 1), st));
  if (gr.dwo) TRY(gemm_bf16(linear_wgrad(da, C, s.ctx, C, rows_all, C, C, gr.dwo, C), side_fork(st, side)));
  {
    // dq32 / dkv32 are accumulated with red.add by the attention kernel: its delta pre-pass zero-fills them (dq32 is [B, T, C]
    // with C == heads * D, so every element is covered) instead of two memset launches
    AttentionBwdDesc a;
    a.zero_fill = 1;
    a.B = (int)d.B; a.heads = (int)d.heads; a.T = (int)d.T; a.Tk = (int)d.Tk; a.D = (int)d.D;
    a.scale = 1.0f / sqrtf((float)d.D);
    a.q = s.q; a.ldq = C; a.q_batch_stride = d.T * C;
    a.k = s.kv; a.v = s.kv + C; a.ldkv = 2 * C; a.kv_batch_stride = d.Tk * 2 * C;
    a.o = s.ctx; a.ldo = C; a.o_batch_stride = d.T * C; a.lse = s.lse;
    a.dO = dctx; a.delta = delta;
    a.dq = dq32; a.ld_dq = C; a.dq_batch_stride = d.T * C;
    a.dk = dkv32; a.dv = dkv32 + C; a.ld_dkv = 2 * C; a.dkv_batch_stride = d.Tk * 2 * C;
    TRY(attention_bwd(a, st));
    // fp32 -> bf16 for the projection dgrad/wgrad GEMMs, fused with the q / k / v bias gradients
    TRY(scale_cast_sum(dq32, 0, id, dq16, 1, id, nullptr, 1, rows_all, (int)C, gr.dbq, nullptr, 0, st));
    if (gr.dbk != nullptr && gr.dbv != nullptr)
      TRY(scale_cast_sum(dkv32, 0, id, dkv16, 1, id, nullptr, 1, rows_kv, (int)(2 * C), gr.dbk, gr.dbv, (int)C, st));
    else
      TRY(scale_cast(dkv32, 0, id, dkv16, 1, id, nullptr, 1, rows_kv, (int)(2 * C), st));
  }
  __nv_bfloat16* dxn1_16 = nullptr;   // SR path's contribution to LN1's output gradient when the dgrad GEMM scatters it itself
  // K | V projections
  if (gr.dwk || gr.dwv) {
    cudaStream_t ws = side_fork(st, side);
    if (gr.dwk && gr.dwv == gr.dwk + C * C) {
      // key.weight / value.weight gradients adjacent in the flat gradient buffer (FusedClipAdam's layout): one [2C, C] wgrad
      TRY(gemm_bf16(linear_wgrad(dkv16, 2 * C, s.kvn, C, rows_kv, 2 * C, C, gr.dwk, C), ws));
    } else {
      if (gr.dwk) TRY(gemm_bf16(linear_wgrad(dkv16, 2 * C, s.kvn, C, rows_kv, C, C, gr.dwk, C), ws));
      if (gr.dwv) TRY(gemm_bf16(linear_wgrad(dkv16 + C, 2 * C, s.kvn, C, rows_kv, C, C, gr.dwv, C), ws));
    }
  }
  if ((gr.dbk == nullptr) != (gr.dbv == nullptr)) {  // only one of the two biases trainable: separate column sums
    if (gr.dbk) TRY(colsum(dkv16, 1, 2 * C, gr.dbk, rows_kv, (int)C, st));
    if (gr.dbv) TRY(colsum(dkv16 + C, 1, 2 * C, gr.dbv, rows_kv, (int)C, st));
  }
  if (d.sr > 1) {
    const long long Ksr = d.sr * d.sr * C, rows_red = d.B * d.Nkv;
    __nv_bfloat16* dkvn = sc.take<__nv_bfloat16>(rows_kv * C);
    __nv_bfloat16* dsr = sc.take<__nv_bfloat16>(rows_kv * C);
    __nv_bfloat16* dsr_c = d.P > 0 ? sc.take<__nv_bfloat16>(rows_red * C) : nullptr;
    __nv_bfloat16* A = sc.take<__nv_bfloat16>(rows_red * Ksr);
    __nv_bfloat16* dA = sc.take<__nv_bfloat16>(rows_red * Ksr);
    float* dwsr = sc.take<float>(C * Ksr);
    SFK_REQUIRE(sc.off <= scratch_bytes, "layer_backward: scratch arena too small (%zu < %zu)", scratch_bytes, sc.off);
    TRY(gemm_bf16(linear_dgrad(dkv16, 2 * C, p.wkv, C, rows_kv, 2 * C, C, dkvn, C, 1), st));
    // without prompt rows the LayerNorm's input gradient IS the SR conv's output gradient: its column sums (the conv's bias
    // gradient) come out of the same pass
    const bool fuse_dbsr = d.P == 0 && gr.dbsr != nullptr && gr.dlnsr_g != nullptr;
    if (fuse_dbsr) {
      TRY(layernorm_bwd_colsum(dkvn, s.sr_out, s.mean_sr, s.rstd_sr, p.lnsr_g, dsr, 1, gr.dlnsr_g, gr.dlnsr_b, rows_kv, (int)C, gr.dbsr, st));
    } else {
      TRY(layernorm_bwd(dkvn, 1, s.sr_out, 0, id, s.mean_sr, s.rstd_sr, p.lnsr_g, nullptr, id, dsr, 1, id, gr.dlnsr_g, gr.dlnsr_b,
                        rows_kv, (int)C, st));
    }
    const __nv_bfloat16* dsr_sp = dsr;
    if (d.P > 0) {
      TRY(scale_cast(dsr, 1, kv_spatial_map(d), dsr_c, 1, id, nullptr, 1, rows_red, (int)C, st));
      dsr_sp = dsr_c;
      // prompt rows feed LN1's output gradient directly
      TRY(scale_cast(dsr, 1, front_map(d.P, d.Nkv), dxn1, 0, front_map(d.P, d.N), nullptr, 1, d.B * d.P, (int)C, st));
    }
    const PatchView pv_out = sr_patch_view(d);
    if (d.P == 0 && g_sr_patch_store && patch_view_ok(pv_out, 32) && (Ksr % 64) == 0) {
      // k == stride: the conv's input gradient is a permutation of dA, so the GEMM's TMA stores scatter it straight into the
      // [B, H, W, C] gradient (bf16) - no dA buffer, no col2im pass; the Q-projection dgrad below adds it as a bf16 residual
      dxn1_16 = sc.take<__nv_bfloat16>(rows_all * C);
      SFK_REQUIRE(sc.off <= scratch_bytes, "layer_backward: scratch arena too small (%zu < %zu)", scratch_bytes, sc.off);
      GemmDesc g = linear_dgrad(dsr_sp, C, p.wsr, Ksr, rows_red, C, Ksr, dxn1_16, Ksr, 1);
      g.out_patch = pv_out;
      TRY(gemm_bf16(g, st));
    } else {
      TRY(gemm_bf16(linear_dgrad(dsr_sp, C, p.wsr, Ksr, rows_red, C, Ksr, dA, Ksr, 1), st));
      TRY(col2im_nhwc(dA, dxn1, (int)d.B, (int)d.H, (int)d.W, (int)C, (int)d.sr, (int)d.sr, 0, d.T, d.P, 0, st));
    }
    if (gr.dwsr) {
      cudaStream_t ws = side_fork(st, side);  // dsr_sp is final here; patch gather + wgrad + layout fix all run off-path
      cudaError_t e = cudaMemsetAsync(dwsr, 0, (size_t)C * Ksr * sizeof(float), ws);
      SFK_REQUIRE(e == cudaSuccess, "layer_backward: memset failed");
      const PatchView pv = sr_patch_view(d);
      if (patch_view_ok(pv, 64)) {
        GemmDesc g = linear_wgrad(dsr_sp, C, s.xn1 + d.P * C, Ksr, rows_red, C, Ksr, dwsr, Ksr);
        g.b_patch = pv;
        TRY(gemm_bf16(g, ws));
      } else {
        TRY(im2col_nhwc(s.xn1, A, (int)d.B, (int)d.H, (int)d.W, (int)C, (int)d.sr, (int)d.sr, 0, d.T, d.P, ws));
        TRY(gemm_bf16(linear_wgrad(dsr_sp, C, A, Ksr, rows_red, C, Ksr, dwsr, Ksr), ws));
      }
      TRY(conv_wgrad_ohwi_to_oihw(dwsr, gr.dwsr, (int)C, (int)C, (int)d.sr, (int)d.sr, Ksr, ws));
    }
    if (gr.dbsr && !fuse_dbsr) TRY(colsum(dsr_sp, 1, C, gr.dbsr, rows_red, (int)C, st));
  } else {
    TRY(gemm_bf16(linear_dgrad(dkv16, 2 * C, p.wkv, C, rows_kv, 2 * C, C, dxn1, C, 0), st));
  }
  // Q projection: dxn1_total = dq · Wq + dxn1
  {
    GemmDesc g = linear_dgrad(dq16, C, p.wq, C, rows_all, C, C, dxn1t, C, 1);
    if (dxn1_16 != nullptr) { g.residual = reinterpret_cast<const float*>(dxn1_16); g.res_bf16 = 1; }
    else g.residual = dxn1;
    g.ldr = C;
    TRY(gemm_bf16(g, st));
  }
  if (gr.dwq) TRY(gemm_bf16(linear_wgrad(dq16, C, s.xn1, C, rows_all, C, C, gr.dwq, C), side_fork(st, side)));
  // LN1 backward + residual gradient
  if (d.P == 0 && gr.next_dmlp != nullptr && gr.dln1_g != nullptr && gr.dln1_b != nullptr) {
    // dx is the preceding layer's output gradient: the same pass emits its drop-path-scaled bf16 copy and dense2 bias gradient
    TRY(layernorm_bwd_aux(dxn1t, xcat, s.mean1, s.rstd1, p.ln1_g, dx1, dx, gr.dln1_g, gr.dln1_b, rows_all, (int)C, gr.next_dmlp,
                          gr.next_drop_scale, (int)d.N, gr.next_db2, st));
  } else if (d.P == 0) {
    SFK_REQUIRE(gr.next_dmlp == nullptr, "layer_backward: next_dmlp needs trainable LayerNorm-1 parameters");
    TRY(layernorm_bwd(dxn1t, 1, xcat, 0, id, s.mean1, s.rstd1, p.ln1_g, dx1, id, dx, 0, id, gr.dln1_g, gr.dln1_b, rows_all, (int)C,
                      st));
  } else {
    SFK_REQUIRE(gr.next_dmlp == nullptr, "layer_backward: next_dmlp is only available without prompt / cls rows");
    float* dxcat = sc.take<float>(rows_all * C);
    SFK_REQUIRE(sc.off <= scratch_bytes, "layer_backward: scratch arena too small (%zu < %zu)", scratch_bytes, sc.off);
    TRY(layernorm_bwd(dxn1t, 1, xcat, 0, id, s.mean1, s.rstd1, p.ln1_g, dx1, id, dxcat, 0, id, gr.dln1_g, gr.dln1_b, rows_all,
                      (int)C, st));
    TRY(scale_cast(dxcat, 0, spatial_map(d), dx, 0, id, nullptr, 1, rows_sp, (int)C, st));
    if (dfront != nullptr) TRY(scale_cast(dxcat, 0, front_map(d.P, d.N), dfront, 0, id, nullptr, 1, d.B * d.P, (int)C, st));
  }
  TRY(side_end(side, parity & 1));
  return 0;
}

}  // namespace sfk

extern "C" {

size_t sfk_layer_saved_bytes(const sfk_layer_dims* dims) { return sfk::layer_saved_bytes(*dims); }
size_t sfk_layer_scratch_bytes(const sfk_layer_dims* dims) { return sfk::layer_scratch_bytes(*dims); }

int sfk_layer_fwd(const sfk_layer_dims* dims, const sfk_layer_params* params, const float* x, const float* front, float* out,
                  float* front_out, void* saved, size_t saved_bytes, void* scratch, size_t scratch_bytes, sfk_stream_t stream) {
  return sfk::layer_forward(*dims, *params, x, front, out, front_out, saved, saved_bytes, scratch, scratch_bytes,
                            static_cast<cudaStream_t>(stream));
}

int sfk_layer_bwd(const sfk_layer_dims* dims, const sfk_layer_params* params, const sfk_layer_grads* grads, const float* x,
                  const float* dout, const float* dfront_out, float* dx, float* dfront, void* saved, size_t saved_bytes,
                  void* scratch, size_t scratch_bytes, sfk_stream_t stream, sfk_stream_t wgrad_stream, int parity) {
  return sfk::layer_backward(*dims, *params, *grads, x, dout, dfront_out, dx, dfront, saved, saved_bytes, scratch, scratch_bytes,
                             static_cast<cudaStream_t>(stream), static_cast<cudaStream_t>(wgrad_stream), parity);
}

}  // extern "C"
