// Channels-last (token-major) convolution helpers of the SegFormer path:
//   * depthwise 3x3 (+bias) (+exact-erf GELU)                      MS:283-294, MS:310-313
//   * its backward (recompute pre-activation, dpre, dW, db, dgrad)
//   * patch gather ("im2col") for the overlapping patch-embed conv MS:123-131 and for the k==stride SR conv MS:171-173,
//     and the matching scatter/gather for their input gradients.
// All kernels move 64/128-bit vectors with the channel dimension innermost (fully coalesced).
#include "common.cuh"
#include "sfk_internal.h"
#include "colsum_slots.cuh"
#include "../../include/sfk.h"

namespace sfk {

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, stride 1, pad 1.  x: bf16 [B,H,W,C].  wt: fp32 [9][C] (tap-major copy of weight[C,1,3,3]).
// Per-thread channel sums (bias / filter gradients) are reduced across the block in shared memory and committed through
// the slotted low-contention column-sum path (colsum_slots.cuh).
// ------------------------------------------------------------------------------------------------

// 128-bit shared-memory read the compiler may not hoist out of the row loop (loop-invariant filter taps would otherwise
// be promoted back into registers, which is exactly what keeping them in shared memory is meant to avoid)
__device__ __forceinline__ float4 lds128_nohoist(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

// ------------------------------------------------------------------------------------------------
// Rolling-window depthwise 3x3 (replaces the per-position kernels above on the hot path).
// ncu on the per-position kernels: issue bound (~49 instructions per output element), not HBM bound.  This version cuts the
// instruction count ~2.5x:
//   * a thread owns 4 channels x 4 pixels (along W) and WALKS DOWN the image: every input row is loaded and unpacked once
//     and scattered into the three output rows it touches (accumulators rotate statically, loop unrolled by 3), so the
//     bf16->fp32 unpack cost is (SEG+2)/SEG = 1.5 per output instead of 4.5;
//   * all fp32 math is packed (fma.rn.f32x2 -> FFMA2, two lanes per issue slot): taps, bias, GELU polynomial;
//   * the 9 taps of the thread's 4 channels live in registers (18 packed pairs), no shared-memory reads in the loop;
//   * the next input row is prefetched (6 independent 8-byte loads) while the current one is consumed.
// A warp = 32 channel quads of one (image, row chunk, 4-pixel segment): every load/store instruction covers 256 contiguous bytes.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// q(x) = P(t) * t * exp(-x^2/2), t = 1/(1 + p|x|/sqrt2): erf(|x|/sqrt2) = 1 - q  (Abramowitz-Stegun 7.1.26, |err| <= 1.5e-7).
// Returns q, |x| and e = exp(-x^2/2) for a packed pair.
__device__ __forceinline__ p2 erfc_pair(p2 x, p2& ax, p2& e) {
  uint32_t xl, xh;
  p2_split_u(x, xl, xh);
  ax = p2_make_u(xl & 0x7fffffffu, xh & 0x7fffffffu);
  const p2 m = p2_mul(p2_mul(x, x), p2_splat(-0.72134752044448170368f));  // -x^2/2 * log2(e)
  float ml, mh, dl, dh;
  p2_split(m, ml, mh);
  e = p2_make(ex2_approx(ml), ex2_approx(mh));
  const p2 d = p2_fma(ax, p2_splat(0.3275911f * 0.70710678118654752440f), p2_splat(1.0f));
  p2_split(d, dl, dh);
  const p2 t = p2_make(rcp_approx(dl), rcp_approx(dh));
  p2 p = p2_fma(p2_splat(1.061405429f), t, p2_splat(-1.453152027f));
  p = p2_fma(p, t, p2_splat(1.421413741f));
  p = p2_fma(p, t, p2_splat(-0.284496736f));
  p = p2_fma(p, t, p2_splat(0.254829592f));
  return p2_mul(p2_mul(p, t), e);
}
// gelu(x) = 0.5 x (1 + erf(x/sqrt2)) = relu(x) - 0.5 |x| q(x)
__device__ __forceinline__ p2 gelu_pair(p2 x) {
  p2 ax, e;
  const p2 q = erfc_pair(x, ax, e);
  const p2 w = p2_mul(ax, q);
  const p2 s = p2_add(x, ax);  // 2 relu(x)
  return p2_fma(w, p2_splat(-0.5f), p2_mul(s, p2_splat(0.5f)));
}
// gelu'(x) = Phi(x) + x phi(x),  Phi(x) = 0.5 + copysign(0.5 (1 - q), x),  phi(x) = e / sqrt(2 pi)
__device__ __forceinline__ p2 gelu_grad_pair(p2 x) {
  p2 ax, e;
  const p2 q = erfc_pair(x, ax, e);
  const p2 u = p2_fma(q, p2_splat(-0.5f), p2_splat(0.5f));  // 0.5 (1 - q) >= 0
  uint32_t ul, uh, xl, xh;
  p2_split_u(u, ul, uh);
  p2_split_u(x, xl, xh);
  const p2 us = p2_make_u(ul | (xl & 0x80000000u), uh | (xh & 0x80000000u));
  const p2 phi_x = p2_mul(p2_mul(x, p2_splat(0.39894228040143267794f)), e);
  return p2_add(p2_add(us, p2_splat(0.5f)), phi_x);
}

static constexpr int DWR_SEG = 4;  // output pixels per thread along W
static constexpr int DWR_VEC = 4;  // channels per thread

// loads one (SEG+2)-pixel input row for this thread's 4 channels; `rp` points at pixel (x0-1) of that row (never dereferenced
// where the row / column is outside the image: zeros there)
__device__ __forceinline__ void dwr_load_row(uint2 (&row)[DWR_SEG + 2], const __nv_bfloat16* rp, bool vy, unsigned colmask, int C) {
#pragma unroll
  for (int j = 0; j < DWR_SEG + 2; ++j)
    row[j] = (vy && ((colmask >> j) & 1u)) ? __ldg(reinterpret_cast<const uint2*>(rp + (unsigned)(j * C))) : make_uint2(0, 0);
}

// scatter one unpacked input row into the three output rows it feeds: a_old gets filter row 2, a_mid row 1, a_new row 0
#define DWR_SCATTER(a_old, a_mid, a_new, cur)                                                                \
  {                                                                                                       \
    _Pragma("unroll") for (int j = 0; j < DWR_SEG + 2; ++j) {                                             \
      const p2 f0 = p2_from_bf16x2(cur[j].x), f1 = p2_from_bf16x2(cur[j].y);                              \
      _Pragma("unroll") for (int tt = 0; tt < 3; ++tt) {                                                  \
        const int o = j - tt;                                                                             \
        if (o >= 0 && o < DWR_SEG) {                                                                      \
          a_old[o][0] = p2_fma(wk[6 + tt][0], f0, a_old[o][0]);                                           \
          a_old[o][1] = p2_fma(wk[6 + tt][1], f1, a_old[o][1]);                                           \
          a_mid[o][0] = p2_fma(wk[3 + tt][0], f0, a_mid[o][0]);                                           \
          a_mid[o][1] = p2_fma(wk[3 + tt][1], f1, a_mid[o][1]);                                           \
          a_new[o][0] = p2_fma(wk[tt][0], f0, a_new[o][0]);                                               \
          a_new[o][1] = p2_fma(wk[tt][1], f1, a_new[o][1]);                                               \
        }                                                                                                 \
      }                                                                                                   \
    }                                                                                                     \
  }

// mode 0: y = gelu(conv(x) + bias)     mode 1: y = conv_flipped(x)  (input gradient of the conv); colsum (+)= sum_pixels y
template <int MODE>
__global__ void __launch_bounds__(256, 2) dwconv3x3_roll_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                                                                const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                                float* __restrict__ colsum, int B, int H, int W, int C, int ns) {
  pdl_sync();
  constexpr int SEG = DWR_SEG, VEC = DWR_VEC;
  __shared__ float sh[32 * VEC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 32 + lane) * VEC;
  const bool lane_ok = c0 < C;
  float csum[1][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) csum[0][v] = 0.f;
  if (lane_ok) {
    p2 wk[9][2], bs[2];
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(wt + (size_t)(MODE == 1 ? 8 - k : k) * C + c0));
      wk[k][0] = p2_make(w4.x, w4.y);
      wk[k][1] = p2_make(w4.z, w4.w);
    }
    if (MODE == 0) {
      const float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + c0));
      bs[0] = p2_make(b4.x, b4.y);
      bs[1] = p2_make(b4.z, b4.w);
    } else {
      bs[0] = bs[1] = p2_splat(0.f);
    }
    p2 cs0 = p2_splat(0.f), cs1 = p2_splat(0.f);
    // work split: the (image, 4-pixel segment) columns of this channel block form one long list of rows; every warp of
    // the grid takes an equal contiguous range of it (at most two partial columns), so the GPU is filled evenly in ONE
    // wave whatever B, H, W are
    const unsigned nseg = (unsigned)(W + SEG - 1) / SEG;
    const unsigned long long total = (unsigned long long)B * nseg * H;
    const unsigned gwarp = blockIdx.y * 8 + warp, nwarp = gridDim.y * 8;
    unsigned pos = (unsigned)(gwarp * total / nwarp);
    const unsigned pos_end = (unsigned)((gwarp + 1) * total / nwarp);
    while (pos < pos_end) {
      const unsigned col = pos / (unsigned)H;
      const int y0 = (int)(pos - col * (unsigned)H);
      int y1 = y0 + (int)(pos_end - pos);
      if (y1 > H) y1 = H;
      pos += (unsigned)(y1 - y0);
      const int b = (int)(col / nseg);
      const int x0 = (int)(col % nseg) * SEG;
      p2 a0[SEG][2], a1[SEG][2], a2[SEG][2];
#pragma unroll
      for (int o = 0; o < SEG; ++o) {
        a0[o][0] = a0[o][1] = a1[o][0] = a1[o][1] = p2_splat(0.f);
      }
      uint2 rb0[SEG + 2], rb1[SEG + 2], rb2[SEG + 2];  // row buffers: step k consumes rb(k%3) and prefetches rb((k+1)%3)
      unsigned colmask = 0;
#pragma unroll
      for (int j = 0; j < SEG + 2; ++j) colmask |= (x0 + j - 1 >= 0 && x0 + j - 1 < W) ? (1u << j) : 0u;
      const ptrdiff_t rstride = (ptrdiff_t)W * C;
      // rp: pixel (x0-1) of the NEXT input row to load;  yp: pixel x0 of the next output row to store
      const __nv_bfloat16* rp = x + ((ptrdiff_t)(b * H + y0 - 1) * W + (x0 - 1)) * C + c0;
      __nv_bfloat16* yp = y + ((ptrdiff_t)(b * H + y0) * W + x0) * C + c0;
      dwr_load_row(rb0, rp, y0 - 1 >= 0, colmask, C);
      rp += rstride;
      // one step: input row ry is in `cur`; completes output row ry-1 (a_old)
#define DWR_STEP(a_old, a_mid, a_new, cur, nxt, ry)                                                       \
  {                                                                                                       \
    dwr_load_row(nxt, rp, (ry) < y1 && (ry) + 1 < H, colmask, C);                                         \
    rp += rstride;                                                                                        \
    _Pragma("unroll") for (int o = 0; o < SEG; ++o) { a_new[o][0] = bs[0]; a_new[o][1] = bs[1]; }         \
    DWR_SCATTER(a_old, a_mid, a_new, cur)                                                                 \
    if ((ry) - 1 >= y0 && (ry) - 1 < y1) {                                                                \
      _Pragma("unroll") for (int o = 0; o < SEG; ++o) {                                                   \
        if ((colmask >> (o + 1)) & 1u) {                                                                  \
          p2 r0 = a_old[o][0], r1 = a_old[o][1];                                                          \
          if (MODE == 0) { r0 = gelu_pair(r0); r1 = gelu_pair(r1); }                                      \
          else { cs0 = p2_add(cs0, r0); cs1 = p2_add(cs1, r1); }                                          \
          *reinterpret_cast<uint2*>(yp + (unsigned)(o * C)) = make_uint2(p2_to_bf16x2(r0), p2_to_bf16x2(r1)); \
        }                                                                                                 \
      }                                                                                                   \
      yp += rstride;                                                                                      \
    }                                                                                                     \
  }
      // steps are padded to a multiple of 3 (no control-flow merge inside the rotation): the extra steps load nothing
      // (zeros) and store nothing
      for (int ry = y0 - 1; ry <= y1; ry += 3) {
        DWR_STEP(a0, a1, a2, rb0, rb1, ry)
        DWR_STEP(a1, a2, a0, rb1, rb2, ry + 1)
        DWR_STEP(a2, a0, a1, rb2, rb0, ry + 2)
      }
#undef DWR_STEP
    }
    if (MODE == 1) {
      p2_split(cs0, csum[0][0], csum[0][1]);
      p2_split(cs1, csum[0][2], csum[0][3]);
    }
  }
  if (MODE == 1 && colsum != nullptr) dw_block_commit<VEC, 1>(csum, sh, nullptr, 0, colsum, c0, C, lane_ok, ns);
}

// grid = (channel blocks, row-range blocks): exactly one wave (<= resident_per_sm blocks per SM), each warp >= ~6 rows
static dim3 dwr_grid(int B, int H, int W, int C, int vec, int resident_per_sm) {
  const int gx = (C / vec + 31) / 32;
  const long long total = (long long)B * ((W + DWR_SEG - 1) / DWR_SEG) * H;
  long long gy = (long long)num_sms() * resident_per_sm / gx;
  const long long gy_rows = (total + 6 * 8 - 1) / (6 * 8);
  if (gy > gy_rows) gy = gy_rows;
  if (gy < 1) gy = 1;
  return dim3(gx, (unsigned)gy);
}

int dwconv3x3(const void* x, const float* wt, const float* bias, void* y, float* colsum, int B, int H, int W, int C, int mode,
              cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0, "dwconv3x3: C=%d must be a multiple of 8", C);
  SFK_REQUIRE((long long)B * H * W * C < (1LL << 31), "dwconv3x3: tensor too large for 32-bit offsets");
  if ((long long)B * H * W == 0) return 0;
  const dim3 grid = dwr_grid(B, H, W, C, DWR_VEC, 2);
  if (mode == 0)
    launch_k(dwconv3x3_roll_kernel<0>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, bias, (__nv_bfloat16*)y, nullptr, B, H, W,
             C, 1);
  else
    launch_k(dwconv3x3_roll_kernel<1>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, nullptr, (__nv_bfloat16*)y, colsum, B, H, W,
             C, colsum_num_slots((long long)grid.x * 32 * DWR_VEC, grid.y));
  return check_launch("dwconv3x3_roll_kernel");
}

// Rolling-window backward stage 1 (same walk as dwconv3x3_roll_kernel): recompute pre = conv(x)+bias row by row, when
// an output row completes: dpre = dg * gelu'(pre) (stored bf16), db += dpre, dW[ky][kx] += dpre * x(window).
// Filter taps come from shared memory ([tap][lane][VEC], conflict-free vector reads) to leave registers for the packed
// gradient accumulators; three packed input rows stay live for the dW products of the row that just completed.
// VEC = channels per thread: 2 keeps the kernel at <= 128 registers (2 blocks = 16 warps per SM; the latency of the
// row / dg prefetches is then hidden), 4 halves the load instructions but needs ~250 registers (8 warps per SM).
template <int VEC> struct DwrVec;
template <> struct DwrVec<4> {
  typedef uint2 T;
  static __device__ __forceinline__ T zero() { return make_uint2(0, 0); }
  static __device__ __forceinline__ uint32_t word(const T& v, int q) { return q == 0 ? v.x : v.y; }
  static __device__ __forceinline__ T make(const p2* r) { return make_uint2(p2_to_bf16x2(r[0]), p2_to_bf16x2(r[1])); }
  static __device__ __forceinline__ void lds(const float* p, p2* w) {
    const float4 v = lds128_nohoist(p);
    w[0] = p2_make(v.x, v.y);
    w[1] = p2_make(v.z, v.w);
  }
};
template <> struct DwrVec<2> {
  typedef uint32_t T;
  static __device__ __forceinline__ T zero() { return 0u; }
  static __device__ __forceinline__ uint32_t word(const T& v, int) { return v; }
  static __device__ __forceinline__ T make(const p2* r) { return p2_to_bf16x2(r[0]); }
  static __device__ __forceinline__ void lds(const float* p, p2* w) {
    asm volatile("ld.shared.b64 %0, [%1];" : "=l"(w[0]) : "r"(smem_u32(p)));
  }
};

template <int VEC>
__global__ void __launch_bounds__(256, VEC == 2 ? 2 : 1)
dwconv3x3_gelu_bwd_roll_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt, const float* __restrict__ bias,
                               const __nv_bfloat16* __restrict__ dg, __nv_bfloat16* __restrict__ dpre, float* __restrict__ dwt,
                               float* __restrict__ dbias, int B, int H, int W, int C, int ns, int dw_kstride, int dw_cstride) {
  pdl_sync();
  constexpr int SEG = DWR_SEG, NP = VEC / 2;
  typedef DwrVec<VEC> V;
  typedef typename V::T RowT;
  __shared__ __align__(16) float sh[10 * 32 * VEC];   // filter taps during the main loop, block reduction afterwards
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = blockIdx.x * 32 * VEC;
  const int c0 = cblk + lane * VEC;
  const bool lane_ok = c0 < C;
  for (int i = threadIdx.x; i < 9 * 32 * VEC; i += 256) {
    const int k = i / (32 * VEC), j = i % (32 * VEC);
    sh[i] = (cblk + j < C) ? wt[(long long)k * C + cblk + j] : 0.f;
  }
  __syncthreads();
  float acc[10][VEC];  // acc[0..8] = dW taps, acc[9] = dbias (unpacked at the end for the block commit)
#pragma unroll
  for (int k = 0; k < 10; ++k)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
  if (lane_ok) {
    p2 gw[10][NP];
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
      for (int q = 0; q < NP; ++q) gw[k][q] = p2_splat(0.f);
    p2 bs[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) bs[q] = p2_make(__ldg(bias + c0 + 2 * q), __ldg(bias + c0 + 2 * q + 1));
    const float* swl = sh + lane * VEC;
    const unsigned nseg = (unsigned)(W + SEG - 1) / SEG;
    const unsigned long long total = (unsigned long long)B * nseg * H;
    const unsigned gwarp = blockIdx.y * 8 + warp, nwarp = gridDim.y * 8;
    unsigned pos = (unsigned)(gwarp * total / nwarp);
    const unsigned pos_end = (unsigned)((gwarp + 1) * total / nwarp);
    while (pos < pos_end) {
      const unsigned col = pos / (unsigned)H;
      const int y0 = (int)(pos - col * (unsigned)H);
      int y1 = y0 + (int)(pos_end - pos);
      if (y1 > H) y1 = H;
      pos += (unsigned)(y1 - y0);
      const int b = (int)(col / nseg);
      const int x0 = (int)(col % nseg) * SEG;
      p2 a0[SEG][NP], a1[SEG][NP], a2[SEG][NP];
#pragma unroll
      for (int o = 0; o < SEG; ++o)
#pragma unroll
        for (int q = 0; q < NP; ++q) a0[o][q] = a1[o][q] = p2_splat(0.f);
      RowT rb0[SEG + 2], rb1[SEG + 2], rb2[SEG + 2], nxt[SEG + 2], g[SEG], gn[SEG];
#pragma unroll
      for (int j = 0; j < SEG + 2; ++j) rb1[j] = rb2[j] = V::zero();
#pragma unroll
      for (int o = 0; o < SEG; ++o) g[o] = V::zero();
      unsigned colmask = 0;
#pragma unroll
      for (int j = 0; j < SEG + 2; ++j) colmask |= (x0 + j - 1 >= 0 && x0 + j - 1 < W) ? (1u << j) : 0u;
      const ptrdiff_t rstride = (ptrdiff_t)W * C;
      const __nv_bfloat16* rp = x + ((ptrdiff_t)(b * H + y0 - 1) * W + (x0 - 1)) * C + c0;   // next input row to load
      const __nv_bfloat16* gp = dg + ((ptrdiff_t)(b * H + y0 - 1) * W + x0) * C + c0;         // dg row of the next step's input row - 1
      __nv_bfloat16* op = dpre + ((ptrdiff_t)(b * H + y0) * W + x0) * C + c0;                  // next dpre row to store
#pragma unroll
      for (int j = 0; j < SEG + 2; ++j)
        rb0[j] = (y0 - 1 >= 0 && ((colmask >> j) & 1u)) ? __ldg(reinterpret_cast<const RowT*>(rp + (unsigned)(j * C))) : V::zero();
      rp += rstride;
      // one step: input row ry is in r_cur (r_p1 = row ry-1, r_p2 = row ry-2) and g = dg row ry-1; completes output row
      // ry-1 (a_old).  Prefetched for the next step: input row ry+1 (nxt) and dg row ry (gn).
#define DWB_DW_ROW(row, kbase)                                                                            \
  {                                                                                                       \
    _Pragma("unroll") for (int j = 0; j < SEG + 2; ++j) {                                                 \
      _Pragma("unroll") for (int q = 0; q < NP; ++q) {                                                    \
        const p2 fq = p2_from_bf16x2(V::word(row[j], q));                                                 \
        _Pragma("unroll") for (int tt = 0; tt < 3; ++tt) {                                                \
          const int o = j - tt;                                                                           \
          if (o >= 0 && o < SEG) gw[(kbase) + tt][q] = p2_fma(d[o][q], fq, gw[(kbase) + tt][q]);          \
        }                                                                                                 \
      }                                                                                                   \
    }                                                                                                     \
  }
#define DWB_STEP(a_old, a_mid, a_new, r_p2, r_p1, r_cur, ry)                                              \
  {                                                                                                       \
    const bool emit = (ry) - 1 >= y0 && (ry) - 1 < y1;                                                    \
    const bool more = (ry) < y1 && (ry) + 1 < H;                                                          \
    const bool gmore = (ry) >= y0 && (ry) < y1;                                                           \
    _Pragma("unroll") for (int j = 0; j < SEG + 2; ++j)                                                   \
      nxt[j] = (more && ((colmask >> j) & 1u)) ? __ldg(reinterpret_cast<const RowT*>(rp + (unsigned)(j * C))) : V::zero(); \
    _Pragma("unroll") for (int o = 0; o < SEG; ++o)                                                       \
      gn[o] = (gmore && ((colmask >> (o + 1)) & 1u)) ? __ldg(reinterpret_cast<const RowT*>(gp + (unsigned)(o * C))) : V::zero(); \
    rp += rstride;                                                                                        \
    gp += rstride;                                                                                        \
    _Pragma("unroll") for (int o = 0; o < SEG; ++o)                                                       \
      _Pragma("unroll") for (int q = 0; q < NP; ++q) a_new[o][q] = bs[q];                                 \
    {                                                                                                     \
      p2 f[SEG + 2][NP];                                                                                  \
      _Pragma("unroll") for (int j = 0; j < SEG + 2; ++j)                                                 \
        _Pragma("unroll") for (int q = 0; q < NP; ++q) f[j][q] = p2_from_bf16x2(V::word(r_cur[j], q));    \
      _Pragma("unroll") for (int tt = 0; tt < 3; ++tt) {                                                  \
        p2 wo[NP], wm[NP], wn[NP];                                                                        \
        V::lds(swl + (6 + tt) * 32 * VEC, wo);                                                            \
        V::lds(swl + (3 + tt) * 32 * VEC, wm);                                                            \
        V::lds(swl + tt * 32 * VEC, wn);                                                                  \
        _Pragma("unroll") for (int o = 0; o < SEG; ++o)                                                   \
          _Pragma("unroll") for (int q = 0; q < NP; ++q) {                                                \
            a_old[o][q] = p2_fma(wo[q], f[o + tt][q], a_old[o][q]);                                       \
            a_mid[o][q] = p2_fma(wm[q], f[o + tt][q], a_mid[o][q]);                                       \
            a_new[o][q] = p2_fma(wn[q], f[o + tt][q], a_new[o][q]);                                       \
          }                                                                                               \
      }                                                                                                   \
    }                                                                                                     \
    if (emit) {                                                                                           \
      p2 d[SEG][NP];                                                                                      \
      _Pragma("unroll") for (int o = 0; o < SEG; ++o) {                                                   \
        _Pragma("unroll") for (int q = 0; q < NP; ++q) {                                                  \
          d[o][q] = p2_mul(p2_from_bf16x2(V::word(g[o], q)), gelu_grad_pair(a_old[o][q]));                \
          gw[9][q] = p2_add(gw[9][q], d[o][q]);                                                           \
        }                                                                                                 \
        if ((colmask >> (o + 1)) & 1u) *reinterpret_cast<RowT*>(op + (unsigned)(o * C)) = V::make(d[o]);  \
      }                                                                                                   \
      DWB_DW_ROW(r_p2, 0)                                                                                 \
      DWB_DW_ROW(r_p1, 3)                                                                                 \
      DWB_DW_ROW(r_cur, 6)                                                                                \
      op += rstride;                                                                                      \
    }                                                                                                     \
    _Pragma("unroll") for (int j = 0; j < SEG + 2; ++j) r_p2[j] = nxt[j];                                 \
    _Pragma("unroll") for (int o = 0; o < SEG; ++o) g[o] = gn[o];                                         \
  }
      for (int ry = y0 - 1; ry <= y1; ry += 3) {
        DWB_STEP(a0, a1, a2, rb1, rb2, rb0, ry)
        DWB_STEP(a1, a2, a0, rb2, rb0, rb1, ry + 1)
        DWB_STEP(a2, a0, a1, rb0, rb1, rb2, ry + 2)
      }
#undef DWB_STEP
#undef DWB_DW_ROW
    }
#pragma unroll
    for (int k = 0; k < 10; ++k)
#pragma unroll
      for (int q = 0; q < NP; ++q) p2_split(gw[k][q], acc[k][2 * q], acc[k][2 * q + 1]);
  }
  if (dwt != nullptr) {
    __syncthreads();  // everyone is done with the filter taps in sh
    dw_block_commit<VEC, 10>(acc, sh, dwt, (long long)dw_kstride, dbias, c0, C, lane_ok, ns, dw_cstride);
  }
}

static int g_dw_bwd_vec = 4;  // measured on B200 (tools/check_ops.py dw_perf): 4 is 5-25% faster than 2 on the MiT-b5 shapes
void set_dw_bwd_vec(int v) { g_dw_bwd_vec = (v == 2) ? 2 : 4; }

// dw_oihw == 0: dwt is a tap-major [9][C] accumulator (public ABI); 1: dwt is the parameter gradient itself, [C, 1, 3, 3]
int dwconv3x3_gelu_bwd(const void* x, const float* wt, const float* bias, const void* dg, void* dpre, float* dwt,
                       float* dbias, int B, int H, int W, int C, cudaStream_t stream, int dw_oihw) {
  const int dw_kstride = dw_oihw ? 1 : C, dw_cstride = dw_oihw ? 9 : 1;
  SFK_REQUIRE(C % 8 == 0, "dwconv3x3_gelu_bwd: C=%d must be a multiple of 8", C);
  SFK_REQUIRE((dwt == nullptr) == (dbias == nullptr), "dwconv3x3_gelu_bwd: pass both or neither of dwt / dbias");
  SFK_REQUIRE((long long)B * H * W * C < (1LL << 31), "dwconv3x3_gelu_bwd: tensor too large for 32-bit offsets");
  if ((long long)B * H * W == 0) return 0;
  if (g_dw_bwd_vec == 4) {
    const dim3 grid = dwr_grid(B, H, W, C, 4, 1);
    launch_k(dwconv3x3_gelu_bwd_roll_kernel<4>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, bias,
             (const __nv_bfloat16*)dg, (__nv_bfloat16*)dpre, dwt, dbias, B, H, W, C,
             colsum_num_slots((long long)grid.x * 10 * 32 * 4, grid.y), dw_kstride, dw_cstride);
  } else {
    const dim3 grid = dwr_grid(B, H, W, C, 2, 2);
    launch_k(dwconv3x3_gelu_bwd_roll_kernel<2>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, bias,
             (const __nv_bfloat16*)dg, (__nv_bfloat16*)dpre, dwt, dbias, B, H, W, C,
             colsum_num_slots((long long)grid.x * 10 * 32 * 2, grid.y), dw_kstride, dw_cstride);
  }
  return check_launch("dwconv3x3_gelu_bwd_roll_kernel");
}

// ------------------------------------------------------------------------------------------------
// Patch gather for a KxK / stride S / pad P conv on a channels-last bf16 map:
//   A[m, (ky*K+kx)*C + c] = x[b, oy*S+ky-P, ox*S+kx-P, c]  (0 outside), m = (b*Ho+oy)*Wo+ox
// x rows may live inside a longer token sequence: token t of sample b is at row b*tok_stride + tok_offset + t.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ A,
                                                          int B, int H, int W, int C, int K, int S, int P, int Ho, int Wo,
                                                          long long tok_stride, long long tok_offset) {
  pdl_sync();
  // one block per output row (b, oy); threads walk (ox, tap, 8-channel chunk) with 32-bit arithmetic
  const int C8 = C >> 3, KK = K * K;
  const int oy = blockIdx.x % Ho;
  const int b = blockIdx.x / Ho;
  const int per_pix = KK * C8;
  const int n = Wo * per_pix;
  const __nv_bfloat16* xb = x + ((size_t)b * tok_stride + tok_offset) * C;
  __nv_bfloat16* Ab = A + ((size_t)b * Ho + oy) * Wo * (size_t)per_pix * 8;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int ox = i / per_pix, r = i - ox * per_pix;
    const int tap = r / C8, c8 = r - tap * C8;
    const int ky = tap / K, kx = tap - ky * K;
    const int sy = oy * S + ky - P, sx = ox * S + kx - P;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = *reinterpret_cast<const uint4*>(xb + ((size_t)sy * W + sx) * C + c8 * 8);
    *reinterpret_cast<uint4*>(Ab + (size_t)i * 8) = v;
  }
}

int im2col_nhwc(const void* x, void* A, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                long long tok_offset, cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0, "im2col_nhwc: C=%d must be a multiple of 8", C);
  const int Ho = (H + 2 * P - K) / S + 1, Wo = (W + 2 * P - K) / S + 1;
  if ((long long)B * Ho * Wo == 0) return 0;
  SFK_REQUIRE((long long)Wo * K * K * (C / 8) < (1LL << 31), "im2col_nhwc: row too large");
  launch_k(im2col_nhwc_kernel, dim3((unsigned)(B * Ho)), dim3(256), 0, stream, (const __nv_bfloat16*)x, (__nv_bfloat16*)A, B, H, W, C, K, S,
           P, Ho, Wo, tok_stride, tok_offset);
  return check_launch("im2col_nhwc_kernel");
}

// Stage-1 gather from the fp32 NCHW image: A[m, c*K*K + ky*K + kx] (OIHW-flatten order), K padded to Kp with zeros.
// One block per output row (b, oy); a thread produces 8 consecutive k of one output pixel (one 16-byte store) with 32-bit
// index arithmetic.  (The first version wrote one bf16 per thread behind two 64-bit divisions: 330 us for 80 MB.)
static constexpr int IM2COL_MAX_KP = 1024;   // padded patch length the table kernel supports (stage 1: 3 * 7 * 7 -> 152)
__global__ void __launch_bounds__(256) im2col_nchw_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ A, int B,
                                                              int Cin, int H, int W, int K, int S, int P, int Ho, int Wo,
                                                              int Kp) {
  pdl_sync();
  // per-column table (tap offset inside the sample, ky, kx), built once per block: decoding k -> (c, ky, kx) with two runtime
  // divisions for every one of the 40 M elements made this pass instruction bound (141 us for 130 MB)
  __shared__ int t_off[IM2COL_MAX_KP];
  __shared__ int t_yx[IM2COL_MAX_KP];
  const int oy = blockIdx.x % Ho;
  const int b = blockIdx.x / Ho;
  const int KK = K * K, Kreal = Cin * KK, nch = Kp >> 3;
  for (int k = threadIdx.x; k < Kp; k += 256) {
    int off = 0, yx = -1;
    if (k < Kreal) {
      const int c = k / KK, r = k - c * KK;
      const int ky = r / K, kx = r - ky * K;
      off = (c * H + ky) * W + kx;
      yx = (ky << 16) | kx;
    }
    t_off[k] = off;
    t_yx[k] = yx;
  }
  __syncthreads();
  const int n = Wo * nch;
  const float* xb = x + (size_t)b * Cin * H * W;
  __nv_bfloat16* arow = A + ((size_t)b * Ho + oy) * Wo * Kp;
  const int y0 = oy * S - P;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int ox = i / nch, ch = i - ox * nch;
    const int x0 = ox * S - P;
    const float* xp = xb + y0 * W + x0;   // (formed as an offset: only dereferenced for in-range taps)
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = ch * 8 + j;
      const int yx = t_yx[k];
      v[j] = 0.f;
      if (yx >= 0) {
        const int sy = y0 + (yx >> 16), sx = x0 + (yx & 0xffff);
        if (sy >= 0 && sy < H && sx >= 0 && sx < W) v[j] = __ldg(xp + t_off[k]);
      }
    }
    uint4 o;
    o.x = pack_bf16x2(v[0], v[1]); o.y = pack_bf16x2(v[2], v[3]);
    o.z = pack_bf16x2(v[4], v[5]); o.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(arow + (size_t)ox * Kp + ch * 8) = o;
  }
}

int im2col_nchw_f32(const float* x, void* A, int B, int Cin, int H, int W, int K, int S, int P, int Kp,
                    cudaStream_t stream) {
  const int Ho = (H + 2 * P - K) / S + 1, Wo = (W + 2 * P - K) / S + 1;
  SFK_REQUIRE(Kp % 8 == 0 && Kp >= Cin * K * K, "im2col_nchw_f32: padded K=%d must be a multiple of 8 and >= Cin*K*K", Kp);
  SFK_REQUIRE(Kp <= IM2COL_MAX_KP && K < 65536 && (long long)Cin * H * W < (1LL << 31), "im2col_nchw_f32: patch of %d values too long", Kp);
  if ((long long)B * Ho * Wo == 0) return 0;
  launch_k(im2col_nchw_f32_kernel, dim3((unsigned)(B * Ho)), dim3(256), 0, stream, x, (__nv_bfloat16*)A, B, Cin, H, W, K, S, P, Ho, Wo,
           Kp);
  return check_launch("im2col_nchw_f32_kernel");
}

// Input gradient of the patch conv: dx[b, y, x, c] (+)= sum over taps of dA[m(b,oy,ox), tap*C + c].
// Gather formulation (no atomics): every input pixel visits the <= ceil(K/S)^2 patches that contain it.
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const __nv_bfloat16* __restrict__ dA, float* __restrict__ dx, int B,
                                                          int H, int W, int C, int K, int S, int P, int Ho, int Wo,
                                                          long long tok_stride, long long tok_offset, int accumulate) {
  pdl_sync();
  // one block per input row (b, sy); threads walk (sx, 4-channel chunk) with 32-bit arithmetic
  const int C4 = C >> 2, KK = K * K;
  const int sy = blockIdx.x % H;
  const int b = blockIdx.x / H;
  const int n = W * C4;
  float* drow = dx + ((size_t)b * tok_stride + tok_offset + (size_t)sy * W) * C;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int sx = i / C4, c4 = i - sx * C4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int ky = (sy + P) % S; ky < K; ky += S) {
      const int oy = (sy + P - ky) / S;
      if (sy + P - ky < 0 || oy >= Ho) continue;
      for (int kx = (sx + P) % S; kx < K; kx += S) {
        const int ox = (sx + P - kx) / S;
        if (sx + P - kx < 0 || ox >= Wo) continue;
        const size_t m = ((size_t)b * Ho + oy) * Wo + ox;
        const uint2 u = *reinterpret_cast<const uint2*>(dA + (m * KK + ky * K + kx) * C + c4 * 4);
        const float2 a = unpack_bf16x2(u.x), bq = unpack_bf16x2(u.y);
        acc.x += a.x; acc.y += a.y; acc.z += bq.x; acc.w += bq.y;
      }
    }
    float* dst = drow + (size_t)sx * C + c4 * 4;
    if (accumulate) {
      const float4 o = *reinterpret_cast<const float4*>(dst);
      acc.x += o.x; acc.y += o.y; acc.z += o.z; acc.w += o.w;
    }
    *reinterpret_cast<float4*>(dst) = acc;
  }
}

int col2im_nhwc(const void* dA, float* dx, int B, int H, int W, int C, int K, int S, int P, long long tok_stride,
                long long tok_offset, int accumulate, cudaStream_t stream) {
  SFK_REQUIRE(C % 4 == 0, "col2im_nhwc: C=%d must be a multiple of 4", C);
  const int Ho = (H + 2 * P - K) / S + 1, Wo = (W + 2 * P - K) / S + 1;
  if ((long long)B * H * W == 0) return 0;
  launch_k(col2im_nhwc_kernel, dim3((unsigned)(B * H)), dim3(256), 0, stream, (const __nv_bfloat16*)dA, dx, B, H, W, C, K, S, P, Ho, Wo,
           tok_stride, tok_offset, accumulate);
  return check_launch("col2im_nhwc_kernel");
}

// ------------------------------------------------------------------------------------------------
// Weight re-layouts for the bf16 shadow copies (run once per optimiser step)
// ------------------------------------------------------------------------------------------------
// OIHW fp32 -> [O, (ky,kx,c)] bf16 (channels-last patch order), row pitch ldo
__global__ void conv_weight_ohwi_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ out, int O, int I, int KH,
                                        int KW, long long ldo) {
  pdl_sync();
  const long long total = (long long)O * I * KH * KW;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int c = (int)(idx % I);
  long long t = idx / I;
  const int kx = (int)(t % KW);
  t /= KW;
  const int ky = (int)(t % KH);
  const long long o = t / KH;
  out[o * ldo + ((long long)ky * KW + kx) * I + c] = __float2bfloat16(w[((o * I + c) * KH + ky) * KW + kx]);
}

int conv_weight_ohwi(const float* w, void* out, int O, int I, int KH, int KW, long long ldo, cudaStream_t stream) {
  const long long total = (long long)O * I * KH * KW;
  if (total == 0) return 0;
  launch_k(conv_weight_ohwi_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, w, (__nv_bfloat16*)out, O, I, KH, KW, ldo);
  return check_launch("conv_weight_ohwi_kernel");
}

// inverse for gradients: dW_oihw[o,c,ky,kx] += g[o, (ky,kx,c)] (fp32 -> fp32)
__global__ void conv_wgrad_ohwi_to_oihw_kernel(const float* __restrict__ g, float* __restrict__ dw, int O, int I, int KH, int KW,
                                               long long ldg) {
  pdl_sync();
  const unsigned total = (unsigned)O * (unsigned)I * (unsigned)(KH * KW);   // < 2^31 (host check); 32-bit divisions
  const unsigned idx = blockIdx.x * 256u + threadIdx.x;
  if (idx >= total) return;
  const unsigned kx = idx % (unsigned)KW;
  unsigned t = idx / (unsigned)KW;
  const unsigned ky = t % (unsigned)KH;
  t /= (unsigned)KH;
  const unsigned c = t % (unsigned)I;
  const unsigned o = t / (unsigned)I;
  dw[idx] += g[(long long)o * ldg + (ky * (unsigned)KW + kx) * (unsigned)I + c];
}

int conv_wgrad_ohwi_to_oihw(const float* g, float* dw, int O, int I, int KH, int KW, long long ldg, cudaStream_t stream) {
  const long long total = (long long)O * I * KH * KW;
  if (total == 0) return 0;
  SFK_REQUIRE(total < (1LL << 31), "conv_wgrad_ohwi_to_oihw: weight tensor too large");
  launch_k(conv_wgrad_ohwi_to_oihw_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, g, dw, O, I, KH, KW, ldg);
  return check_launch("conv_wgrad_ohwi_to_oihw_kernel");
}

// generic 2-D strided transpose-free copy fp32 -> bf16 with row pitches (for weight shadows incl. K padding)
__global__ void cast_2d_kernel(const float* __restrict__ src, long long lds, __nv_bfloat16* __restrict__ dst, long long ldd,
                               long long rows, int cols, int cols_padded) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= rows * cols_padded) return;
  const int c = (int)(idx % cols_padded);
  const long long r = idx / cols_padded;
  dst[r * ldd + c] = __float2bfloat16(c < cols ? src[r * lds + c] : 0.f);
}

int cast_2d(const float* src, long long lds, void* dst, long long ldd, long long rows, int cols, int cols_padded,
            cudaStream_t stream) {
  if (rows * cols_padded == 0) return 0;
  launch_k(cast_2d_kernel, dim3((unsigned)((rows * cols_padded + 255) / 256)), dim3(256), 0, stream, src, lds, (__nv_bfloat16*)dst, ldd, rows, cols,
                                                                                  cols_padded);
  return check_launch("cast_2d_kernel");
}

// [C,1,3,3] fp32 -> tap-major [9][C] fp32 and back (gradient accumulate)
__global__ void dw_weight_to_tap_kernel(const float* __restrict__ w, float* __restrict__ wt, int C) {
  pdl_sync();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= 9 * C) return;
  const int k = idx / C, c = idx % C;
  wt[idx] = w[c * 9 + k];
}
__global__ void dw_wgrad_from_tap_kernel(const float* __restrict__ dwt, float* __restrict__ dw, int C) {
  pdl_sync();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= 9 * C) return;
  const int c = idx / 9, k = idx % 9;
  dw[idx] += dwt[k * C + c];
}
int dw_weight_to_tap(const float* w, float* wt, int C, cudaStream_t stream) {
  launch_k(dw_weight_to_tap_kernel, dim3((9 * C + 255) / 256), dim3(256), 0, stream, w, wt, C);
  return check_launch("dw_weight_to_tap_kernel");
}
// All layout permutations of a model in ONE launch: entry e (blockIdx.y) is either an SR-conv weight (kind 0: OIHW fp32 ->
// [O, (ky,kx,c)] bf16) or a depthwise weight (kind 1: [C,1,3,3] fp32 -> tap-major [9][C] fp32).  The per-layer versions above
// cost ~100 tiny launches (and ctypes calls) per forward pass after every optimiser / EMA step.
static_assert(sizeof(::sfk_layout_entry) == 48, "sfk_layout_entry is mirrored as a 48-byte record by the Python host code");
__global__ void __launch_bounds__(256) weight_layouts_batched_kernel(const ::sfk_layout_entry* __restrict__ table, int n) {
  pdl_sync();
  const ::sfk_layout_entry e = table[blockIdx.y];
  // 32-bit index arithmetic (a weight tensor has far fewer than 2^31 elements; the host checks): the first version decoded a
  // 64-bit flat index with five long-long divisions per element and took 126 us per encoder
  const unsigned total = e.kind == 0 ? (unsigned)e.O * (unsigned)e.I * (unsigned)(e.KH * e.KW) : 9u * (unsigned)e.O;
  const unsigned I = (unsigned)e.I, KW = (unsigned)e.KW, KH = (unsigned)e.KH, O = (unsigned)e.O;
  for (unsigned idx = blockIdx.x * 256u + threadIdx.x; idx < total; idx += gridDim.x * 256u) {
    if (e.kind == 0) {
      const unsigned c = idx % I;
      unsigned t = idx / I;
      const unsigned kx = t % KW;
      t /= KW;
      const unsigned ky = t % KH;
      const unsigned o = t / KH;
      reinterpret_cast<__nv_bfloat16*>(e.dst)[(long long)o * e.ld + (ky * KW + kx) * I + c] =
          __float2bfloat16(reinterpret_cast<const float*>(e.src)[((o * I + c) * KH + ky) * KW + kx]);
    } else {
      const unsigned k = idx / O, c = idx - k * O;
      reinterpret_cast<float*>(e.dst)[idx] = reinterpret_cast<const float*>(e.src)[c * 9u + k];
    }
  }
}
int weight_layouts_batched(const void* table_dev, int n, cudaStream_t stream) {
  if (n <= 0) return 0;
  SFK_REQUIRE(n <= 65535, "weight_layouts_batched: too many entries");
  launch_k(weight_layouts_batched_kernel, dim3(32, (unsigned)n), dim3(256), 0, stream, static_cast<const ::sfk_layout_entry*>(table_dev), n);
  return check_launch("weight_layouts_batched_kernel");
}

int dw_wgrad_from_tap(const float* dwt, float* dw, int C, cudaStream_t stream) {
  launch_k(dw_wgrad_from_tap_kernel, dim3((9 * C + 255) / 256), dim3(256), 0, stream, dwt, dw, C);
  return check_launch("dw_wgrad_from_tap_kernel");
}

}  // namespace sfk
