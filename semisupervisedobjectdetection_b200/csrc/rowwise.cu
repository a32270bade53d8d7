// Bandwidth-bound row-wise kernels: LayerNorm fwd/bwd, scaled casts, column sums, token concat/split.
// One warp per row, 128-bit loads, warp-shuffle reductions (no shared memory needed for C <= 1024).
//
// Reference call sites: nn.LayerNorm at MS:133,174,325,333,418 (eps = 1e-5, biased variance, affine) and the
// autograd gradients of those; bias gradients of every nn.Linear/Conv2d on the path.
#include "common.cuh"
#include "sfk_internal.h"
#include "colsum_slots.cuh"

namespace sfk {

static constexpr int LN_CHUNK_LIMIT = 8;  // float4 chunks per lane -> C <= 1024

// sfk_set_option("ln_bwd_ring", 0|1): cp.async row-ring LayerNorm backward (default) or the register-only kernel
int g_ln_bwd_ring = 1;

static bool map_is_identity(const RowMap& m) { return m.period == 0 && m.offset == 0; }

__device__ __forceinline__ long long map_row(long long row, const RowMap& m) {
  if (m.period > 0) row += (row / m.period) * m.extra;
  return row + m.offset;
}

template <typename T> struct Ld4;
template <> struct Ld4<float> {
  static __device__ __forceinline__ float4 ld(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static __device__ __forceinline__ void st(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
};
template <> struct Ld4<__nv_bfloat16> {
  static __device__ __forceinline__ float4 ld(const __nv_bfloat16* p) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    return make_float4(a.x, a.y, b.x, b.y);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, float4 v) {
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// ------------------------------------------------------------------------------------------------
// LayerNorm forward
// ------------------------------------------------------------------------------------------------
// A warp owns rows warp, warp+8, ... of its block's row range: gamma / beta live in registers for the whole range and the next
// row's loads are issued before the current row's two reductions (ncu on the one-row-per-warp version: 11.4 us for a 31.5 MB
// pass, long-scoreboard bound, blocks lived for a single memory round trip).
template <typename TIn, typename TOut, int LN_MAX_CHUNKS>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const TIn* __restrict__ x, RowMap in_map, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, TOut* __restrict__ y,
                                                     float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                                     long long rows, int C, float eps, int rows_per_block) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = C >> 2;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  long long row_end = row0 + rows_per_block;
  if (row_end > rows) row_end = rows;
  long long row = row0 + warp;
  if (row >= row_end) return;
  float4 g[LN_MAX_CHUNKS], bt[LN_MAX_CHUNKS], v[LN_MAX_CHUNKS], vn[LN_MAX_CHUNKS];
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
    const int ci = lane + 32 * k;
    const bool ok = ci < nchunk;
    g[k] = ok ? __ldg(reinterpret_cast<const float4*>(gamma) + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
    bt[k] = ok ? __ldg(reinterpret_cast<const float4*>(beta) + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
    v[k] = ok ? Ld4<TIn>::ld(x + map_row(row, in_map) * C + ci * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const float invC = 1.0f / (float)C;
  for (; row < row_end; row += 8) {
    const long long nrow = row + 8;
    if (nrow < row_end) {
      const TIn* xn = x + map_row(nrow, in_map) * C;
#pragma unroll
      for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
        const int ci = lane + 32 * k;
        vn[k] = ci < nchunk ? Ld4<TIn>::ld(xn + ci * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) s += v[k].x + v[k].y + v[k].z + v[k].w;   // lanes past C hold zeros
    const float mean = warp_sum(s) * invC;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunk) {
        const float a = v[k].x - mean, b = v[k].y - mean, c = v[k].z - mean, d = v[k].w - mean;
        q += a * a + b * b + c * c + d * d;
      }
    }
    const float rstd = rsqrtf(warp_sum(q) * invC + eps);
    TOut* yr = y + row * C;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunk) {
        float4 o;
        o.x = (v[k].x - mean) * rstd * g[k].x + bt[k].x;
        o.y = (v[k].y - mean) * rstd * g[k].y + bt[k].y;
        o.z = (v[k].z - mean) * rstd * g[k].z + bt[k].z;
        o.w = (v[k].w - mean) * rstd * g[k].w + bt[k].w;
        Ld4<TOut>::st(yr + ci * 4, o);
      }
    }
    if (lane == 0) {
      if (mean_out) mean_out[row] = mean;
      if (rstd_out) rstd_out[row] = rstd;
    }
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) v[k] = vn[k];
  }
}

int layernorm_fwd(const void* x, int x_bf16, RowMap in_map, const float* gamma, const float* beta, void* y, int y_bf16,
                  float* mean, float* rstd, long long rows, int C, float eps, cudaStream_t stream) {
  SFK_REQUIRE(C % 4 == 0 && C <= 4 * 32 * LN_CHUNK_LIMIT, "layernorm: C=%d unsupported (need C %% 4 == 0, C <= 1024)", C);
  if (rows == 0) return 0;
  // ~3 resident blocks per SM (8 for narrow rows, where a row is only 256-512 bytes in flight per warp); each warp walks
  // rows_per_block / 8 rows
  const long long slots = (long long)num_sms() * (C <= 128 ? 8 : 3);
  int rows_per_block = (int)((rows + slots - 1) / slots);
  rows_per_block = (rows_per_block + 7) / 8 * 8;
  const unsigned grid = (unsigned)((rows + rows_per_block - 1) / rows_per_block);
#define LAUNCH_N(TI, TO, NCH) launch_k(ln_fwd_kernel<TI, TO, NCH>, dim3(grid), dim3(256), 0, stream, (const TI*)x, in_map, gamma, beta, (TO*)y, mean, rstd, rows, C, eps, rows_per_block)
#define LAUNCH(TI, TO)                                   \
  do {                                                   \
    if (C <= 128) LAUNCH_N(TI, TO, 1);                   \
    else if (C <= 256) LAUNCH_N(TI, TO, 2);              \
    else if (C <= 384) LAUNCH_N(TI, TO, 3);              \
    else if (C <= 512) LAUNCH_N(TI, TO, 4);              \
    else LAUNCH_N(TI, TO, 8);                            \
  } while (0)
  if (x_bf16 && y_bf16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (x_bf16) LAUNCH(__nv_bfloat16, float);
  else if (y_bf16) LAUNCH(float, __nv_bfloat16);
  else LAUNCH(float, float);
#undef LAUNCH
#undef LAUNCH_N
  return check_launch("ln_fwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward:  dx = rstd * (dy*g - mean_c(dy*g) - xhat * mean_c(dy*g*xhat))  (+ dres)
//                      dgamma += sum_rows dy * xhat ; dbeta += sum_rows dy
// ------------------------------------------------------------------------------------------------
// AUX: additionally emit aux = bf16(dx * aux_scale[row / aux_rps]) and aux_colsum[c] += sum_rows of the scaled fp32 values —
// the drop-path-scaled bf16 copy (and bias gradient) the next GEMMs of the layer backward consume, which used to be a
// separate scale_cast pass over dx.
template <typename TDy, typename TX, typename TDx, int LN_MAX_CHUNKS, bool AUX>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const TDy* __restrict__ dy, const TX* __restrict__ x, RowMap x_map,
                                                     const float* __restrict__ mean, const float* __restrict__ rstd,
                                                     const float* __restrict__ gamma, const float* __restrict__ dres,
                                                     RowMap dres_map, TDx* __restrict__ dx, RowMap dx_map,
                                                     float* __restrict__ dgamma, float* __restrict__ dbeta, long long rows,
                                                     int C, int rows_per_block, int ns, __nv_bfloat16* __restrict__ aux,
                                                     const float* __restrict__ aux_scale, int aux_rps,
                                                     float* __restrict__ aux_colsum) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = C >> 2;
  float4 gacc[LN_MAX_CHUNKS], bacc[LN_MAX_CHUNKS], g[LN_MAX_CHUNKS], cacc[AUX ? LN_MAX_CHUNKS : 1];
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
    gacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    bacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (AUX) cacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const int ci = lane + 32 * k;
    g[k] = ci < nchunk ? __ldg(reinterpret_cast<const float4*>(gamma) + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  for (int i = warp; i < rows_per_block; i += 8) {
    const long long row = row0 + i;
    if (row >= rows) break;
    const TDy* dyr = dy + row * C;
    const TX* xr = x + map_row(row, x_map) * C;
    const float* rr = dres ? dres + map_row(row, dres_map) * C : nullptr;
    const float mu = mean[row], rs = rstd[row];
    float4 dyv[LN_MAX_CHUNKS], xh[LN_MAX_CHUNKS], rv[LN_MAX_CHUNKS];
    float s1 = 0.f, s2 = 0.f;
    // all loads of the row (dy, x and the residual gradient) are issued before the reductions: one memory latency per row
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunk) {
        dyv[k] = Ld4<TDy>::ld(dyr + ci * 4);
        xh[k] = Ld4<TX>::ld(xr + ci * 4);
        rv[k] = rr ? *reinterpret_cast<const float4*>(rr + ci * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunk) {
        xh[k] = make_float4((xh[k].x - mu) * rs, (xh[k].y - mu) * rs, (xh[k].z - mu) * rs, (xh[k].w - mu) * rs);
        const float a = dyv[k].x * g[k].x, b = dyv[k].y * g[k].y, c = dyv[k].z * g[k].z, d = dyv[k].w * g[k].w;
        s1 += a + b + c + d;
        s2 += a * xh[k].x + b * xh[k].y + c * xh[k].z + d * xh[k].w;
        gacc[k].x += dyv[k].x * xh[k].x; gacc[k].y += dyv[k].y * xh[k].y;
        gacc[k].z += dyv[k].z * xh[k].z; gacc[k].w += dyv[k].w * xh[k].w;
        bacc[k].x += dyv[k].x; bacc[k].y += dyv[k].y; bacc[k].z += dyv[k].z; bacc[k].w += dyv[k].w;
      }
    }
    s1 = warp_sum(s1) / (float)C;
    s2 = warp_sum(s2) / (float)C;
    TDx* dxr = dx + map_row(row, dx_map) * C;
    float asc = 1.f;
    if (AUX && aux_scale != nullptr) asc = aux_scale[row / aux_rps];
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunk) {
        float4 o;
        o.x = rs * (dyv[k].x * g[k].x - s1 - xh[k].x * s2) + rv[k].x;
        o.y = rs * (dyv[k].y * g[k].y - s1 - xh[k].y * s2) + rv[k].y;
        o.z = rs * (dyv[k].z * g[k].z - s1 - xh[k].z * s2) + rv[k].z;
        o.w = rs * (dyv[k].w * g[k].w - s1 - xh[k].w * s2) + rv[k].w;
        Ld4<TDx>::st(dxr + ci * 4, o);
        if (AUX) {
          o.x *= asc; o.y *= asc; o.z *= asc; o.w *= asc;
          Ld4<__nv_bfloat16>::st(aux + row * C + ci * 4, o);
          cacc[k].x += o.x; cacc[k].y += o.y; cacc[k].z += o.z; cacc[k].w += o.w;
        }
      }
    }
  }
  if (AUX && aux_colsum != nullptr) {
    // same cross-warp reduce + slotted commit as below, on the third per-column partial
    extern __shared__ float red[];
    float* my = red + warp * 2 * C;
#pragma unroll
    for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
      const int ci = lane + 32 * k;
      if (ci < nchunk) *reinterpret_cast<float4*>(my + ci * 4) = cacc[k];
    }
    __syncthreads();
    float* tot = red + 8 * 2 * C;
    for (int c = threadIdx.x; c < C; c += 256) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) t += red[w * 2 * C + c];
      tot[c] = t;
    }
    __syncthreads();
    // slot columns [2C, 3C): disjoint from the dgamma / dbeta commit below (blocks reach the two commits at different times)
    colsum_commit(tot, 2 * C, C, 3 * C, ns, blockIdx.x, gridDim.x, 1, [&](int c, float sum) { aux_colsum[c - 2 * C] += sum; });
    __syncthreads();
  }
  if (dgamma == nullptr) return;
  // cross-warp reduction of the per-column partials, then one atomic per column per block
  extern __shared__ float red[];  // [8][2*C]
  float* my = red + warp * 2 * C;
#pragma unroll
  for (int k = 0; k < LN_MAX_CHUNKS; ++k) {
    const int ci = lane + 32 * k;
    if (ci < nchunk) {
      *reinterpret_cast<float4*>(my + ci * 4) = gacc[k];
      *reinterpret_cast<float4*>(my + C + ci * 4) = bacc[k];
    }
  }
  __syncthreads();
  float* tot = red + 8 * 2 * C;  // [2*C] block totals
  for (int c = threadIdx.x; c < 2 * C; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w * 2 * C + c];
    tot[c] = t;
  }
  __syncthreads();
  colsum_commit(tot, 0, 2 * C, AUX ? 3 * C : 2 * C, ns, blockIdx.x, gridDim.x, 0, [&](int c, float sum) {
    if (c < C) dgamma[c] += sum;
    else dbeta[c - C] += sum;
  });
}

// ------------------------------------------------------------------------------------------------
// LayerNorm backward, fast path (identity row maps, bf16 dy, fp32 x, fp32 residual gradient or none): the rows a warp walks
// are prefetched RING rows ahead into shared memory with cp.async, so the memory latency of a row is hidden behind the
// arithmetic of the previous ones without spending registers on it.  ncu on the register-only kernel (16384 x 320):
// 458 instructions per row, 7.2 cycles of long-scoreboard stall per issued instruction with 4 warps per scheduler, 45 us for
// an 84 MB pass.  Every lane copies and later reads only its own chunks, so cp.async.wait_group is the only synchronisation.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <typename TDx, int NCH, bool AUX, bool RES>
__global__ void __launch_bounds__(256, 2) ln_bwd_ring_kernel(const __nv_bfloat16* __restrict__ dy, const float* __restrict__ x,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ dres,
                                                             TDx* __restrict__ dx, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                             unsigned rows, unsigned C, unsigned rows_per_block, int ns,
                                                             __nv_bfloat16* __restrict__ aux, const float* __restrict__ aux_scale,
                                                             unsigned aux_rps, float* __restrict__ aux_colsum) {
  pdl_sync();
  constexpr int RING = NCH <= 3 ? 3 : 2;   // rows in flight per warp (92 / 82 KB of shared memory per block: two blocks per SM)
  constexpr uint32_t ROW_SLOT = NCH * 1280;   // bytes per WARP per ring slot: x [NCH][32 lanes][16 B], dres likewise, dy [NCH][32][8 B]
  extern __shared__ __align__(16) uint8_t ring_raw[];
  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned nchunk = C >> 2;
  // per-warp ring, lane-interleaved so that the 32 lanes of a copy / load touch consecutive 16-byte (8-byte) words
  const uint32_t my_ring = smem_u32(ring_raw) + warp * (RING * ROW_SLOT);
  const uint32_t off_x = lane * 16, off_r = NCH * 512 + lane * 16, off_d = NCH * 1024 + lane * 8;
  float4 gacc[NCH], bacc[NCH], g[NCH], cacc[AUX ? NCH : 1];
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    gacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    bacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (AUX) cacc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
    const unsigned ci = lane + 32 * k;
    g[k] = ci < nchunk ? __ldg(reinterpret_cast<const float4*>(gamma) + ci) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  const unsigned row0 = blockIdx.x * rows_per_block;
  unsigned row_end = row0 + rows_per_block;
  if (row_end > rows) row_end = rows;
  const float invC = 1.0f / (float)C;
  auto issue = [&](unsigned row, int slot) {
    if (row < row_end) {
      const uint32_t base = my_ring + slot * ROW_SLOT;
      const size_t off = (size_t)row * C;
#pragma unroll
      for (int k = 0; k < NCH; ++k) {
        const unsigned ci = lane + 32 * k;
        if (ci < nchunk) {
          cp_async_16(base + off_x + k * 512, x + off + ci * 4);
          if (RES) cp_async_16(base + off_r + k * 512, dres + off + ci * 4);
          cp_async_8(base + off_d + k * 256, dy + off + ci * 4);
        }
      }
    }
    cp_async_commit();   // one group per row, empty past the end, so the wait count stays uniform
  };
#pragma unroll
  for (int sidx = 0; sidx < RING; ++sidx) issue(row0 + warp + 8 * sidx, sidx);
  int slot = 0;
  for (unsigned row = row0 + warp; row < row_end; row += 8) {
    cp_async_wait<RING - 1>();
    const uint32_t base = my_ring + slot * ROW_SLOT;
    const float mu = mean[row], rs = rstd[row];
    float4 dyv[NCH], xh[NCH], rv[NCH];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const unsigned ci = lane + 32 * k;
      if (ci < nchunk) {
        const uint4 xv = ld_shared_v4(base + off_x + k * 512);
        xh[k] = make_float4((__uint_as_float(xv.x) - mu) * rs, (__uint_as_float(xv.y) - mu) * rs, (__uint_as_float(xv.z) - mu) * rs,
                            (__uint_as_float(xv.w) - mu) * rs);
        if (RES) {
          const uint4 r4 = ld_shared_v4(base + off_r + k * 512);
          rv[k] = make_float4(__uint_as_float(r4.x), __uint_as_float(r4.y), __uint_as_float(r4.z), __uint_as_float(r4.w));
        } else {
          rv[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        uint2 d2;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(d2.x), "=r"(d2.y) : "r"(base + off_d + k * 256) : "memory");
        const float2 da = unpack_bf16x2(d2.x), db = unpack_bf16x2(d2.y);
        dyv[k] = make_float4(da.x, da.y, db.x, db.y);
        const float a = dyv[k].x * g[k].x, b = dyv[k].y * g[k].y, c = dyv[k].z * g[k].z, d = dyv[k].w * g[k].w;
        s1 += (a + b) + (c + d);
        s2 += (a * xh[k].x + b * xh[k].y) + (c * xh[k].z + d * xh[k].w);
        gacc[k].x = fmaf(dyv[k].x, xh[k].x, gacc[k].x); gacc[k].y = fmaf(dyv[k].y, xh[k].y, gacc[k].y);
        gacc[k].z = fmaf(dyv[k].z, xh[k].z, gacc[k].z); gacc[k].w = fmaf(dyv[k].w, xh[k].w, gacc[k].w);
        bacc[k].x += dyv[k].x; bacc[k].y += dyv[k].y; bacc[k].z += dyv[k].z; bacc[k].w += dyv[k].w;
      }
    }
    // this slot's data is in registers: refill it with the row RING steps ahead
    issue(row + 8 * RING, slot);
    slot = slot + 1 == RING ? 0 : slot + 1;
    s1 = warp_sum(s1) * invC;
    s2 = warp_sum(s2) * invC;
    float asc = 1.f;
    if (AUX && aux_scale != nullptr) asc = aux_scale[row / aux_rps];
    TDx* dxr = dx + (size_t)row * C;
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
      const unsigned ci = lane + 32 * k;
      if (ci < nchunk) {
        float4 o;
        o.x = rs * (dyv[k].x * g[k].x - s1 - xh[k].x * s2) + rv[k].x;
        o.y = rs * (dyv[k].y * g[k].y - s1 - xh[k].y * s2) + rv[k].y;
        o.z = rs * (dyv[k].z * g[k].z - s1 - xh[k].z * s2) + rv[k].z;
        o.w = rs * (dyv[k].w * g[k].w - s1 - xh[k].w * s2) + rv[k].w;
        Ld4<TDx>::st(dxr + ci * 4, o);
        if (AUX) {
          o.x *= asc; o.y *= asc; o.z *= asc; o.w *= asc;
          if (aux != nullptr) Ld4<__nv_bfloat16>::st(aux + (size_t)row * C + ci * 4, o);
          cacc[k].x += o.x; cacc[k].y += o.y; cacc[k].z += o.z; cacc[k].w += o.w;
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();   // the ring is dead: its shared memory now carries the cross-warp column reductions
  // ONE reduction + commit for all column sums of the pass: [dgamma | dbeta | aux column sums].  (Two separate commits meant two
  // device-scope fences, two ticket rounds and two last-block folds at the tail of a kernel that is 18 us long for the
  // 4096-row SR LayerNorm.)
  const bool csum = AUX && aux_colsum != nullptr;
  if (dgamma == nullptr && !csum) return;
  const unsigned W3 = csum ? 3 * C : 2 * C;
  float* red = reinterpret_cast<float*>(ring_raw);   // [8][W3] partials + [W3] totals  (9 * W3 floats <= ring size, checked on the host)
  float* my = red + warp * W3;
#pragma unroll
  for (int k = 0; k < NCH; ++k) {
    const unsigned ci = lane + 32 * k;
    if (ci < nchunk) {
      *reinterpret_cast<float4*>(my + ci * 4) = gacc[k];
      *reinterpret_cast<float4*>(my + C + ci * 4) = bacc[k];
      if (AUX) {
        if (csum) *reinterpret_cast<float4*>(my + 2 * C + ci * 4) = cacc[k];
      }
    }
  }
  __syncthreads();
  float* tot = red + 8 * W3;
  for (unsigned c = threadIdx.x; c < W3; c += 256) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w * W3 + c];
    tot[c] = t;
  }
  __syncthreads();
  colsum_commit(tot, 0, (int)W3, AUX ? 3 * C : 2 * C, ns, blockIdx.x, gridDim.x, 0, [&](int c, float sum) {
    if (c < (int)C) { if (dgamma != nullptr) dgamma[c] += sum; }
    else if (c < 2 * (int)C) { if (dbeta != nullptr) dbeta[c - C] += sum; }
    else aux_colsum[c - 2 * C] += sum;
  });
}

template <typename TDx, int NCH>
static int ln_bwd_ring_launch(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const float* dres,
                              void* dx, float* dgamma, float* dbeta, long long rows, int C, void* aux, const float* aux_scale,
                              int aux_rps, float* aux_colsum, cudaStream_t stream) {
  const long long slots = (long long)num_sms() * 2;
  int rows_per_block = (int)((rows + slots - 1) / slots);
  rows_per_block = (rows_per_block + 7) / 8 * 8;
  const long long grid = (rows + rows_per_block - 1) / rows_per_block;
  size_t smem = (size_t)8 * (NCH <= 3 ? 3 : 2) * NCH * 1280;
  const bool want_aux = aux != nullptr || aux_colsum != nullptr;
  const size_t red_bytes = (size_t)9 * (want_aux ? 3 : 2) * C * sizeof(float);
  if (smem < red_bytes) smem = red_bytes;
  const int ns = colsum_num_slots((want_aux ? 3 : 2) * C, grid, colsum_bank(stream));
#define LAUNCH_RING(AUXF, RESF)                                                                                               \
  do {                                                                                                                         \
    cudaFuncSetAttribute(ln_bwd_ring_kernel<TDx, NCH, AUXF, RESF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
    launch_k(ln_bwd_ring_kernel<TDx, NCH, AUXF, RESF>, dim3((unsigned)grid), dim3(256), smem, stream, (const __nv_bfloat16*)dy, \
             (const float*)x, mean, rstd, gamma, dres, (TDx*)dx, dgamma, dbeta, (unsigned)rows, (unsigned)C,                   \
             (unsigned)rows_per_block, ns, (__nv_bfloat16*)aux, aux_scale, (unsigned)(aux_rps > 0 ? aux_rps : 1), aux_colsum); \
  } while (0)
  if (want_aux) { if (dres) LAUNCH_RING(true, true); else LAUNCH_RING(true, false); }
  else { if (dres) LAUNCH_RING(false, true); else LAUNCH_RING(false, false); }
#undef LAUNCH_RING
  return check_launch("ln_bwd_ring_kernel");
}

static int layernorm_bwd_impl(const void* dy, int dy_bf16, const void* x, int x_bf16, RowMap x_map, const float* mean,
                              const float* rstd, const float* gamma, const float* dres, RowMap dres_map, void* dx, int dx_bf16,
                              RowMap dx_map, float* dgamma, float* dbeta, long long rows, int C, void* aux, const float* aux_scale,
                              int aux_rps, float* aux_colsum, cudaStream_t stream) {
  SFK_REQUIRE(C % 4 == 0 && C <= 4 * 32 * LN_CHUNK_LIMIT, "layernorm_bwd: C=%d unsupported", C);
  if (rows == 0) return 0;
  const bool ring_ok = g_ln_bwd_ring && dy_bf16 && !x_bf16 && map_is_identity(x_map) && map_is_identity(dres_map) &&
                       map_is_identity(dx_map) && C <= 512 && rows * (long long)C < (1LL << 31) && rows >= 256;
  SFK_REQUIRE(aux != nullptr || aux_colsum == nullptr || (ring_ok && dgamma != nullptr),
              "layernorm_bwd: column sums of dx without the bf16 copy need the ring kernel (ask layernorm_bwd_colsum)");
  if (ring_ok && (aux == nullptr || (!dx_bf16 && dgamma != nullptr))) {
#define RING_N(NCH)                                                                                                           \
  (dx_bf16 ? ln_bwd_ring_launch<__nv_bfloat16, NCH>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C, aux, aux_scale, \
                                                    aux_rps, aux_colsum, stream)                                                \
           : ln_bwd_ring_launch<float, NCH>(dy, x, mean, rstd, gamma, dres, dx, dgamma, dbeta, rows, C, aux, aux_scale, aux_rps, \
                                            aux_colsum, stream))
    if (C <= 128) return RING_N(1);
    if (C <= 256) return RING_N(2);
    if (C <= 384) return RING_N(3);
    return RING_N(4);
#undef RING_N
  }
  // one even wave: two resident blocks per SM (register-limited), every block the same number of rows (multiple of the
  // 8 warps).  ncu: a fixed 64 rows / block left 40 of the 148 SMs with half the work of the others at 16384 rows.
  const long long slots = (long long)num_sms() * 2;
  int rows_per_block = (int)((rows + slots - 1) / slots);
  rows_per_block = (rows_per_block + 7) / 8 * 8;
  const long long grid = (rows + rows_per_block - 1) / rows_per_block;
  const size_t smem = (size_t)9 * 2 * C * sizeof(float);
  const int ns = colsum_num_slots((aux != nullptr ? 3 : 2) * C, grid, colsum_bank(stream));
#define LAUNCH_NA(TD, TX, TO, NCH, AUXF)                                                                                \
  if (smem > 48 * 1024) cudaFuncSetAttribute(ln_bwd_kernel<TD, TX, TO, NCH, AUXF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  launch_k(ln_bwd_kernel<TD, TX, TO, NCH, AUXF>, dim3((unsigned)grid), dim3(256), smem, stream, (const TD*)dy, (const TX*)x, x_map, mean, rstd, gamma, \
                                                                        dres, dres_map, (TO*)dx, dx_map, dgamma, dbeta, rows, \
                                                                        C, rows_per_block, ns, (__nv_bfloat16*)aux, aux_scale, aux_rps, aux_colsum)
#define LAUNCH_N(TD, TX, TO, NCH) LAUNCH_NA(TD, TX, TO, NCH, false)
#define LAUNCH(TD, TX, TO)                               \
  do {                                                   \
    if (C <= 128) { LAUNCH_N(TD, TX, TO, 1); }           \
    else if (C <= 256) { LAUNCH_N(TD, TX, TO, 2); }      \
    else if (C <= 384) { LAUNCH_N(TD, TX, TO, 3); }      \
    else if (C <= 512) { LAUNCH_N(TD, TX, TO, 4); }      \
    else { LAUNCH_N(TD, TX, TO, 8); }                    \
  } while (0)
  const int key = (dy_bf16 ? 4 : 0) | (x_bf16 ? 2 : 0) | (dx_bf16 ? 1 : 0);
  if (aux != nullptr) {
    SFK_REQUIRE(key == 4, "layernorm_bwd: the fused bf16 copy is built for bf16 dy / fp32 x / fp32 dx only");
    SFK_REQUIRE(dgamma != nullptr, "layernorm_bwd: the fused bf16 copy needs the parameter-gradient path (shared reduction buffers)");
    if (C <= 128) { LAUNCH_NA(__nv_bfloat16, float, float, 1, true); }
    else if (C <= 256) { LAUNCH_NA(__nv_bfloat16, float, float, 2, true); }
    else if (C <= 384) { LAUNCH_NA(__nv_bfloat16, float, float, 3, true); }
    else if (C <= 512) { LAUNCH_NA(__nv_bfloat16, float, float, 4, true); }
    else { LAUNCH_NA(__nv_bfloat16, float, float, 8, true); }
    return check_launch("ln_bwd_kernel");
  }
  switch (key) {
    case 0: LAUNCH(float, float, float); break;
    case 1: LAUNCH(float, float, __nv_bfloat16); break;
    case 2: LAUNCH(float, __nv_bfloat16, float); break;
    case 3: LAUNCH(float, __nv_bfloat16, __nv_bfloat16); break;
    case 4: LAUNCH(__nv_bfloat16, float, float); break;
    case 5: LAUNCH(__nv_bfloat16, float, __nv_bfloat16); break;
    case 6: LAUNCH(__nv_bfloat16, __nv_bfloat16, float); break;
    default: LAUNCH(__nv_bfloat16, __nv_bfloat16, __nv_bfloat16); break;
  }
#undef LAUNCH
#undef LAUNCH_N
#undef LAUNCH_NA
  return check_launch("ln_bwd_kernel");
}

int layernorm_bwd(const void* dy, int dy_bf16, const void* x, int x_bf16, RowMap x_map, const float* mean,
                  const float* rstd, const float* gamma, const float* dres, RowMap dres_map, void* dx, int dx_bf16,
                  RowMap dx_map, float* dgamma, float* dbeta, long long rows, int C, cudaStream_t stream) {
  return layernorm_bwd_impl(dy, dy_bf16, x, x_bf16, x_map, mean, rstd, gamma, dres, dres_map, dx, dx_bf16, dx_map, dgamma, dbeta,
                            rows, C, nullptr, nullptr, 1, nullptr, stream);
}

// same, also writing aux = bf16(dx * aux_scale[row / aux_rps]) [rows, C] (identity row map) and aux_colsum += column sums
// LayerNorm backward (bf16 dy, fp32 x) whose dx column sums are wanted as well (a bias gradient upstream): fused into the ring
// kernel when it applies, else the plain backward followed by the column-sum kernel.
int layernorm_bwd_colsum(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, void* dx,
                         int dx_bf16, float* dgamma, float* dbeta, long long rows, int C, float* dx_colsum, cudaStream_t stream) {
  RowMap id;
  const bool fuse = g_ln_bwd_ring && C <= 512 && rows * (long long)C < (1LL << 31) && rows >= 256 && dgamma != nullptr && dx_colsum != nullptr;
  if (fuse)
    return layernorm_bwd_impl(dy, 1, x, 0, id, mean, rstd, gamma, nullptr, id, dx, dx_bf16, id, dgamma, dbeta, rows, C, nullptr, nullptr, 1,
                              dx_colsum, stream);
  int rc = layernorm_bwd_impl(dy, 1, x, 0, id, mean, rstd, gamma, nullptr, id, dx, dx_bf16, id, dgamma, dbeta, rows, C, nullptr, nullptr, 1,
                              nullptr, stream);
  if (rc || dx_colsum == nullptr) return rc;
  return colsum(dx, dx_bf16, C, dx_colsum, rows, C, stream);
}

int layernorm_bwd_aux(const void* dy, const void* x, const float* mean, const float* rstd, const float* gamma, const float* dres,
                      void* dx, float* dgamma, float* dbeta, long long rows, int C, void* aux, const float* aux_scale, int aux_rps,
                      float* aux_colsum, cudaStream_t stream) {
  RowMap id;
  return layernorm_bwd_impl(dy, 1, x, 0, id, mean, rstd, gamma, dres, id, dx, 0, id, dgamma, dbeta, rows, C, aux, aux_scale, aux_rps,
                            aux_colsum, stream);
}

// ------------------------------------------------------------------------------------------------
// y[out_map(row), :] = x[in_map(row), :] * scale[row / rows_per_scale]  with dtype conversion (drop-path backward, casts),
// optionally fused with the column sums of y (bias gradients):  colsum0[c] += sum_rows y[row, c] for c < split,
// colsum1[c - split] += ... for c >= split.
// Mapping: every thread owns one fixed 4-column chunk and walks rows, so the column sums live in registers; a warp reads
// 32 consecutive chunks (512 B) of one row.
// ------------------------------------------------------------------------------------------------
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256) scale_cast_kernel(const TIn* __restrict__ x, RowMap in_map, TOut* __restrict__ y,
                                                         RowMap out_map, const float* __restrict__ scale, int rows_per_scale,
                                                         long long rows, int C4, int cwb, int rows_per_block,
                                                         float* __restrict__ colsum0, float* __restrict__ colsum1, int split,
                                                         int ns) {
  pdl_sync();
  __shared__ __align__(16) float sh[256 * 4];
  const int lanes_r = 256 / cwb;                 // row lanes per block
  const int rl = threadIdx.x / cwb;              // this thread's row lane
  const int c4 = blockIdx.x * cwb + (threadIdx.x % cwb);
  const bool active = rl < lanes_r && c4 < C4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (active) {
    const long long r0 = (long long)blockIdx.y * rows_per_block;
    long long r1 = r0 + rows_per_block;
    if (r1 > rows) r1 = rows;
    const long long ldc = (long long)C4 * 4;
    // 4 rows per iteration: four independent 16-byte loads in flight per thread
    long long row = r0 + rl;
    for (; row + 3LL * lanes_r < r1; row += 4LL * lanes_r) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = Ld4<TIn>::ld(x + map_row(row + (long long)u * lanes_r, in_map) * ldc + c4 * 4);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long ru = row + (long long)u * lanes_r;
        if (scale) {
          const float sc = scale[ru / rows_per_scale];
          v[u].x *= sc; v[u].y *= sc; v[u].z *= sc; v[u].w *= sc;
        }
        Ld4<TOut>::st(y + map_row(ru, out_map) * ldc + c4 * 4, v[u]);
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    for (; row < r1; row += lanes_r) {
      float4 v = Ld4<TIn>::ld(x + map_row(row, in_map) * ldc + c4 * 4);
      if (scale) {
        const float sc = scale[row / rows_per_scale];
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      }
      Ld4<TOut>::st(y + map_row(row, out_map) * ldc + c4 * 4, v);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (colsum0 == nullptr) return;
  *reinterpret_cast<float4*>(&sh[threadIdx.x * 4]) = acc;
  __syncthreads();
  __shared__ __align__(16) float colsh[256 * 4];
  if (threadIdx.x < cwb) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < lanes_r; ++r) {
      const float4 u = *reinterpret_cast<const float4*>(&sh[(r * cwb + threadIdx.x) * 4]);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    *reinterpret_cast<float4*>(&colsh[threadIdx.x * 4]) = t;
  }
  __syncthreads();
  const int v0 = blockIdx.x * cwb * 4;
  int nv = cwb * 4;
  if (v0 + nv > C4 * 4) nv = C4 * 4 - v0;
  colsum_commit(colsh, v0, nv, C4 * 4, ns, blockIdx.y, gridDim.y, blockIdx.x, [&](int c, float sum) {
    float* dst = c < split ? colsum0 + c : colsum1 + (c - split);
    *dst += sum;
  });
}

// Fast path of the above for identity row maps (every call of the prompt-free step): 32-bit indexing, the per-sample scale
// looked up without a division in the common case (a block's rows span at most two samples), 8 independent 16-byte loads in
// flight per thread, 64 registers so that four blocks share an SM.  ncu on the generic kernel (stage-3 drop-path cast,
// 16384 x 320): 86 instructions per 16-byte chunk (64-bit row maps and divisions per row), 24 % active warps, 22 us for a
// 31.5 MB pass.
template <typename TIn, typename TOut>
__global__ void __launch_bounds__(256, 4) scale_cast_fast_kernel(const TIn* __restrict__ x, TOut* __restrict__ y,
                                                                 const float* __restrict__ scale, unsigned rows_per_scale,
                                                                 unsigned rows, unsigned C4, unsigned lanes_r, unsigned rows_per_block,
                                                                 float* __restrict__ colsum0, float* __restrict__ colsum1, int split,
                                                                 int ns) {
  pdl_sync();
  __shared__ __align__(16) float sh[256 * 4];
  const unsigned rl = threadIdx.x / C4;              // this thread's row lane
  const unsigned c4 = threadIdx.x - rl * C4;         // its fixed 4-column chunk (gridDim.x == 1: C4 <= 256)
  const unsigned r0 = blockIdx.y * rows_per_block;
  unsigned r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (rl < lanes_r) {
    const unsigned ldc = C4 * 4;
    const unsigned s_lo = scale ? r0 / rows_per_scale : 0u;
    const unsigned r_switch = scale ? (s_lo + 1u) * rows_per_scale : 0xffffffffu;   // first row of the next sample
    const float sc_lo = scale ? scale[s_lo] : 1.f;
    const float sc_hi = (scale && r_switch < r1) ? scale[s_lo + 1u] : sc_lo;
    const bool two_only = !scale || (r1 - r0) <= rows_per_scale;                    // block touches at most two samples
    const TIn* xp = x + (size_t)c4 * 4;
    TOut* yp = y + (size_t)c4 * 4;
    unsigned row = r0 + rl;
    constexpr int U = 8;
    for (; row + (U - 1) * lanes_r < r1; row += U * lanes_r) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) v[u] = Ld4<TIn>::ld(xp + (size_t)(row + u * lanes_r) * ldc);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned ru = row + u * lanes_r;
        if (scale) {
          const float sc = two_only ? (ru < r_switch ? sc_lo : sc_hi) : scale[ru / rows_per_scale];
          v[u].x *= sc; v[u].y *= sc; v[u].z *= sc; v[u].w *= sc;
        }
        Ld4<TOut>::st(yp + (size_t)ru * ldc, v[u]);
        acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
      }
    }
    for (; row < r1; row += lanes_r) {
      float4 v = Ld4<TIn>::ld(xp + (size_t)row * ldc);
      if (scale) {
        const float sc = two_only ? (row < r_switch ? sc_lo : sc_hi) : scale[row / rows_per_scale];
        v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc;
      }
      Ld4<TOut>::st(yp + (size_t)row * ldc, v);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  if (colsum0 == nullptr) return;
  *reinterpret_cast<float4*>(&sh[threadIdx.x * 4]) = acc;
  __syncthreads();
  __shared__ __align__(16) float colsh[256 * 4];
  if (threadIdx.x < C4) {
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (unsigned r = 0; r < lanes_r; ++r) {
      const float4 u = *reinterpret_cast<const float4*>(&sh[(r * C4 + threadIdx.x) * 4]);
      t.x += u.x; t.y += u.y; t.z += u.z; t.w += u.w;
    }
    *reinterpret_cast<float4*>(&colsh[threadIdx.x * 4]) = t;
  }
  __syncthreads();
  colsum_commit(colsh, 0, (int)C4 * 4, (int)C4 * 4, ns, blockIdx.y, gridDim.y, 0, [&](int c, float sum) {
    float* dst = c < split ? colsum0 + c : colsum1 + (c - split);
    *dst += sum;
  });
}


int scale_cast_sum(const void* x, int x_bf16, RowMap in_map, void* y, int y_bf16, RowMap out_map, const float* scale,
                   int rows_per_scale, long long rows, int C, float* colsum0, float* colsum1, int split, cudaStream_t stream) {
  SFK_REQUIRE(C % 4 == 0 && split % 4 == 0, "scale_cast: C=%d and split must be multiples of 4", C);
  if (rows == 0) return 0;
  if (map_is_identity(in_map) && map_is_identity(out_map) && C / 4 <= 256 && rows * (long long)C < (1LL << 31) && rows >= 64) {
    const unsigned C4f = (unsigned)(C / 4);
    const unsigned lanes = 256u / C4f;
    const unsigned threads = lanes * C4f;
    // ~4 blocks per SM, every block a multiple of the row lanes
    long long gy = (long long)num_sms() * 4;
    long long rpb = (rows + gy - 1) / gy;
    rpb = (rpb + lanes - 1) / lanes * lanes;
    if (rpb < (long long)lanes) rpb = lanes;
    gy = (rows + rpb - 1) / rpb;
    if (rows_per_scale <= 0) rows_per_scale = 1;
    if (colsum0 != nullptr && colsum1 == nullptr) { colsum1 = colsum0; split = C; }
    SFK_REQUIRE(colsum0 == nullptr || C <= COLSUM_SLOT_FLOATS, "scale_cast: C=%d too wide for fused sums", C);
    const int nsf = colsum_num_slots(C, gy, colsum_bank(stream));
    dim3 gridf(1, (unsigned)gy);
#define LAUNCHF(TI, TO)                                                                                                        \
  launch_k(scale_cast_fast_kernel<TI, TO>, dim3(gridf), dim3(threads), 0, stream, (const TI*)x, (TO*)y, scale, (unsigned)rows_per_scale, \
           (unsigned)rows, C4f, lanes, (unsigned)rpb, colsum0, colsum1, split, nsf)
    if (x_bf16 && y_bf16) LAUNCHF(__nv_bfloat16, __nv_bfloat16);
    else if (x_bf16) LAUNCHF(__nv_bfloat16, float);
    else if (y_bf16) LAUNCHF(float, __nv_bfloat16);
    else LAUNCHF(float, float);
#undef LAUNCHF
    return check_launch("scale_cast_fast_kernel");
  }
  const int C4 = C / 4;
  const int cwb = C4 < 256 ? C4 : 256;
  const int gx = (C4 + cwb - 1) / cwb;
  const int lanes_r = 256 / cwb;
  // with fused column sums every block pays a fence + slot commit: use few, long-lived blocks (2 per SM)
  long long gy = ((long long)num_sms() * (colsum0 != nullptr ? 2 : 8) + gx - 1) / gx;
  long long rpb = (rows + gy - 1) / gy;
  if (rpb < 4LL * lanes_r) rpb = 4LL * lanes_r;
  gy = (rows + rpb - 1) / rpb;
  if (rows_per_scale <= 0) rows_per_scale = 1;
  if (colsum0 != nullptr && colsum1 == nullptr) { colsum1 = colsum0; split = C; }
  SFK_REQUIRE(colsum0 == nullptr || (gx <= COLSUM_MAX_GX && C <= COLSUM_SLOT_FLOATS), "scale_cast: C=%d too wide for fused sums", C);
  const int ns = colsum_num_slots(C, gy, colsum_bank(stream));
  dim3 grid(gx, (unsigned)gy);
#define LAUNCH(TI, TO)                                                                                                       \
  launch_k(scale_cast_kernel<TI, TO>, dim3(grid), dim3(256), 0, stream, (const TI*)x, in_map, (TO*)y, out_map, scale, rows_per_scale, rows, C4, \
                                                      cwb, (int)rpb, colsum0, colsum1, split, ns)
  if (x_bf16 && y_bf16) LAUNCH(__nv_bfloat16, __nv_bfloat16);
  else if (x_bf16) LAUNCH(__nv_bfloat16, float);
  else if (y_bf16) LAUNCH(float, __nv_bfloat16);
  else LAUNCH(float, float);
#undef LAUNCH
  return check_launch("scale_cast_kernel");
}

int scale_cast(const void* x, int x_bf16, RowMap in_map, void* y, int y_bf16, RowMap out_map, const float* scale,
               int rows_per_scale, long long rows, int C, cudaStream_t stream) {
  return scale_cast_sum(x, x_bf16, in_map, y, y_bf16, out_map, scale, rows_per_scale, rows, C, nullptr, nullptr, 0, stream);
}

// ------------------------------------------------------------------------------------------------
// out[c] += sum_rows x[row, c]   (bias gradients).  Block = 256 columns x ROWS rows, one atomic per column.
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) colsum_kernel(const T* __restrict__ x, long long ld, float* __restrict__ out,
                                                     long long rows, int C, int rows_per_block) {
  pdl_sync();
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float a0 = 0.f, a1 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    if constexpr (sizeof(T) == 2) {
      const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * ld + c));
      a0 += v.x; a1 += v.y;
    } else {
      const float2 v = *reinterpret_cast<const float2*>(x + r * ld + c);
      a0 += v.x; a1 += v.y;
    }
  }
  atomicAdd(out + c, a0);
  atomicAdd(out + c + 1, a1);
}

int colsum(const void* x, int x_bf16, long long ld, float* out, long long rows, int C, cudaStream_t stream) {
  SFK_REQUIRE(C % 2 == 0 && ld % 2 == 0, "colsum: C and ld must be even");
  if (rows == 0) return 0;
  const int gx = (C / 2 + 127) / 128;
  long long gy = (long long)num_sms() * 4 / gx;
  if (gy < 1) gy = 1;
  int rpb = (int)((rows + gy - 1) / gy);
  if (rpb < 32) rpb = 32;
  gy = (rows + rpb - 1) / rpb;
  dim3 grid(gx, (unsigned)gy);
  if (x_bf16) launch_k(colsum_kernel<__nv_bfloat16>, dim3(grid), dim3(128), 0, stream, (const __nv_bfloat16*)x, ld, out, rows, C, rpb);
  else launch_k(colsum_kernel<float>, dim3(grid), dim3(128), 0, stream, (const float*)x, ld, out, rows, C, rpb);
  return check_launch("colsum_kernel");
}

// ------------------------------------------------------------------------------------------------
// Token concat: out[b, 0:P, :] = front (shared [P,C] when front_batch_stride == 0, else per-sample),
//               out[b, P:P+N, :] = x[b, :, :]                      (MS:446-460)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) concat_front_kernel(const float* __restrict__ front, long long front_batch_stride,
                                                           const float* __restrict__ x, float* __restrict__ out, int B, int P,
                                                           int N, int C4) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long total = (long long)B * (P + N) * C4;
  if (idx >= total) return;
  const int c = (int)(idx % C4);
  const long long rowg = idx / C4;
  const int t = (int)(rowg % (P + N));
  const long long b = rowg / (P + N);
  float4 v;
  if (t < P) v = *reinterpret_cast<const float4*>(front + b * front_batch_stride + ((long long)t * C4 + c) * 4);
  else v = *reinterpret_cast<const float4*>(x + ((b * N + (t - P)) * C4 + c) * 4);
  *reinterpret_cast<float4*>(out + idx * 4) = v;
}

int concat_front(const float* front, long long front_batch_stride, const float* x, float* out, int B, int P, int N, int C,
                 cudaStream_t stream) {
  SFK_REQUIRE(C % 4 == 0, "concat_front: C must be a multiple of 4");
  const long long total = (long long)B * (P + N) * (C / 4);
  if (total == 0) return 0;
  launch_k(concat_front_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, front, front_batch_stride, x, out, B, P, N, C / 4);
  return check_launch("concat_front_kernel");
}

// dfront[P,C] (+)= sum_b g[b, 0:P, :] (shared prompt rows: MS:451 repeat -> sum over batch)
__global__ void front_grad_sum_kernel(const float* __restrict__ g, float* __restrict__ dfront, int B, int P, int T, int C) {
  pdl_sync();
  const int idx = blockIdx.x * 256 + threadIdx.x;
  if (idx >= P * C) return;
  const int t = idx / C, c = idx % C;
  float s = 0.f;
  for (int b = 0; b < B; ++b) s += g[((long long)b * T + t) * C + c];
  dfront[idx] += s;
}

int front_grad_sum(const float* g, float* dfront, int B, int P, int T, int C, cudaStream_t stream) {
  if (P * C == 0) return 0;
  launch_k(front_grad_sum_kernel, dim3((P * C + 255) / 256), dim3(256), 0, stream, g, dfront, B, P, T, C);
  return check_launch("front_grad_sum_kernel");
}

}  // namespace sfk
