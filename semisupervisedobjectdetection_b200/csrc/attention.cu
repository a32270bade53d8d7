// Fused attention for SegFormer's efficient self-attention (MS:207-226):  O = softmax(Q K^T / sqrt(d)) V
// with the (optionally prompt/cls-extended) key/value sequence, never materialising the score tensor.
//
// Both directions run on tcgen05 tensor cores with accumulators in TMEM and operands staged by TMA:
//   forward : CTA = 128 query rows of one (sample, head). Per 128-row K/V chunk:  S = Q K^T (TMEM) -> 128 softmax
//             threads, one per row (online softmax, exp2) -> P (bf16, 128B-swizzled smem) -> O_c = P V (TMEM)
//             -> rescale-accumulate in registers.  Two CTAs are resident per SM so one CTA's softmax overlaps the
//             other's MMAs.
//   backward: CTA = one 128-row K/V block of one (sample, head) and a range of query tiles. Per query tile:
//             S^T = K Q^T and dP^T = V dO^T (TMEM) -> threads (one per K/V row) form P^T and dS^T (bf16 smem) ->
//             dV += P^T dO, dK += dS^T Q (TMEM accumulators across tiles), dQ_tile = dS K (TMEM) -> red.add into dQ.
//             No score/probability tensor ever touches HBM; P is recomputed from the saved log-sum-exp.
// Head split is the plain reshape of MS:176-179: head j owns channels [j*d, (j+1)*d).
#include "common.cuh"
#include "sfk_internal.h"

namespace sfk {

static constexpr int ATT_BM = 128;   // query rows per tile
static constexpr int ATT_BN = 128;   // key/value rows per chunk

template <int D> struct AttCfg {
  static constexpr int RB = D * 2;                 // bytes per operand row
  static constexpr uint32_t LT = (RB == 128) ? 2u : 4u;  // UMMA layout type: SWIZZLE_128B / SWIZZLE_64B
  static constexpr int TILE = 128 * RB;            // bytes of a [128 x D] tile
  static constexpr int KSTEPS = D / 16;
};

struct AttFwdParams {
  int T, Tk, heads;
  float scale_log2;  // (1/sqrt(d)) * log2(e)
  __nv_bfloat16* o;
  long long ldo;     // elements between consecutive tokens of o
  long long o_batch_stride;
  float* lse;        // [B, heads, T] in the log2 domain: m*scale_log2 + log2(sum)
};

__device__ __forceinline__ void red_add_f32_g(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32_g(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// 2^x on the MUFU pipe in one instruction (ex2.approx.ftz: 2 ulp, flushes denormals - the softmax numerators it produces are
// rounded to bf16 anyway).  exp2f() compiled to ~8 instructions per element here (range scaling for denormal results).
__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write 8 bf16 (one 16-byte chunk) of row r / column chunk j (8 columns each) into a [128 x 64] SWIZZLE_128B tile
__device__ __forceinline__ void st_swizzled_chunk(uint8_t* tile, int r, int j, uint4 v) {
  *reinterpret_cast<uint4*>(tile + r * 128 + ((j ^ (r & 7)) << 4)) = v;
}

template <int D>
__global__ void __launch_bounds__(160, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                const __grid_constant__ CUtensorMap tmap_v, const AttFwdParams p) {
  pdl_launch_dependents();
  using Cfg = AttCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t sQ = base, sK = sQ + Cfg::TILE, sV = sK + Cfg::TILE, sP = sV + Cfg::TILE;
  uint8_t* sP_gen = gen + 3 * Cfg::TILE;
  const uint32_t bars = sP + 32768;
  const uint32_t b_qfull = bars, b_kfull = bars + 8, b_vfull = bars + 16, b_sfull = bars + 24, b_pready = bars + 32,
                 b_ofull = bars + 40, b_kfree = bars + 48, b_vfree = bars + 56, tmem_slot = bars + 64;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + 3 * Cfg::TILE + 32768 + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, head = blockIdx.y, batch = blockIdx.z;
  const int nchunks = (p.Tk + ATT_BN - 1) / ATT_BN;

  if (threadIdx.x == 128) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(b_qfull, 1); mbar_init(b_kfull, 1); mbar_init(b_vfull, 1); mbar_init(b_sfull, 1);
    mbar_init(b_pready, 128); mbar_init(b_ofull, 1); mbar_init(b_kfree, 1); mbar_init(b_vfree, 1);
    mbar_fence_init();
  }
  if (warp == 4) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  pdl_wait();  // setup above overlaps the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const uint32_t tS = tmem, tO = tmem + 128;

  if (warp == 4) {
    if (lane == 0) {
      // ---------------- control thread: TMA + MMA issue ----------------
      const uint32_t idesc_s = umma_idesc_bf16(128, ATT_BN, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, D, 0, 1);
      mbar_arrive_expect_tx(b_qfull, Cfg::TILE);
      tma_load_3d(sQ, &tmap_q, b_qfull, head * D, q0, batch);
      mbar_arrive_expect_tx(b_kfull, Cfg::TILE);
      tma_load_3d(sK, &tmap_k, b_kfull, head * D, 0, batch);
      mbar_arrive_expect_tx(b_vfull, Cfg::TILE);
      tma_load_3d(sV, &tmap_v, b_vfull, head * D, 0, batch);
      mbar_wait(b_qfull, 0);
      for (int c = 0; c < nchunks; ++c) {
        const uint32_t ph = c & 1;
        mbar_wait(b_kfull, ph);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
          umma_bf16(tS, umma_smem_desc_lt(sQ + kk * 32, 16, 8 * Cfg::RB, Cfg::LT),
                    umma_smem_desc_lt(sK + kk * 32, 16, 8 * Cfg::RB, Cfg::LT), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(b_sfull);
        umma_commit(b_kfree);
        if (c + 1 < nchunks) {
          mbar_wait(b_kfree, ph);
          mbar_arrive_expect_tx(b_kfull, Cfg::TILE);
          tma_load_3d(sK, &tmap_k, b_kfull, head * D, (c + 1) * ATT_BN, batch);
        }
        mbar_wait(b_pready, ph);
        mbar_wait(b_vfull, ph);
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks)
          umma_bf16(tO, umma_smem_desc_lt(sP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2u),
                    umma_smem_desc_lt(sV + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_o,
                    ks > 0 ? 1u : 0u);
        umma_commit(b_ofull);
        umma_commit(b_vfree);
        if (c + 1 < nchunks) {
          mbar_wait(b_vfree, ph);
          mbar_arrive_expect_tx(b_vfull, Cfg::TILE);
          tma_load_3d(sV, &tmap_v, b_vfull, head * D, (c + 1) * ATT_BN, batch);
        }
      }
    }
  } else {
    // ---------------- softmax threads: one query row each ----------------
    const int r = threadIdx.x;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    float m_run = -INFINITY, l_run = 0.f;
    float acc[D];
#pragma unroll
    for (int j = 0; j < D; ++j) acc[j] = 0.f;
    for (int c = 0; c < nchunks; ++c) {
      const uint32_t ph = c & 1;
      const int valid = min(ATT_BN, p.Tk - c * ATT_BN);
      mbar_wait(b_sfull, ph);
      tc_fence_after();
      const bool full = valid == ATT_BN;   // only the last chunk of a ragged K/V length needs column masking
      float cmax = -INFINITY;
#pragma unroll
      for (int cc = 0; cc < ATT_BN; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_base + cc, v);
        tmem_ld_wait();
        if (full) {
          float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};   // four independent chains instead of one of 32
#pragma unroll
          for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
          cmax = fmaxf(cmax, fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (cc + j < valid) cmax = fmaxf(cmax, __uint_as_float(v[j]));
        }
      }
      const float m_new = fmaxf(m_run, cmax);
      const float alpha = fast_exp2((m_run - m_new) * p.scale_log2);
      const float moff = m_new * p.scale_log2;
      float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int cc = 0; cc < ATT_BN; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tS + lane_base + cc, v);
        tmem_ld_wait();
        float e[32];
        if (full) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -moff));
            s4[j & 3] += e[j];
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            e[j] = (cc + j < valid) ? fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -moff)) : 0.f;
            s4[j & 3] += e[j];
          }
        }
        uint8_t* tile = sP_gen + (cc >> 6) * 16384;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 o;
          o.x = pack_bf16x2(e[j8 * 8 + 0], e[j8 * 8 + 1]);
          o.y = pack_bf16x2(e[j8 * 8 + 2], e[j8 * 8 + 3]);
          o.z = pack_bf16x2(e[j8 * 8 + 4], e[j8 * 8 + 5]);
          o.w = pack_bf16x2(e[j8 * 8 + 6], e[j8 * 8 + 7]);
          st_swizzled_chunk(tile, r, ((cc & 63) >> 3) + j8, o);
        }
      }
      const float sum = (s4[0] + s4[1]) + (s4[2] + s4[3]);
      l_run = l_run * alpha + sum;
      m_run = m_new;
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(b_pready);
      mbar_wait(b_ofull, ph);
      tc_fence_after();
#pragma unroll
      for (int cc = 0; cc < D; cc += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tO + lane_base + cc, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cc + j] = fmaf(acc[cc + j], alpha, __uint_as_float(v[j]));
      }
      tc_fence_before();
    }
    const int t = q0 + r;
    if (t < p.T) {
      const float inv = 1.f / l_run;
      __nv_bfloat16* op = p.o + batch * p.o_batch_stride + (long long)t * p.ldo + head * D;
#pragma unroll
      for (int j = 0; j < D; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(acc[j] * inv, acc[j + 1] * inv);
        o.y = pack_bf16x2(acc[j + 2] * inv, acc[j + 3] * inv);
        o.z = pack_bf16x2(acc[j + 4] * inv, acc[j + 5] * inv);
        o.w = pack_bf16x2(acc[j + 6] * inv, acc[j + 7] * inv);
        *reinterpret_cast<uint4*>(op + j) = o;
      }
      if (p.lse) p.lse[((long long)batch * p.heads + head) * p.T + t] = m_run * p.scale_log2 + log2f(l_run);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// ------------------------------------------------------------------------------------------------
// Forward fast path for Tk <= 256 (every stage of the 512x512 path without prompt tokens: the spatial reduction always
// leaves 16 x 16 = 256 key/value rows).  The whole score row fits tensor memory, so there is no online rescaling:
//   S[128 x 256] = Q K^T in ONE accumulator (256 TMEM columns), 256 softmax threads (two per query row, 128 columns each;
//   row maximum and row sum exchanged through shared memory), P[128 x 256] (bf16, four swizzled 64-column tiles) written
//   over the Q and K tiles (dead once S is complete), O = P V with K = 256 into the first D columns of the same TMEM
//   allocation, 256 threads convert and store O.
// 97 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM, one's softmax under the other's loads and MMAs.
// Measured motivation (profiles/r2): the chunked kernel above ran one thread per row through two dependent 128-column
// chunks (S MMA -> softmax -> P V MMA -> rescale, twice) and took 25-30 us for the 640 tiles of a stage-3 layer, ~15 % of
// either roofline.
// ------------------------------------------------------------------------------------------------
__device__ int g_attn_dbg_dev = 0;   // timing experiments only (sfk_set_option("attn_dbg"))
__device__ long long g_attn_trace[16];   // bit 1 of attn_dbg: CTA (0,0,0) stores clock64 at its phase boundaries
#define ATT_TRACE(slot)                                                                                   \
  do {                                                                                                    \
    if ((g_attn_dbg_dev & 2) && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) g_attn_trace[slot] = clock64(); \
  } while (0)
template <int D>
__global__ void __launch_bounds__(288, 2)
attn_fwd256_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_k,
                   const __grid_constant__ CUtensorMap tmap_v, const AttFwdParams p) {
  pdl_launch_dependents();
  if (threadIdx.x == 0) ATT_TRACE(0);
  using Cfg = AttCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // [P: 64 KB, overlaying Q (TILE) and K (2 TILE)] [V: 2 TILE] [barriers, exchange buffers]
  const uint32_t sP = base, sQ = base, sK = base + Cfg::TILE, sV = base + 65536;
  uint8_t* sP_gen = gen;
  const uint32_t bars = sV + 2 * Cfg::TILE;
  const uint32_t b_qk = bars, b_v = bars + 8, b_sfull = bars + 16, b_pready = bars + 24, b_ofull = bars + 32, tmem_slot = bars + 40;
  uint8_t* bars_gen = gen + 65536 + 2 * Cfg::TILE;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(bars_gen + 40);
  float* xmax = reinterpret_cast<float*>(bars_gen + 64);          // [2][128] row maxima of the two column halves
  float* xsum = xmax + 256;                                       // [2][128] row sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * ATT_BM, head = blockIdx.y, batch = blockIdx.z;

  if (threadIdx.x == 256) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(b_qk, 1); mbar_init(b_v, 1); mbar_init(b_sfull, 1); mbar_init(b_pready, 256); mbar_init(b_ofull, 1);
    mbar_fence_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 256); tmem_relinquish(); }
  pdl_wait();  // setup above overlaps the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  if (threadIdx.x == 0) ATT_TRACE(1);

  if (warp == 8) {
    if (lane == 0) {
      // ---------------- control thread: TMA + MMA issue ----------------
      mbar_arrive_expect_tx(b_qk, 3 * Cfg::TILE);
      tma_load_3d(sQ, &tmap_q, b_qk, head * D, q0, batch);
      tma_load_3d(sK, &tmap_k, b_qk, head * D, 0, batch);        // box = 256 rows; rows >= Tk arrive as zeros
      mbar_arrive_expect_tx(b_v, 2 * Cfg::TILE);
      tma_load_3d(sV, &tmap_v, b_v, head * D, 0, batch);
      const uint32_t idesc_s = umma_idesc_bf16(128, 256, 0, 0);
      const uint32_t idesc_o = umma_idesc_bf16(128, D, 0, 1);
      mbar_wait(b_qk, 0);
      ATT_TRACE(2);
      tc_fence_after();
#pragma unroll
      for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
        umma_bf16(tmem, umma_smem_desc_lt(sQ + kk * 32, 16, 8 * Cfg::RB, Cfg::LT),
                  umma_smem_desc_lt(sK + kk * 32, 16, 8 * Cfg::RB, Cfg::LT), idesc_s, kk > 0 ? 1u : 0u);
      umma_commit(b_sfull);
      mbar_wait(b_pready, 0);     // all 256 threads have read S and written P (over Q / K)
      ATT_TRACE(5);
      mbar_wait(b_v, 0);
      tc_fence_after();
#pragma unroll
      for (int ks = 0; ks < 16; ++ks)
        umma_bf16(tmem, umma_smem_desc_lt(sP + (ks >> 2) * 16384 + (ks & 3) * 32, 16, 1024, 2u),
                  umma_smem_desc_lt(sV + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_o, ks > 0 ? 1u : 0u);
      umma_commit(b_ofull);
    }
  } else {
    // ---------------- softmax threads: query row r, column half h ----------------
    const int qd = warp & 3, h = warp >> 2;
    const int r = qd * 32 + lane;
    const uint32_t lane_base = (uint32_t)(qd * 32) << 16;
    const int valid = min(128, max(0, p.Tk - h * 128));   // valid key columns of this half
    const bool full = valid == 128;
    mbar_wait(b_sfull, 0);
    tc_fence_after();
    if (threadIdx.x == 0) ATT_TRACE(3);
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (g_attn_dbg_dev & 1) m4[0] = 8.f;   // experiment (wrong results): skip the row-maximum pass
    else
#pragma unroll
    for (int cc = 0; cc < 128; cc += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + lane_base + h * 128 + cc, v);
      tmem_ld_wait();
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; ++j) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (cc + j < valid) m4[j & 3] = fmaxf(m4[j & 3], __uint_as_float(v[j]));
      }
    }
    const float m_half = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    xmax[h * 128 + r] = m_half;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    if (threadIdx.x == 0) ATT_TRACE(4);
    const float m_row = fmaxf(m_half, xmax[(h ^ 1) * 128 + r]);
    const float moff = m_row * p.scale_log2;
    float s4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int cc = 0; cc < 128; cc += 32) {
      uint32_t v[32];
      tmem_ld_32x32(tmem + lane_base + h * 128 + cc, v);
      tmem_ld_wait();
      float e[32];
      if (full) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          e[j] = fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -moff));
          s4[j & 3] += e[j];
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          e[j] = (cc + j < valid) ? fast_exp2(fmaf(__uint_as_float(v[j]), p.scale_log2, -moff)) : 0.f;
          s4[j & 3] += e[j];
        }
      }
      uint8_t* tile = sP_gen + (h * 2 + (cc >> 6)) * 16384;
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        uint4 o;
        o.x = pack_bf16x2(e[j8 * 8 + 0], e[j8 * 8 + 1]);
        o.y = pack_bf16x2(e[j8 * 8 + 2], e[j8 * 8 + 3]);
        o.z = pack_bf16x2(e[j8 * 8 + 4], e[j8 * 8 + 5]);
        o.w = pack_bf16x2(e[j8 * 8 + 6], e[j8 * 8 + 7]);
        st_swizzled_chunk(tile, r, ((cc & 63) >> 3) + j8, o);
      }
    }
    xsum[h * 128 + r] = (s4[0] + s4[1]) + (s4[2] + s4[3]);
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(b_pready);
    mbar_wait(b_ofull, 0);
    tc_fence_after();
    if (threadIdx.x == 0) ATT_TRACE(6);
    // every thread arrived on b_pready before the P V MMAs were issued, so both halves' sums are visible here
    const float l_row = xsum[r] + xsum[128 + r];
    const float inv = 1.f / l_row;
    constexpr int OC = D / 2;   // output columns per thread
    uint32_t v[OC];
    if (OC == 32) tmem_ld_32x32(tmem + lane_base + h * OC, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
    else tmem_ld_32x16(tmem + lane_base + h * OC, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
    tmem_ld_wait();
    const int t = q0 + r;
    if (t < p.T) {
      __nv_bfloat16* op = p.o + batch * p.o_batch_stride + (long long)t * p.ldo + head * D + h * OC;
#pragma unroll
      for (int j = 0; j < OC; j += 8) {
        uint4 o;
        o.x = pack_bf16x2(__uint_as_float(v[j]) * inv, __uint_as_float(v[j + 1]) * inv);
        o.y = pack_bf16x2(__uint_as_float(v[j + 2]) * inv, __uint_as_float(v[j + 3]) * inv);
        o.z = pack_bf16x2(__uint_as_float(v[j + 4]) * inv, __uint_as_float(v[j + 5]) * inv);
        o.w = pack_bf16x2(__uint_as_float(v[j + 6]) * inv, __uint_as_float(v[j + 7]) * inv);
        *reinterpret_cast<uint4*>(op + j) = o;
      }
      if (p.lse && h == 0) p.lse[((long long)batch * p.heads + head) * p.T + t] = moff + log2f(l_row);
    }
    tc_fence_before();
  }
  __syncthreads();
  if (threadIdx.x == 0) ATT_TRACE(7);
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 256); }
}

// sfk_set_option("attn_fwd256", 0|1): single-pass forward kernel for Tk <= 256 (default 1)
int g_attn_fwd256 = 1;
int set_attn_dbg(int v) { return cudaMemcpyToSymbol(g_attn_dbg_dev, &v, sizeof(int)) == cudaSuccess ? 0 : -1; }
int get_attn_trace(long long* out16) { return cudaMemcpyFromSymbol(out16, g_attn_trace, 16 * sizeof(long long)) == cudaSuccess ? 0 : -1; }

template <int D>
static int attn_fwd256_launch(const AttentionDesc& a, cudaStream_t stream) {
  using Cfg = AttCfg<D>;
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_3d_bf16(&tq, a.q, (uint64_t)a.heads * D, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.ldq * 2,
                             (uint64_t)a.q_batch_stride * 2, D, ATT_BM, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tk, a.k, (uint64_t)a.heads * D, (uint64_t)a.Tk, (uint64_t)a.B, (uint64_t)a.ldkv * 2,
                         (uint64_t)a.kv_batch_stride * 2, D, 256, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tv, a.v, (uint64_t)a.heads * D, (uint64_t)a.Tk, (uint64_t)a.B, (uint64_t)a.ldkv * 2,
                         (uint64_t)a.kv_batch_stride * 2, D, 256, 1, Cfg::RB);
  if (rc) return rc;
  AttFwdParams p;
  p.T = a.T; p.Tk = a.Tk; p.heads = a.heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.o = (__nv_bfloat16*)a.o; p.ldo = a.ldo; p.o_batch_stride = a.o_batch_stride; p.lse = a.lse;
  const size_t smem = 65536 + 2 * Cfg::TILE + 64 + 2048 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd256_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SFK_REQUIRE(e == cudaSuccess, "attention_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((a.T + ATT_BM - 1) / ATT_BM, a.heads, a.B);
  launch_k(attn_fwd256_kernel<D>, dim3(grid), dim3(288), smem, stream, tq, tk, tv, p);
  return check_launch("attn_fwd256_kernel");
}

template <int D>
static int attn_fwd_launch(const AttentionDesc& a, cudaStream_t stream) {
  using Cfg = AttCfg<D>;
  CUtensorMap tq, tk, tv;
  int rc = make_tmap_3d_bf16(&tq, a.q, (uint64_t)a.heads * D, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.ldq * 2,
                             (uint64_t)a.q_batch_stride * 2, D, ATT_BM, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tk, a.k, (uint64_t)a.heads * D, (uint64_t)a.Tk, (uint64_t)a.B, (uint64_t)a.ldkv * 2,
                         (uint64_t)a.kv_batch_stride * 2, D, ATT_BN, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tv, a.v, (uint64_t)a.heads * D, (uint64_t)a.Tk, (uint64_t)a.B, (uint64_t)a.ldkv * 2,
                         (uint64_t)a.kv_batch_stride * 2, D, ATT_BN, 1, Cfg::RB);
  if (rc) return rc;
  AttFwdParams p;
  p.T = a.T; p.Tk = a.Tk; p.heads = a.heads;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.o = (__nv_bfloat16*)a.o; p.ldo = a.ldo; p.o_batch_stride = a.o_batch_stride; p.lse = a.lse;
  const size_t smem = 3 * Cfg::TILE + 32768 + 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SFK_REQUIRE(e == cudaSuccess, "attention_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  dim3 grid((a.T + ATT_BM - 1) / ATT_BM, a.heads, a.B);
  launch_k(attn_fwd_kernel<D>, dim3(grid), dim3(160), smem, stream, tq, tk, tv, p);
  return check_launch("attn_fwd_kernel");
}

int attention_fwd(const AttentionDesc& a, cudaStream_t stream) {
  SFK_REQUIRE(a.D == 64 || a.D == 32, "attention: head dim %d unsupported (32 or 64)", a.D);
  SFK_REQUIRE(a.T > 0 && a.Tk > 0 && a.B > 0 && a.heads > 0, "attention: empty problem");
  SFK_REQUIRE(a.B <= 65535 && a.heads <= 65535, "attention: batch/heads exceed grid limits");
  SFK_REQUIRE(a.ldo % 8 == 0 && (reinterpret_cast<uintptr_t>(a.o) & 15) == 0, "attention: output must be 16-byte aligned");
  if (g_attn_fwd256 && a.Tk <= 256) return a.D == 64 ? attn_fwd256_launch<64>(a, stream) : attn_fwd256_launch<32>(a, stream);
  return a.D == 64 ? attn_fwd_launch<64>(a, stream) : attn_fwd_launch<32>(a, stream);
}

// ------------------------------------------------------------------------------------------------
// delta[b, h, t] = sum_d dO[b,t,h*D+d] * O[b,t,h*D+d]
// ------------------------------------------------------------------------------------------------
// Optionally also zero-fills the gradient accumulators of the main kernel (dq: same (b, t, h, chunk) geometry as this pass; dk|dv:
// a flat block, grid-stride) - two memset launches per layer less.
__global__ void __launch_bounds__(256) attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dO,
                                                         float* __restrict__ delta, int B, int T, int heads, int D,
                                                         long long ld, long long batch_stride, float* __restrict__ zq,
                                                         long long zq_ld, long long zq_batch_stride, float4* __restrict__ zkv,
                                                         long long zkv_n4) {
  pdl_sync();
  if (zkv != nullptr) {
    const long long nthreads = (long long)gridDim.x * gridDim.y * 256;
    for (long long j = ((long long)blockIdx.y * gridDim.x + blockIdx.x) * 256 + threadIdx.x; j < zkv_n4; j += nthreads)
      zkv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  // one thread per (t, h, 8-element chunk) of sample blockIdx.y; reduce across the D/8 chunk lanes with shuffles.  32-bit index
  // arithmetic (the first version decoded a flat 64-bit index with four long-long divisions per thread).
  const unsigned chunks = (unsigned)D / 8;                      // 8 | 4: a power of two
  const unsigned per_sample = (unsigned)T * (unsigned)heads * chunks;
  const unsigned i = blockIdx.x * 256u + threadIdx.x;
  const long long b = blockIdx.y;
  const unsigned th = i / chunks, ch = i % chunks;
  const unsigned t = th / (unsigned)heads, h = th - t * (unsigned)heads;
  float s = 0.f;
  if (i < per_sample) {
    const long long off = b * batch_stride + (long long)t * ld + h * D + ch * 8;
    const uint4 a = *reinterpret_cast<const uint4*>(o + off);
    const uint4 g = *reinterpret_cast<const uint4*>(dO + off);
    if (zq != nullptr) {
      float4* z = reinterpret_cast<float4*>(zq + b * zq_batch_stride + (long long)t * zq_ld + h * D + ch * 8);
      z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
      z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* gp = reinterpret_cast<const uint32_t*>(&g);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 x = unpack_bf16x2(ap[k]), y = unpack_bf16x2(gp[k]);
      s += x.x * y.x + x.y * y.y;
    }
  }
  for (unsigned o2 = chunks >> 1; o2 > 0; o2 >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o2);
  if (i < per_sample && ch == 0) delta[(b * heads + h) * T + t] = s;
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
// sfk_set_option("attn_bwd_persist", 0|1): one persistent CTA per SM walking an equal share of the (batch, head, K/V block,
// query tile) units (default) instead of one CTA per (K/V block, query range)
int g_attn_bwd_persist = 1;

struct AttBwdParams {
  int T, Tk, heads, batches;
  int q_tiles_per_cta, q_tiles;
  float scale, scale_log2;
  const float* lse;    // [B, heads, T] log2 domain
  const float* delta;  // [B, heads, T]
  float* dq;           // fp32 [B, T, ld_dq] accumulated with red.add (pre-zeroed)
  long long ld_dq, dq_batch_stride;
  float* dk;           // fp32, accumulated with red.add (pre-zeroed)
  float* dv;
  long long ld_dkv, dkv_batch_stride;
};

template <int D>
__global__ void __launch_bounds__(288, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                const AttBwdParams p) {
  pdl_launch_dependents();
  using Cfg = AttCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // tiles: K, V (resident), Q[2], dO[2] (double buffered), P^T (32 KB), dS^T (32 KB), lse/delta [2][2][128] floats
  const uint32_t sK = base, sV = sK + Cfg::TILE;
  const uint32_t sQ = sV + Cfg::TILE;             // 2 buffers
  const uint32_t sDO = sQ + 2 * Cfg::TILE;        // 2 buffers
  const uint32_t sPT = sDO + 2 * Cfg::TILE;
  const uint32_t sDS = sPT + 32768;
  const uint32_t off_stats = 6 * Cfg::TILE + 65536;
  uint8_t* sPT_gen = gen + 6 * Cfg::TILE;
  uint8_t* sDS_gen = sPT_gen + 32768;
  float* stats = reinterpret_cast<float*>(gen + off_stats);  // [buf][0: lse, 1: delta][128]
  const uint32_t bars = base + off_stats + 2048;
  const uint32_t b_kvfull = bars, b_qfull = bars + 8 /*2*/, b_spfull = bars + 24, b_pdready = bars + 32, b_dqfull = bars + 40,
                 b_tilefree = bars + 48 /*2*/, b_dqfree = bars + 64, b_statsfree = bars + 72 /*2*/, b_final = bars + 88,
                 tmem_slot = bars + 96;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + off_stats + 2048 + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kvblk = blockIdx.x % ((p.Tk + ATT_BN - 1) / ATT_BN);
  const int qsplit = blockIdx.x / ((p.Tk + ATT_BN - 1) / ATT_BN);
  const int head = blockIdx.y, batch = blockIdx.z;
  const int kv0 = kvblk * ATT_BN;
  const int tile0 = qsplit * p.q_tiles_per_cta;
  const int ntiles = min(p.q_tiles_per_cta, p.q_tiles - tile0);

  if (threadIdx.x == 256) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(b_kvfull, 1);
    mbar_init(b_qfull, 1); mbar_init(b_qfull + 8, 1);
    mbar_init(b_spfull, 1); mbar_init(b_pdready, 256); mbar_init(b_dqfull, 1);
    mbar_init(b_tilefree, 1); mbar_init(b_tilefree + 8, 1);
    mbar_init(b_dqfree, 256);
    mbar_init(b_statsfree, 256); mbar_init(b_statsfree + 8, 256);
    mbar_init(b_final, 1);
    mbar_fence_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();  // setup above overlaps the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 256 + D, tDQ = tmem + 256 + 2 * D;

  if (warp == 8) {
    if (lane == 0 && ntiles > 0) {
      // ---------------- control thread ----------------
      const uint32_t idesc_s = umma_idesc_bf16(128, ATT_BM, 0, 0);   // S^T / dP^T : [kv x q]
      const uint32_t idesc_kv = umma_idesc_bf16(128, D, 0, 1);       // dV / dK    : A = P^T|dS^T (K-major), B = dO|Q (MN-major)
      const uint32_t idesc_dq = umma_idesc_bf16(128, D, 1, 1);       // dQ         : A = dS^T as MN-major, B = K (MN-major)
      mbar_arrive_expect_tx(b_kvfull, 2 * Cfg::TILE);
      tma_load_3d(sK, &tmap_k, b_kvfull, head * D, kv0, batch);
      tma_load_3d(sV, &tmap_v, b_kvfull, head * D, kv0, batch);
      // prefetch first query tile
      mbar_arrive_expect_tx(b_qfull, 2 * Cfg::TILE);
      tma_load_3d(sQ, &tmap_q, b_qfull, head * D, tile0 * ATT_BM, batch);
      tma_load_3d(sDO, &tmap_do, b_qfull, head * D, tile0 * ATT_BM, batch);
      mbar_wait(b_kvfull, 0);
      for (int i = 0; i < ntiles; ++i) {
        const int buf = i & 1;
        const uint32_t ph_buf = (i >> 1) & 1;   // phase of per-buffer barriers
        const uint32_t ph = i & 1;              // phase of once-per-tile barriers
        // prefetch the next tile into the other buffer once the MMAs that read it (tile i-1) have retired
        if (i + 1 < ntiles) {
          const int nb = buf ^ 1;
          if (i >= 1) mbar_wait(b_tilefree + 8 * nb, ((i - 1) >> 1) & 1);
          mbar_arrive_expect_tx(b_qfull + 8 * nb, 2 * Cfg::TILE);
          tma_load_3d(sQ + nb * Cfg::TILE, &tmap_q, b_qfull + 8 * nb, head * D, (tile0 + i + 1) * ATT_BM, batch);
          tma_load_3d(sDO + nb * Cfg::TILE, &tmap_do, b_qfull + 8 * nb, head * D, (tile0 + i + 1) * ATT_BM, batch);
        }
        mbar_wait(b_qfull + 8 * buf, ph_buf);
        // S^T/dP^T of the previous tile must have been consumed: implied by pd_ready(i-1), waited below last iteration
        tc_fence_after();
        const uint32_t q_s = sQ + buf * Cfg::TILE, do_s = sDO + buf * Cfg::TILE;
#pragma unroll
        for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
          umma_bf16(tST, umma_smem_desc_lt(sK + kk * 32, 16, 8 * Cfg::RB, Cfg::LT),
                    umma_smem_desc_lt(q_s + kk * 32, 16, 8 * Cfg::RB, Cfg::LT), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
          umma_bf16(tDPT, umma_smem_desc_lt(sV + kk * 32, 16, 8 * Cfg::RB, Cfg::LT),
                    umma_smem_desc_lt(do_s + kk * 32, 16, 8 * Cfg::RB, Cfg::LT), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(b_spfull);
        mbar_wait(b_pdready, ph);
        if (i >= 1) mbar_wait(b_dqfree, (i - 1) & 1);   // dQ TMEM of tile i-1 drained
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < ATT_BM / 16; ++ks) {
          const uint32_t a_off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_bf16(tDV, umma_smem_desc_lt(sPT + a_off, 16, 1024, 2u),
                    umma_smem_desc_lt(do_s + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_kv,
                    (i > 0 || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < ATT_BM / 16; ++ks) {
          const uint32_t a_off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_bf16(tDK, umma_smem_desc_lt(sDS + a_off, 16, 1024, 2u),
                    umma_smem_desc_lt(q_s + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_kv,
                    (i > 0 || ks > 0) ? 1u : 0u);
        }
        // dQ[q, d] = sum_kv dS[q, kv] K[kv, d]: A = dS^T tile read as MN-major (m = q contiguous, k = kv rows)
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks)
          umma_bf16(tDQ, umma_smem_desc_lt(sDS + ks * 2048, 16384, 1024, 2u),
                    umma_smem_desc_lt(sK + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_dq,
                    ks > 0 ? 1u : 0u);
        umma_commit(b_dqfull);
        umma_commit(b_tilefree + 8 * buf);
        if (i == ntiles - 1) umma_commit(b_final);
      }
    }
  } else if (ntiles > 0) {
    // ---------------- 256 math threads: threads r and r + 128 share K/V row kv0+r (S^T, dP^T, dK, dV) and query row r of
    // dQ; each takes one half of the columns (two warps per TMEM lane quadrant) ----------------
    const int r = threadIdx.x & 127;
    const int hhalf = threadIdx.x >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const bool kv_ok = (kv0 + r) < p.Tk;
    const long long stat_base = ((long long)batch * p.heads + head) * p.T;
    for (int i = 0; i < ntiles; ++i) {
      const int buf = i & 1;
      const uint32_t ph = i & 1;
      const int t0 = (tile0 + i) * ATT_BM;
      // stage lse / delta of this query tile (thread r loads query t0 + r)
      if (i >= 2) mbar_wait(b_statsfree + 8 * buf, ((i - 2) >> 1) & 1);
      {
        const int t = t0 + r;
        if (hhalf == 0) stats[buf * 256 + r] = t < p.T ? p.lse[stat_base + t] : INFINITY;
        else stats[buf * 256 + 128 + r] = t < p.T ? p.delta[stat_base + t] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(b_spfull, ph);
      tc_fence_after();
#pragma unroll 1
      for (int cc = hhalf * (ATT_BM / 2); cc < (hhalf + 1) * (ATT_BM / 2); cc += 32) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32(tST + lane_base + cc, s);
        tmem_ld_32x32(tDPT + lane_base + cc, dp);
        tmem_ld_wait();
        float pv[32], dsv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float l2 = stats[buf * 256 + cc + j];
          const float dl = stats[buf * 256 + 128 + cc + j];
          const float pr = kv_ok ? fast_exp2(fmaf(__uint_as_float(s[j]), p.scale_log2, -l2)) : 0.f;
          pv[j] = pr;
          dsv[j] = pr * (__uint_as_float(dp[j]) - dl) * p.scale;
        }
        uint8_t* tp = sPT_gen + (cc >> 6) * 16384;
        uint8_t* td = sDS_gen + (cc >> 6) * 16384;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 o;
          o.x = pack_bf16x2(pv[j8 * 8 + 0], pv[j8 * 8 + 1]); o.y = pack_bf16x2(pv[j8 * 8 + 2], pv[j8 * 8 + 3]);
          o.z = pack_bf16x2(pv[j8 * 8 + 4], pv[j8 * 8 + 5]); o.w = pack_bf16x2(pv[j8 * 8 + 6], pv[j8 * 8 + 7]);
          st_swizzled_chunk(tp, r, ((cc & 63) >> 3) + j8, o);
          o.x = pack_bf16x2(dsv[j8 * 8 + 0], dsv[j8 * 8 + 1]); o.y = pack_bf16x2(dsv[j8 * 8 + 2], dsv[j8 * 8 + 3]);
          o.z = pack_bf16x2(dsv[j8 * 8 + 4], dsv[j8 * 8 + 5]); o.w = pack_bf16x2(dsv[j8 * 8 + 6], dsv[j8 * 8 + 7]);
          st_swizzled_chunk(td, r, ((cc & 63) >> 3) + j8, o);
        }
      }
      mbar_arrive(b_statsfree + 8 * buf);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(b_pdready);
      // dQ tile -> global (thread r = query row t0 + r)
      mbar_wait(b_dqfull, ph);
      tc_fence_after();
      {
        const int t = t0 + r;
        float* dst = p.dq + batch * p.dq_batch_stride + (long long)t * p.ld_dq + head * D;
#pragma unroll
        for (int cc = hhalf * 32; cc < D; cc += 64) {
          uint32_t v[32];
          tmem_ld_32x32(tDQ + lane_base + cc, v);
          tmem_ld_wait();
          if (t < p.T) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4_f32_g(dst + cc + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(b_dqfree);
    }
    // dK / dV of this K/V block
    mbar_wait(b_final, 0);
    tc_fence_after();
    {
      float* dkp = p.dk + batch * p.dkv_batch_stride + (long long)(kv0 + r) * p.ld_dkv + head * D;
      float* dvp = p.dv + batch * p.dkv_batch_stride + (long long)(kv0 + r) * p.ld_dkv + head * D;
#pragma unroll
      for (int cc = hhalf * 32; cc < D; cc += 64) {
        uint32_t a[32], b[32];
        tmem_ld_32x32(tDK + lane_base + cc, a);
        tmem_ld_32x32(tDV + lane_base + cc, b);
        tmem_ld_wait();
        if (kv_ok) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            red_add_v4_f32_g(dkp + cc + j, __uint_as_float(a[j]), __uint_as_float(a[j + 1]), __uint_as_float(a[j + 2]),
                             __uint_as_float(a[j + 3]));
            red_add_v4_f32_g(dvp + cc + j, __uint_as_float(b[j]), __uint_as_float(b[j + 1]), __uint_as_float(b[j + 2]),
                             __uint_as_float(b[j + 3]));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int D>
__global__ void __launch_bounds__(288, 1)
attn_bwd_persist_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_do,
                const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_v,
                const AttBwdParams p) {
  pdl_launch_dependents();
  using Cfg = AttCfg<D>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  // tiles: K, V (resident), Q[2], dO[2] (double buffered), P^T (32 KB), dS^T (32 KB), lse/delta [2][2][128] floats
  const uint32_t sK = base, sV = sK + Cfg::TILE;
  const uint32_t sQ = sV + Cfg::TILE;             // 2 buffers
  const uint32_t sDO = sQ + 2 * Cfg::TILE;        // 2 buffers
  const uint32_t sPT = sDO + 2 * Cfg::TILE;
  const uint32_t sDS = sPT + 32768;
  const uint32_t off_stats = 6 * Cfg::TILE + 65536;
  uint8_t* sPT_gen = gen + 6 * Cfg::TILE;
  uint8_t* sDS_gen = sPT_gen + 32768;
  float* stats = reinterpret_cast<float*>(gen + off_stats);  // [buf][0: lse, 1: delta][128]
  const uint32_t bars = base + off_stats + 2048;
  const uint32_t b_kvfull = bars, b_qfull = bars + 8 /*2*/, b_spfull = bars + 24, b_pdready = bars + 32, b_dqfull = bars + 40,
                 b_tilefree = bars + 48 /*2*/, b_dqfree = bars + 64, b_statsfree = bars + 72 /*2*/, b_final = bars + 88,
                 b_kvdrained = bars + 96, tmem_slot = bars + 104;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(gen + off_stats + 2048 + 104);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // persistent even split: the (batch, head, K/V block, query tile) units, query tile fastest, form one list; CTA c owns the
  // contiguous range [c U / G, (c+1) U / G).  Consecutive units with the same K/V block form a segment: K/V stay in shared
  // memory and dK/dV in tensor memory across it, then flush by atomics (they are atomics anyway: several CTAs share a block).
  const int kv_blocks = (p.Tk + ATT_BN - 1) / ATT_BN;
  const long long U = (long long)p.batches * p.heads * kv_blocks * p.q_tiles;
  const long long u_begin = U * blockIdx.x / gridDim.x, u_end = U * (blockIdx.x + 1) / gridDim.x;
  const int ntiles = (int)(u_end - u_begin);
  // (U < 2^31, checked on the host: 32-bit divisions, and only for a CTA's FIRST unit - the walk to the next unit is an
  // increment with carries; six long-long divisions per unit in every thread were ~0.4 us on each tile's critical path)
  auto decode = [&](long long u, int& qt, int& kvb, int& hd, int& bt) {
    const unsigned uu = (unsigned)u;
    qt = (int)(uu % (unsigned)p.q_tiles);
    unsigned t = uu / (unsigned)p.q_tiles;
    kvb = (int)(t % (unsigned)kv_blocks);
    t /= (unsigned)kv_blocks;
    hd = (int)(t % (unsigned)p.heads);
    bt = (int)(t / (unsigned)p.heads);
  };
  auto next_unit = [&](int qt, int kvb, int hd, int bt, int& nqt, int& nkvb, int& nhd, int& nbt) {
    nqt = qt + 1; nkvb = kvb; nhd = hd; nbt = bt;
    if (nqt == p.q_tiles) {
      nqt = 0;
      if (++nkvb == kv_blocks) {
        nkvb = 0;
        if (++nhd == p.heads) { nhd = 0; ++nbt; }
      }
    }
  };

  if (threadIdx.x == 256) {
    tma_prefetch_desc(&tmap_q); tma_prefetch_desc(&tmap_do); tma_prefetch_desc(&tmap_k); tma_prefetch_desc(&tmap_v);
    mbar_init(b_kvfull, 1);
    mbar_init(b_qfull, 1); mbar_init(b_qfull + 8, 1);
    mbar_init(b_spfull, 1); mbar_init(b_pdready, 256); mbar_init(b_dqfull, 1);
    mbar_init(b_tilefree, 1); mbar_init(b_tilefree + 8, 1);
    mbar_init(b_dqfree, 256);
    mbar_init(b_statsfree, 256); mbar_init(b_statsfree + 8, 256);
    mbar_init(b_final, 1);
    mbar_init(b_kvdrained, 256);
    mbar_fence_init();
  }
  if (warp == 8) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  pdl_wait();  // setup above overlaps the previous kernel's tail
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot_gen;
  const uint32_t tST = tmem, tDPT = tmem + 128, tDV = tmem + 256, tDK = tmem + 256 + D, tDQ = tmem + 256 + 2 * D;

  if (warp == 8) {
    if (lane == 0 && ntiles > 0) {
      // ---------------- control thread ----------------
      const uint32_t idesc_s = umma_idesc_bf16(128, ATT_BM, 0, 0);   // S^T / dP^T : [kv x q]
      const uint32_t idesc_kv = umma_idesc_bf16(128, D, 0, 1);       // dV / dK    : A = P^T|dS^T (K-major), B = dO|Q (MN-major)
      const uint32_t idesc_dq = umma_idesc_bf16(128, D, 1, 1);       // dQ         : A = dS^T as MN-major, B = K (MN-major)
      int qt, kvb, head, batch;
      decode(u_begin, qt, kvb, head, batch);
      mbar_arrive_expect_tx(b_kvfull, 2 * Cfg::TILE);
      tma_load_3d(sK, &tmap_k, b_kvfull, head * D, kvb * ATT_BN, batch);
      tma_load_3d(sV, &tmap_v, b_kvfull, head * D, kvb * ATT_BN, batch);
      // prefetch first query tile
      mbar_arrive_expect_tx(b_qfull, 2 * Cfg::TILE);
      tma_load_3d(sQ, &tmap_q, b_qfull, head * D, qt * ATT_BM, batch);
      tma_load_3d(sDO, &tmap_do, b_qfull, head * D, qt * ATT_BM, batch);
      mbar_wait(b_kvfull, 0);
      int seg = 0;           // segment index (K/V block changes)
      bool seg_first = true; // first tile of the current segment: dV / dK start from zero
      for (int i = 0; i < ntiles; ++i) {
        const int buf = i & 1;
        const uint32_t ph_buf = (i >> 1) & 1;   // phase of per-buffer barriers
        const uint32_t ph = i & 1;              // phase of once-per-tile barriers
        // coordinates of the next unit; does it start a new segment?
        int nqt = 0, nkvb = kvb, nhead = head, nbatch = batch;
        const bool has_next = i + 1 < ntiles;
        if (has_next) next_unit(qt, kvb, head, batch, nqt, nkvb, nhead, nbatch);
        const bool seg_last = !has_next || nkvb != kvb || nhead != head || nbatch != batch;
        // prefetch the next tile into the other buffer once the MMAs that read it (tile i-1) have retired
        if (has_next) {
          const int nb = buf ^ 1;
          if (i >= 1) mbar_wait(b_tilefree + 8 * nb, ((i - 1) >> 1) & 1);
          mbar_arrive_expect_tx(b_qfull + 8 * nb, 2 * Cfg::TILE);
          tma_load_3d(sQ + nb * Cfg::TILE, &tmap_q, b_qfull + 8 * nb, nhead * D, nqt * ATT_BM, nbatch);
          tma_load_3d(sDO + nb * Cfg::TILE, &tmap_do, b_qfull + 8 * nb, nhead * D, nqt * ATT_BM, nbatch);
        }
        mbar_wait(b_qfull + 8 * buf, ph_buf);
        tc_fence_after();
        const uint32_t q_s = sQ + buf * Cfg::TILE, do_s = sDO + buf * Cfg::TILE;
#pragma unroll
        for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
          umma_bf16(tST, umma_smem_desc_lt(sK + kk * 32, 16, 8 * Cfg::RB, Cfg::LT),
                    umma_smem_desc_lt(q_s + kk * 32, 16, 8 * Cfg::RB, Cfg::LT), idesc_s, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < Cfg::KSTEPS; ++kk)
          umma_bf16(tDPT, umma_smem_desc_lt(sV + kk * 32, 16, 8 * Cfg::RB, Cfg::LT),
                    umma_smem_desc_lt(do_s + kk * 32, 16, 8 * Cfg::RB, Cfg::LT), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(b_spfull);
        mbar_wait(b_pdready, ph);
        if (i >= 1) mbar_wait(b_dqfree, (i - 1) & 1);   // dQ TMEM of tile i-1 drained
        if (seg_first && seg >= 1) mbar_wait(b_kvdrained, (seg - 1) & 1);   // previous segment's dK / dV read out
        tc_fence_after();
#pragma unroll
        for (int ks = 0; ks < ATT_BM / 16; ++ks) {
          const uint32_t a_off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_bf16(tDV, umma_smem_desc_lt(sPT + a_off, 16, 1024, 2u),
                    umma_smem_desc_lt(do_s + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_kv,
                    (!seg_first || ks > 0) ? 1u : 0u);
        }
#pragma unroll
        for (int ks = 0; ks < ATT_BM / 16; ++ks) {
          const uint32_t a_off = (ks >> 2) * 16384 + (ks & 3) * 32;
          umma_bf16(tDK, umma_smem_desc_lt(sDS + a_off, 16, 1024, 2u),
                    umma_smem_desc_lt(q_s + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_kv,
                    (!seg_first || ks > 0) ? 1u : 0u);
        }
        // dQ[q, d] = sum_kv dS[q, kv] K[kv, d]: A = dS^T tile read as MN-major (m = q contiguous, k = kv rows)
#pragma unroll
        for (int ks = 0; ks < ATT_BN / 16; ++ks)
          umma_bf16(tDQ, umma_smem_desc_lt(sDS + ks * 2048, 16384, 1024, 2u),
                    umma_smem_desc_lt(sK + ks * 16 * Cfg::RB, 128 * Cfg::RB, 8 * Cfg::RB, Cfg::LT), idesc_dq,
                    ks > 0 ? 1u : 0u);
        umma_commit(b_dqfull);
        umma_commit(b_tilefree + 8 * buf);
        seg_first = false;
        if (seg_last) {
          umma_commit(b_final);
          if (has_next) {
            // next segment: K / V may be overwritten once every MMA of this one has retired
            mbar_wait(b_final, seg & 1);
            mbar_arrive_expect_tx(b_kvfull, 2 * Cfg::TILE);
            tma_load_3d(sK, &tmap_k, b_kvfull, nhead * D, nkvb * ATT_BN, nbatch);
            tma_load_3d(sV, &tmap_v, b_kvfull, nhead * D, nkvb * ATT_BN, nbatch);
            mbar_wait(b_kvfull, (seg + 1) & 1);
            ++seg;
            seg_first = true;
          }
        }
        qt = nqt; kvb = nkvb; head = nhead; batch = nbatch;
      }
    }
  } else if (ntiles > 0) {
    // ---------------- 256 math threads: threads r and r + 128 share K/V row kv0+r (S^T, dP^T, dK, dV) and query row r of
    // dQ; each takes one half of the columns (two warps per TMEM lane quadrant) ----------------
    const int r = threadIdx.x & 127;
    const int hhalf = threadIdx.x >> 7;
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    int qt, kvb, head, batch;
    decode(u_begin, qt, kvb, head, batch);
    int seg = 0;
    for (int i = 0; i < ntiles; ++i) {
      const int buf = i & 1;
      const uint32_t ph = i & 1;
      int nqt = 0, nkvb = kvb, nhead = head, nbatch = batch;
      const bool has_next = i + 1 < ntiles;
      if (has_next) next_unit(qt, kvb, head, batch, nqt, nkvb, nhead, nbatch);
      const bool seg_last = !has_next || nkvb != kvb || nhead != head || nbatch != batch;
      const int kv0 = kvb * ATT_BN;
      const bool kv_ok = (kv0 + r) < p.Tk;
      const long long stat_base = ((long long)batch * p.heads + head) * p.T;
      const int t0 = qt * ATT_BM;
      // stage lse / delta of this query tile (thread r loads query t0 + r)
      if (i >= 2) mbar_wait(b_statsfree + 8 * buf, ((i - 2) >> 1) & 1);
      {
        const int t = t0 + r;
        if (hhalf == 0) stats[buf * 256 + r] = t < p.T ? p.lse[stat_base + t] : INFINITY;
        else stats[buf * 256 + 128 + r] = t < p.T ? p.delta[stat_base + t] : 0.f;
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(b_spfull, ph);
      tc_fence_after();
#pragma unroll 1
      for (int cc = hhalf * (ATT_BM / 2); cc < (hhalf + 1) * (ATT_BM / 2); cc += 32) {
        uint32_t s[32], dp[32];
        tmem_ld_32x32(tST + lane_base + cc, s);
        tmem_ld_32x32(tDPT + lane_base + cc, dp);
        tmem_ld_wait();
        float pv[32], dsv[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float l2 = stats[buf * 256 + cc + j];
          const float dl = stats[buf * 256 + 128 + cc + j];
          const float pr = kv_ok ? fast_exp2(fmaf(__uint_as_float(s[j]), p.scale_log2, -l2)) : 0.f;
          pv[j] = pr;
          dsv[j] = pr * (__uint_as_float(dp[j]) - dl) * p.scale;
        }
        uint8_t* tp = sPT_gen + (cc >> 6) * 16384;
        uint8_t* td = sDS_gen + (cc >> 6) * 16384;
#pragma unroll
        for (int j8 = 0; j8 < 4; ++j8) {
          uint4 o;
          o.x = pack_bf16x2(pv[j8 * 8 + 0], pv[j8 * 8 + 1]); o.y = pack_bf16x2(pv[j8 * 8 + 2], pv[j8 * 8 + 3]);
          o.z = pack_bf16x2(pv[j8 * 8 + 4], pv[j8 * 8 + 5]); o.w = pack_bf16x2(pv[j8 * 8 + 6], pv[j8 * 8 + 7]);
          st_swizzled_chunk(tp, r, ((cc & 63) >> 3) + j8, o);
          o.x = pack_bf16x2(dsv[j8 * 8 + 0], dsv[j8 * 8 + 1]); o.y = pack_bf16x2(dsv[j8 * 8 + 2], dsv[j8 * 8 + 3]);
          o.z = pack_bf16x2(dsv[j8 * 8 + 4], dsv[j8 * 8 + 5]); o.w = pack_bf16x2(dsv[j8 * 8 + 6], dsv[j8 * 8 + 7]);
          st_swizzled_chunk(td, r, ((cc & 63) >> 3) + j8, o);
        }
      }
      mbar_arrive(b_statsfree + 8 * buf);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(b_pdready);
      // dQ tile -> global (thread r = query row t0 + r)
      mbar_wait(b_dqfull, ph);
      tc_fence_after();
      {
        const int t = t0 + r;
        float* dst = p.dq + batch * p.dq_batch_stride + (long long)t * p.ld_dq + head * D;
#pragma unroll
        for (int cc = hhalf * 32; cc < D; cc += 64) {
          uint32_t v[32];
          tmem_ld_32x32(tDQ + lane_base + cc, v);
          tmem_ld_wait();
          if (t < p.T) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4_f32_g(dst + cc + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                               __uint_as_float(v[j + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(b_dqfree);
      if (seg_last) {
        // dK / dV of this segment's K/V block
        mbar_wait(b_final, seg & 1);
        tc_fence_after();
        float* dkp = p.dk + batch * p.dkv_batch_stride + (long long)(kv0 + r) * p.ld_dkv + head * D;
        float* dvp = p.dv + batch * p.dkv_batch_stride + (long long)(kv0 + r) * p.ld_dkv + head * D;
#pragma unroll
        for (int cc = hhalf * 32; cc < D; cc += 64) {
          uint32_t a[32], b[32];
          tmem_ld_32x32(tDK + lane_base + cc, a);
          tmem_ld_32x32(tDV + lane_base + cc, b);
          tmem_ld_wait();
          if (kv_ok) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              red_add_v4_f32_g(dkp + cc + j, __uint_as_float(a[j]), __uint_as_float(a[j + 1]), __uint_as_float(a[j + 2]),
                               __uint_as_float(a[j + 3]));
              red_add_v4_f32_g(dvp + cc + j, __uint_as_float(b[j]), __uint_as_float(b[j + 1]), __uint_as_float(b[j + 2]),
                               __uint_as_float(b[j + 3]));
            }
          }
        }
        tc_fence_before();
        mbar_arrive(b_kvdrained);
        ++seg;
      }
      qt = nqt; kvb = nkvb; head = nhead; batch = nbatch;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 8) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int D>
static int attn_bwd_launch(const AttentionBwdDesc& a, cudaStream_t stream) {
  using Cfg = AttCfg<D>;
  CUtensorMap tq, tdo, tk, tv;
  int rc = make_tmap_3d_bf16(&tq, a.q, (uint64_t)a.heads * D, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.ldq * 2,
                             (uint64_t)a.q_batch_stride * 2, D, ATT_BM, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tdo, a.dO, (uint64_t)a.heads * D, (uint64_t)a.T, (uint64_t)a.B, (uint64_t)a.ldo * 2,
                         (uint64_t)a.o_batch_stride * 2, D, ATT_BM, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tk, a.k, (uint64_t)a.heads * D, (uint64_t)a.Tk, (uint64_t)a.B, (uint64_t)a.ldkv * 2,
                         (uint64_t)a.kv_batch_stride * 2, D, ATT_BN, 1, Cfg::RB);
  if (rc) return rc;
  rc = make_tmap_3d_bf16(&tv, a.v, (uint64_t)a.heads * D, (uint64_t)a.Tk, (uint64_t)a.B, (uint64_t)a.ldkv * 2,
                         (uint64_t)a.kv_batch_stride * 2, D, ATT_BN, 1, Cfg::RB);
  if (rc) return rc;

  // delta = rowsum(dO * O)
  {
    const long long total = (long long)a.T * a.heads * (D / 8);   // threads per sample; grid.y = sample
    SFK_REQUIRE(total < (1LL << 31), "attention_bwd: sample too large for the delta pass");
    float* zq = nullptr;
    float4* zkv = nullptr;
    long long zkv_n4 = 0;
    if (a.zero_fill) {
      SFK_REQUIRE(a.dv == a.dk + a.ld_dkv / 2 && a.dkv_batch_stride == (long long)a.Tk * a.ld_dkv &&
                  (reinterpret_cast<uintptr_t>(a.dk) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.dq) & 15) == 0,
                  "attention_bwd: zero_fill needs dk|dv as one contiguous 16-byte aligned [B, Tk, ld_dkv] block");
      zq = a.dq;
      zkv = reinterpret_cast<float4*>(a.dk);
      zkv_n4 = (long long)a.B * a.Tk * a.ld_dkv / 4;
    }
    launch_k(attn_delta_kernel, dim3((unsigned)((total + 255) / 256), (unsigned)a.B), dim3(256), 0, stream, (const __nv_bfloat16*)a.o, (const __nv_bfloat16*)a.dO,
                                                                           a.delta, a.B, a.T, a.heads, D, a.ldo,
                                                                           a.o_batch_stride, zq, a.ld_dq, a.dq_batch_stride, zkv, zkv_n4);
    rc = check_launch("attn_delta_kernel");
    if (rc) return rc;
  }
  AttBwdParams p;
  p.T = a.T; p.Tk = a.Tk; p.heads = a.heads;
  p.q_tiles = (a.T + ATT_BM - 1) / ATT_BM;
  const int kv_blocks = (a.Tk + ATT_BN - 1) / ATT_BN;
  const long long base_ctas = (long long)kv_blocks * a.heads * a.B;
  int qsplits = (int)((2LL * num_sms() + base_ctas - 1) / base_ctas);
  if (qsplits < 1) qsplits = 1;
  if (qsplits > p.q_tiles) qsplits = p.q_tiles;
  p.q_tiles_per_cta = (p.q_tiles + qsplits - 1) / qsplits;
  qsplits = (p.q_tiles + p.q_tiles_per_cta - 1) / p.q_tiles_per_cta;
  p.scale = a.scale;
  p.scale_log2 = a.scale * 1.4426950408889634f;
  p.lse = a.lse; p.delta = a.delta;
  p.dq = a.dq; p.ld_dq = a.ld_dq; p.dq_batch_stride = a.dq_batch_stride;
  p.dk = a.dk; p.dv = a.dv; p.ld_dkv = a.ld_dkv; p.dkv_batch_stride = a.dkv_batch_stride;
  const size_t smem = 6 * Cfg::TILE + 65536 + 2048 + 128 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    SFK_REQUIRE(e == cudaSuccess, "attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  p.batches = a.B;
  SFK_REQUIRE((long long)a.B * a.heads * kv_blocks * p.q_tiles < (1LL << 31), "attention_bwd: too many work units");
  if (g_attn_bwd_persist) {
    static bool attr2 = false;
    if (!attr2) {
      cudaError_t e = cudaFuncSetAttribute(attn_bwd_persist_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      SFK_REQUIRE(e == cudaSuccess, "attention_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      attr2 = true;
    }
    const long long units = base_ctas * p.q_tiles;
    const int grid_p = (int)(units < num_sms() ? units : num_sms());
    launch_k(attn_bwd_persist_kernel<D>, dim3(grid_p), dim3(288), smem, stream, tq, tdo, tk, tv, p);
    return check_launch("attn_bwd_persist_kernel");
  }
  dim3 grid(kv_blocks * qsplits, a.heads, a.B);
  launch_k(attn_bwd_kernel<D>, dim3(grid), dim3(288), smem, stream, tq, tdo, tk, tv, p);
  return check_launch("attn_bwd_kernel");
}

int attention_bwd(const AttentionBwdDesc& a, cudaStream_t stream) {
  SFK_REQUIRE(a.D == 64 || a.D == 32, "attention_bwd: head dim %d unsupported (32 or 64)", a.D);
  SFK_REQUIRE(a.T > 0 && a.Tk > 0 && a.B > 0 && a.heads > 0, "attention_bwd: empty problem");
  SFK_REQUIRE(a.B <= 65535 && a.heads <= 65535, "attention_bwd: batch/heads exceed grid limits");
  SFK_REQUIRE(a.ld_dq % 4 == 0 && a.ld_dkv % 4 == 0, "attention_bwd: gradient pitches must be multiples of 4");
  return a.D == 64 ? attn_bwd_launch<64>(a, stream) : attn_bwd_launch<32>(a, stream);
}

}  // namespace sfk
