// tcgen05 / TMA / TMEM GEMM for sm_100a:   D[M,N] = alpha * (A · B^T) (+ bias) (* row_scale) (+ residual)
//
// One persistent CTA per SM, 6 warps:
//   warp 0      : TMA producer   (cp.async.bulk.tensor -> 128B-swizzled smem ring, mbarrier complete_tx)
//   warp 1      : MMA issuer     (one elected lane issues tcgen05.mma, accumulators in TMEM, double buffered)
//   warps 2..5  : epilogue       (tcgen05.ld -> registers -> bias / drop-path scale / residual -> global)
//
// Operand layouts ("major-ness") are runtime options so that the same kernel serves
//   forward  Linear      : A = X  [M,K] K-major,           B = W  [N,K] K-major
//   dgrad                : A = dY [M,N'] K-major,          B = W  [N',K'] used as [K'(n), N'(k)]  MN-major
//   wgrad (split-K)      : A = dY^T (stored [tokens,N'])   MN-major, B = X^T (stored [tokens,K']) MN-major
// which removes every explicit transpose from the backward pass.
//
// Reference semantics served by this kernel: nn.Linear / 1x1 Conv2d / k==stride Conv2d call sites in
// segformer_source/models/segformer/modeling_segformer.py:163-165,171-173,236,301,307,713,738 (forward) and their
// autograd-derived gradients.
#include "common.cuh"
#include "sfk_internal.h"

namespace sfk {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int MAX_BLOCK_N = 256;
static constexpr int TMEM_COLS = 512;  // 2 accumulator buffers x 256 fp32 columns
static constexpr int GEMM_THREADS = 192;
static constexpr int MAX_STAGES = 8;
// behind the barrier block: the current tile's bias slice (MAX_BLOCK_N floats), staged by the epilogue before it waits for
// the accumulator
static constexpr uint32_t EPI_BIAS_OFF = 16 * MAX_STAGES + 64;
static constexpr uint32_t EPI_BIAS_BYTES = 256 * 4;

// runtime switch (sfk_set_option("gemm_cluster2", 0|1|2)): CTA pairs (cta_group::2).  Two CTAs of a 2-cluster drive one
// 256 x N MMA per k-step; each loads its own 128 A rows and HALF of the B tile, so the L2 -> shared-memory stream per
// output element drops by a third and each SM's shared memory takes half of B.
//   0 = never, 1 = large problems only (default), 2 = whenever the shape allows.
// Measured on B200 inside the MiT-b5 step (per-launch CUDA events, profiles/): the three decode-head GEMMs
// (262144 x 768 x 3072 and its two gradients) gain 12-15 % (1.24 -> 1.05 ms); every K <= 1280 / M <= 65536 shape LOSES
// 5-25 % (half as many schedulable units, no two-CTAs-per-SM mode, TMEM taken after the grid dependency resolves).
int g_gemm_cluster2 = 1;
// small-K GEMMs: 2 resident CTAs per SM (<= 110 KB smem, 256 TMEM columns each) so one CTA's epilogue / pipeline fill
// overlaps the other's main loop
int g_gemm_occ2 = 1;

// optional in-kernel timeline (debug): CTA 0 stores globaltimer (ns) at phase boundaries into this device buffer
unsigned long long* g_gemm_trace = nullptr;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GEMM_TRACE(slot)                                                     \
  do {                                                                       \
    if (p.trace != nullptr && blockIdx.x == 0) p.trace[slot] = gtime();      \
  } while (0)

struct GemmKernelParams {
  int M, N, K;
  int block_n;
  int a_mn, b_mn;
  int splits;
  int k_blocks;  // ceil(K / BLOCK_K)
  int m_tiles, n_tiles;
  int stages;
  uint32_t stage_bytes;  // A + B bytes per stage
  // epilogue
  float alpha;
  const float* bias;
  const float* residual;
  long long ldr;
  RowMap res_map;
  RowMap out_map;
  const float* row_scale;
  int rows_per_scale;
  void* out;
  long long ldo;
  int out_bf16;
  int atomic_add;
  int relu;
  unsigned long long* trace;
  int tmem_cols;   // 512 (1 CTA / SM) or 256 (2 CTAs / SM); accumulator buffer a lives at column a * tmem_cols / 2
  // patch-view operands (rank-5 TMA): rows per sample, ox-row length, k-blocks per ky
  int a_patch, pa_rows_b, pa_wo, pa_kspan;
  int b_patch, pb_rows_b, pb_wo, pb_span;   // pb_span = sr*C (elements per ky)
  int cluster2;  // CTA-pair mode: launched as 2-CTA clusters, one cta_group::2 MMA per pair (rows 256, B split across the pair)
};

__device__ __forceinline__ long long gemm_map_row(long long row, const RowMap& m) {
  if (m.period > 0) row += (row / m.period) * m.extra;
  return row + m.offset;
}
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// PAIR = 1: CTA-pair kernel (must be launched in 2-CTA clusters: it contains cta_group::2 instructions)
template <int PAIR>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const GemmKernelParams p) {
  pdl_launch_dependents();  // let the next kernel of the stream get scheduled; it waits for our completion itself
  if (threadIdx.x == 0) GEMM_TRACE(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-B alignment is required by SWIZZLE_128B atoms; dynamic smem base is only 16-B aligned by contract.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;
  const uint32_t bars_off = p.stages * p.stage_bytes;
  // barrier block: full[S], empty[S], tmem_full[2], tmem_empty[2], tmem_ptr
  const uint32_t bar_full = smem_base + bars_off;
  const uint32_t bar_empty = bar_full + 8 * MAX_STAGES;
  const uint32_t bar_tfull = bar_empty + 8 * MAX_STAGES;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bars_off + 16 * MAX_STAGES + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, PAIR ? 8 : 4);   // pair: both CTAs' epilogue warps report to the leader
    }
    mbar_fence_init();
  }
  // pair kernels take their (pair-collective) TMEM allocation only after the grid dependency has resolved: a dependent
  // launched early under PDL that sits on both SMs' tensor memory while it waits was observed to deadlock against
  // kernels of other streams
  if (PAIR) pdl_wait();
  if (warp == 1) {
    if (PAIR) { tmem_alloc_pair(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish_pair(); }
    else { tmem_alloc(tmem_slot, (uint32_t)p.tmem_cols); tmem_relinquish(); }
  }
  // barrier init / TMEM allocation / descriptor prefetch above overlap the previous kernel's tail (PDL);
  // nothing before this point touches global memory produced by it.
  if (!PAIR) pdl_wait();
  tc_fence_before();
  __syncthreads();
  if (PAIR) cluster_sync_all();  // peer barriers are initialised before any remote arrive / commit targets them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;
  if (threadIdx.x == 0) GEMM_TRACE(1);

  // work units: (row tile | row-tile pair, column tile, k split); in cluster mode the two CTAs of a pair walk the same
  // unit sequence in lockstep, CTA rank r owning row tile 2*pair + r
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  const int m_groups = PAIR ? (p.m_tiles + 1) / 2 : p.m_tiles;
  const int tiles_mn = m_groups * p.n_tiles;
  const int total_units = tiles_mn * p.splits;
  const int unit0 = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int unit_stride = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int unit = unit0; unit < total_units; unit += unit_stride) {
        const int split = unit % p.splits;
        const int tile = unit / p.splits;
        const int n_blk = tile % p.n_tiles;
        const int m_blk = PAIR ? 2 * (tile / p.n_tiles) + (int)crank : tile / p.n_tiles;
        const int kb0 = (int)(((long long)split * p.k_blocks) / p.splits);
        const int kb1 = (int)(((long long)(split + 1) * p.k_blocks) / p.splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sb = sa + a_bytes;
          if (!PAIR) {
            const uint32_t full = bar_full + 8 * stage;
            mbar_arrive_expect_tx(full, p.stage_bytes);
            if (unit == unit0 && kb == kb0) GEMM_TRACE(2);
            if (p.a_patch) {
              // implicit im2col: 128 patch rows = (128 / Wo) ox-rows of one sample; this k-block = 64 of the (kx, c) run of one ky
              const int m0 = m_blk * BLOCK_M;
              const int bi = m0 / p.pa_rows_b, oy0 = (m0 - bi * p.pa_rows_b) / p.pa_wo;
              tma_load_5d(sa, &tmap_a, full, (kb % p.pa_kspan) * BLOCK_K, kb / p.pa_kspan, 0, oy0, bi);
            } else if (!p.a_mn) {
              tma_load_2d(sa, &tmap_a, full, kb * BLOCK_K, m_blk * BLOCK_M);
            } else {
              // stored [K, M] (M contiguous): 2 boxes of [64 m, 64 k]
              tma_load_2d(sa, &tmap_a, full, m_blk * BLOCK_M, kb * BLOCK_K);
              tma_load_2d(sa + 8192, &tmap_a, full, m_blk * BLOCK_M + 64, kb * BLOCK_K);
            }
            if (p.b_patch) {
              // MN-major B = patch matrix [K = patch rows, N = (ky, kx, c)]: boxes of [64 n, 64 patch rows]
              const int k0 = kb * BLOCK_K;
              const int bi = k0 / p.pb_rows_b, oy0 = (k0 - bi * p.pb_rows_b) / p.pb_wo;
              for (int j = 0; j < p.block_n / 64; ++j) {
                const int n = n_blk * p.block_n + j * 64;
                tma_load_5d(sb + j * 8192, &tmap_b, full, n % p.pb_span, n / p.pb_span, 0, oy0, bi);
              }
            } else if (!p.b_mn) {
              tma_load_2d(sb, &tmap_b, full, kb * BLOCK_K, n_blk * p.block_n);
            } else {
              for (int j = 0; j < p.block_n / 64; ++j)
                tma_load_2d(sb + j * 8192, &tmap_b, full, n_blk * p.block_n + j * 64, kb * BLOCK_K);
            }
          } else {
            // pair: both CTAs' bytes are counted by the LEADER's full barrier (the MMA issuer lives there); this CTA
            // brings its own 128 A rows and its half of the B tile
            const uint32_t full = mapa_shared(bar_full + 8 * stage, 0);
            if (crank == 0) mbar_arrive_expect_tx(bar_full + 8 * stage, 2 * p.stage_bytes);
            if (unit == unit0 && kb == kb0) GEMM_TRACE(2);
            if (!p.a_mn) {
              tma_load_2d_pair(sa, &tmap_a, full, kb * BLOCK_K, m_blk * BLOCK_M);
            } else {
              tma_load_2d_pair(sa, &tmap_a, full, m_blk * BLOCK_M, kb * BLOCK_K);
              tma_load_2d_pair(sa + 8192, &tmap_a, full, m_blk * BLOCK_M + 64, kb * BLOCK_K);
            }
            const int half = p.block_n / 2;
            const int nb0 = n_blk * p.block_n + (int)crank * half;
            if (!p.b_mn) {
              tma_load_2d_pair(sb, &tmap_b, full, kb * BLOCK_K, nb0);   // box = [64 k, block_n / 2 rows]
            } else {
              for (int j = 0; j < half / 64; ++j)
                tma_load_2d_pair(sb + j * 8192, &tmap_b, full, nb0 + j * 64, kb * BLOCK_K);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (pair mode: the leader CTA issues for both) =====================
    const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, p.block_n, p.a_mn, p.b_mn);
    const bool issuer = !PAIR || crank == 0;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase[2] = {0, 0};
    for (int unit = unit0; issuer && unit < total_units; unit += unit_stride) {
      const int split = unit % p.splits;
      const int kb0 = (int)(((long long)split * p.k_blocks) / p.splits);
      const int kb1 = (int)(((long long)(split + 1) * p.k_blocks) / p.splits);
      // wait for the epilogue to have drained this accumulator buffer
      mbar_wait(bar_tempty + 8 * acc, acc_phase[acc] ^ 1);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + acc * (p.tmem_cols >> 1);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(bar_full + 8 * stage, phase);
        tc_fence_after();
        if (lane == 0 && unit == unit0 && kb == kb0) GEMM_TRACE(3);
        if (lane == 0) {
          const uint32_t sa = smem_base + stage * p.stage_bytes;
          const uint32_t sb = sa + a_bytes;
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            const uint64_t adesc = p.a_mn ? umma_smem_desc(sa + kk * 2048, 8192, 1024)
                                          : umma_smem_desc(sa + kk * 32, 16, 1024);
            const uint64_t bdesc = p.b_mn ? umma_smem_desc(sb + kk * 2048, 8192, 1024)
                                          : umma_smem_desc(sb + kk * 32, 16, 1024);
            if (PAIR) umma_bf16_pair(tmem_d, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            else umma_bf16(tmem_d, adesc, bdesc, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          if (PAIR) {
            umma_commit_pair(bar_empty + 8 * stage, 3);                       // both producers of the pair
            if (kb == kb1 - 1) umma_commit_pair(bar_tfull + 8 * acc, 3);      // both epilogues
          } else {
            umma_commit(bar_empty + 8 * stage);                   // smem slot reusable once these MMAs retire
            if (kb == kb1 - 1) umma_commit(bar_tfull + 8 * acc);  // accumulator complete
          }
          if (kb == kb1 - 1 && unit == unit0) GEMM_TRACE(4);
        }
        __syncwarp();
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      acc_phase[acc] ^= 1;
      acc ^= 1;
    }
  } else {
    // ===================== epilogue warps =====================
    const int q = warp & 3;  // TMEM lane quadrant this warp may access
    int acc = 0;
    int trace_tile = 0;
    uint32_t acc_phase[2] = {0, 0};
    for (int unit = unit0; unit < total_units; unit += unit_stride) {
      const int tile = unit / p.splits;
      const int n_blk = tile % p.n_tiles;
      const int m_blk = PAIR ? 2 * (tile / p.n_tiles) + (int)crank : tile / p.n_tiles;
      const int n0 = n_blk * p.block_n;
      // ncu: the bias LDGs issued after the TMEM load were the largest stall of the epilogue.  The tile's bias slice goes to
      // shared memory while the accumulator is still being produced; published by the epilogue-only named barrier below
      // and released for the next tile by the one at the end of the tile.
      if (p.bias != nullptr && warp == 2) {
        float* sbias_gen = reinterpret_cast<float*>(smem_gen + bars_off + EPI_BIAS_OFF);
        for (int j = lane; j < p.block_n; j += 32) sbias_gen[j] = (n0 + j < p.N) ? __ldg(p.bias + n0 + j) : 0.f;
      }
      const uint32_t sbias = smem_base + bars_off + EPI_BIAS_OFF;
      mbar_wait(bar_tfull + 8 * acc, acc_phase[acc]);
      tc_fence_after();
      if (p.bias != nullptr) asm volatile("bar.sync 1, 128;" ::: "memory");
      if (threadIdx.x == 64 && unit == unit0) GEMM_TRACE(5);
      if (threadIdx.x == 64 && trace_tile < 28) GEMM_TRACE(8 + 2 * trace_tile);
      const long long row = (long long)m_blk * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M;
      float rscale = p.alpha;
      if (p.row_scale != nullptr && row_ok) rscale *= p.row_scale[row / p.rows_per_scale];
      const long long orow = gemm_map_row(row, p.out_map);
      for (int c = 0; c < p.block_n; c += 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * (p.tmem_cols >> 1) + c, r);
        tmem_ld_wait();
        const int col0 = n0 + c;
        if (!row_ok || col0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const bool full_chunk = (col0 + 32 <= p.N);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint4 b4 = ld_shared_v4(sbias + (uint32_t)(c + j) * 4);   // broadcast read (zeros past N)
            v[j] += __uint_as_float(b4.x); v[j + 1] += __uint_as_float(b4.y);
            v[j + 2] += __uint_as_float(b4.z); v[j + 3] += __uint_as_float(b4.w);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= rscale;
        if (p.residual != nullptr) {
          const float* rp = p.residual + gemm_map_row(row, p.res_map) * p.ldr + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              const float4 r4 = *reinterpret_cast<const float4*>(rp + j);
              v[j] += r4.x; v[j + 1] += r4.y; v[j + 2] += r4.z; v[j + 3] += r4.w;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) v[j] += rp[j];
          }
        }
        if (p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        if (p.atomic_add) {
          float* op = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4) red_add_v4_f32(op + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) red_add_f32(op + j, v[j]);
          }
        } else if (p.out_bf16) {
          __nv_bfloat16* op = reinterpret_cast<__nv_bfloat16*>(p.out) + orow * p.ldo + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
              uint4 o;
              o.x = pack_bf16x2(v[j], v[j + 1]);
              o.y = pack_bf16x2(v[j + 2], v[j + 3]);
              o.z = pack_bf16x2(v[j + 4], v[j + 5]);
              o.w = pack_bf16x2(v[j + 6], v[j + 7]);
              *reinterpret_cast<uint4*>(op + j) = o;
            }
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) op[j] = __float2bfloat16(v[j]);
          }
        } else {
          float* op = reinterpret_cast<float*>(p.out) + orow * p.ldo + col0;
          if (full_chunk) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(op + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32; ++j)
              if (col0 + j < p.N) op[j] = v[j];
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_shared(bar_tempty + 8 * acc, 0));   // the leader's MMA issuer waits for both CTAs
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      if (threadIdx.x == 64 && unit == unit0) GEMM_TRACE(6);
      if (threadIdx.x == 64 && trace_tile < 28) GEMM_TRACE(9 + 2 * trace_tile);
      ++trace_tile;
      if (p.bias != nullptr) asm volatile("bar.sync 1, 128;" ::: "memory");  // everyone is done with this tile's bias slice
      acc_phase[acc] ^= 1;
      acc ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) GEMM_TRACE(7);
  if (PAIR) cluster_sync_all();  // the peer may still read our smem / arrive on our barriers until it is done
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_pair(tmem_base, (uint32_t)p.tmem_cols);
    else tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
// Tile-width heuristic.  These GEMMs are mostly bound by the L2 -> shared-memory fill rate of one SM (small K, weights
// re-read by every row tile), so per k-block time ~ bytes loaded ~ (BLOCK_M + bn) and the whole problem costs
//   waves(bn) * (BLOCK_M + bn) * k_blocks + per-tile overhead.
// Split-K (wgrad) launches fill the machine through `splits`, so they simply take the widest tile that does not waste
// more than ~20 % of its columns on padding: wide tiles double the flop/byte of the operand stream.
static int pick_block_n(int N, int b_mn, int M, int atomic_add) {
  const int cands_k[] = {256, 224, 192, 160, 128, 96, 64, 32};
  const int cands_mn[] = {256, 192, 128, 64};
  const int* cands = b_mn ? cands_mn : cands_k;
  const int nc = b_mn ? 4 : 8;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int sms = num_sms();
  if (atomic_add) {
    for (int i = 0; i < nc; ++i) {
      const int bn = cands[i];
      const int n_tiles = (N + bn - 1) / bn;
      if ((double)n_tiles * bn <= 1.2 * N || i == nc - 1) return bn;
    }
  }
  int best = cands[nc - 1];
  double best_cost = 1e30;
  for (int i = 0; i < nc; ++i) {
    const int bn = cands[i];
    const int n_tiles = (N + bn - 1) / bn;
    const long long tiles = (long long)m_tiles * n_tiles;
    const long long waves = (tiles + sms - 1) / sms;
    const double cost = (double)waves * (BLOCK_M + bn + 24);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// CTA-pair mode: a pair owns 256 rows; per k-block each CTA streams (BLOCK_M + bn / 2) operand rows.  Returns 0 when no
// candidate covers N with <= 25 % padding (then the single-CTA kernel is used).
static int pick_block_n_pair(int N, int b_mn, int M, int atomic_add) {
  const int cands_k[] = {256, 224, 192, 160, 128, 96, 64};
  const int cands_mn[] = {256, 128};
  const int* cands = b_mn ? cands_mn : cands_k;
  const int nc = b_mn ? 2 : 7;
  const int m_pairs = (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M);
  const int slots = num_sms() / 2;
  int best = 0;
  double best_cost = 1e30;
  for (int i = 0; i < nc; ++i) {
    const int bn = cands[i];
    const int n_tiles = (N + bn - 1) / bn;
    if ((double)n_tiles * bn > 1.25 * N) continue;
    if (atomic_add) return bn;   // split-K fills the machine: widest acceptable tile
    const long long tiles = (long long)m_pairs * n_tiles;
    const long long waves = (tiles + slots - 1) / slots;
    const double cost = (double)waves * (BLOCK_M + bn / 2 + 24);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; }
  }
  return best;
}

// Host-side launch plan: tile width, residency mode (1 CTA/SM, 2 CTAs/SM, CTA pairs), split-K count, pipeline depth, grid.
// Pure arithmetic on the problem shape (no device access beyond the cached SM count), so tests can exercise it without a GPU
// through sfk_gemm_plan().
static int gemm_make_plan(const GemmDesc& g, GemmKernelParams& p, int& grid, size_t& smem_bytes) {
  SFK_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  SFK_REQUIRE(!(g.atomic_add && g.out_bf16), "gemm: atomic accumulation needs an fp32 output");
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_mn = g.a_mn ? 1 : 0;
  p.b_mn = g.b_mn ? 1 : 0;
  p.k_blocks = (g.K + BLOCK_K - 1) / BLOCK_K;
  // CTA pairs (cta_group::2) when enabled and the shape allows it
  int pair_bn = 0;
  const bool a_patch = g.a_patch.enabled != 0, b_patch = g.b_patch.enabled != 0;
  if (a_patch) {
    SFK_REQUIRE(!g.a_mn && patch_view_ok(g.a_patch, BLOCK_M), "gemm: A patch view not tileable (Wo=%d Ho=%d)", g.a_patch.Wo, g.a_patch.Ho);
    SFK_REQUIRE(g.K == g.a_patch.sr * g.a_patch.sr * g.a_patch.C && g.M == g.a_patch.B * g.a_patch.Ho * g.a_patch.Wo,
                "gemm: A patch view does not match M=%d K=%d", g.M, g.K);
  }
  if (b_patch) {
    SFK_REQUIRE(g.b_mn && patch_view_ok(g.b_patch, 64), "gemm: B patch view not tileable (Wo=%d Ho=%d)", g.b_patch.Wo, g.b_patch.Ho);
    SFK_REQUIRE(g.N == g.b_patch.sr * g.b_patch.sr * g.b_patch.C && g.K == g.b_patch.B * g.b_patch.Ho * g.b_patch.Wo,
                "gemm: B patch view does not match N=%d K=%d", g.N, g.K);
  }
  const bool pair_wanted = !a_patch && !b_patch &&
                           (g_gemm_cluster2 == 2 || (g_gemm_cluster2 == 1 && (double)g.M * g.N * g.K >= 1.0e11 && g.K >= 1024));
  if (pair_wanted && g.M > BLOCK_M && (num_sms() % 2 == 0)) {
    if (g.block_n > 0) {
      if (g.block_n % 32 == 0 && (!p.b_mn || g.block_n % 128 == 0)) pair_bn = g.block_n;
    } else {
      pair_bn = pick_block_n_pair(g.N, p.b_mn, g.M, g.atomic_add);
    }
  }
  p.cluster2 = pair_bn > 0 ? 1 : 0;
  // latency-bound regime (few k-blocks per tile): prefer two co-resident CTAs per SM over wide tiles
  const bool occ2 = !p.cluster2 && g_gemm_occ2 && g.block_n <= 0 && !g.atomic_add && p.k_blocks <= 8 && g.N >= 64;
  p.block_n = pair_bn > 0 ? pair_bn : (g.block_n > 0 ? g.block_n : pick_block_n(g.N, p.b_mn, g.M, g.atomic_add));
  if (occ2 && p.block_n > 128) {
    // widest tile <= 128 columns with the least padding
    const int c_k[] = {128, 96, 64}, c_mn[] = {128, 64};
    const int* c = p.b_mn ? c_mn : c_k;
    const int nc2 = p.b_mn ? 2 : 3;
    int best = c[0];
    long long best_cols = 1LL << 60;
    for (int i = 0; i < nc2; ++i) {
      const long long cols = (long long)((g.N + c[i] - 1) / c[i]) * c[i];
      if (cols < best_cols) { best_cols = cols; best = c[i]; }
    }
    p.block_n = best;
  }
  p.tmem_cols = (occ2 && p.block_n <= 128) ? 256 : TMEM_COLS;
  SFK_REQUIRE(p.block_n % 32 == 0 && p.block_n >= 32 && p.block_n <= MAX_BLOCK_N, "gemm: bad block_n %d", p.block_n);
  SFK_REQUIRE(!p.b_mn || p.block_n % 64 == 0, "gemm: MN-major B needs block_n %% 64 == 0");
  p.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
  p.n_tiles = (g.N + p.block_n - 1) / p.block_n;
  int splits = g.splits;
  if (splits <= 0) {
    splits = 1;
    if (g.atomic_add) {
      const long long tiles = (long long)(p.cluster2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles;
      const int sms = p.cluster2 ? num_sms() / 2 : num_sms();   // schedulable units: CTAs or CTA pairs
      // Split-K count by a round model instead of "fill 2 x SMs": cost = rounds x (k-blocks per split + fixed per-unit
      // cost), rounds = ceil(tiles x splits / SMs).  The old rule gave e.g. 20 tiles x 15 splits = 300 units = 3 rounds
      // on 148 SMs for the 1280 x 320 x 16384 wgrad; 7 splits = 140 units = 1 round does the same work in ~60 % of
      // the time with half the atomic traffic.  The fixed cost (prologue + fp32 red.add epilogue of a 128 x bn tile)
      // is worth ~8 k-blocks.
      const int kOver = 8;
      const int max_s = p.k_blocks < 2 * sms ? p.k_blocks : 2 * sms;
      long long best_cost = -1;
      for (int sp = 1; sp <= max_s; ++sp) {
        const long long per = (p.k_blocks + sp - 1) / sp;
        const long long rounds = (tiles * sp + sms - 1) / sms;
        const long long cost = rounds * (per + kOver);
        if (best_cost < 0 || cost < best_cost) { best_cost = cost; splits = sp; }
      }
    }
  }
  if (splits > p.k_blocks) splits = p.k_blocks;
  if (splits < 1) splits = 1;
  SFK_REQUIRE(splits == 1 || g.atomic_add, "gemm: split-K requires atomic accumulation");
  p.splits = splits;
  p.stage_bytes = BLOCK_M * BLOCK_K * 2 + (p.cluster2 ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;   // per CTA
  const int smem_budget = (p.tmem_cols == 256 ? 110 * 1024 : 227 * 1024) - 2048 - (int)EPI_BIAS_BYTES;
  int stages = smem_budget / (int)p.stage_bytes;
  if (stages > MAX_STAGES) stages = MAX_STAGES;
  if (stages > p.k_blocks + 1 && p.tmem_cols == 256) stages = p.k_blocks + 1;
  SFK_REQUIRE(stages >= 2, "gemm: not enough shared memory for 2 stages");
  p.stages = stages;
  p.alpha = g.alpha;
  p.bias = g.bias;
  p.residual = g.residual;
  p.ldr = g.ldr;
  p.res_map = g.res_map;
  p.out_map = g.out_map;
  p.row_scale = g.row_scale;
  p.rows_per_scale = g.rows_per_scale > 0 ? g.rows_per_scale : 1;
  p.out = g.out;
  p.ldo = g.ldo;
  p.out_bf16 = g.out_bf16;
  p.atomic_add = g.atomic_add;
  p.relu = g.relu;
  p.trace = g_gemm_trace;

  p.a_patch = a_patch ? 1 : 0;
  p.pa_rows_b = a_patch ? g.a_patch.Ho * g.a_patch.Wo : 1;
  p.pa_wo = a_patch ? g.a_patch.Wo : 1;
  p.pa_kspan = a_patch ? g.a_patch.sr * g.a_patch.C / BLOCK_K : 1;
  p.b_patch = b_patch ? 1 : 0;
  p.pb_rows_b = b_patch ? g.b_patch.Ho * g.b_patch.Wo : 1;
  p.pb_wo = b_patch ? g.b_patch.Wo : 1;
  p.pb_span = b_patch ? g.b_patch.sr * g.b_patch.C : 1;
  smem_bytes = (size_t)p.stages * p.stage_bytes + EPI_BIAS_OFF + EPI_BIAS_BYTES + 1024;
  const long long units = (long long)(p.cluster2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles * p.splits;
  if (p.cluster2) {
    const long long clusters = units < num_sms() / 2 ? units : num_sms() / 2;
    grid = (int)(2 * clusters);
  } else {
    const long long slots = (long long)num_sms() * (p.tmem_cols == 256 ? 2 : 1);
    grid = (int)(units < slots ? units : slots);
  }
  return 0;
}

int gemm_plan(const GemmDesc& g, int out[8]) {
  GemmKernelParams p;
  int grid = 0;
  size_t smem = 0;
  const int rc = gemm_make_plan(g, p, grid, smem);
  if (rc) return rc;
  out[0] = p.block_n; out[1] = p.splits; out[2] = p.stages; out[3] = grid;
  out[4] = p.cluster2; out[5] = p.tmem_cols; out[6] = (int)smem; out[7] = p.m_tiles * p.n_tiles;
  return 0;
}

int gemm_bf16(const GemmDesc& g, cudaStream_t stream) {
  SFK_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  SFK_REQUIRE(g.a != nullptr && g.b != nullptr && g.out != nullptr, "gemm: null pointer");
  SFK_REQUIRE((reinterpret_cast<uintptr_t>(g.a) & 15) == 0 && (reinterpret_cast<uintptr_t>(g.b) & 15) == 0,
              "gemm: operands must be 16-byte aligned");
  SFK_REQUIRE((g.lda % 8) == 0 && (g.ldb % 8) == 0, "gemm: lda/ldb must be multiples of 8 elements (lda=%lld ldb=%lld)",
              g.lda, g.ldb);
  const int esz = g.out_bf16 ? 2 : 4;
  SFK_REQUIRE((reinterpret_cast<uintptr_t>(g.out) & 15) == 0 && ((g.ldo * esz) % 16) == 0,
              "gemm: output must be 16-byte aligned with a 16-byte multiple row pitch");
  SFK_REQUIRE(!(g.atomic_add && g.out_bf16), "gemm: atomic accumulation needs an fp32 output");
  if (g.residual) SFK_REQUIRE((reinterpret_cast<uintptr_t>(g.residual) & 15) == 0 && (g.ldr % 4) == 0,
                              "gemm: residual must be 16-byte aligned");
  if (g.bias) SFK_REQUIRE((reinterpret_cast<uintptr_t>(g.bias) & 15) == 0, "gemm: bias must be 16-byte aligned");

  GemmKernelParams p;
  int grid = 0;
  size_t smem = 0;
  {
    const int prc = gemm_make_plan(g, p, grid, smem);
    if (prc) return prc;
  }
  const bool a_patch = p.a_patch != 0, b_patch = p.b_patch != 0;
  CUtensorMap ta, tb;
  int rc;
  auto patch_map = [&](CUtensorMap* t, const void* base, const PatchView& v, int rows) {
    const long long W = (long long)v.Wo * v.sr;
    const uint64_t dims[5] = {(uint64_t)v.sr * v.C, (uint64_t)v.sr, (uint64_t)v.Wo, (uint64_t)v.Ho, (uint64_t)v.B};
    const uint64_t strides[4] = {(uint64_t)W * v.C * 2, (uint64_t)v.sr * v.C * 2, (uint64_t)v.sr * W * v.C * 2,
                                 (uint64_t)v.batch_stride * 2};
    const uint32_t box[5] = {64, 1, (uint32_t)v.Wo, (uint32_t)(rows / v.Wo), 1};
    return make_tmap_5d_bf16(t, base, dims, strides, box);
  };
  if (a_patch) rc = patch_map(&ta, g.a, g.a_patch, BLOCK_M);
  else if (!p.a_mn) rc = make_tmap_2d_bf16(&ta, g.a, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda * 2, BLOCK_K, BLOCK_M);
  else         rc = make_tmap_2d_bf16(&ta, g.a, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda * 2, 64, BLOCK_K);
  if (rc) return rc;
  if (b_patch) rc = patch_map(&tb, g.b, g.b_patch, 64);
  else if (!p.b_mn) rc = make_tmap_2d_bf16(&tb, g.b, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldb * 2, BLOCK_K,
                                      p.cluster2 ? p.block_n / 2 : p.block_n);
  else         rc = make_tmap_2d_bf16(&tb, g.b, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldb * 2, 64, BLOCK_K);
  if (rc) return rc;

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    SFK_REQUIRE(e == cudaSuccess, "gemm: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  prof_gemm_shape(g.M, g.N, g.K, (g.a_mn ? 1 : 0) | (g.b_mn ? 2 : 0) | (g.atomic_add ? 4 : 0) | (p.splits << 8) | ((long long)p.block_n << 20));
  prof_gemm_begin(2.0 * (double)g.M * (double)g.N * (double)g.K, stream);
  if (p.cluster2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    (void)cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<1>, ta, tb, p);
  } else {
    launch_k(gemm_bf16_tcgen05_kernel<0>, dim3(grid), dim3(GEMM_THREADS), smem, stream, ta, tb, p);
  }
  prof_gemm_end(stream);
  return check_launch("gemm_bf16_tcgen05_kernel");
}

}  // namespace sfk
