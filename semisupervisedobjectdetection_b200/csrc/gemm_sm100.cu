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
#include <cstring>

#include "common.cuh"
#include "sfk_internal.h"

namespace sfk {

static constexpr int BLOCK_M = 128;
static constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
static constexpr int UMMA_K = 16;
static constexpr int MAX_BLOCK_N = 256;
static constexpr int TMEM_COLS = 512;  // 2 accumulator buffers x 256 fp32 columns
static constexpr int GEMM_EPI_WARPS = 8;   // two per TMEM lane quadrant: a lone warp per scheduler cannot hide its own latencies
static constexpr int GEMM_THREADS = 64 + 32 * GEMM_EPI_WARPS;
static constexpr int MAX_STAGES = 12;
// barrier block (behind the operand buffers): full[S], empty[S], tmem_full[2], tmem_empty[2], tmem slot, B-region free,
// (8 spare barriers); then the current tile's bias slice (MAX_BLOCK_N floats), staged by the epilogue
// before it waits for the accumulator
static constexpr uint32_t BAR_MISC_OFF = 16 * MAX_STAGES;
static constexpr uint32_t EPI_BIAS_OFF = BAR_MISC_OFF + 128;
static constexpr uint32_t EPI_BIAS_BYTES = 2 * 256 * 4;   // two slots: the next tile's slice is staged while this tile is read
static constexpr uint32_t EPI_WARP_BUF = 4096;   // one staged chunk of one epilogue warp: 32 rows x 128 B (or 32 x 64 B)

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
int g_gemm_occ2 = 0;

// TMA-store epilogue (sfk_set_option("gemm_epi_tma", 0|1), default 1): each of the 8 epilogue warps stages its 32 rows of a
// 32-column (f32) or 64-column (bf16) chunk in swizzled shared memory and one lane issues a TMA tile store (reduce-add for
// split-K); the two warps of a TMEM lane quadrant take alternate chunks.  The legacy epilogue (per-thread row-wise global
// stores) remains for remapped rows (prompt tokens).
// Measured motivation (profiles/gemm_r2.md): with 4 epilogue warps (one per scheduler, nothing to switch to) a 32-column
// chunk cost ~1100 cycles of dependent-issue latency and the epilogue, not the MMA, paced every K <= 1280 GEMM (tensor pipe
// 16-20 %); row-wise fp32 stores made the residual variants ~2x the bf16 ones (24.2 vs 13.5 us at 16384 x 320 x 320).
int g_gemm_epi_tma = 1;
// staging buffers per epilogue warp (sfk_set_option("gemm_epi_bufs", 1|2), default 2; falls back to 1 where the operand ring
// would otherwise drop below three stages)
int g_gemm_epi_bufs = 2;
// experiment: cap on the operand ring depth (sfk_set_option("gemm_max_stages", n); 0 = no cap)
int g_gemm_max_stages = 0;
// per-CTA rotated k-block order (sfk_set_option("gemm_krot", 0|1))
int g_gemm_krot = 0;   // measured on the MiT-b5 shapes: no effect (19.51 vs 19.52 ms per step), so off
// B-stationary main loop (sfk_set_option("gemm_bstat", 0|1), default 1) for K <= 512 without split-K: a CTA owns a contiguous,
// n-major range of output tiles, keeps the [block_n x K] weight tile resident in shared memory and streams only A.  These
// GEMMs (M >> N, K) were bound by the L2 -> shared-memory stream, 2/3 of which was the weight tile re-read for every row tile.
int g_gemm_bstat = 1;
// experiments (sfk_set_option("gemm_dbg"), wrong results, timing only): bit 0 = skip the TMA stores, bit 1 = skip the proxy
// fence, bit 2 = no operand loads, bit 3 = no MMAs, bit 4 = no tiles at all
int g_gemm_dbg = 0;

// optional in-kernel timeline (debug): CTA 0 stores globaltimer (ns) at phase boundaries into this device buffer
unsigned long long* g_gemm_trace = nullptr;

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  // SM cycle counter: %globaltimer reads cost ~0.7 us each and distorted the very timeline they were recording.  All
  // trace slots are written by CTA 0 (one SM), so cycle counts are mutually comparable.
  asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
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
  int units;                  // streaming mode: (row tiles | row-tile pairs) x column tiles x splits
  int step_s, step_n, step_m; // mixed-radix digits of the per-CTA unit stride (grid | grid / 2) over (splits, n_tiles)
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
  // shared-memory layout (offsets from the 1024-aligned base): [epilogue staging][A/B ring][B region][barriers + bias]
  uint32_t ring_off, bars_off;
  int bstat;               // B-stationary: ring holds A only (a_bytes per stage), B tile lives at breg_off (k_blocks x b_kb_bytes)
  uint32_t breg_off, b_kb_bytes;
  int epi_tma;             // TMA-store epilogue (kernel template EPI: 1 = f32 chunks of 32 columns, 2 = bf16 chunks of 64 [+ a 32 tail])
  uint32_t epi_off;        // staging area: 8 warps x epi_bufs x 4 KB
  int epi_bufs;            // staging buffers per epilogue warp (1 | 2)
  int krot;                // rotate each CTA's k-block order (streaming mode)
  int res_bf16;            // residual rows are bf16 (EPI == 2 only): one 128-byte fetch covers a whole 64-column chunk
  int out_patch, po_rows_b, po_wo, po_span;   // patch-view store (EPI == 2): patches per sample, ox-row length, sr*C
  uint32_t res_off;        // residual transpose buffers: 8 warps x 4 KB (behind the staging area; only with a residual)
  int do_scale;            // alpha != 1 or per-row scales present
  int epi_warps;           // 8 (320-thread CTA, one per SM) | 4 (192-thread CTA: two CTAs per SM, row-wise epilogue)
  int dbg;
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

// Work schedule.  Streaming mode: units (tile, k split) dealt round-robin, column tile fastest.  B-stationary mode: the
// n-major tile sequence t = n_blk * m_tiles + m_blk is cut into gridDim.x contiguous, equally long ranges, so a CTA changes
// its weight tile at most a few times.
// The walk is division-free per tile: the producer, the MMA issuer and the epilogue warps each run this iterator in a single
// thread's (or warp's) instruction stream, and with 5-k-block tiles the ~10 integer divisions (two of them 64-bit) of the
// first version cost 0.65 us between two tiles of the issuer and 0.8 us before the first TMA (in-kernel timeline, profiles/r2).
// The host splits the stride of the round-robin deal into its mixed-radix digits (step_s, step_n, step_m).
struct TileIter {
  int m_grp, n_blk, split;   // current unit: row tile (row-tile PAIR in pair mode), column tile, k split
  int left;                  // units left including the current one
};
__device__ __forceinline__ TileIter tile_iter_begin(const GemmKernelParams& p, int pair) {
  TileIter it;
  if (p.bstat) {
    const unsigned total = (unsigned)p.m_tiles * (unsigned)p.n_tiles;   // host guarantees total * gridDim.x < 2^32
    const unsigned begin = total * blockIdx.x / gridDim.x, end = total * (blockIdx.x + 1u) / gridDim.x;
    it.n_blk = (int)(begin / (unsigned)p.m_tiles);
    it.m_grp = (int)(begin - (unsigned)it.n_blk * (unsigned)p.m_tiles);
    it.split = 0;
    it.left = (int)(end - begin);
  } else {
    const int unit0 = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    const int step = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    it.split = unit0 % p.splits;
    const int tile = unit0 / p.splits;
    it.n_blk = tile % p.n_tiles;
    it.m_grp = tile / p.n_tiles;
    it.left = unit0 < p.units ? (p.units - 1 - unit0) / step + 1 : 0;
  }
  return it;
}
__device__ __forceinline__ void tile_iter_next(const GemmKernelParams& p, TileIter& it) {
  --it.left;
  if (p.bstat) {
    if (++it.m_grp == p.m_tiles) { it.m_grp = 0; ++it.n_blk; }
  } else {
    it.split += p.step_s;
    if (it.split >= p.splits) { it.split -= p.splits; ++it.n_blk; }
    it.n_blk += p.step_n;
    if (it.n_blk >= p.n_tiles) { it.n_blk -= p.n_tiles; ++it.m_grp; }
    it.m_grp += p.step_m;
  }
}
__device__ __forceinline__ void tile_k_range(const GemmKernelParams& p, const TileIter& it, int& kb0, int& kb1) {
  if (p.splits == 1) { kb0 = 0; kb1 = p.k_blocks; return; }
  kb0 = (int)((unsigned)it.split * (unsigned)p.k_blocks / (unsigned)p.splits);
  kb1 = (int)((unsigned)(it.split + 1) * (unsigned)p.k_blocks / (unsigned)p.splits);
}

// PAIR = 1: CTA-pair kernel (must be launched in 2-CTA clusters: it contains cta_group::2 instructions)
// EPI: 0 = per-thread row-wise epilogue (any output type, row maps), 1 = staged f32 chunks + TMA store / reduce-add,
//      2 = staged bf16 chunks + TMA store
template <int PAIR, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                         const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_tail,
                         const GemmKernelParams p) {
  pdl_launch_dependents();  // let the next kernel of the stream get scheduled; it waits for our completion itself
  if (threadIdx.x == 0) GEMM_TRACE(0);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // 1024-B alignment is required by SWIZZLE_128B atoms; dynamic smem base is only 16-B aligned by contract.
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));

  const uint32_t a_bytes = BLOCK_M * BLOCK_K * 2;
  const uint32_t ring = smem_base + p.ring_off;
  const uint32_t bars_off = p.bars_off;
  // barrier block: full[S], empty[S], tmem_full[2], tmem_empty[2], tmem_ptr, bfree
  const uint32_t bar_full = smem_base + bars_off;
  const uint32_t bar_empty = bar_full + 8 * MAX_STAGES;
  const uint32_t bar_tfull = bar_full + BAR_MISC_OFF;
  const uint32_t bar_tempty = bar_tfull + 16;
  const uint32_t tmem_slot = bar_tempty + 16;
  const uint32_t bar_bfree = tmem_slot + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + bars_off + BAR_MISC_OFF + 32);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0) {
    // one lane per barrier pair / descriptor instead of one thread walking all of them
    if (lane < p.stages) {
      mbar_init(bar_full + 8 * lane, 1);
      mbar_init(bar_empty + 8 * lane, 1);
    } else if (lane < p.stages + 2) {
      const int a = lane - p.stages;
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, (PAIR ? 2 : 1) * p.epi_warps);   // pair: both CTAs' epilogue warps report to the leader
    } else if (lane == p.stages + 2) {
      mbar_init(bar_bfree, 1);
    } else if (lane == 31) {
      tma_prefetch_desc(&tmap_a);
      tma_prefetch_desc(&tmap_b);
      if (EPI != 0) { tma_prefetch_desc(&tmap_out); tma_prefetch_desc(&tmap_tail); }
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

  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;
  TileIter it0 = tile_iter_begin(p, PAIR);
  if (p.dbg & 16) it0.left = 0;   // experiment: no work at all (launch + prologue + teardown floor)
  const uint32_t ring_stage_bytes = p.bstat ? a_bytes : p.stage_bytes;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      int cur_n = -1;             // B-stationary: column tile currently resident in the B region
      uint32_t bfree_phase = 0;
      int ptile = 0;
      for (TileIter it = it0; it.left > 0; tile_iter_next(p, it)) {
        const int m_blk = PAIR ? 2 * it.m_grp + (int)crank : it.m_grp, n_blk = it.n_blk;
        if (ptile < 24) GEMM_TRACE(64 + ptile);   // producer starts this tile (debug timeline)
        ++ptile;
        int kb0, kb1;
        tile_k_range(p, it, kb0, kb1);
        bool load_b = true;
        if (p.bstat) {
          load_b = n_blk != cur_n;
          if (load_b && cur_n >= 0) {   // every MMA that reads the old weight tile must have retired
            mbar_wait(bar_bfree, bfree_phase);
            bfree_phase ^= 1;
          }
          cur_n = n_blk;
        }
        // k-blocks are walked from a per-CTA rotated start (the sum does not care about the order): CTAs that share a
        // weight tile and start together would otherwise all request the same few L2 lines at the same moment
        const int krot = p.krot ? (int)((PAIR ? blockIdx.x >> 1 : blockIdx.x) % (unsigned)(kb1 - kb0)) : 0;
        for (int ki = kb0; ki < kb1; ++ki) {
          int kb = ki + krot;
          if (kb >= kb1) kb -= kb1 - kb0;
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t sa = ring + stage * ring_stage_bytes;
          const uint32_t sb = p.bstat ? smem_base + p.breg_off + (uint32_t)kb * p.b_kb_bytes : sa + a_bytes;
          if (p.dbg & 4) {   // experiment: no operand loads (timing of the MMA / epilogue side alone; wrong results)
            if (!PAIR || crank == 0) mbar_arrive(bar_full + 8 * stage);
          } else if (!PAIR) {
            const uint32_t full = bar_full + 8 * stage;
            mbar_arrive_expect_tx(full, p.bstat ? a_bytes + (load_b ? p.b_kb_bytes : 0u) : p.stage_bytes);
            if (ptile == 1 && ki == kb0) GEMM_TRACE(2);
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
            if (!load_b) {
              // weight tile already resident
            } else if (p.b_patch) {
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
            if (ptile == 1 && ki == kb0) GEMM_TRACE(2);
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
    // ONE thread runs the whole loop.  Its instruction stream is the pacing resource of short-K GEMMs: the first version
    // rebuilt both shared-memory descriptors (runtime layout selects, 64-bit packing, R2UR broadcasts under ELECT) for every
    // tcgen05.mma, ~160 dependent instructions = ~880 cycles per k-block against 320 cycles of tensor work (profiles/
    // gemm_r2.md).  Now the descriptors of stage 0 / k-step 0 are built once and every MMA's pair is two 64-bit adds.
    const bool issuer = !PAIR || crank == 0;
    if (lane == 0 && issuer) {
      const uint32_t idesc = umma_idesc_bf16(PAIR ? 2 * BLOCK_M : BLOCK_M, p.block_n, p.a_mn, p.b_mn);
      const uint32_t b_base = p.bstat ? smem_base + p.breg_off : ring + a_bytes;
      const uint64_t adesc0 = p.a_mn ? umma_smem_desc(ring, 8192, 1024) : umma_smem_desc(ring, 16, 1024);
      const uint64_t bdesc0 = p.b_mn ? umma_smem_desc(b_base, 8192, 1024) : umma_smem_desc(b_base, 16, 1024);
      // descriptor address fields count 16-byte units: next k-step = +32 B (K-major) or +2048 B (MN-major)
      const uint64_t a_kstep = p.a_mn ? 128u : 2u, b_kstep = p.b_mn ? 128u : 2u;
      const uint64_t a_sstep = ring_stage_bytes >> 4;
      const uint64_t b_step = p.bstat ? (p.b_kb_bytes >> 4) : (ring_stage_bytes >> 4);   // per k-block (B region) | per stage (ring)
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      int mtile = 0;
      uint32_t acc_phase = 0;   // bit a = phase of accumulator buffer a (a register, not an indexed local array)
      for (TileIter it = it0; it.left > 0; tile_iter_next(p, it)) {
        int kb0, kb1;
        tile_k_range(p, it, kb0, kb1);
        // wait for the epilogue to have drained this accumulator buffer
        mbar_wait(bar_tempty + 8 * acc, ((acc_phase >> acc) & 1u) ^ 1u);
        tc_fence_after();
        if (mtile < 24) GEMM_TRACE(96 + mtile);    // accumulator buffer free (debug timeline)
        const uint32_t tmem_d = tmem_base + acc * (p.tmem_cols >> 1);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint64_t ad = adesc0 + (uint64_t)stage * a_sstep;
          const uint64_t bd = bdesc0 + (uint64_t)(p.bstat ? kb : stage) * b_step;
          if (p.dbg & 8) {   // experiment: no MMAs (timing of the operand stream alone; wrong results)
            if (PAIR) { umma_commit_pair(bar_empty + 8 * stage, 3); if (kb == kb1 - 1) umma_commit_pair(bar_tfull + 8 * acc, 3); }
            else { umma_commit(bar_empty + 8 * stage); if (kb == kb1 - 1) umma_commit(bar_tfull + 8 * acc); }
          } else if (PAIR) {
            umma_bf16_pair(tmem_d, ad, bd, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
            for (int kk = 1; kk < BLOCK_K / UMMA_K; ++kk) umma_bf16_pair(tmem_d, ad + kk * a_kstep, bd + kk * b_kstep, idesc, 1u);
            umma_commit_pair(bar_empty + 8 * stage, 3);                       // both producers of the pair
            if (kb == kb1 - 1) umma_commit_pair(bar_tfull + 8 * acc, 3);      // both epilogues
          } else {
            umma_bf16(tmem_d, ad, bd, idesc, kb > kb0 ? 1u : 0u);
#pragma unroll
            for (int kk = 1; kk < BLOCK_K / UMMA_K; ++kk) umma_bf16(tmem_d, ad + kk * a_kstep, bd + kk * b_kstep, idesc, 1u);
            umma_commit(bar_empty + 8 * stage);                   // smem slot reusable once these MMAs retire
            if (kb == kb1 - 1) umma_commit(bar_tfull + 8 * acc);  // accumulator complete
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (mtile == 0) { GEMM_TRACE(3); GEMM_TRACE(4); }
        if (mtile < 24) GEMM_TRACE(160 + mtile);   // last k-block of the tile issued
        ++mtile;
        if (p.bstat && it.left > 1 && it.m_grp + 1 == p.m_tiles)
          umma_commit(bar_bfree);   // last tile of this column tile: the producer may overwrite the B region once these retire
        acc_phase ^= 1u << acc;
        acc ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue warps (2..9) =====================
    const int q = warp & 3;              // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;    // which of the quadrant's two warps: takes every other chunk
    const int wpq = p.epi_warps >> 2;    // epilogue warps per quadrant (1 | 2)
    const int epi_threads = 32 * p.epi_warps;
    int acc = 0;
    int trace_tile = 0;
    uint32_t acc_phase = 0;   // bit a = phase of accumulator buffer a
    // staging buffers of this warp (1 or 2 x 4 KB): with two, the TMA store of a chunk reads its buffer while the next chunk
    // is converted into the other one (ncu, profiles/r2: the wait for the previous store's shared-memory read was the
    // largest single stall of the epilogue warps)
    const uint32_t stage_buf0 = smem_base + p.epi_off + (uint32_t)(warp - 2) * (uint32_t)p.epi_bufs * EPI_WARP_BUF;
    const uint32_t res_buf = smem_base + p.res_off + (uint32_t)(warp - 2) * EPI_WARP_BUF;   // residual transpose buffer
    uint32_t stage_sel = 0;
    const bool has_bias = p.bias != nullptr;
    const bool has_res = p.residual != nullptr;
    const bool do_scale = p.do_scale != 0;
    // bias slices live in two shared-memory slots: while a tile reads one, warp 2 stages the NEXT tile's slice (only when
    // its column tile differs) into the other; the named barrier at the end of every tile publishes it.  (ncu: bias LDGs
    // issued after the TMEM load, then a per-tile staging step between two barriers, were the largest epilogue stalls.)
    uint32_t bias_slot = 0;
    float* sbias_gen = reinterpret_cast<float*>(smem_gen + bars_off + EPI_BIAS_OFF);
    if (has_bias && it0.left > 0) {
      const int n_blk = it0.n_blk;
      if (warp == 2)
        for (int j = lane; j < p.block_n; j += 32) sbias_gen[j] = (n_blk * p.block_n + j < p.N) ? __ldg(p.bias + n_blk * p.block_n + j) : 0.f;
      asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");
    }
    for (TileIter it = it0; it.left > 0; tile_iter_next(p, it)) {
      const int m_blk = PAIR ? 2 * it.m_grp + (int)crank : it.m_grp, n_blk = it.n_blk;
      const int n0 = n_blk * p.block_n;
      const int row0 = m_blk * BLOCK_M + q * 32;
      const uint32_t sbias = smem_base + bars_off + EPI_BIAS_OFF + bias_slot * 1024u;
      // next tile's bias slice: loads issued now (registers), stored to the other slot after this tile's chunks
      int n_next = n_blk;
      if (it.left > 1) {
        TileIter nx = it;
        tile_iter_next(p, nx);
        n_next = nx.n_blk;
      }
      const bool bias_reload = has_bias && n_next != n_blk;
      float bias_next[MAX_BLOCK_N / 32];
      if (bias_reload && warp == 2) {
#pragma unroll
        for (int j = 0; j < MAX_BLOCK_N / 32; ++j) {
          const int cc = j * 32 + lane;
          bias_next[j] = (cc < p.block_n && n_next * p.block_n + cc < p.N) ? __ldg(p.bias + n_next * p.block_n + cc) : 0.f;
        }
      }
      // residual tile: read with COALESCED 16-byte loads (8 lanes per row, 4 rows per instruction), transposed to the
      // thread-per-row TMEM layout through a swizzled 4 KB buffer per warp, and fetched one 32-column half ahead of its use.
      // (The first version let each thread read its own row: 32 cache lines per load instruction, 2560 L1 tag cycles per
      // 128 x 160 tile - the epilogue of the residual GEMMs alone took 6.6 us more than the bf16 ones, profiles/r2.)
      constexpr int CWR = EPI == 2 ? 64 : 32;
      // first row of this warp's 32 rows in the residual / output tensors (row maps with period % 32 == 0 keep them consecutive)
      const long long res_row0 = gemm_map_row((long long)row0, p.res_map);
      const int out_row0 = (int)gemm_map_row((long long)row0, p.out_map);
      float4 rr[8];
      const int res_r = lane >> 3, res_p = lane & 7;   // row within a 4-row group / 16-byte piece of the 128-byte row segment
      const bool res16 = EPI == 2 && p.res_bf16 != 0;   // bf16 residual: a 128-byte row segment is a whole 64-column chunk
      auto res_fetch = [&](int c0f, int hf) {
        const int wf = (EPI == 2 && c0f + CWR > p.block_n) ? 32 : CWR;
        const int cc = res16 ? 8 * res_p : 32 * hf + 4 * res_p;
        const bool cok = cc < wf && n0 + c0f + cc < p.N;
        const long long off = (res_row0 + res_r) * p.ldr + n0 + c0f + cc;
        const char* base = reinterpret_cast<const char*>(p.residual) + off * (res16 ? 2 : 4);
        const long long rstep = 4 * p.ldr * (res16 ? 2 : 4);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          rr[j] = (cok && row0 + res_r + 4 * j < p.M) ? __ldg(reinterpret_cast<const float4*>(base + j * rstep))
                                                      : make_float4(0.f, 0.f, 0.f, 0.f);
      };
      const int c0_first = (wpq == 2 ? (half ^ (trace_tile & 1)) : 0) * CWR;
      if (EPI != 0 && has_res && c0_first < p.block_n) res_fetch(c0_first, 0);
      mbar_wait(bar_tfull + 8 * acc, (acc_phase >> acc) & 1u);
      tc_fence_after();
      if (threadIdx.x == 64 && trace_tile == 0) GEMM_TRACE(5);
      if (threadIdx.x == 64 && trace_tile < 28) GEMM_TRACE(8 + 2 * trace_tile);
      const long long row = (long long)row0 + lane;
      const bool row_ok = row < p.M;
      float rscale = p.alpha;
      if (p.row_scale != nullptr && row_ok) rscale *= p.row_scale[(unsigned)row / (unsigned)p.rows_per_scale];   // M < 2^31: 32-bit division
      const p2 rs2 = p2_splat(rscale);
      const uint32_t tmem_acc = tmem_base + ((uint32_t)(q * 32) << 16) + acc * (p.tmem_cols >> 1);
      if (EPI != 0) {
        // ---- staged epilogue: TMEM -> registers -> swizzled shared memory -> TMA store / reduce-add ----
        constexpr int CW = EPI == 2 ? 64 : 32;               // columns per chunk (a bf16 tile may end in a 32-column chunk)
        for (int c0 = (wpq == 2 ? (half ^ (trace_tile & 1)) : 0) * CW; c0 < p.block_n; c0 += wpq * CW) {
          const int w = (EPI == 2 && c0 + CW > p.block_n) ? 32 : CW;   // this chunk's width
          uint32_t r[CW];
          __syncwarp();
          tmem_ld_32x32(tmem_acc + c0, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
          if (CW == 64 && w == 64) tmem_ld_32x32(tmem_acc + c0 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[CW - 32]));
          // the TMA store that last read the staging buffer about to be overwritten must be done with it
          if (lane == 0) {
            if (p.epi_bufs == 2) tma_store_wait_read<1>();
            else tma_store_wait_read<0>();
          }
          __syncwarp();
          tmem_ld_wait();
          const uint32_t stage_buf = stage_buf0 + stage_sel * EPI_WARP_BUF;
          stage_sel ^= (uint32_t)(p.epi_bufs - 1);
          const uint32_t row_bytes = (EPI == 2 && w == 32) ? 64u : 128u;   // bf16 tail chunk: 64-byte rows (SWIZZLE_64B)
          const uint32_t xr = row_bytes == 128u ? (uint32_t)(lane & 7) : (uint32_t)((lane >> 1) & 3);
          const uint32_t dst = stage_buf + (uint32_t)lane * row_bytes;
          // 32 columns at a time: few enough live registers that the 8 bias LDS.128 (and the 8 residual LDG.128) of a step are
          // issued back to back instead of one round trip per FADD2 (measured: 750 cycles of bias adds per 64-column chunk)
#pragma unroll
          for (int h = 0; h < CW / 32; ++h) {
            if (32 * h < w) {
              p2 v[16];
#pragma unroll
              for (int j = 0; j < 16; ++j) v[j] = p2_make_u(r[32 * h + 2 * j], r[32 * h + 2 * j + 1]);
              if (has_bias) {
                uint4 b4[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b4[j] = ld_shared_v4(sbias + (uint32_t)(c0 + 32 * h + 4 * j) * 4);   // broadcast reads
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  v[2 * j] = p2_add(v[2 * j], p2_make_u(b4[j].x, b4[j].y));
                  v[2 * j + 1] = p2_add(v[2 * j + 1], p2_make_u(b4[j].z, b4[j].w));
                }
              }
              if (do_scale) {
#pragma unroll
                for (int j = 0; j < 16; ++j) v[j] = p2_mul(v[j], rs2);
              }
              if (has_res) {
                if (!res16 || h == 0) {
                  __syncwarp();   // the previous reads of the transpose buffer are done
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    const uint32_t rw = (uint32_t)(4 * j + res_r);
                    st_shared_v4(res_buf + rw * 128u + ((((uint32_t)res_p) ^ (rw & 7u)) << 4), __float_as_uint(rr[j].x),
                                 __float_as_uint(rr[j].y), __float_as_uint(rr[j].z), __float_as_uint(rr[j].w));
                  }
                  __syncwarp();
                  // fetch the next 128-byte row segments of this tile (this chunk's second half, or the warp's next chunk)
                  if (!res16 && h + 1 < CW / 32 && 32 * (h + 1) < w) res_fetch(c0, h + 1);
                  else if (c0 + wpq * CW < p.block_n) res_fetch(c0 + wpq * CW, 0);
                }
                if (res16) {
                  uint4 r4[4];
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    r4[j] = ld_shared_v4(res_buf + (uint32_t)lane * 128u + ((((uint32_t)(4 * h + j)) ^ ((uint32_t)lane & 7u)) << 4));
#pragma unroll
                  for (int j = 0; j < 4; ++j) {
                    v[4 * j] = p2_add(v[4 * j], p2_from_bf16x2(r4[j].x));
                    v[4 * j + 1] = p2_add(v[4 * j + 1], p2_from_bf16x2(r4[j].y));
                    v[4 * j + 2] = p2_add(v[4 * j + 2], p2_from_bf16x2(r4[j].z));
                    v[4 * j + 3] = p2_add(v[4 * j + 3], p2_from_bf16x2(r4[j].w));
                  }
                } else {
                  uint4 r4[8];
#pragma unroll
                  for (int j = 0; j < 8; ++j) r4[j] = ld_shared_v4(res_buf + (uint32_t)lane * 128u + ((((uint32_t)j) ^ ((uint32_t)lane & 7u)) << 4));
#pragma unroll
                  for (int j = 0; j < 8; ++j) {
                    v[2 * j] = p2_add(v[2 * j], p2_make_u(r4[j].x, r4[j].y));
                    v[2 * j + 1] = p2_add(v[2 * j + 1], p2_make_u(r4[j].z, r4[j].w));
                  }
                }
              }
              if (p.relu) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                  float lo, hi;
                  p2_split(v[j], lo, hi);
                  v[j] = p2_make(fmaxf(lo, 0.f), fmaxf(hi, 0.f));
                }
              }
              if (EPI == 2) {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                  st_shared_v4(dst + (((uint32_t)(4 * h + j) ^ xr) << 4), p2_to_bf16x2(v[4 * j]), p2_to_bf16x2(v[4 * j + 1]),
                               p2_to_bf16x2(v[4 * j + 2]), p2_to_bf16x2(v[4 * j + 3]));
              } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  uint32_t a0, a1, a2, a3;
                  p2_split_u(v[2 * j], a0, a1);
                  p2_split_u(v[2 * j + 1], a2, a3);
                  st_shared_v4(dst + (((uint32_t)j ^ xr) << 4), a0, a1, a2, a3);
                }
              }
            }
          }
          if (!(p.dbg & 2)) fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            // boxes reaching past M or N are clipped by the TMA unit
            if (p.dbg & 1) {}   // experiment: no stores
            else if (row0 >= p.M) {}   // mapped tensors are taller than M: rows past the problem must not be written
            else if (EPI == 1 && p.atomic_add) tma_reduce_add_2d(&tmap_out, stage_buf, n0 + c0, out_row0);
            else if (EPI == 2 && p.out_patch) {
              // scatter through the patch view: these 32 patch rows are (32 / Wo) ox-rows of one sample, the 64 columns one
              // (kx, c) run of a single ky
              const int n = n0 + c0;
              const int bi = row0 / p.po_rows_b, oy0 = (row0 - bi * p.po_rows_b) / p.po_wo;
              tma_store_5d(&tmap_out, stage_buf, n % p.po_span, n / p.po_span, 0, oy0, bi);
            }
            else if (EPI == 2 && w == 32) tma_store_2d(&tmap_tail, stage_buf, n0 + c0, out_row0);
            else tma_store_2d(&tmap_out, stage_buf, n0 + c0, out_row0);
            tma_store_commit();
          }
        }
      } else {
      const long long orow = gemm_map_row(row, p.out_map);
      for (int c = (wpq == 2 ? (half ^ (trace_tile & 1)) : 0) * 32; c < p.block_n; c += wpq * 32) {
        uint32_t r[32];
        __syncwarp();
        tmem_ld_32x32(tmem_acc + c, r);
        tmem_ld_wait();
        const int col0 = n0 + c;
        if (!row_ok || col0 >= p.N) continue;
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        const bool full_chunk = (col0 + 32 <= p.N);
        if (has_bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const uint4 b4 = ld_shared_v4(sbias + (uint32_t)(c + j) * 4);   // broadcast read (zeros past N)
            v[j] += __uint_as_float(b4.x); v[j + 1] += __uint_as_float(b4.y);
            v[j + 2] += __uint_as_float(b4.z); v[j + 3] += __uint_as_float(b4.w);
          }
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] *= rscale;
        if (has_res) {
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
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) mbar_arrive_cluster(mapa_shared(bar_tempty + 8 * acc, 0));   // the leader's MMA issuer waits for both CTAs
        else mbar_arrive(bar_tempty + 8 * acc);
      }
      if (threadIdx.x == 64 && trace_tile == 0) GEMM_TRACE(6);
      if (threadIdx.x == 64 && trace_tile < 28) GEMM_TRACE(9 + 2 * trace_tile);
      ++trace_tile;
      if (bias_reload) {
        if (warp == 2) {
#pragma unroll
          for (int j = 0; j < MAX_BLOCK_N / 32; ++j)
            if (j * 32 + lane < p.block_n) sbias_gen[(bias_slot ^ 1u) * 256u + j * 32 + lane] = bias_next[j];
        }
        bias_slot ^= 1u;
        asm volatile("bar.sync 1, %0;" ::"r"(epi_threads) : "memory");  // next slice published; (two tiles later) this slot is free again
      }
      acc_phase ^= 1u << acc;
      acc ^= 1;
    }
    // staged chunks must have left shared memory before the CTA retires
    if (EPI != 0 && lane == 0) tma_store_wait_read<0>();
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

// B-stationary tile width: the [bn x K] weight tile must leave room for >= 3 A stages; cost = tiles per CTA x (bn + fixed
// per-tile overhead) with the contiguous even split (every CTA gets floor or ceil of tiles / CTAs).
static int pick_block_n_bstat(int N, int b_mn, int M, int k_blocks, int smem_fixed, int* stages_out) {
  const int cands_k[] = {256, 224, 192, 160, 128, 96, 64, 32};
  const int cands_mn[] = {256, 192, 128, 64};
  const int* cands = b_mn ? cands_mn : cands_k;
  const int nc = b_mn ? 4 : 8;
  const int m_tiles = (M + BLOCK_M - 1) / BLOCK_M;
  const int sms = num_sms();
  int best = 0, best_stages = 0;
  double best_cost = 1e30;
  for (int i = 0; i < nc; ++i) {
    const int bn = cands[i];
    const long long b_bytes = (long long)bn * k_blocks * 128;
    long long stages = (227 * 1024 - smem_fixed - b_bytes) / (BLOCK_M * BLOCK_K * 2);
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 3) continue;
    const int n_tiles = (N + bn - 1) / bn;
    const long long tiles = (long long)m_tiles * n_tiles;
    const long long per_cta = (tiles + sms - 1) / sms;
    const double cost = (double)per_cta * (bn + 40);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = bn; best_stages = (int)stages; }
  }
  *stages_out = best_stages;
  return best;
}

// Host-side launch plan: tile width, residency mode (1 CTA/SM, 2 CTAs/SM, CTA pairs, B-stationary), epilogue mode, split-K
// count, pipeline depth, shared-memory layout, grid.  Pure arithmetic on the problem shape (no device access beyond the
// cached SM count), so tests can exercise it without a GPU through sfk_gemm_plan().
static int gemm_make_plan(const GemmDesc& g, GemmKernelParams& p, int& grid, size_t& smem_bytes) {
  SFK_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "gemm: empty problem M=%d N=%d K=%d", g.M, g.N, g.K);
  SFK_REQUIRE(!(g.atomic_add && g.out_bf16), "gemm: atomic accumulation needs an fp32 output");
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_mn = g.a_mn ? 1 : 0;
  p.b_mn = g.b_mn ? 1 : 0;
  p.k_blocks = (g.K + BLOCK_K - 1) / BLOCK_K;
  p.m_tiles = (g.M + BLOCK_M - 1) / BLOCK_M;
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
  const bool out_patch = g.out_patch.enabled != 0;
  if (out_patch) {
    SFK_REQUIRE(g.out_bf16 && !g.atomic_add && patch_view_ok(g.out_patch, 32), "gemm: output patch view needs a bf16 output and whole ox-rows per 32 patches (Wo=%d Ho=%d)",
                g.out_patch.Wo, g.out_patch.Ho);
    SFK_REQUIRE(g.N == g.out_patch.sr * g.out_patch.sr * g.out_patch.C && g.M == g.out_patch.B * g.out_patch.Ho * g.out_patch.Wo,
                "gemm: output patch view does not match M=%d N=%d", g.M, g.N);
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

  // epilogue mode: TMA tile stores need identity row maps (prompt-token remaps keep the per-thread path)
  // Row maps (prompt / cls rows in front of every sample's spatial tokens) whose period is a multiple of 32 keep each epilogue
  // warp's 32 rows inside one sample: the mapped rows are still 32 consecutive rows, so tile stores and the coalesced residual
  // fetch only need the mapped index of their first row.  Other maps keep the per-thread path.
  auto map_ok = [](const RowMap& m) { return m.period == 0 ? true : (m.period % 32 == 0 && m.extra >= 0); };
  const bool maps_identity = g.out_map.period == 0 && g.out_map.offset == 0 && g.res_map.period == 0 && g.res_map.offset == 0;
  bool epi_tma = g_gemm_epi_tma && map_ok(g.out_map) && map_ok(g.res_map) && (maps_identity || g.M % 32 == 0) &&
                 !(g.out_patch.enabled && !maps_identity);
  int epi_bufs = epi_tma ? (g_gemm_epi_bufs >= 2 ? 2 : 1) : 1;
  const uint32_t res_bytes = (epi_tma && g.residual != nullptr) ? (uint32_t)GEMM_EPI_WARPS * EPI_WARP_BUF : 0u;
  uint32_t stage_bytes_epi = epi_tma ? (uint32_t)GEMM_EPI_WARPS * (uint32_t)epi_bufs * EPI_WARP_BUF + res_bytes : 0u;
  const uint32_t bars_bytes = EPI_BIAS_OFF + EPI_BIAS_BYTES;

  // B-stationary main loop: small K, no split-K, plain operands
  p.bstat = 0;
  int bstat_bn = 0, bstat_stages = 0;
  // (measured per MiT-b5 shape, profiles/r2: with K <= 128 the resident tile is small and the A ring deep - the loop wins by
  // 3-10 %; from K = 320 on the up-front tile load and the one-tile-deep ring lose 5-30 % against streaming wide tiles, so
  // automatic selection stops at two k-blocks; an explicit block_n still selects it up to K = 512)
  if (g_gemm_bstat && !p.cluster2 && !a_patch && !b_patch && (!g.atomic_add || g.splits == 1) && g.splits <= 1 &&
      p.k_blocks <= (g.block_n > 0 ? 8 : 2) && (long long)p.m_tiles * ((g.N + 31) / 32) * num_sms() < (1LL << 32)) {
    const int fixed = (int)(stage_bytes_epi + bars_bytes + 1024u);
    if (g.block_n > 0) {
      const long long b_bytes = (long long)g.block_n * p.k_blocks * 128;
      long long st = (227 * 1024 - fixed - b_bytes) / (BLOCK_M * BLOCK_K * 2);
      if (st > MAX_STAGES) st = MAX_STAGES;
      if (st >= 3 && g.block_n % 32 == 0 && (!p.b_mn || g.block_n % 64 == 0)) { bstat_bn = g.block_n; bstat_stages = (int)st; }
    } else {
      bstat_bn = pick_block_n_bstat(g.N, p.b_mn, g.M, p.k_blocks, fixed, &bstat_stages);
    }
    if (bstat_bn > 0) p.bstat = 1;
  }

  p.block_n = p.bstat ? bstat_bn : pair_bn > 0 ? pair_bn : (g.block_n > 0 ? g.block_n : pick_block_n(g.N, p.b_mn, g.M, g.atomic_add));
  // "light" CTAs (sfk_set_option("gemm_occ2", 1)): 192 threads, <= 110 KB of shared memory, 256 TMEM columns, row-wise epilogue,
  // so that two CTAs - of this launch or of a GEMM on another stream - share an SM
  const bool light = g_gemm_occ2 && !p.bstat && !p.cluster2 && g.block_n <= 0 && !g.atomic_add && p.k_blocks <= 8 && g.N >= 64;
  if (light && p.block_n > 128) {
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
  p.epi_warps = light ? 4 : GEMM_EPI_WARPS;
  p.tmem_cols = light ? 256 : TMEM_COLS;
  SFK_REQUIRE(p.block_n % 32 == 0 && p.block_n >= 32 && p.block_n <= MAX_BLOCK_N, "gemm: bad block_n %d", p.block_n);
  SFK_REQUIRE(!p.b_mn || p.block_n % 64 == 0, "gemm: MN-major B needs block_n %% 64 == 0");
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

  // shared-memory layout: [epilogue staging][A/B ring][B region][barriers + bias]
  if (light) { epi_tma = false; epi_bufs = 1; stage_bytes_epi = 0; }
  p.stage_bytes = BLOCK_M * BLOCK_K * 2 + (p.cluster2 ? p.block_n / 2 : p.block_n) * BLOCK_K * 2;   // per CTA
  if (epi_tma && epi_bufs == 2 && !p.bstat &&
      (227 * 1024 - 1024 - (int)bars_bytes - (int)stage_bytes_epi) / (int)p.stage_bytes < 3) {
    epi_bufs = 1;   // keep the operand ring at least three deep
    stage_bytes_epi = (uint32_t)GEMM_EPI_WARPS * EPI_WARP_BUF + res_bytes;
  }
  p.epi_tma = epi_tma ? 1 : 0;
  p.epi_bufs = epi_bufs;
  p.krot = (g_gemm_krot && !p.bstat) ? 1 : 0;
  p.res_bf16 = g.res_bf16 ? 1 : 0;
  SFK_REQUIRE(!g.res_bf16 || g.residual == nullptr || (epi_tma && g.out_bf16), "gemm: a bf16 residual needs the staged bf16-output epilogue");
  p.out_patch = out_patch ? 1 : 0;
  p.po_rows_b = out_patch ? g.out_patch.Ho * g.out_patch.Wo : 1;
  p.po_wo = out_patch ? g.out_patch.Wo : 1;
  p.po_span = out_patch ? g.out_patch.sr * g.out_patch.C : 1;
  SFK_REQUIRE(!out_patch || (epi_tma && p.block_n % 64 == 0), "gemm: output patch view needs the staged epilogue and 64-column chunks (block_n=%d)", p.block_n);
  p.epi_off = 0;
  p.res_off = (uint32_t)GEMM_EPI_WARPS * (uint32_t)epi_bufs * EPI_WARP_BUF;
  p.dbg = g_gemm_dbg;
  p.do_scale = (g.alpha != 1.0f || g.row_scale != nullptr) ? 1 : 0;
  p.ring_off = epi_tma ? stage_bytes_epi : 0u;
  p.b_kb_bytes = (uint32_t)p.block_n * BLOCK_K * 2;
  int stages;
  if (p.bstat) {
    stages = bstat_stages;
    p.breg_off = p.ring_off + (uint32_t)stages * BLOCK_M * BLOCK_K * 2;
    p.bars_off = p.breg_off + (uint32_t)p.k_blocks * p.b_kb_bytes;
  } else {
    const int budget = (light ? 110 : 227) * 1024 - 1024 - (int)bars_bytes - (int)p.ring_off;
    stages = budget / (int)p.stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (light && stages > p.k_blocks + 1) stages = p.k_blocks + 1;
    p.breg_off = 0;
    p.bars_off = p.ring_off + (uint32_t)stages * p.stage_bytes;
  }
  if (g_gemm_max_stages >= 2 && stages > g_gemm_max_stages) {
    stages = g_gemm_max_stages;
    if (p.bstat) { p.breg_off = p.ring_off + (uint32_t)stages * BLOCK_M * BLOCK_K * 2; p.bars_off = p.breg_off + (uint32_t)p.k_blocks * p.b_kb_bytes; }
    else p.bars_off = p.ring_off + (uint32_t)stages * p.stage_bytes;
  }
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
  smem_bytes = (size_t)p.bars_off + bars_bytes + 1024;
  SFK_REQUIRE(smem_bytes <= (size_t)227 * 1024, "gemm: shared-memory plan of %zu bytes exceeds 227 KB", smem_bytes);
  const long long units = (long long)(p.cluster2 ? (p.m_tiles + 1) / 2 : p.m_tiles) * p.n_tiles * p.splits;
  SFK_REQUIRE(units < (1LL << 30), "gemm: %lld work units exceed the scheduler's range", units);
  if (p.cluster2) {
    const long long clusters = units < num_sms() / 2 ? units : num_sms() / 2;
    grid = (int)(2 * clusters);
  } else {
    const long long slots = (long long)num_sms() * (light ? 2 : 1);
    grid = (int)(units < slots ? units : slots);
    if (g.max_ctas > 0 && grid > g.max_ctas) grid = g.max_ctas;
  }
  // per-CTA unit stride of the round-robin deal, in mixed-radix digits over (splits, n_tiles): TileIter adds them with carries
  p.units = (int)units;
  {
    const int step = p.cluster2 ? grid / 2 : grid;
    p.step_s = step % p.splits;
    const int t = step / p.splits;
    p.step_n = t % p.n_tiles;
    p.step_m = t / p.n_tiles;
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
  out[4] = p.cluster2 | (p.bstat << 1) | (p.epi_tma << 2); out[5] = p.tmem_cols; out[6] = (int)smem; out[7] = p.m_tiles * p.n_tiles;
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
  if (g.residual) SFK_REQUIRE((reinterpret_cast<uintptr_t>(g.residual) & 15) == 0 && (g.ldr % (g.res_bf16 ? 8 : 4)) == 0,
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

  CUtensorMap tout, ttail;
  memset(&tout, 0, sizeof(tout));
  memset(&ttail, 0, sizeof(ttail));
  const int epi = !p.epi_tma ? 0 : (g.out_bf16 ? 2 : 1);
  // rows of the output tensor as the tile stores see it: the mapped index of the last row + 1
  long long out_rows = g.M;
  if (g.out_map.period > 0 || g.out_map.offset != 0) {
    const long long last = (long long)g.M - 1;
    out_rows = last + (g.out_map.period > 0 ? (last / g.out_map.period) * g.out_map.extra : 0) + g.out_map.offset + 1;
  }
  if (epi == 1) {
    rc = make_tmap_2d_any(&tout, g.out, 4, (uint64_t)g.N, (uint64_t)out_rows, (uint64_t)g.ldo * 4, 32, 32, 128);
    if (rc) return rc;
    ttail = tout;
  } else if (epi == 2 && p.out_patch) {
    const PatchView& v = g.out_patch;
    const long long W = (long long)v.Wo * v.sr;
    const uint64_t dims[5] = {(uint64_t)v.sr * v.C, (uint64_t)v.sr, (uint64_t)v.Wo, (uint64_t)v.Ho, (uint64_t)v.B};
    const uint64_t strides[4] = {(uint64_t)W * v.C * 2, (uint64_t)v.sr * v.C * 2, (uint64_t)v.sr * W * v.C * 2,
                                 (uint64_t)v.batch_stride * 2};
    const uint32_t box[5] = {64, 1, (uint32_t)v.Wo, (uint32_t)(32 / v.Wo), 1};
    rc = make_tmap_5d_bf16(&tout, g.out, dims, strides, box);
    if (rc) return rc;
    ttail = tout;
  } else if (epi == 2) {
    rc = make_tmap_2d_any(&tout, g.out, 2, (uint64_t)g.N, (uint64_t)out_rows, (uint64_t)g.ldo * 2, 64, 32, 128);
    if (rc) return rc;
    rc = make_tmap_2d_any(&ttail, g.out, 2, (uint64_t)g.N, (uint64_t)out_rows, (uint64_t)g.ldo * 2, 32, 32, 64);
    if (rc) return rc;
  }

  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<0, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
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
    if (epi == 1) (void)cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<1, 1>, ta, tb, tout, ttail, p);
    else if (epi == 2) (void)cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<1, 2>, ta, tb, tout, ttail, p);
    else (void)cudaLaunchKernelEx(&cfg, gemm_bf16_tcgen05_kernel<1, 0>, ta, tb, tout, ttail, p);
  } else if (epi == 1) {
    launch_k(gemm_bf16_tcgen05_kernel<0, 1>, dim3(grid), dim3(GEMM_THREADS), smem, stream, ta, tb, tout, ttail, p);
  } else if (epi == 2) {
    launch_k(gemm_bf16_tcgen05_kernel<0, 2>, dim3(grid), dim3(GEMM_THREADS), smem, stream, ta, tb, tout, ttail, p);
  } else {
    launch_k(gemm_bf16_tcgen05_kernel<0, 0>, dim3(grid), dim3(64 + 32 * p.epi_warps), smem, stream, ta, tb, tout, ttail, p);
  }
  prof_gemm_end(stream);
  return check_launch("gemm_bf16_tcgen05_kernel");
}

}  // namespace sfk
