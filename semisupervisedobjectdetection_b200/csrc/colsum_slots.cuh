// Low-contention column reductions for kernels whose blocks each hold per-column partial sums (bias / LayerNorm-affine /
// depthwise-weight gradients).  Plain atomicAdd of every block into one destination serialises ~40 ns per same-address
// atomic at the L2 (measured: 1171 blocks -> 48 us).  Instead:
//   1. every block adds its partials into one of NS (<= 8) slot copies   g_slots[slot][v]   (contention / NS),
//   2. the last block to finish (ticket counter per blockIdx.x) sums the NS copies, adds them into the real destinations
//      with plain read-modify-write, and zeroes the slots + ticket again (self-cleaning static device memory).
// Invariant the host code keeps: kernels that use the slot area (LayerNorm / depthwise / cast-and-sum BACKWARD kernels and
// sfk_colsum) are only ever enqueued on the stream that runs backward; the side streams of the step carry other kernels
// (teacher forward: no column sums; weight-gradient stream: GEMMs, patch gathers, layout fixes).  Every kernel starts with
// griddepcontrol.wait, so on that one stream the static slot area is never shared by two running kernels.  Two models
// running backward concurrently on different streams of ONE process are not supported (one process per GPU).
// The one sanctioned exception is the layer's auxiliary stream (layer.cu: the K/V chain of backward next to the dq cast): the
// area has two BANKS, kernels launched on an auxiliary stream use bank 1.  The bank travels in bits 16+ of the `ns` argument
// every slot-using kernel already has (colsum_num_slots(V, sharers, bank) packs it, colsum_commit unpacks it).
#pragma once
#include "common.cuh"

namespace sfk {

static constexpr int COLSUM_SLOT_FLOATS = 1 << 18;   // 1 MiB of static device memory
static constexpr int COLSUM_MAX_GX = 64;

static constexpr int COLSUM_BANK_FLOATS = COLSUM_SLOT_FLOATS / 2;

static __device__ float g_colsum_slots[COLSUM_SLOT_FLOATS];
static __device__ unsigned int g_colsum_ticket[2 * COLSUM_MAX_GX];

// host helper: number of slot copies for V virtual columns and `sharers` blocks per column
inline int colsum_num_slots(long long V, long long sharers, int bank = 0) {
  long long ns = COLSUM_BANK_FLOATS / (V > 0 ? V : 1);
  // 8 copies: ~300-600 blocks per launch leave 40-75 same-address atomics per slot word (spread over the kernel's duration),
  // while the last block's fold - L2 round trips at the very end of the kernel, nothing left to overlap them - shrinks 8x
  // against the first version's 64 (measured: stage-3 layer backward 552 -> 500 us, eight slot-using kernels per layer)
  if (ns > 8) ns = 8;
  if (ns > sharers) ns = sharers;
  if (ns < 1) ns = 1;
  return (int)ns | ((bank & 1) << 16);
}

// Device: called by ALL threads of the block (contains __syncthreads).
//   sh      : block's partial sums for virtual columns [v0, v0 + nv)   (shared memory, already complete + synced)
//   V       : total number of virtual columns of this launch; ns slot copies laid out [ns][V]
//   share_id: index of this block among the blocks sharing these columns (e.g. blockIdx.y); sharers = their count
//   ticket  : which ticket counter (e.g. blockIdx.x)
//   write   : functor (int v, float sum) called once per virtual column of [v0, v0+nv) by the last block
template <typename Write>
__device__ __forceinline__ void colsum_commit(const float* sh, int v0, int nv, int V, int ns, unsigned share_id, unsigned sharers,
                                              int ticket, Write write) {
  __shared__ unsigned s_last;
  const int bank = ns >> 16;
  ns &= 0xffff;
  float* bank_slots = g_colsum_slots + (size_t)bank * COLSUM_BANK_FLOATS;
  ticket += bank * COLSUM_MAX_GX;
  float* slot = bank_slots + (size_t)(share_id % ns) * V + v0;
  for (int i = threadIdx.x; i < nv; i += blockDim.x) atomicAdd(slot + i, sh[i]);
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(&g_colsum_ticket[ticket], 1u) == sharers - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int i = threadIdx.x; i < nv; i += blockDim.x) {
    float* p = bank_slots + v0 + i;
    float part[4] = {0.f, 0.f, 0.f, 0.f};   // four independent chains: the loads of a column are all in flight together
    int s = 0;
    for (; s + 4 <= ns; s += 4) {
#pragma unroll
      for (int u = 0; u < 4; ++u) part[u] += __ldcg(p + (size_t)(s + u) * V);
    }
    for (; s < ns; ++s) part[0] += __ldcg(p + (size_t)s * V);
    for (s = 0; s < ns; ++s) __stcg(p + (size_t)s * V, 0.f);
    write(v0 + i, (part[0] + part[1]) + (part[2] + part[3]));
  }
  if (threadIdx.x == 0) g_colsum_ticket[ticket] = 0u;
}

// block-level reduction of per-thread channel partials (8 warps -> shared memory), then the slotted low-contention commit:
// virtual column v = k * C + channel  (k-th output array)
template <int VEC, int NV>
__device__ __forceinline__ void dw_block_commit(float (&acc)[NV][VEC], float* sh, float* out_base, long long out_stride,
                                                float* out_last, int c0, int C, bool lane_ok, int ns, int out_cstride = 1) {
  const int lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < NV * 32 * VEC; i += 256) sh[i] = 0.f;
  __syncthreads();
  if (lane_ok) {
#pragma unroll
    for (int k = 0; k < NV; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) atomicAdd(&sh[(k * 32 + lane) * VEC + v], acc[k][v]);
  }
  __syncthreads();
  const int cbase = c0 - lane * VEC;  // first channel of this block (same for every lane)
  int ncol = C - cbase;
  if (ncol > 32 * VEC) ncol = 32 * VEC;
  // slots are laid out [ns][NV][ncol-wide window per blockIdx.x]: use V = gridDim.x * NV * 32 * VEC virtual columns
  const int V = gridDim.x * NV * 32 * VEC;
  const int v0 = blockIdx.x * NV * 32 * VEC;
  colsum_commit(sh, v0, NV * 32 * VEC, V, ns, blockIdx.y, gridDim.y, blockIdx.x, [&](int v, float sum) {
    const int local = v - v0;
    const int k = local / (32 * VEC), j = local % (32 * VEC);
    if (j < ncol) {
      // arrays 0..NV-2: element (k, channel) at out_base[k * out_stride + channel * out_cstride] (tap-major [9][C] with
      // (C, 1), the parameter's own [C, 1, 3, 3] layout with (1, 9)); the last array is a plain [C] vector
      if (k == NV - 1) out_last[cbase + j] += sum;
      else out_base[(long long)k * out_stride + (long long)(cbase + j) * out_cstride] += sum;
    }
  });
}


}  // namespace sfk
