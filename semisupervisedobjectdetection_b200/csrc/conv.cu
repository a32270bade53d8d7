// Channels-last (token-major) convolution helpers of the SegFormer path:
//   * depthwise 3x3 (+bias) (+exact-erf GELU)                      MS:283-294, MS:310-313
//   * its backward (recompute pre-activation, dpre, dW, db, dgrad)
//   * patch gather ("im2col") for the overlapping patch-embed conv MS:123-131 and for the k==stride SR conv MS:171-173,
//     and the matching scatter/gather for their input gradients.
// All kernels move 64/128-bit vectors with the channel dimension innermost (fully coalesced).
#include "common.cuh"
#include "sfk_internal.h"
#include "colsum_slots.cuh"

namespace sfk {

// ------------------------------------------------------------------------------------------------
// depthwise 3x3, stride 1, pad 1.  x: bf16 [B,H,W,C].  wt: fp32 [9][C] (tap-major copy of weight[C,1,3,3]).
// Mapping: a warp = 32 consecutive channel groups (VEC channels each) of ONE (b, y, 4-pixel segment) position, so every
// load/store instruction of a warp is one contiguous 256/512-byte run; the 8 warps of a block walk different positions.
// Each thread produces SEG consecutive outputs along W, loading its 3 x (SEG+2) window once (4.5 loads per output).
// Per-thread channel sums (bias gradients) are reduced across the block in shared memory -> one atomic per channel/block.
// ------------------------------------------------------------------------------------------------
static constexpr int DW_SEG = 4;

// erf with |err| <= 1.5e-7 (Abramowitz-Stegun 7.1.26) from one rcp and one ex2; `e` = exp(-x*x) is passed in so the
// GELU derivative can reuse it for the Gaussian pdf.
__device__ __forceinline__ float erf_from_exp(float ax, float e) {
  const float t = __fdividef(1.0f, fmaf(0.3275911f, ax, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  return 1.0f - p * t * e;
}
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float e = __expf(-z * z);
  const float er = copysignf(erf_from_exp(fabsf(z), e), x);
  return 0.5f * x * (1.0f + er);
}
__device__ __forceinline__ float gelu_grad_fast(float x) {
  const float z = x * 0.70710678118654752440f;
  const float e = __expf(-z * z);  // == exp(-x^2/2)
  const float er = copysignf(erf_from_exp(fabsf(z), e), x);
  return fmaf(x * 0.39894228040143267794f, e, 0.5f * (1.0f + er));
}

// 128-bit shared-memory read the compiler may not hoist out of the position loop (loop-invariant filter taps would
// otherwise be promoted back into ~80 registers, which is exactly what keeping them in shared memory is meant to avoid)
__device__ __forceinline__ float4 lds128_nohoist(const float* p) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(smem_u32(p)));
  return v;
}

template <int VEC> struct VecIO;
template <> struct VecIO<8> {
  typedef uint4 packed;
  static __device__ __forceinline__ packed ldp(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
  static __device__ __forceinline__ packed zero() { return make_uint4(0, 0, 0, 0); }
  static __device__ __forceinline__ void unpack(const packed& u, float* f) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) { unpack(ldp(p), f); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    *reinterpret_cast<uint4*>(p) = u;
  }
};
template <> struct VecIO<4> {
  typedef uint2 packed;
  static __device__ __forceinline__ packed ldp(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  static __device__ __forceinline__ packed zero() { return make_uint2(0, 0); }
  static __device__ __forceinline__ void unpack(const packed& u, float* f) {
    float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y;
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float* f) { unpack(ldp(p), f); }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float* f) {
    uint2 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    *reinterpret_cast<uint2*>(p) = u;
  }
};

// block-level reduction of per-thread channel partials (8 warps -> shared memory), then the slotted low-contention commit:
// virtual column v = k * C + channel  (k-th output array)
template <int VEC, int NV>
__device__ __forceinline__ void dw_block_commit(float (&acc)[NV][VEC], float* sh, float* out_base, long long out_stride,
                                                float* out_last, int c0, int C, bool lane_ok, int ns) {
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
      float* dst = (k == NV - 1) ? out_last : out_base + (long long)k * out_stride;
      dst[cbase + j] += sum;
    }
  });
}

// mode 0: y = gelu(conv(x) + bias)     mode 1: y = conv_flipped(x)  (input gradient of the conv); colsum (+)= sum_pixels y
// ncu (profiles/): the first versions were (a) latency bound — loads consumed one by one — then (b) issue bound — 64-bit
// index division and per-(tap,output) re-unpacking.  This version: 8 channels x 2 pixels per thread, the 3x4 window is
// loaded up front (12 independent 16-byte loads), every window vector is unpacked once and feeds all taps that see it,
// filter taps / bias come from shared memory ([tap][lane][8], conflict-free 128-bit reads), 32-bit index arithmetic.
static constexpr int DWF_SEG = 4;
static constexpr int DWB_SEG = 2;

template <int MODE>
__global__ void __launch_bounds__(256, 1) dwconv3x3_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                                                           const float* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                                                           float* __restrict__ colsum, int B, int H, int W, int C,
                                                           int pos_per_block, int ns) {
  pdl_sync();
  constexpr int VEC = 8;
  __shared__ __align__(16) float sw[10 * 32 * VEC];   // 9 taps + bias
  __shared__ float sh[32 * VEC];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cblk = blockIdx.x * 32 * VEC;
  const int c0 = cblk + lane * VEC;
  const bool lane_ok = c0 < C;
  for (int i = threadIdx.x; i < 10 * 32 * VEC; i += 256) {
    const int k = i / (32 * VEC), j = i % (32 * VEC);
    float v = 0.f;
    if (cblk + j < C) {
      if (k < 9) v = wt[(MODE == 1 ? 8 - k : k) * C + cblk + j];
      else if (MODE == 0) v = bias[cblk + j];
    }
    sw[i] = v;
  }
  __syncthreads();
  const unsigned nseg = (unsigned)(W + DWF_SEG - 1) / DWF_SEG;
  const unsigned npos = (unsigned)B * H * nseg;
  const unsigned p0 = blockIdx.y * (unsigned)pos_per_block;
  unsigned p1 = p0 + pos_per_block;
  if (p1 > npos) p1 = npos;
  float csum[1][VEC];
#pragma unroll
  for (int v = 0; v < VEC; ++v) csum[0][v] = 0.f;
  if (lane_ok) {
    const float* swl = sw + lane * VEC;
    for (unsigned pos = p0 + warp; pos < p1; pos += 8) {
      const unsigned seg = pos % nseg, t = pos / nseg;
      const int yy = (int)(t % (unsigned)H), b = (int)(t / (unsigned)H);
      const int x0 = (int)seg * DWF_SEG;
      uint4 win[3][DWF_SEG + 2];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int sy = yy + r - 1;
        const bool vy = sy >= 0 && sy < H;
        const unsigned rowoff = ((unsigned)(b * H + sy) * W) * C + c0;
#pragma unroll
        for (int j = 0; j < DWF_SEG + 2; ++j) {
          const int sx = x0 + j - 1;
          win[r][j] = (vy && sx >= 0 && sx < W) ? *reinterpret_cast<const uint4*>(x + (size_t)(rowoff + (unsigned)sx * C))
                                                : make_uint4(0, 0, 0, 0);
        }
      }
      const unsigned outoff = ((unsigned)(b * H + yy) * W + x0) * C + c0;
      // minimum-instruction order: every window vector is unpacked once and feeds the (up to 3) outputs that see it
      float acc[DWF_SEG][VEC];
      {
        const float4 b0 = lds128_nohoist(swl + 9 * 32 * VEC);
        const float4 b1 = lds128_nohoist(swl + 9 * 32 * VEC + 4);
#pragma unroll
        for (int o = 0; o < DWF_SEG; ++o) {
          acc[o][0] = b0.x; acc[o][1] = b0.y; acc[o][2] = b0.z; acc[o][3] = b0.w;
          acc[o][4] = b1.x; acc[o][5] = b1.y; acc[o][6] = b1.z; acc[o][7] = b1.w;
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float wr[3][VEC];
#pragma unroll
        for (int tt = 0; tt < 3; ++tt) {
          const float4 w0 = lds128_nohoist(swl + (r * 3 + tt) * 32 * VEC);
          const float4 w1 = lds128_nohoist(swl + (r * 3 + tt) * 32 * VEC + 4);
          wr[tt][0] = w0.x; wr[tt][1] = w0.y; wr[tt][2] = w0.z; wr[tt][3] = w0.w;
          wr[tt][4] = w1.x; wr[tt][5] = w1.y; wr[tt][6] = w1.z; wr[tt][7] = w1.w;
        }
#pragma unroll
        for (int j = 0; j < DWF_SEG + 2; ++j) {
          float f[VEC];
          VecIO<VEC>::unpack(win[r][j], f);
#pragma unroll
          for (int tt = 0; tt < 3; ++tt) {
            const int o = j - tt;
            if (o >= 0 && o < DWF_SEG) {
#pragma unroll
              for (int v = 0; v < VEC; ++v) acc[o][v] = fmaf(wr[tt][v], f[v], acc[o][v]);
            }
          }
        }
      }
#pragma unroll
      for (int o = 0; o < DWF_SEG; ++o) {
        if (x0 + o >= W) break;
        if (MODE == 0) {
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[o][v] = gelu_fast(acc[o][v]);
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) csum[0][v] += acc[o][v];
        }
        VecIO<VEC>::st(y + (size_t)(outoff + (unsigned)o * C), acc[o]);
      }
    }
  }
  if (MODE == 1 && colsum != nullptr) dw_block_commit<VEC, 1>(csum, sh, nullptr, 0, colsum, c0, C, lane_ok, ns);
}

static void dw_grid(int C, int vec, long long npos, dim3& grid, int& ppb) {
  // npos = number of (b, y, segment) positions; one warp handles 32 channel groups of one position
  const int gx = (C / vec + 31) / 32;
  long long gy = ((long long)num_sms() * 8 + gx - 1) / gx;   // ~8 blocks per SM in flight across the grid
  if (gy < 1) gy = 1;
  long long per = (npos + gy - 1) / gy;
  if (per < 8) per = 8;
  per = (per + 7) / 8 * 8;
  gy = (npos + per - 1) / per;
  ppb = (int)per;
  grid = dim3(gx, (unsigned)gy);
}

int dwconv3x3(const void* x, const float* wt, const float* bias, void* y, float* colsum, int B, int H, int W, int C, int mode,
              cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0, "dwconv3x3: C=%d must be a multiple of 8", C);
  SFK_REQUIRE((long long)B * H * W * C < (1LL << 31), "dwconv3x3: tensor too large for 32-bit offsets");
  const long long npos = (long long)B * H * ((W + DWF_SEG - 1) / DWF_SEG);
  if (npos == 0) return 0;
  dim3 grid;
  int ppb;
  dw_grid(C, 8, npos, grid, ppb);
  if (mode == 0)
    launch_k(dwconv3x3_kernel<0>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, bias, (__nv_bfloat16*)y, nullptr, B, H, W, C, ppb, 1);
  else
    launch_k(dwconv3x3_kernel<1>, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, nullptr, (__nv_bfloat16*)y, colsum, B, H, W, C, ppb,
             colsum_num_slots((long long)grid.x * 32 * 8, grid.y));
  return check_launch("dwconv3x3_kernel");
}

// Backward stage 1: recompute pre = conv(x)+bias, dpre = dg * gelu'(pre); accumulate dW[9][C] and db[C].
// Window kept packed (bf16), filter taps in shared memory, unpack-on-use: no spills at 128 registers (2 blocks / SM).
__global__ void __launch_bounds__(256, 2) dwconv3x3_gelu_bwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ wt,
                                                                    const float* __restrict__ bias,
                                                                    const __nv_bfloat16* __restrict__ dg,
                                                                    __nv_bfloat16* __restrict__ dpre, float* __restrict__ dwt,
                                                                    float* __restrict__ dbias, int B, int H, int W, int C,
                                                                    int pos_per_block, int ns) {
  pdl_sync();
  constexpr int VEC = 4;
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
  const unsigned nseg = (unsigned)(W + DWB_SEG - 1) / DWB_SEG;
  const unsigned npos = (unsigned)B * H * nseg;
  const unsigned p0 = blockIdx.y * (unsigned)pos_per_block;
  unsigned p1 = p0 + pos_per_block;
  if (p1 > npos) p1 = npos;
  float acc[10][VEC], bs[VEC];  // acc[0..8] = dW taps, acc[9] = dbias
#pragma unroll
  for (int k = 0; k < 10; ++k)
#pragma unroll
    for (int v = 0; v < VEC; ++v) acc[k][v] = 0.f;
  if (lane_ok) {
    {
      const float4 a = __ldg(reinterpret_cast<const float4*>(bias + c0));
      bs[0] = a.x; bs[1] = a.y; bs[2] = a.z; bs[3] = a.w;
    }
    const float* swl = sh + lane * VEC;
    for (unsigned pos = p0 + warp; pos < p1; pos += 8) {
      const unsigned seg = pos % nseg, t = pos / nseg;
      const int yy = (int)(t % (unsigned)H), b = (int)(t / (unsigned)H);
      const int x0 = (int)seg * DWB_SEG;
      VecIO<VEC>::packed win[3][DWB_SEG + 2];
      VecIO<VEC>::packed gin[DWB_SEG];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int sy = yy + r - 1;
        const bool vy = sy >= 0 && sy < H;
        const unsigned rowoff = ((unsigned)(b * H + sy) * W) * C + c0;
#pragma unroll
        for (int j = 0; j < DWB_SEG + 2; ++j) {
          const int sx = x0 + j - 1;
          win[r][j] = (vy && sx >= 0 && sx < W) ? VecIO<VEC>::ldp(x + (size_t)(rowoff + (unsigned)sx * C)) : VecIO<VEC>::zero();
        }
      }
      const unsigned off0 = ((unsigned)(b * H + yy) * W + x0) * C + c0;
#pragma unroll
      for (int o = 0; o < DWB_SEG; ++o)
        gin[o] = (x0 + o < W) ? VecIO<VEC>::ldp(dg + (size_t)(off0 + (unsigned)o * C)) : VecIO<VEC>::zero();
#pragma unroll
      for (int o = 0; o < DWB_SEG; ++o) {
        if (x0 + o >= W) break;
        float pre[VEC];
#pragma unroll
        for (int v = 0; v < VEC; ++v) pre[v] = bs[v];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          const float4 w4 = lds128_nohoist(swl + k * 32 * VEC);
          float f[VEC];
          VecIO<VEC>::unpack(win[k / 3][o + k % 3], f);
          pre[0] = fmaf(w4.x, f[0], pre[0]); pre[1] = fmaf(w4.y, f[1], pre[1]);
          pre[2] = fmaf(w4.z, f[2], pre[2]); pre[3] = fmaf(w4.w, f[3], pre[3]);
        }
        float g[VEC], d[VEC];
        VecIO<VEC>::unpack(gin[o], g);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          d[v] = g[v] * gelu_grad_fast(pre[v]);
          acc[9][v] += d[v];
        }
        VecIO<VEC>::st(dpre + (size_t)(off0 + (unsigned)o * C), d);
#pragma unroll
        for (int k = 0; k < 9; ++k) {
          float f[VEC];
          VecIO<VEC>::unpack(win[k / 3][o + k % 3], f);
#pragma unroll
          for (int v = 0; v < VEC; ++v) acc[k][v] = fmaf(d[v], f[v], acc[k][v]);
        }
      }
    }
  }
  if (dwt != nullptr) {
    __syncthreads();  // everyone is done with the filter taps in sh
    dw_block_commit<VEC, 10>(acc, sh, dwt, (long long)C, dbias, c0, C, lane_ok, ns);
  }
}

int dwconv3x3_gelu_bwd(const void* x, const float* wt, const float* bias, const void* dg, void* dpre, float* dwt,
                       float* dbias, int B, int H, int W, int C, cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0, "dwconv3x3_gelu_bwd: C=%d must be a multiple of 8", C);
  SFK_REQUIRE((dwt == nullptr) == (dbias == nullptr), "dwconv3x3_gelu_bwd: pass both or neither of dwt / dbias");
  SFK_REQUIRE((long long)B * H * W * C < (1LL << 31), "dwconv3x3_gelu_bwd: tensor too large for 32-bit offsets");
  const long long npos = (long long)B * H * ((W + DWB_SEG - 1) / DWB_SEG);
  if (npos == 0) return 0;
  dim3 grid;
  int ppb;
  dw_grid(C, 4, npos, grid, ppb);
  launch_k(dwconv3x3_gelu_bwd_kernel, dim3(grid), dim3(256), 0, stream, (const __nv_bfloat16*)x, wt, bias, (const __nv_bfloat16*)dg,
                                                      (__nv_bfloat16*)dpre, dwt, dbias, B, H, W, C, ppb,
           colsum_num_slots((long long)grid.x * 10 * 32 * 4, grid.y));
  return check_launch("dwconv3x3_gelu_bwd_kernel");
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
__global__ void __launch_bounds__(256) im2col_nchw_f32_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ A, int B,
                                                              int Cin, int H, int W, int K, int S, int P, int Ho, int Wo,
                                                              int Kp) {
  pdl_sync();
  const long long total = (long long)B * Ho * Wo * Kp;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int k = (int)(idx % Kp);
  const long long m = idx / Kp;
  float v = 0.f;
  if (k < Cin * K * K) {
    const int c = k / (K * K), r = k % (K * K);
    const int ky = r / K, kx = r % K;
    const int ox = (int)(m % Wo);
    const int oy = (int)((m / Wo) % Ho);
    const long long b = m / ((long long)Wo * Ho);
    const int sy = oy * S + ky - P, sx = ox * S + kx - P;
    if (sy >= 0 && sy < H && sx >= 0 && sx < W) v = x[((b * Cin + c) * H + sy) * W + sx];
  }
  A[idx] = __float2bfloat16(v);
}

int im2col_nchw_f32(const float* x, void* A, int B, int Cin, int H, int W, int K, int S, int P, int Kp,
                    cudaStream_t stream) {
  const int Ho = (H + 2 * P - K) / S + 1, Wo = (W + 2 * P - K) / S + 1;
  const long long total = (long long)B * Ho * Wo * Kp;
  if (total == 0) return 0;
  launch_k(im2col_nchw_f32_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, (__nv_bfloat16*)A, B, Cin, H, W, K, S, P, Ho,
                                                                              Wo, Kp);
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
  const long long total = (long long)O * I * KH * KW;
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int kx = (int)(idx % KW);
  long long t = idx / KW;
  const int ky = (int)(t % KH);
  t /= KH;
  const int c = (int)(t % I);
  const long long o = t / I;
  dw[idx] += g[o * ldg + ((long long)ky * KW + kx) * I + c];
}

int conv_wgrad_ohwi_to_oihw(const float* g, float* dw, int O, int I, int KH, int KW, long long ldg, cudaStream_t stream) {
  const long long total = (long long)O * I * KH * KW;
  if (total == 0) return 0;
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
int dw_wgrad_from_tap(const float* dwt, float* dw, int C, cudaStream_t stream) {
  launch_k(dw_wgrad_from_tap_kernel, dim3((9 * C + 255) / 256), dim3(256), 0, stream, dwt, dw, C);
  return check_launch("dw_wgrad_from_tap_kernel");
}

}  // namespace sfk
