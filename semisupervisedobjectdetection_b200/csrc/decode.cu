// All-MLP decode head kernels (MS:752-782), channels-last:
//   * bilinear upsample (align_corners=False) of a per-stage projection into its channel slice of the concat buffer
//     (MS:769-774, concat order c4,c3,c2,c1) and the transposed gather for its gradient
//   * BatchNorm2d batch statistics / running-stat update (MS:744,775; eps 1e-5, momentum 0.1)
//   * BN apply + ReLU + Dropout + 1x1 classifier fused into one pass over the fuse-conv output (MS:775-780),
//     and the two-pass backward (per-channel reductions, then dx).
#include "common.cuh"
#include "sfk_internal.h"
#include "colsum_slots.cuh"

namespace sfk {

// PyTorch area_pixel_compute_source_index(align_corners=False): src = scale*(dst+0.5)-0.5, clamped at 0
__device__ __forceinline__ void bilinear_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - i0;
}

// out[b, y, x, ch_off + c] = bilinear(in[b, :, :, c]);  in: bf16 [B,h,w,C], out: bf16 [B,H,W,ld_out]
// One block per output row (b, y): the vertical taps are computed once per block, threads walk (x, 8-channel chunk) with
// 32-bit index arithmetic (the first version spent most of its instructions on 64-bit div/mod per element).
__global__ void __launch_bounds__(256) upsample_concat_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                                              int B, int h, int w, int C, int H, int W, long long ld_out,
                                                              int ch_off) {
  pdl_sync();
  const int C8 = C >> 3;
  const int y = blockIdx.x % H;
  const int b = blockIdx.x / H;
  const bool ident = (h == H && w == W);
  int y0 = y, y1 = y;
  float ly = 0.f;
  if (!ident) bilinear_src(y, (float)h / (float)H, h, y0, y1, ly);
  const float sx = (float)w / (float)W;
  const __nv_bfloat16* row0 = in + ((size_t)b * h + y0) * w * C;
  const __nv_bfloat16* row1 = in + ((size_t)b * h + y1) * w * C;
  __nv_bfloat16* orow = out + ((size_t)b * H + y) * W * ld_out + ch_off;
  const int n = W * C8;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int x = i / C8, c8 = i - x * C8;
    __nv_bfloat16* op = orow + (size_t)x * ld_out + c8 * 8;
    if (ident) {
      *reinterpret_cast<uint4*>(op) = *reinterpret_cast<const uint4*>(row0 + (size_t)x * C + c8 * 8);
      continue;
    }
    int x0, x1;
    float lx;
    bilinear_src(x, sx, w, x0, x1, lx);
    const float w00 = (1.f - ly) * (1.f - lx), w01 = (1.f - ly) * lx, w10 = ly * (1.f - lx), w11 = ly * lx;
    const uint4 a = *reinterpret_cast<const uint4*>(row0 + (size_t)x0 * C + c8 * 8);
    const uint4 bq = *reinterpret_cast<const uint4*>(row0 + (size_t)x1 * C + c8 * 8);
    const uint4 c = *reinterpret_cast<const uint4*>(row1 + (size_t)x0 * C + c8 * 8);
    const uint4 d = *reinterpret_cast<const uint4*>(row1 + (size_t)x1 * C + c8 * 8);
    const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* bp = reinterpret_cast<const uint32_t*>(&bq);
    const uint32_t* cp = reinterpret_cast<const uint32_t*>(&c);
    const uint32_t* dp = reinterpret_cast<const uint32_t*>(&d);
    uint4 o;
    uint32_t* o32 = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 fa = unpack_bf16x2(ap[k]), fb = unpack_bf16x2(bp[k]), fc = unpack_bf16x2(cp[k]), fd = unpack_bf16x2(dp[k]);
      o32[k] = pack_bf16x2(w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x, w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y);
    }
    *reinterpret_cast<uint4*>(op) = o;
  }
}

int upsample_concat(const void* in, void* out, int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off,
                    cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0 && ld_out % 8 == 0 && ch_off % 8 == 0, "upsample_concat: channel counts must be multiples of 8");
  const long long total = (long long)B * H * W * (C / 8);
  if (total == 0) return 0;
  launch_k(upsample_concat_kernel, dim3((unsigned)(B * H)), dim3(256), 0, stream, (const __nv_bfloat16*)in, (__nv_bfloat16*)out, B, h, w, C,
           H, W, ld_out, ch_off);
  return check_launch("upsample_concat_kernel");
}

// Folded decode head (algebraic shortcut): bilinear upsampling commutes with the 1x1 convs, so the fuse conv of
// the concatenated, upsampled projections equals the SUM of per-stage projections (folded weights) upsampled to the
// stage-1 grid.  out[b,y,x,:] = z0[b,y,x,:] + sum_{i=1..3} bilinear(z_i)[b,y,x,:] + bias[:]   (bf16 in / out, fp32 math)
// One block per output row (b, y); the vertical taps of the three coarser levels are computed once per block.
struct UpLevel {
  const __nv_bfloat16* z;
  int h, w;
};
__global__ void __launch_bounds__(256) upsample_add4_kernel(const __nv_bfloat16* __restrict__ z0, UpLevel l1, UpLevel l2, UpLevel l3,
                                                            const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B,
                                                            int H, int W, int C) {
  pdl_sync();
  const int C8 = C >> 3;
  const int y = blockIdx.x % H;
  const int b = blockIdx.x / H;
  const UpLevel lv[3] = {l1, l2, l3};
  const __nv_bfloat16* r0[3];
  const __nv_bfloat16* r1[3];
  float ly[3], sx[3];
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    int y0, y1;
    bilinear_src(y, (float)lv[k].h / (float)H, lv[k].h, y0, y1, ly[k]);
    r0[k] = lv[k].z + ((size_t)b * lv[k].h + y0) * lv[k].w * C;
    r1[k] = lv[k].z + ((size_t)b * lv[k].h + y1) * lv[k].w * C;
    sx[k] = (float)lv[k].w / (float)W;
  }
  const __nv_bfloat16* zrow = z0 + ((size_t)b * H + y) * W * C;
  __nv_bfloat16* orow = out + ((size_t)b * H + y) * W * C;
  const int n = W * C8;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int x = i / C8, c8 = i - x * C8;
    float acc[8];
    {
      const uint4 a = *reinterpret_cast<const uint4*>(zrow + (size_t)x * C + c8 * 8);
      const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + c8 * 8 + 4));
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 f = unpack_bf16x2(ap[k]);
        acc[2 * k] = f.x + bb[2 * k];
        acc[2 * k + 1] = f.y + bb[2 * k + 1];
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int x0, x1;
      float lx;
      bilinear_src(x, sx[k], lv[k].w, x0, x1, lx);
      const float w00 = (1.f - ly[k]) * (1.f - lx), w01 = (1.f - ly[k]) * lx, w10 = ly[k] * (1.f - lx), w11 = ly[k] * lx;
      const uint4 a = *reinterpret_cast<const uint4*>(r0[k] + (size_t)x0 * C + c8 * 8);
      const uint4 bq = *reinterpret_cast<const uint4*>(r0[k] + (size_t)x1 * C + c8 * 8);
      const uint4 c = *reinterpret_cast<const uint4*>(r1[k] + (size_t)x0 * C + c8 * 8);
      const uint4 d = *reinterpret_cast<const uint4*>(r1[k] + (size_t)x1 * C + c8 * 8);
      const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
      const uint32_t* bp = reinterpret_cast<const uint32_t*>(&bq);
      const uint32_t* cp = reinterpret_cast<const uint32_t*>(&c);
      const uint32_t* dp = reinterpret_cast<const uint32_t*>(&d);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 fa = unpack_bf16x2(ap[j]), fb = unpack_bf16x2(bp[j]), fc = unpack_bf16x2(cp[j]), fd = unpack_bf16x2(dp[j]);
        acc[2 * j] += w00 * fa.x + w01 * fb.x + w10 * fc.x + w11 * fd.x;
        acc[2 * j + 1] += w00 * fa.y + w01 * fb.y + w10 * fc.y + w11 * fd.y;
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(orow + (size_t)x * C + c8 * 8) = o;
  }
}

// Separable version (default): a block owns one output row (b, y) and a 64-channel slab.  Phase 1 interpolates the two
// contributing rows of each coarse level VERTICALLY once (fp32, shared memory: (w1 + w2 + w3) x 64 values); phase 2 walks
// the output pixels and takes two horizontal taps per level from shared memory.  The direct kernel above recomputed the
// four-tap weights and unpacked 12 bf16 vectors per output vector (~250 instructions per 8 channels: issue bound at 436 us
// for the 16 x 128 x 128 x 768 tile batch, 1.9 TB/s); this one runs in 326 us (2.5 TB/s, ncu: still 66 % issue-slot busy -
// the per-pixel source-index arithmetic is the next thing to hoist).
// Shared-memory pixel layout: [8 chunks x floats 0..3][8 chunks x floats 4..7], so the 8 threads of a pixel read 128
// contiguous bytes per LDS.128.
static constexpr int UPS_CS = 64;
// sfk_set_option("upsample_sep", 0|1): separable shared-memory upsample-add / transposed upsampling (default) or the direct kernels
int g_upsample_sep = 1;
__global__ void __launch_bounds__(256) upsample_add4_sep_kernel(const __nv_bfloat16* __restrict__ z0, UpLevel l1, UpLevel l2, UpLevel l3,
                                                                const float* __restrict__ bias, __nv_bfloat16* __restrict__ out, int B,
                                                                int H, int W, int C) {
  pdl_sync();
  extern __shared__ __align__(16) float vrow[];
  constexpr int CS = UPS_CS, K8 = CS / 8;
  // grid = (channel slabs, rows): consecutive blocks are the slabs of ONE row, so the row's 128-byte pieces are requested
  // together (with rows fastest every block touched one line per 1.5 KB and the kernel sat at 2.5 TB/s of scattered lines)
  const int y = blockIdx.y % H;
  const int b = blockIdx.y / H;
  const int cb = blockIdx.x * CS;
  const int off1 = l1.w, off2 = l1.w + l2.w, wtot = l1.w + l2.w + l3.w;
  constexpr int PX = 256 / K8;
  const int c8 = threadIdx.x % K8;
  const bool c_ok = cb + c8 * 8 < C;
  const __nv_bfloat16* zrow = z0 + ((size_t)b * H + y) * W * C + cb + c8 * 8;
  // ---- phase 1: vertical interpolation of every coarse pixel of the two source rows ----
  for (int i = threadIdx.x; i < wtot * K8; i += 256) {
    const int xa = i / K8, c8 = i - xa * K8;   // (shadows the phase-2 chunk index: same value, 256 % K8 == 0)
    const int k = xa < off1 ? 0 : (xa < off2 ? 1 : 2);
    const int xc = xa - (k == 0 ? 0 : (k == 1 ? off1 : off2));
    // (selects instead of an indexed UpLevel array: a runtime index would put the array in local memory)
    const __nv_bfloat16* zk = k == 0 ? l1.z : (k == 1 ? l2.z : l3.z);
    const int hk = k == 0 ? l1.h : (k == 1 ? l2.h : l3.h), wk = k == 0 ? l1.w : (k == 1 ? l2.w : l3.w);
    int y0, y1;
    float ly;
    bilinear_src(y, (float)hk / (float)H, hk, y0, y1, ly);
    const __nv_bfloat16* p0 = zk + (((size_t)b * hk + y0) * wk + xc) * C + cb + c8 * 8;
    const __nv_bfloat16* p1 = zk + (((size_t)b * hk + y1) * wk + xc) * C + cb + c8 * 8;
    uint4 a = make_uint4(0, 0, 0, 0), c = make_uint4(0, 0, 0, 0);
    if (cb + c8 * 8 < C) {
      a = *reinterpret_cast<const uint4*>(p0);
      c = *reinterpret_cast<const uint4*>(p1);
    }
    const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
    const uint32_t* cp = reinterpret_cast<const uint32_t*>(&c);
    float v[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 fa = unpack_bf16x2(ap[j]), fc = unpack_bf16x2(cp[j]);
      v[2 * j] = fa.x + ly * (fc.x - fa.x);
      v[2 * j + 1] = fa.y + ly * (fc.y - fa.y);
    }
    float* dst = vrow + (size_t)xa * CS + c8 * 4;
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(dst + CS / 2) = make_float4(v[4], v[5], v[6], v[7]);
  }
  __syncthreads();
  // ---- phase 2: output pixels; thread = (pixel lane, 8-channel chunk) ----
  if (!c_ok) return;
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + cb + c8 * 8));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(bias + cb + c8 * 8 + 4));
  const float sx1 = (float)l1.w / (float)W, sx2 = (float)l2.w / (float)W, sx3 = (float)l3.w / (float)W;
  __nv_bfloat16* orow = out + ((size_t)b * H + y) * W * C + cb + c8 * 8;
  for (int x = threadIdx.x / K8; x < W; x += PX) {
    const uint4 a = *reinterpret_cast<const uint4*>(zrow + (size_t)x * C);
    const uint32_t* ap = reinterpret_cast<const uint32_t*>(&a);
    float acc[8];
    {
      const float2 f0 = unpack_bf16x2(ap[0]), f1 = unpack_bf16x2(ap[1]), f2 = unpack_bf16x2(ap[2]), f3 = unpack_bf16x2(ap[3]);
      acc[0] = f0.x + b0.x; acc[1] = f0.y + b0.y; acc[2] = f1.x + b0.z; acc[3] = f1.y + b0.w;
      acc[4] = f2.x + b1.x; acc[5] = f2.y + b1.y; acc[6] = f3.x + b1.z; acc[7] = f3.y + b1.w;
    }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      int x0, x1;
      float lx;
      bilinear_src(x, k == 0 ? sx1 : (k == 1 ? sx2 : sx3), k == 0 ? l1.w : (k == 1 ? l2.w : l3.w), x0, x1, lx);
      const float* t0 = vrow + ((k == 0 ? 0 : (k == 1 ? off1 : off2)) + x0) * CS + c8 * 4;
      const float* t1 = t0 + (x1 - x0) * CS;
      const float4 p0 = *reinterpret_cast<const float4*>(t0), p1 = *reinterpret_cast<const float4*>(t0 + CS / 2);
      const float4 q0 = *reinterpret_cast<const float4*>(t1), q1 = *reinterpret_cast<const float4*>(t1 + CS / 2);
      const float w0 = 1.f - lx;
      acc[0] += w0 * p0.x + lx * q0.x; acc[1] += w0 * p0.y + lx * q0.y; acc[2] += w0 * p0.z + lx * q0.z; acc[3] += w0 * p0.w + lx * q0.w;
      acc[4] += w0 * p1.x + lx * q1.x; acc[5] += w0 * p1.y + lx * q1.y; acc[6] += w0 * p1.z + lx * q1.z; acc[7] += w0 * p1.w + lx * q1.w;
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(orow + (size_t)x * C) = o;
  }
}


int upsample_add4(const void* z0, const void* z1, int h1, int w1, const void* z2, int h2, int w2, const void* z3, int h3, int w3,
                  const float* bias, void* out, int B, int H, int W, int C, cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0, "upsample_add4: C=%d must be a multiple of 8", C);
  SFK_REQUIRE(h1 > 0 && w1 > 0 && h2 > 0 && w2 > 0 && h3 > 0 && w3 > 0, "upsample_add4: empty level");
  if ((long long)B * H * W == 0) return 0;
  UpLevel l1{(const __nv_bfloat16*)z1, h1, w1}, l2{(const __nv_bfloat16*)z2, h2, w2}, l3{(const __nv_bfloat16*)z3, h3, w3};
  const size_t sep_smem = (size_t)(w1 + w2 + w3) * UPS_CS * sizeof(float);
  if (g_upsample_sep && sep_smem <= 200 * 1024 && (long long)B * H <= 65535) {
    static size_t attr_bytes = 48 * 1024;
    if (sep_smem > attr_bytes) {
      cudaError_t e = cudaFuncSetAttribute(upsample_add4_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_smem);
      SFK_REQUIRE(e == cudaSuccess, "upsample_add4: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      attr_bytes = sep_smem;
    }
    launch_k(upsample_add4_sep_kernel, dim3((unsigned)((C + UPS_CS - 1) / UPS_CS), (unsigned)(B * H)), dim3(256), sep_smem, stream,
             (const __nv_bfloat16*)z0, l1, l2, l3, bias, (__nv_bfloat16*)out, B, H, W, C);
    return check_launch("upsample_add4_sep_kernel");
  }
  launch_k(upsample_add4_kernel, dim3((unsigned)(B * H)), dim3(256), 0, stream, (const __nv_bfloat16*)z0, l1, l2, l3, bias,
           (__nv_bfloat16*)out, B, H, W, C);
  return check_launch("upsample_add4_kernel");
}

// Transposed bilinear: din[b, yi, xi, c] = sum over hi-res (y, x) of weight(y,x -> yi,xi) * dout[b, y, x, ch_off + c].
// Gather formulation (no atomics, deterministic). One block per low-res row (b, yi): the contributing hi-res rows and their
// vertical weights are found once per block (shared memory), threads walk (xi, 8-channel chunk) with 32-bit arithmetic.
__global__ void __launch_bounds__(256) upsample_concat_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                                  __nv_bfloat16* __restrict__ din, int B, int h, int w, int C,
                                                                  int H, int W, long long ld_out, int ch_off) {
  pdl_sync();
  __shared__ int s_rows[64];
  __shared__ float s_wy[64];
  __shared__ int s_nrows;
  const int C8 = C >> 3;
  const int yi = blockIdx.x % h;
  const int b = blockIdx.x / h;
  const bool ident = (h == H && w == W);
  if (threadIdx.x == 0) {
    int n = 0;
    if (!ident) {
      const float sy = (float)h / (float)H, fy = (float)H / (float)h;
      int ylo = (int)floorf((yi - 1 + 0.5f) * fy - 0.5f) - 1, yhi = (int)ceilf((yi + 1 + 0.5f) * fy - 0.5f) + 1;
      if (ylo < 0) ylo = 0;
      if (yhi > H - 1) yhi = H - 1;
      for (int y = ylo; y <= yhi && n < 64; ++y) {
        int y0, y1;
        float ly;
        bilinear_src(y, sy, h, y0, y1, ly);
        float wy = 0.f;
        if (y0 == yi) wy += 1.f - ly;
        if (y1 == yi) wy += ly;
        if (wy != 0.f) { s_rows[n] = y; s_wy[n] = wy; ++n; }
      }
    }
    s_nrows = n;
  }
  __syncthreads();
  const int nrows = s_nrows;
  const float sx = (float)w / (float)W, fx = (float)W / (float)w;
  const int n = w * C8;
  for (int i = threadIdx.x; i < n; i += 256) {
    const int xi = i / C8, c8 = i - xi * C8;
    __nv_bfloat16* dst = din + (((size_t)b * h + yi) * w + xi) * C + c8 * 8;
    if (ident) {
      *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(dout + (((size_t)b * H + yi) * W + xi) * ld_out + ch_off + c8 * 8);
      continue;
    }
    int xlo = (int)floorf((xi - 1 + 0.5f) * fx - 0.5f) - 1, xhi = (int)ceilf((xi + 1 + 0.5f) * fx - 0.5f) + 1;
    if (xlo < 0) xlo = 0;
    if (xhi > W - 1) xhi = W - 1;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int x = xlo; x <= xhi; ++x) {
      int x0, x1;
      float lx;
      bilinear_src(x, sx, w, x0, x1, lx);
      float wx = 0.f;
      if (x0 == xi) wx += 1.f - lx;
      if (x1 == xi) wx += lx;
      if (wx == 0.f) continue;
      const __nv_bfloat16* col = dout + ((size_t)b * H * W + x) * ld_out + ch_off + c8 * 8;
      for (int r = 0; r < nrows; ++r) {
        const float wt = wx * s_wy[r];
        const uint4 v = *reinterpret_cast<const uint4*>(col + (size_t)s_rows[r] * W * ld_out);
        const uint32_t* vp = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(vp[k]);
          acc[2 * k] = fmaf(wt, f.x, acc[2 * k]);
          acc[2 * k + 1] = fmaf(wt, f.y, acc[2 * k + 1]);
        }
      }
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(dst) = o;
  }
}

// Separable version of the transposed bilinear gather (default for real down-scaling): a block owns one low-res row (b, yi)
// and a 64-channel slab.  Phase 1 reduces the contributing hi-res rows VERTICALLY (thread = (hi-res pixel, 8-channel chunk),
// fp32 registers, four independent 16-byte loads per row step), phase 2 reduces HORIZONTALLY out of shared memory.  The direct
// kernel above walks the whole (2f x 2f) window per output: (2f)^2 16-byte loads and weight products per 8 outputs - 258 /
// 285 / 330 us for the three levels of the 16 x 128 x 128 x 768 batch (1.2-1.6 TB/s).
__global__ void __launch_bounds__(256) upsample_bwd_sep_kernel(const __nv_bfloat16* __restrict__ dout, __nv_bfloat16* __restrict__ din,
                                                               int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off) {
  pdl_sync();
  extern __shared__ __align__(16) float tmp[];   // [W][CS]: per pixel [8 chunks x floats 0..3][8 chunks x floats 4..7]
  __shared__ int s_rows[64];
  __shared__ float s_wy[64];
  __shared__ int s_nrows;
  constexpr int CS = UPS_CS, K8 = CS / 8, PX = 256 / K8;
  const int yi = blockIdx.y % h;
  const int b = blockIdx.y / h;
  const int cb = blockIdx.x * CS;
  if (threadIdx.x == 0) {
    int n = 0;
    const float sy = (float)h / (float)H, fy = (float)H / (float)h;
    int ylo = (int)floorf((yi - 1 + 0.5f) * fy - 0.5f) - 1, yhi = (int)ceilf((yi + 1 + 0.5f) * fy - 0.5f) + 1;
    if (ylo < 0) ylo = 0;
    if (yhi > H - 1) yhi = H - 1;
    for (int y = ylo; y <= yhi && n < 64; ++y) {
      int y0, y1;
      float ly;
      bilinear_src(y, sy, h, y0, y1, ly);
      float wy = 0.f;
      if (y0 == yi) wy += 1.f - ly;
      if (y1 == yi) wy += ly;
      if (wy != 0.f) { s_rows[n] = y; s_wy[n] = wy; ++n; }
    }
    s_nrows = n;
  }
  __syncthreads();
  const int nrows = s_nrows;
  const int c8 = threadIdx.x % K8;
  const bool c_ok = cb + c8 * 8 < C;
  // ---- phase 1: tmp[x][c] = sum_rows wy * dout[row][x][c] ----
  for (int xb = threadIdx.x / K8; xb < W; xb += 4 * PX) {
    float acc[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[u][k] = 0.f;
    const __nv_bfloat16* base = dout + ((size_t)b * H * W + xb) * ld_out + ch_off + cb + c8 * 8;
    for (int r = 0; r < nrows; ++r) {
      const float wy = s_wy[r];
      const __nv_bfloat16* rowp = base + (size_t)s_rows[r] * W * ld_out;
      uint4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (c_ok && xb + u * PX < W) ? *reinterpret_cast<const uint4*>(rowp + (size_t)(u * PX) * ld_out) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const uint32_t* vp = reinterpret_cast<const uint32_t*>(&v[u]);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float2 f = unpack_bf16x2(vp[k]);
          acc[u][2 * k] = fmaf(wy, f.x, acc[u][2 * k]);
          acc[u][2 * k + 1] = fmaf(wy, f.y, acc[u][2 * k + 1]);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int x = xb + u * PX;
      if (x < W) {
        float* dst = tmp + (size_t)x * CS + c8 * 4;
        *reinterpret_cast<float4*>(dst) = make_float4(acc[u][0], acc[u][1], acc[u][2], acc[u][3]);
        *reinterpret_cast<float4*>(dst + CS / 2) = make_float4(acc[u][4], acc[u][5], acc[u][6], acc[u][7]);
      }
    }
  }
  __syncthreads();
  if (!c_ok) return;
  // ---- phase 2: din[xi][c] = sum_x wx * tmp[x][c] ----
  const float sx = (float)w / (float)W, fx = (float)W / (float)w;
  for (int xi = threadIdx.x / K8; xi < w; xi += PX) {
    int xlo = (int)floorf((xi - 1 + 0.5f) * fx - 0.5f) - 1, xhi = (int)ceilf((xi + 1 + 0.5f) * fx - 0.5f) + 1;
    if (xlo < 0) xlo = 0;
    if (xhi > W - 1) xhi = W - 1;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int x = xlo; x <= xhi; ++x) {
      int x0, x1;
      float lx;
      bilinear_src(x, sx, w, x0, x1, lx);
      float wx = 0.f;
      if (x0 == xi) wx += 1.f - lx;
      if (x1 == xi) wx += lx;
      if (wx == 0.f) continue;
      const float* t = tmp + (size_t)x * CS + c8 * 4;
      const float4 p0 = *reinterpret_cast<const float4*>(t), p1 = *reinterpret_cast<const float4*>(t + CS / 2);
      acc[0] = fmaf(wx, p0.x, acc[0]); acc[1] = fmaf(wx, p0.y, acc[1]); acc[2] = fmaf(wx, p0.z, acc[2]); acc[3] = fmaf(wx, p0.w, acc[3]);
      acc[4] = fmaf(wx, p1.x, acc[4]); acc[5] = fmaf(wx, p1.y, acc[5]); acc[6] = fmaf(wx, p1.z, acc[6]); acc[7] = fmaf(wx, p1.w, acc[7]);
    }
    uint4 o;
    o.x = pack_bf16x2(acc[0], acc[1]); o.y = pack_bf16x2(acc[2], acc[3]);
    o.z = pack_bf16x2(acc[4], acc[5]); o.w = pack_bf16x2(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(din + (((size_t)b * h + yi) * w + xi) * C + cb + c8 * 8) = o;
  }
}

int upsample_concat_bwd(const void* dout, void* din, int B, int h, int w, int C, int H, int W, long long ld_out, int ch_off,
                        cudaStream_t stream) {
  SFK_REQUIRE(C % 8 == 0 && ld_out % 8 == 0 && ch_off % 8 == 0, "upsample_concat_bwd: channel counts must be multiples of 8");
  const long long total = (long long)B * h * w * (C / 8);
  if (total == 0) return 0;
  SFK_REQUIRE(H <= 16 * h + 32 || h == H, "upsample_concat_bwd: scale factor too large");
  const size_t sep_smem = (size_t)W * UPS_CS * sizeof(float);
  if (g_upsample_sep && !(h == H && w == W) && sep_smem <= 200 * 1024 && (long long)B * h <= 65535) {
    static size_t attr_bytes = 48 * 1024;
    if (sep_smem > attr_bytes) {
      cudaError_t e = cudaFuncSetAttribute(upsample_bwd_sep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sep_smem);
      SFK_REQUIRE(e == cudaSuccess, "upsample_concat_bwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      attr_bytes = sep_smem;
    }
    launch_k(upsample_bwd_sep_kernel, dim3((unsigned)((C + UPS_CS - 1) / UPS_CS), (unsigned)(B * h)), dim3(256), sep_smem, stream,
             (const __nv_bfloat16*)dout, (__nv_bfloat16*)din, B, h, w, C, H, W, ld_out, ch_off);
    return check_launch("upsample_bwd_sep_kernel");
  }
  launch_k(upsample_concat_bwd_kernel, dim3((unsigned)(B * h)), dim3(256), 0, stream, (const __nv_bfloat16*)dout, (__nv_bfloat16*)din, B, h, w,
           C, H, W, ld_out, ch_off);
  return check_launch("upsample_concat_bwd_kernel");
}

// out[b*N + t, :] (bf16) += vec[b, :] (f32)     — cls_proj broadcast add (MS:720-723)
__global__ void __launch_bounds__(256) add_sample_bias_kernel(__nv_bfloat16* __restrict__ out, const float* __restrict__ vec,
                                                              long long rows, int rows_per_sample, int C) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  const int C2 = C >> 1;
  if (idx >= rows * C2) return;
  const int c2 = (int)(idx % C2);
  const long long row = idx / C2;
  const float2 v = *reinterpret_cast<const float2*>(vec + (row / rows_per_sample) * C + c2 * 2);
  uint32_t* p = reinterpret_cast<uint32_t*>(out + row * C + c2 * 2);
  const float2 o = unpack_bf16x2(*p);
  *p = pack_bf16x2(o.x + v.x, o.y + v.y);
}
int add_sample_bias(void* out, const float* vec, long long rows, int rows_per_sample, int C, cudaStream_t stream) {
  SFK_REQUIRE(C % 2 == 0, "add_sample_bias: C must be even");
  if (rows == 0) return 0;
  launch_k(add_sample_bias_kernel, dim3((unsigned)((rows * (C / 2) + 255) / 256)), dim3(256), 0, stream, (__nv_bfloat16*)out, vec, rows,
                                                                                       rows_per_sample, C);
  return check_launch("add_sample_bias_kernel");
}

// ------------------------------------------------------------------------------------------------
// BatchNorm batch statistics over rows of x [M, C] (bf16): partial[blk][0][c] = sum, partial[blk][1][c] = sum of squares
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) bn_partial_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ partial,
                                                         long long M, int C, int rows_per_block) {
  pdl_sync();
  const int c = (blockIdx.x * 128 + threadIdx.x) * 2;
  if (c >= C) return;
  const long long r0 = (long long)blockIdx.y * rows_per_block;
  long long r1 = r0 + rows_per_block;
  if (r1 > M) r1 = M;
  float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
  for (long long r = r0; r < r1; ++r) {
    const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(x + r * C + c));
    s0 += v.x; s1 += v.y;
    q0 = fmaf(v.x, v.x, q0); q1 = fmaf(v.y, v.y, q1);
  }
  float* p = partial + (long long)blockIdx.y * 2 * C;
  p[c] = s0; p[c + 1] = s1;
  p[C + c] = q0; p[C + c + 1] = q1;
}

// mean / rstd (biased variance) + running-stat update with the unbiased variance (nn.BatchNorm2d train mode)
// Block = 32 channels x 8 row lanes: each lane adds every 8th partial row in fp64, the lanes are combined through shared memory.
// (One thread per channel walking all ~400 partial rows with dependent fp64 adds took 66 us per launch.)
__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partial, int nblocks, long long M, int C, float eps,
                                                          float momentum, float* __restrict__ mean, float* __restrict__ rstd,
                                                          float* __restrict__ running_mean, float* __restrict__ running_var) {
  pdl_sync();
  __shared__ double sh_s[8][32], sh_q[8][32];
  const int lc = threadIdx.x & 31, lr = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lc;
  double s = 0.0, q = 0.0;
  if (c < C) {
    for (int b = lr; b < nblocks; b += 8) {
      s += (double)partial[(long long)b * 2 * C + c];
      q += (double)partial[(long long)b * 2 * C + C + c];
    }
  }
  sh_s[lr][lc] = s;
  sh_q[lr][lc] = q;
  __syncthreads();
  if (lr != 0 || c >= C) return;
#pragma unroll
  for (int j = 1; j < 8; ++j) { s += sh_s[j][lc]; q += sh_q[j][lc]; }
  const double mu = s / (double)M;
  double var = q / (double)M - mu * mu;
  if (var < 0.0) var = 0.0;
  mean[c] = (float)mu;
  rstd[c] = (float)(1.0 / sqrt(var + (double)eps));
  if (running_mean != nullptr) {
    const double unbiased = M > 1 ? var * (double)M / (double)(M - 1) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mu;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
  }
}

size_t bn_stats_workspace_bytes(long long M, int C) {
  long long gy = (long long)num_sms() * 8 / ((C / 2 + 127) / 128);
  if (gy < 1) gy = 1;
  if (gy > M) gy = M > 0 ? M : 1;
  return (size_t)gy * 2 * C * sizeof(float);
}

int bn_batch_stats(const void* x, long long M, int C, float eps, float momentum, float* mean, float* rstd, float* running_mean,
                   float* running_var, float* workspace, size_t workspace_bytes, cudaStream_t stream) {
  SFK_REQUIRE(C % 2 == 0, "bn_batch_stats: C must be even");
  SFK_REQUIRE(M > 0, "bn_batch_stats: empty batch");
  const int gx = (C / 2 + 127) / 128;
  long long gy = (long long)num_sms() * 8 / gx;
  if (gy < 1) gy = 1;
  if (gy > M) gy = M;
  SFK_REQUIRE(workspace_bytes >= (size_t)gy * 2 * C * sizeof(float), "bn_batch_stats: workspace too small");
  const int rpb = (int)((M + gy - 1) / gy);
  gy = (M + rpb - 1) / rpb;
  launch_k(bn_partial_kernel, dim3(dim3(gx, (unsigned)gy)), dim3(128), 0, stream, (const __nv_bfloat16*)x, workspace, M, C, rpb);
  int rc = check_launch("bn_partial_kernel");
  if (rc) return rc;
  launch_k(bn_finalize_kernel, dim3((C + 31) / 32), dim3(256), 0, stream, workspace, (int)gy, M, C, eps, momentum, mean, rstd, running_mean,
                                                          running_var);
  return check_launch("bn_finalize_kernel");
}

// eval mode: mean = running_mean, rstd = 1/sqrt(running_var + eps)
__global__ void bn_eval_stats_kernel(const float* __restrict__ rm, const float* __restrict__ rv, float eps, float* __restrict__ mean,
                                     float* __restrict__ rstd, int C) {
  pdl_sync();
  const int c = blockIdx.x * 128 + threadIdx.x;
  if (c >= C) return;
  mean[c] = rm[c];
  rstd[c] = 1.f / sqrtf(rv[c] + eps);
}
int bn_eval_stats(const float* rm, const float* rv, float eps, float* mean, float* rstd, int C, cudaStream_t stream) {
  launch_k(bn_eval_stats_kernel, dim3((C + 127) / 128), dim3(128), 0, stream, rm, rv, eps, mean, rstd, C);
  return check_launch("bn_eval_stats_kernel");
}

// counter-based dropout mask (regenerated identically in backward)
// 8 consecutive elements at once: two 64-bit hashes -> eight 16-bit uniforms (the per-element 64-bit mix above cost ~25
// instructions per element and dominated the decode-head tail kernels).  m[e] = keep ? 1/(1-p) : 0; element index `elem0`
// must be a multiple of 8.  Regenerated identically in backward.
__device__ __forceinline__ unsigned long long hash64(unsigned long long seed, unsigned long long idx) {
  unsigned long long z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ void dropout_mult8(unsigned long long seed, unsigned long long elem0, unsigned thr16, float keep_scale,
                                              float (&m)[8]) {
  const unsigned long long z0 = hash64(seed, elem0 >> 2), z1 = hash64(seed, (elem0 >> 2) + 1);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    m[e] = ((unsigned)(z0 >> (16 * e)) & 0xffffu) >= thr16 ? keep_scale : 0.f;
    m[4 + e] = ((unsigned)(z1 >> (16 * e)) & 0xffffu) >= thr16 ? keep_scale : 0.f;
  }
}
__device__ __forceinline__ unsigned dropout_thr16(float p) { return (unsigned)(p * 65536.0f); }

static constexpr int CLS_MAX_LABELS = 8;

// y = relu(bn(x)) * dropmask ; logits[b, l, pix] = sum_c y[c] * Wc[l, c] + bc[l].
// Per-channel affine (a = rstd*gamma, b = beta - mean*a) and the classifier rows are staged once per block in shared
// memory; every warp then streams rows with 16-byte loads (8 channels per lane per step).
template <int DUMMY>
__global__ void __launch_bounds__(256) bn_relu_cls_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                              const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, const float* __restrict__ wc,
                                                              const float* __restrict__ bc, float* __restrict__ logits,
                                                              long long M, int C, int NL, int pix_per_sample, float drop_p,
                                                              unsigned long long seed, int rows_per_block) {
  pdl_sync();
  extern __shared__ __align__(16) float shf[];  // a[C], b[C], wc[NL][C]
  float* sa = shf;
  float* sb = shf + C;
  float* sw = shf + 2 * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float a = rstd[c] * gamma[c];
    sa[c] = a;
    sb[c] = beta[c] - mean[c] * a;
  }
  for (int i = threadIdx.x; i < NL * C; i += 256) sw[i] = wc[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int nchunk = C >> 3;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  for (int i = warp; i < rows_per_block; i += 8) {
    const long long row = row0 + i;
    if (row >= M) break;
    float acc[CLS_MAX_LABELS];
#pragma unroll
    for (int l = 0; l < CLS_MAX_LABELS; ++l) acc[l] = 0.f;
    for (int ch = lane; ch < nchunk; ch += 32) {
      const int c = ch * 8;
      const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + c);
      const uint32_t* up = reinterpret_cast<const uint32_t*>(&u);
      float y[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float2 v = unpack_bf16x2(up[k]);
        y[2 * k] = fmaxf(fmaf(v.x, sa[c + 2 * k], sb[c + 2 * k]), 0.f);
        y[2 * k + 1] = fmaxf(fmaf(v.y, sa[c + 2 * k + 1], sb[c + 2 * k + 1]), 0.f);
      }
      if (drop_p > 0.f) {
        float dm[8];
        dropout_mult8(seed, (unsigned long long)row * C + c, dropout_thr16(drop_p), keep_scale, dm);
#pragma unroll
        for (int k = 0; k < 8; ++k) y[k] *= dm[k];
      }
#pragma unroll
      for (int l = 0; l < CLS_MAX_LABELS; ++l) {
        if (l < NL) {
          const float* wl = sw + l * C + c;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[l] = fmaf(y[k], wl[k], acc[l]);
        }
      }
    }
    const long long b = row / pix_per_sample, pix = row % pix_per_sample;
#pragma unroll
    for (int l = 0; l < CLS_MAX_LABELS; ++l) {
      if (l < NL) {
        const float sres = warp_sum(acc[l]);
        if (lane == 0) logits[(b * NL + l) * pix_per_sample + pix] = sres + bc[l];
      }
    }
  }
}

// num_labels == 1 forward: per-channel affine and classifier row in registers (each lane owns the same <= 4 chunks of 8
// channels for every row), two rows in flight per warp, no shared memory in the loop.
template <int NK, int R>   // NK chunks of 8 channels per lane (C <= 256 * NK), R rows in flight per warp
__global__ void __launch_bounds__(256) bn_relu_cls1_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                               const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ wc,
                                                               const float* __restrict__ bc, float* __restrict__ logits, long long M,
                                                               int C, float drop_p, unsigned long long seed, int rows_per_block) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nchunk = C >> 3;
  float a[NK][8], b[NK][8], w[NK][8];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int ch = lane + 32 * k;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ch * 8 + j;
      const bool ok = ch < nchunk;
      const float aa = ok ? rstd[c] * gamma[c] : 0.f;
      a[k][j] = aa;
      b[k][j] = ok ? beta[c] - mean[c] * aa : 0.f;
      w[k][j] = ok ? wc[c] : 0.f;
    }
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const unsigned thr = dropout_thr16(drop_p);
  const float bias0 = bc[0];
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  long long row_end = row0 + rows_per_block;
  if (row_end > M) row_end = M;
  for (long long rb = row0 + warp; rb < row_end; rb += 8 * R) {
    uint4 u[R][NK];
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const long long row = rb + 8 * q;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int ch = lane + 32 * k;
        u[q][k] = (row < row_end && ch < nchunk) ? __ldg(reinterpret_cast<const uint4*>(x + row * C + ch * 8)) : make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int q = 0; q < R; ++q) {
      const long long row = rb + 8 * q;
      if (row >= row_end) break;
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int ch = lane + 32 * k;
        if (ch >= nchunk) break;
        const uint32_t* up = reinterpret_cast<const uint32_t*>(&u[q][k]);
        float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (drop_p > 0.f) dropout_mult8(seed, (unsigned long long)row * C + ch * 8, thr, keep_scale, dm);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 v2 = unpack_bf16x2(up[j >> 1]);
          const float xv = (j & 1) ? v2.y : v2.x;
          const float yv = fmaxf(fmaf(xv, a[k][j], b[k][j]), 0.f) * dm[j];
          acc = fmaf(yv, w[k][j], acc);
        }
      }
      acc = warp_sum(acc);
      if (lane == 0) logits[row] = acc + bias0;   // one label: logits is [B, 1, pix] == [M]
    }
  }
}

int bn_relu_cls_fwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* wc,
                    const float* bc, float* logits, long long M, int C, int NL, int pix_per_sample, float drop_p,
                    unsigned long long seed, cudaStream_t stream) {
  SFK_REQUIRE(NL >= 1 && NL <= CLS_MAX_LABELS, "bn_relu_cls: num_labels=%d unsupported (1..%d)", NL, CLS_MAX_LABELS);
  SFK_REQUIRE(C % 8 == 0, "bn_relu_cls: C must be a multiple of 8");
  if (M == 0) return 0;
  if (NL == 1 && C <= 1024) {
    const long long slots = (long long)num_sms();   // register-heavy: one resident block per SM, one wave
    int rpb1 = (int)((M + slots - 1) / slots);
    rpb1 = (rpb1 + 31) / 32 * 32;
    const long long g1 = (M + rpb1 - 1) / rpb1;
    if (C <= 256)
      launch_k(bn_relu_cls1_fwd_kernel<1, 4>, dim3((unsigned)g1), dim3(256), 0, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
               bc, logits, M, C, drop_p, seed, rpb1);
    else if (C <= 768)
      launch_k(bn_relu_cls1_fwd_kernel<3, 4>, dim3((unsigned)g1), dim3(256), 0, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
               bc, logits, M, C, drop_p, seed, rpb1);
    else
      launch_k(bn_relu_cls1_fwd_kernel<4, 2>, dim3((unsigned)g1), dim3(256), 0, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
               bc, logits, M, C, drop_p, seed, rpb1);
    return check_launch("bn_relu_cls1_fwd_kernel");
  }
  const size_t smem = (size_t)(2 + NL) * C * sizeof(float);
  SFK_REQUIRE(smem <= 48 * 1024, "bn_relu_cls: C*(2+NL) too large for shared memory");
  int rpb = 64;
  long long grid = (M + rpb - 1) / rpb;
  const long long max_grid = (long long)num_sms() * 8;
  if (grid > max_grid) {
    rpb = (int)((M + max_grid - 1) / max_grid);
    rpb = (rpb + 7) / 8 * 8;
    grid = (M + rpb - 1) / rpb;
  }
  launch_k(bn_relu_cls_fwd_kernel<0>, dim3((unsigned)grid), dim3(256), smem, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
           bc, logits, M, C, NL, pix_per_sample, drop_p, seed, rpb);
  return check_launch("bn_relu_cls_fwd_kernel");
}

// Backward of BN(+batch statistics) -> ReLU -> Dropout -> classifier, two streaming passes over x [M, C] (bf16):
//   pass A: dyb[c] = (sum_l dlogit[l] * Wc[l,c]) * dropmask * [bn(x) > 0]     (gradient at the BN output)
//           sums[0][c] += dyb ; sums[1][c] += dyb * xhat ; dwc[l][c] += dlogit[l] * y[c] ; dbc[l] += dlogit[l]
//           (per-lane register partials over all rows of the block -> shared-memory block reduce -> slotted commit)
//   pass B: dx = gamma*rstd*(dyb - sums0/M - xhat*sums1/M)  (batch stats)   or   gamma*rstd*dyb  (running stats)
// Each lane owns the same 8-channel chunks for every row, so per-channel constants live in shared memory / registers.
static constexpr int CLS_CHUNKS = 4;  // 8-channel chunks per lane -> C <= 1024

template <int LG>  // labels whose dWc partials are kept in registers by this launch
__global__ void __launch_bounds__(256) bn_relu_cls_bwd_a_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ wc,
                                                                const float* __restrict__ dlogits, float* __restrict__ sums,
                                                                float* __restrict__ dwc, float* __restrict__ dbc, long long M, int C,
                                                                int NL, int l0, int pix_per_sample, float drop_p,
                                                                unsigned long long seed, int rows_per_block, int do_stats, int ns) {
  pdl_sync();
  extern __shared__ __align__(16) float shf[];  // mean[C], rstd[C], a[C], b[C], wc[NL][C]; later reused for the block reduce
  float* smean = shf;
  float* srstd = shf + C;
  float* sa = shf + 2 * C;
  float* sb = shf + 3 * C;
  float* sw = shf + 4 * C;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float r = rstd[c], a = r * gamma[c];
    smean[c] = mean[c]; srstd[c] = r; sa[c] = a; sb[c] = beta[c] - mean[c] * a;
  }
  for (int i = threadIdx.x; i < NL * C; i += 256) sw[i] = wc[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int nchunk = C >> 3;
  float a0[CLS_CHUNKS][8], a1[CLS_CHUNKS][8], aw[LG][CLS_CHUNKS][8], adb[LG];
#pragma unroll
  for (int k = 0; k < CLS_CHUNKS; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      a0[k][j] = 0.f; a1[k][j] = 0.f;
#pragma unroll
      for (int l = 0; l < LG; ++l) aw[l][k][j] = 0.f;
    }
#pragma unroll
  for (int l = 0; l < LG; ++l) adb[l] = 0.f;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  for (int i = warp; i < rows_per_block; i += 8) {
    const long long row = row0 + i;
    if (row >= M) break;
    const long long b = row / pix_per_sample, pix = row % pix_per_sample;
    float dl[CLS_MAX_LABELS];
#pragma unroll
    for (int l = 0; l < CLS_MAX_LABELS; ++l) dl[l] = l < NL ? dlogits[(b * NL + l) * pix_per_sample + pix] : 0.f;
#pragma unroll
    for (int l = 0; l < LG; ++l)
      if (l0 + l < NL) adb[l] += dl[l0 + l];
#pragma unroll
    for (int k = 0; k < CLS_CHUNKS; ++k) {
      const int ch = lane + 32 * k;
      if (ch >= nchunk) break;
      const int c = ch * 8;
      const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + c);
      const uint32_t* up = reinterpret_cast<const uint32_t*>(&u);
      float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (drop_p > 0.f) dropout_mult8(seed, (unsigned long long)row * C + c, dropout_thr16(drop_p), keep_scale, dm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 v2 = unpack_bf16x2(up[j >> 1]);
        const float xv = (j & 1) ? v2.y : v2.x;
        const float bn = fmaf(xv, sa[c + j], sb[c + j]);
        const float m = bn > 0.f ? dm[j] : 0.f;
        float dy = 0.f;
#pragma unroll
        for (int l = 0; l < CLS_MAX_LABELS; ++l)
          if (l < NL) dy = fmaf(dl[l], sw[l * C + c + j], dy);
        const float dyb = dy * m;
        const float xh = (xv - smean[c + j]) * srstd[c + j];
        a0[k][j] += dyb;
        a1[k][j] = fmaf(dyb, xh, a1[k][j]);
        const float yv = bn * m;  // relu(bn) * dropout multiplier (m already carries both)
#pragma unroll
        for (int l = 0; l < LG; ++l)
          if (l0 + l < NL) aw[l][k][j] = fmaf(dl[l0 + l], yv, aw[l][k][j]);
      }
    }
  }
  // block reduce: virtual columns v = arr * C + c, arr in {0: sums0, 1: sums1, 2..2+LG: dwc rows}
  __syncthreads();
  const int V = (2 + LG) * C;
  float* red = shf;  // needs V floats <= (4 + NL) * C
  for (int i = threadIdx.x; i < V; i += 256) red[i] = 0.f;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < CLS_CHUNKS; ++k) {
    const int ch = lane + 32 * k;
    if (ch >= nchunk) break;
    const int c = ch * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      atomicAdd(&red[c + j], a0[k][j]);
      atomicAdd(&red[C + c + j], a1[k][j]);
#pragma unroll
      for (int l = 0; l < LG; ++l) atomicAdd(&red[(2 + l) * C + c + j], aw[l][k][j]);
    }
  }
  __syncthreads();
  colsum_commit(red, 0, V, V, ns, blockIdx.x, gridDim.x, 0, [&](int v, float sum) {
    const int arr = v / C, c = v % C;
    if (arr == 0) { if (do_stats) sums[c] += sum; }
    else if (arr == 1) { if (do_stats) sums[C + c] += sum; }
    else if (l0 + arr - 2 < NL) dwc[(l0 + arr - 2) * C + c] += sum;
  });
  if (lane == 0) {
#pragma unroll
    for (int l = 0; l < LG; ++l)
      if (l0 + l < NL) atomicAdd(dbc + l0 + l, adb[l]);
  }
}

__global__ void __launch_bounds__(256) bn_relu_cls_bwd_b_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                                const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, const float* __restrict__ wc,
                                                                const float* __restrict__ dlogits, const float* __restrict__ sums,
                                                                __nv_bfloat16* __restrict__ dx, long long M, int C, int NL,
                                                                int pix_per_sample, float drop_p, unsigned long long seed,
                                                                int train_stats, int rows_per_block) {
  pdl_sync();
  extern __shared__ __align__(16) float shf[];  // mean, rstd, a, b, g (=gamma*rstd), k0, k1, wc[NL][C]
  float* smean = shf;
  float* srstd = shf + C;
  float* sa = shf + 2 * C;
  float* sb = shf + 3 * C;
  float* sk0 = shf + 4 * C;
  float* sk1 = shf + 5 * C;
  float* sw = shf + 6 * C;
  const float invM = 1.f / (float)M;
  for (int c = threadIdx.x; c < C; c += 256) {
    const float r = rstd[c], a = r * gamma[c];
    smean[c] = mean[c]; srstd[c] = r; sa[c] = a; sb[c] = beta[c] - mean[c] * a;
    sk0[c] = train_stats ? sums[c] * invM : 0.f;
    sk1[c] = train_stats ? sums[C + c] * invM : 0.f;
  }
  for (int i = threadIdx.x; i < NL * C; i += 256) sw[i] = wc[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const int nchunk = C >> 3;
  const long long row0 = (long long)blockIdx.x * rows_per_block;
  for (int i = warp; i < rows_per_block; i += 8) {
    const long long row = row0 + i;
    if (row >= M) break;
    const long long b = row / pix_per_sample, pix = row % pix_per_sample;
    float dl[CLS_MAX_LABELS];
#pragma unroll
    for (int l = 0; l < CLS_MAX_LABELS; ++l) dl[l] = l < NL ? dlogits[(b * NL + l) * pix_per_sample + pix] : 0.f;
    for (int ch = lane; ch < nchunk; ch += 32) {
      const int c = ch * 8;
      const uint4 u = *reinterpret_cast<const uint4*>(x + row * C + c);
      const uint32_t* up = reinterpret_cast<const uint32_t*>(&u);
      float o[8];
      float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (drop_p > 0.f) dropout_mult8(seed, (unsigned long long)row * C + c, dropout_thr16(drop_p), keep_scale, dm);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 v2 = unpack_bf16x2(up[j >> 1]);
        const float xv = (j & 1) ? v2.y : v2.x;
        const float bn = fmaf(xv, sa[c + j], sb[c + j]);
        const float m = bn > 0.f ? dm[j] : 0.f;
        float dy = 0.f;
#pragma unroll
        for (int l = 0; l < CLS_MAX_LABELS; ++l)
          if (l < NL) dy = fmaf(dl[l], sw[l * C + c + j], dy);
        const float dyb = dy * m;
        const float xh = (xv - smean[c + j]) * srstd[c + j];
        o[j] = sa[c + j] * (dyb - sk0[c + j] - xh * sk1[c + j]);
      }
      uint4 w;
      w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]);
      w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + row * C + c) = w;
    }
  }
}

// sums: f32 [2][C] workspace (zero-filled here). After the call sums[0] = dbeta, sums[1] = dgamma contributions.
// ---- num_labels == 1 fast path (the reference default): channel-slab kernels --------------------------------------
// ncu on the generic kernels above: 1.29 ms + 1.11 ms for 403 MB of input (0.3-0.7 TB/s) — five scalar shared-memory
// constant reads per element (8-way bank conflicts at a lane stride of 8 floats) and 250 registers of partials.  Here a
// block owns ONE 256-channel slab (blockIdx.x) and a row range (blockIdx.y): each lane keeps the constants of its 8
// channels in registers, a warp still reads 512 contiguous bytes per row, 4 rows are in flight per warp, and with one
// label everything pass A needs reduces to two per-channel sums:
//   g = dlogit * dropmask * [bn > 0];  S0 = sum g;  S1 = sum g * xhat
//   dbeta-sum = wc S0;  dgamma-sum = wc S1;  dwc = sum g * bn = gamma S1 + beta S0
static constexpr int CLS1_ROWS = 4;   // rows in flight per warp

struct Cls1Consts {
  float r[8], mr[8], g[8], bt[8];
};
__device__ __forceinline__ void cls1_load_consts(Cls1Consts& k, const float* mean, const float* rstd, const float* gamma,
                                                 const float* beta, int c0) {
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    k.r[j] = rstd[c0 + j];
    k.mr[j] = -mean[c0 + j] * k.r[j];
    k.g[j] = gamma[c0 + j];
    k.bt[j] = beta[c0 + j];
  }
}

__global__ void __launch_bounds__(256) bn_relu_cls1_bwd_a_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ wc,
                                                                 const float* __restrict__ dlogits, float* __restrict__ sums,
                                                                 float* __restrict__ dwc, float* __restrict__ dbc, long long M, int C,
                                                                 float drop_p, unsigned long long seed, int rows_per_block, int ns) {
  pdl_sync();
  __shared__ float sh[3 * 32 * 8];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 32 + lane) * 8;
  const bool lane_ok = c0 < C;
  float acc[3][8];   // [0] = wc*S0 (dbeta sums), [1] = wc*S1 (dgamma sums), [2] = dwc
#pragma unroll
  for (int k = 0; k < 3; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[k][j] = 0.f;
  float adb = 0.f;
  if (lane_ok) {
    Cls1Consts k;
    cls1_load_consts(k, mean, rstd, gamma, beta, c0);
    const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    const unsigned thr = dropout_thr16(drop_p);
    float s0[8], s1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { s0[j] = 0.f; s1[j] = 0.f; }
    const long long row0 = (long long)blockIdx.y * rows_per_block;
    long long row_end = row0 + rows_per_block;
    if (row_end > M) row_end = M;
    for (long long rb = row0 + warp; rb < row_end; rb += 8 * CLS1_ROWS) {
      uint4 u[CLS1_ROWS];
      float dl[CLS1_ROWS];
#pragma unroll
      for (int q = 0; q < CLS1_ROWS; ++q) {
        const long long row = rb + 8 * q;
        const bool ok = row < row_end;
        u[q] = ok ? __ldg(reinterpret_cast<const uint4*>(x + row * C + c0)) : make_uint4(0, 0, 0, 0);
        dl[q] = ok ? __ldg(dlogits + row) : 0.f;   // one label: dlogits is [B, 1, pix] == [M]
      }
#pragma unroll
      for (int q = 0; q < CLS1_ROWS; ++q) {
        const long long row = rb + 8 * q;
        if (row >= row_end) break;
        adb += dl[q];
        float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
        if (drop_p > 0.f) dropout_mult8(seed, (unsigned long long)row * C + c0, thr, keep_scale, dm);
        const uint32_t* up = reinterpret_cast<const uint32_t*>(&u[q]);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float2 v2 = unpack_bf16x2(up[j >> 1]);
          const float xv = (j & 1) ? v2.y : v2.x;
          const float xh = fmaf(xv, k.r[j], k.mr[j]);
          const float bn = fmaf(xh, k.g[j], k.bt[j]);
          const float g = bn > 0.f ? dl[q] * dm[j] : 0.f;
          s0[j] += g;
          s1[j] = fmaf(g, xh, s1[j]);
        }
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float w = wc[c0 + j];
      acc[0][j] = w * s0[j];
      acc[1][j] = w * s1[j];
      acc[2][j] = fmaf(k.g[j], s1[j], k.bt[j] * s0[j]);
    }
  }
  dw_block_commit<8, 3>(acc, sh, sums, (long long)C, dwc, c0, C, lane_ok, ns);
  if (blockIdx.x == 0 && lane == 0) atomicAdd(dbc, adb);
}

__global__ void __launch_bounds__(256) bn_relu_cls1_bwd_b_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ mean,
                                                                 const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                                 const float* __restrict__ beta, const float* __restrict__ wc,
                                                                 const float* __restrict__ dlogits, const float* __restrict__ sums,
                                                                 __nv_bfloat16* __restrict__ dx, long long M, int C, float drop_p,
                                                                 unsigned long long seed, int train_stats, int rows_per_block) {
  pdl_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * 32 + lane) * 8;
  if (c0 >= C) return;
  Cls1Consts k;
  cls1_load_consts(k, mean, rstd, gamma, beta, c0);
  float a[8], k0[8], k1[8], w[8];
  const float invM = 1.f / (float)M;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    a[j] = k.g[j] * k.r[j];
    k0[j] = train_stats ? sums[c0 + j] * invM : 0.f;
    k1[j] = train_stats ? sums[C + c0 + j] * invM : 0.f;
    w[j] = wc[c0 + j];
  }
  const float keep_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const unsigned thr = dropout_thr16(drop_p);
  const long long row0 = (long long)blockIdx.y * rows_per_block;
  long long row_end = row0 + rows_per_block;
  if (row_end > M) row_end = M;
  for (long long rb = row0 + warp; rb < row_end; rb += 8 * CLS1_ROWS) {
    uint4 u[CLS1_ROWS];
    float dl[CLS1_ROWS];
#pragma unroll
    for (int q = 0; q < CLS1_ROWS; ++q) {
      const long long row = rb + 8 * q;
      const bool ok = row < row_end;
      u[q] = ok ? __ldg(reinterpret_cast<const uint4*>(x + row * C + c0)) : make_uint4(0, 0, 0, 0);
      dl[q] = ok ? __ldg(dlogits + row) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < CLS1_ROWS; ++q) {
      const long long row = rb + 8 * q;
      if (row >= row_end) break;
      float dm[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f};
      if (drop_p > 0.f) dropout_mult8(seed, (unsigned long long)row * C + c0, thr, keep_scale, dm);
      const uint32_t* up = reinterpret_cast<const uint32_t*>(&u[q]);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float2 v2 = unpack_bf16x2(up[j >> 1]);
        const float xv = (j & 1) ? v2.y : v2.x;
        const float xh = fmaf(xv, k.r[j], k.mr[j]);
        const float bn = fmaf(xh, k.g[j], k.bt[j]);
        const float dyb = bn > 0.f ? dl[q] * w[j] * dm[j] : 0.f;
        o[j] = a[j] * (dyb - k0[j] - xh * k1[j]);
      }
      uint4 wv;
      wv.x = pack_bf16x2(o[0], o[1]); wv.y = pack_bf16x2(o[2], o[3]);
      wv.z = pack_bf16x2(o[4], o[5]); wv.w = pack_bf16x2(o[6], o[7]);
      *reinterpret_cast<uint4*>(dx + row * C + c0) = wv;
    }
  }
}

int bn_relu_cls_bwd(const void* x, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* wc,
                    const float* dlogits, float* sums, float* dwc, float* dbc, void* dx, long long M, int C, int NL,
                    int pix_per_sample, float drop_p, unsigned long long seed, int train_stats, cudaStream_t stream) {
  SFK_REQUIRE(NL >= 1 && NL <= CLS_MAX_LABELS, "bn_relu_cls_bwd: num_labels=%d unsupported", NL);
  SFK_REQUIRE(C % 8 == 0 && C <= 256 * CLS_CHUNKS, "bn_relu_cls_bwd: C=%d unsupported (multiple of 8, <= %d)", C, 256 * CLS_CHUNKS);
  if (M == 0) return 0;
  cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)2 * C * sizeof(float), stream);
  SFK_REQUIRE(e == cudaSuccess, "bn_relu_cls_bwd: memset failed: %s", cudaGetErrorString(e));
  if (NL == 1) {
    const int gx = (C / 8 + 31) / 32;
    long long gy = ((long long)num_sms() * 4 + gx - 1) / gx;   // ~4 blocks per SM, long-lived (register partials, one commit each)
    int rpb1 = (int)((M + gy - 1) / gy);
    rpb1 = (rpb1 + 8 * CLS1_ROWS - 1) / (8 * CLS1_ROWS) * (8 * CLS1_ROWS);
    gy = (M + rpb1 - 1) / rpb1;
    const int ns1 = colsum_num_slots((long long)gx * 3 * 32 * 8, gy);
    launch_k(bn_relu_cls1_bwd_a_kernel, dim3(gx, (unsigned)gy), dim3(256), 0, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
             dlogits, sums, dwc, dbc, M, C, drop_p, seed, rpb1, ns1);
    int rc = check_launch("bn_relu_cls1_bwd_a_kernel");
    if (rc) return rc;
    launch_k(bn_relu_cls1_bwd_b_kernel, dim3(gx, (unsigned)gy), dim3(256), 0, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
             dlogits, (const float*)sums, (__nv_bfloat16*)dx, M, C, drop_p, seed, train_stats, rpb1);
    return check_launch("bn_relu_cls1_bwd_b_kernel");
  }
  const long long max_grid = (long long)num_sms() * 2;   // few, long-lived blocks: register partials amortise the commit
  int rpb = (int)((M + max_grid - 1) / max_grid);
  rpb = (rpb + 7) / 8 * 8;
  if (rpb < 8) rpb = 8;
  const long long grid = (M + rpb - 1) / rpb;
  const size_t smem_a = (size_t)(4 + NL) * C * sizeof(float);
  const size_t smem_b = (size_t)(6 + NL) * C * sizeof(float);
  SFK_REQUIRE(smem_b <= 48 * 1024, "bn_relu_cls_bwd: C*(6+NL) too large for shared memory");
  for (int l0 = 0; l0 < NL; l0 += 4) {
    const int lg = (NL - l0) >= 2 || l0 > 0 ? 4 : 1;
    const int ns = colsum_num_slots((long long)(2 + lg) * C, grid);
    if (lg == 1)
      launch_k(bn_relu_cls_bwd_a_kernel<1>, dim3((unsigned)grid), dim3(256), smem_a, stream, (const __nv_bfloat16*)x, mean, rstd, gamma,
               beta, wc, dlogits, sums, dwc, dbc, M, C, NL, l0, pix_per_sample, drop_p, seed, rpb, l0 == 0 ? 1 : 0, ns);
    else
      launch_k(bn_relu_cls_bwd_a_kernel<4>, dim3((unsigned)grid), dim3(256), smem_a, stream, (const __nv_bfloat16*)x, mean, rstd, gamma,
               beta, wc, dlogits, sums, dwc, dbc, M, C, NL, l0, pix_per_sample, drop_p, seed, rpb, l0 == 0 ? 1 : 0, ns);
    int rc = check_launch("bn_relu_cls_bwd_a_kernel");
    if (rc) return rc;
  }
  int rpb_b = 64;
  long long grid_b = (M + rpb_b - 1) / rpb_b;
  const long long max_b = (long long)num_sms() * 8;
  if (grid_b > max_b) {
    rpb_b = (int)((M + max_b - 1) / max_b);
    rpb_b = (rpb_b + 7) / 8 * 8;
    grid_b = (M + rpb_b - 1) / rpb_b;
  }
  launch_k(bn_relu_cls_bwd_b_kernel, dim3((unsigned)grid_b), dim3(256), smem_b, stream, (const __nv_bfloat16*)x, mean, rstd, gamma, beta, wc,
           dlogits, (const float*)sums, (__nv_bfloat16*)dx, M, C, NL, pix_per_sample, drop_p, seed, train_stats, rpb_b);
  return check_launch("bn_relu_cls_bwd_b_kernel");
}

}  // namespace sfk
