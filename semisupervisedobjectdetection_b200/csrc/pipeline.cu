// Callers on either side of the SegFormer hot path:
//   * the few-shot / domain-prompt cosine losses on the cls tokens (main_segformer/segFormer_fewshot_learning.py:185-188,
//     216-226: mean over the batch of nn.CosineSimilarity(dim=1) between two [B, 512] token matrices) with gradients,
//   * the training input pipeline of config.py:27-38 + archaeological_georgia_biostyle_dataloader.py:59-108 on the device:
//     uint8 HWC tile -> RandomCrop(500) -> one of {HorizontalFlip, VerticalFlip, RandomRotate90} -> Normalize(mean 0, std 255,
//     max_pixel_value 1) -> Resize(512, INTER_LINEAR) -> NCHW float32, plus the mask path (INTER_NEAREST, min-max scaling).
//     The random decisions stay on the host (per-sample parameter rows); the kernel is the deterministic part.
#include "common.cuh"
#include "sfk_internal.h"

namespace sfk {

// ------------------------------------------------------------------------------------------------
// mean_i cos(a_i, b_i),  cos = <a,b> / (max(|a|, eps) * max(|b|, eps))     (torch.nn.CosineSimilarity, eps = 1e-8)
// One block; one warp per row.  stats[i] = {<a,b>, |a|, |b|} is kept for the backward pass.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cosine_mean_fwd_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b,
                                                              long long ldb, int n, int d, float eps, float* __restrict__ stats,
                                                              float* __restrict__ out) {
  pdl_sync();
  __shared__ float s_cos[256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float local = 0.f;
  for (int i = warp; i < n; i += 8) {
    float dot = 0.f, na = 0.f, nb = 0.f;
    for (int j = lane; j < d; j += 32) {
      const float x = a[i * lda + j], y = b[i * ldb + j];
      dot = fmaf(x, y, dot);
      na = fmaf(x, x, na);
      nb = fmaf(y, y, nb);
    }
    dot = warp_sum(dot); na = sqrtf(warp_sum(na)); nb = sqrtf(warp_sum(nb));
    if (lane == 0) {
      stats[i * 3 + 0] = dot; stats[i * 3 + 1] = na; stats[i * 3 + 2] = nb;
      local += dot / (fmaxf(na, eps) * fmaxf(nb, eps));
    }
  }
  s_cos[threadIdx.x] = lane == 0 ? local : 0.f;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += s_cos[w * 32];
    out[0] = t / (float)n;
  }
}

// d mean / d a_i = (1/n) (b_i / (|a||b|) - cos_i a_i / |a|^2)   (and symmetrically for b); norms clamped like the forward
__global__ void __launch_bounds__(256) cosine_mean_bwd_kernel(const float* __restrict__ a, long long lda, const float* __restrict__ b,
                                                              long long ldb, int n, int d, float eps, const float* __restrict__ stats,
                                                              const float* __restrict__ dloss, float* __restrict__ da, long long ldda,
                                                              float* __restrict__ db, long long lddb) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)n * d) return;
  const int i = (int)(idx / d), j = (int)(idx - (long long)i * d);
  const float dot = stats[i * 3], na = fmaxf(stats[i * 3 + 1], eps), nb = fmaxf(stats[i * 3 + 2], eps);
  const float g = dloss[0] / (float)n;
  const float inv = 1.f / (na * nb), c = dot * inv;
  const float x = a[i * lda + j], y = b[i * ldb + j];
  // a norm at the clamp has no gradient through the norm term (torch clamps with max(), whose derivative is 0 there)
  const float ta = stats[i * 3 + 1] > eps ? c * x / (na * na) : 0.f;
  const float tb = stats[i * 3 + 2] > eps ? c * y / (nb * nb) : 0.f;
  if (da != nullptr) da[i * ldda + j] = g * (y * inv - ta);
  if (db != nullptr) db[i * lddb + j] = g * (x * inv - tb);
}

int cosine_mean_fwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, float* stats, float* out,
                    cudaStream_t stream) {
  SFK_REQUIRE(n > 0 && d > 0 && n <= 65536, "cosine: bad shape n=%d d=%d", n, d);
  launch_k(cosine_mean_fwd_kernel, dim3(1), dim3(256), 0, stream, a, lda, b, ldb, n, d, 1e-8f, stats, out);
  return check_launch("cosine_mean_fwd_kernel");
}
int cosine_mean_bwd(const float* a, long long lda, const float* b, long long ldb, int n, int d, const float* stats,
                    const float* dloss, float* da, long long ldda, float* db, long long lddb, cudaStream_t stream) {
  SFK_REQUIRE(n > 0 && d > 0, "cosine: bad shape n=%d d=%d", n, d);
  const long long total = (long long)n * d;
  launch_k(cosine_mean_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, a, lda, b, ldb, n, d, 1e-8f, stats,
           dloss, da, ldda, db, lddb);
  return check_launch("cosine_mean_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// Tile augmentation.  params[b] = {crop_y0, crop_x0, op, 0};  op: 0 none, 1 HorizontalFlip, 2 VerticalFlip,
// 3 + k = np.rot90(img, k), k = 0..3.  (ty, tx) indexes the transformed S x S crop; returns the source pixel in the tile.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void aug_source(int ty, int tx, int S, int op, int cy, int cx, int& sy, int& sx) {
  int y = ty, x = tx;
  switch (op) {
    case 1: x = S - 1 - tx; break;                       // cv2.flip(img, 1)
    case 2: y = S - 1 - ty; break;                       // cv2.flip(img, 0)
    case 4: y = tx; x = S - 1 - ty; break;               // np.rot90(img, 1): out[i, j] = in[j, S-1-i]
    case 5: y = S - 1 - ty; x = S - 1 - tx; break;       // k = 2
    case 6: y = S - 1 - tx; x = ty; break;               // k = 3
    default: break;                                      // 0 and 3 (k = 0): identity
  }
  sy = cy + y;
  sx = cx + x;
}

// cv2.resize(INTER_LINEAR) source coordinate (resize.cpp: fx = (dx + 0.5) * scale - 0.5; sx = floor(fx); fx -= sx; clamped)
__device__ __forceinline__ void cv_linear_coord(int dst, double scale, int ssize, int& s0, int& s1, float& w1) {
  float f = (float)((dst + 0.5) * scale - 0.5);
  int s = (int)floorf(f);
  f -= (float)s;
  if (s < 0) { s = 0; f = 0.f; }
  if (s >= ssize - 1) { s = ssize - 1; f = 0.f; }
  s0 = s;
  s1 = s + (s < ssize - 1 ? 1 : 0);
  w1 = f;
}

__global__ void __launch_bounds__(256) tile_augment_kernel(const uint8_t* __restrict__ src, long long batch_stride, long long row_stride,
                                                           int pix_stride, const int* __restrict__ params, int S, float* __restrict__ out,
                                                           int B, int OH, int OW) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * OH * OW) return;
  const int ox = (int)(idx % OW), oy = (int)((idx / OW) % OH), b = (int)(idx / ((long long)OW * OH));
  const int cy = params[b * 4 + 0], cx = params[b * 4 + 1], op = params[b * 4 + 2];
  int y0, y1, x0, x1;
  float wy, wx;
  cv_linear_coord(oy, (double)S / (double)OH, S, y0, y1, wy);
  cv_linear_coord(ox, (double)S / (double)OW, S, x0, x1, wx);
  const uint8_t* base = src + b * batch_stride;
  int sy[4], sx[4];
  aug_source(y0, x0, S, op, cy, cx, sy[0], sx[0]);
  aug_source(y0, x1, S, op, cy, cx, sy[1], sx[1]);
  aug_source(y1, x0, S, op, cy, cx, sy[2], sx[2]);
  aug_source(y1, x1, S, op, cy, cx, sy[3], sx[3]);
  const float inv255 = 1.0f / 255.0f;   // albumentations normalize(): img *= np.reciprocal(std * max_pixel_value) in float32
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    float v[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) v[t] = (float)base[sy[t] * row_stride + (long long)sx[t] * pix_stride + c] * inv255;
    // horizontal pass first, then vertical (the order of cv2's HResize / VResize)
    const float top = v[0] * (1.f - wx) + v[1] * wx;
    const float bot = v[2] * (1.f - wx) + v[3] * wx;
    out[(((long long)b * 3 + c) * OH + oy) * OW + ox] = top * (1.f - wy) + bot * wy;
  }
}

// mask path: INTER_NEAREST (sx = min(floor(dx * scale), S - 1)) of channel 0, written as float; per-sample {min, max} by atomics
// on an int view (values are 0..255, so integer atomics order correctly)
__global__ void __launch_bounds__(256) mask_augment_kernel(const uint8_t* __restrict__ src, long long batch_stride, long long row_stride,
                                                           int pix_stride, const int* __restrict__ params, int S, float* __restrict__ out,
                                                           int* __restrict__ minmax, int B, int OH, int OW) {
  pdl_sync();
  __shared__ int s_min, s_max;
  if (threadIdx.x == 0) { s_min = 255; s_max = 0; }
  __syncthreads();
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix < OH * OW) {
    const int ox = pix % OW, oy = pix / OW;
    int ty = (int)floor(oy * ((double)S / (double)OH)), tx = (int)floor(ox * ((double)S / (double)OW));
    if (ty > S - 1) ty = S - 1;
    if (tx > S - 1) tx = S - 1;
    int sy, sx;
    aug_source(ty, tx, S, params[b * 4 + 2], params[b * 4 + 0], params[b * 4 + 1], sy, sx);
    const int v = src[b * batch_stride + sy * row_stride + (long long)sx * pix_stride];
    out[((long long)b * OH + oy) * OW + ox] = (float)v;
    atomicMin(&s_min, v);
    atomicMax(&s_max, v);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    atomicMin(&minmax[b * 2 + 0], s_min);
    atomicMax(&minmax[b * 2 + 1], s_max);
  }
}

__global__ void mask_minmax_init_kernel(int* __restrict__ minmax, int B) {
  pdl_sync();
  const int b = blockIdx.x * 256 + threadIdx.x;
  if (b < B) { minmax[b * 2 + 0] = 255; minmax[b * 2 + 1] = 0; }
}

// (m - min) / (max - min) per sample (dataloader :85-86; a constant mask gives 0/0 = NaN there as well)
__global__ void __launch_bounds__(256) mask_minmax_scale_kernel(float* __restrict__ m, const int* __restrict__ minmax, int B, long long P) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)B * P) return;
  const int b = (int)(idx / P);
  const float lo = (float)minmax[b * 2 + 0], hi = (float)minmax[b * 2 + 1];
  m[idx] = (m[idx] - lo) / (hi - lo);
}

int tile_augment_u8(const void* src, long long batch_stride, long long row_stride, int pix_stride, int Hs, int Ws, const int* params,
                    int crop, float* out, int B, int OH, int OW, const void* mask, long long m_batch_stride, long long m_row_stride,
                    int m_pix_stride, float* mask_out, int* minmax, cudaStream_t stream) {
  SFK_REQUIRE(B > 0 && OH > 0 && OW > 0 && crop > 0 && crop <= Hs && crop <= Ws && pix_stride >= 3,
              "tile_augment: bad shape (tile %dx%d, crop %d, pixel stride %d)", Hs, Ws, crop, pix_stride);
  SFK_REQUIRE(src != nullptr && out != nullptr && params != nullptr, "tile_augment: null pointer");
  const long long total = (long long)B * OH * OW;
  launch_k(tile_augment_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, (const uint8_t*)src, batch_stride,
           row_stride, pix_stride, params, crop, out, B, OH, OW);
  int rc = check_launch("tile_augment_kernel");
  if (rc || mask == nullptr) return rc;
  SFK_REQUIRE(mask_out != nullptr && minmax != nullptr && m_pix_stride >= 1, "tile_augment: mask output / workspace missing");
  launch_k(mask_minmax_init_kernel, dim3((unsigned)((B + 255) / 256)), dim3(256), 0, stream, minmax, B);
  rc = check_launch("mask_minmax_init_kernel");
  if (rc) return rc;
  launch_k(mask_augment_kernel, dim3((unsigned)((OH * OW + 255) / 256), (unsigned)B), dim3(256), 0, stream, (const uint8_t*)mask,
           m_batch_stride, m_row_stride, m_pix_stride, params, crop, mask_out, minmax, B, OH, OW);
  rc = check_launch("mask_augment_kernel");
  if (rc) return rc;
  launch_k(mask_minmax_scale_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, mask_out, (const int*)minmax, B,
           (long long)OH * OW);
  return check_launch("mask_minmax_scale_kernel");
}

}  // namespace sfk
