// Teacher-student loop kernels: logits -> probabilities, Dice / BCE-valid-mask / CE-ignore / MSE losses with their
// gradients, and the pseudo-label threshold + confidence statistics.
//   SFM:117-125  bilinear x4 upsample (align_corners=False) + sigmoid
//   LOSS:16-95   diceCoeff / SegmentationLoss ('dice', 'dice_argmax', 'mse')
//   MS:853-868   in-model BCEWithLogits * valid_mask (.mean over all pixels) / CrossEntropy(ignore_index=255)
//   TS:24-56     threshold_pseudo_masks
// Reductions are two-stage (per-block partials, then one finishing block) so results are run-to-run deterministic.
#include "common.cuh"
#include "sfk_internal.h"

namespace sfk {

__device__ __forceinline__ void bil_src(int dst, float scale, int in_size, int& i0, int& i1, float& l1) {
  float src = scale * (dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  if (i0 > in_size - 1) i0 = in_size - 1;
  i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
  l1 = src - i0;
}

template <int NV>
__device__ __forceinline__ void block_reduce(float (&v)[NV], float* sh /* [NV][8] */) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    v[k] = warp_sum(v[k]);
    if (lane == 0) sh[k * 8 + warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 32) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float t = lane < (blockDim.x >> 5) ? sh[k * 8 + lane] : 0.f;
      v[k] = warp_sum(t);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// y[p, Y, X] = act(bilinear(x[p, :, :]))      planes p = B*C (NCHW, fp32)
// ------------------------------------------------------------------------------------------------
// IDX = unsigned when the tensor has fewer than 2^31 elements (every case of this path): the flat-index decode is then three
// 32-bit divisions instead of three long-long ones per element
template <typename IDX>
__global__ void __launch_bounds__(256) upsample_nchw_kernel(const float* __restrict__ x, float* __restrict__ y, long long planes,
                                                            int h, int w, int H, int W, int act) {
  pdl_sync();
  const IDX idx = (IDX)blockIdx.x * 256 + threadIdx.x;
  if ((long long)idx >= planes * H * W) return;
  const int X = (int)(idx % (IDX)W);
  const int Y = (int)((idx / (IDX)W) % (IDX)H);
  const long long p = (long long)(idx / ((IDX)W * (IDX)H));
  int y0, y1, x0, x1;
  float ly, lx;
  bil_src(Y, (float)h / (float)H, h, y0, y1, ly);
  bil_src(X, (float)w / (float)W, w, x0, x1, lx);
  const float* b = x + p * h * w;
  float v = (1.f - ly) * ((1.f - lx) * b[y0 * w + x0] + lx * b[y0 * w + x1]) +
            ly * ((1.f - lx) * b[y1 * w + x0] + lx * b[y1 * w + x1]);
  if (act == 1) v = 1.f / (1.f + expf(-v));
  y[idx] = v;
}

int upsample_nchw_fwd(const float* x, float* y, long long planes, int h, int w, int H, int W, int act, cudaStream_t stream) {
  const long long total = planes * H * W;
  if (total == 0) return 0;
  if (total < (1LL << 31))
    launch_k(upsample_nchw_kernel<unsigned>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, y, planes, h, w, H, W, act);
  else
    launch_k(upsample_nchw_kernel<long long>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, x, y, planes, h, w, H, W, act);
  return check_launch("upsample_nchw_kernel");
}

// dx[p, yi, xi] = sum_{Y,X} w(Y,X -> yi,xi) * dy[p,Y,X] * (act ? y(1-y) : 1)
template <typename IDX>
__global__ void __launch_bounds__(256) upsample_nchw_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ yact,
                                                                float* __restrict__ dx, long long planes, int h, int w, int H,
                                                                int W, int accumulate) {
  pdl_sync();
  const IDX idx = (IDX)blockIdx.x * 256 + threadIdx.x;
  if ((long long)idx >= planes * h * w) return;
  const int xi = (int)(idx % (IDX)w);
  const int yi = (int)((idx / (IDX)w) % (IDX)h);
  const long long p = (long long)(idx / ((IDX)w * (IDX)h));
  const float sy = (float)h / (float)H, sx = (float)w / (float)W;
  const float fy = (float)H / (float)h, fx = (float)W / (float)w;
  int ylo = (int)floorf((yi - 1 + 0.5f) * fy - 0.5f) - 1, yhi = (int)ceilf((yi + 1 + 0.5f) * fy - 0.5f) + 1;
  int xlo = (int)floorf((xi - 1 + 0.5f) * fx - 0.5f) - 1, xhi = (int)ceilf((xi + 1 + 0.5f) * fx - 0.5f) + 1;
  if (ylo < 0) ylo = 0;
  if (xlo < 0) xlo = 0;
  if (yhi > H - 1) yhi = H - 1;
  if (xhi > W - 1) xhi = W - 1;
  float acc = 0.f;
  for (int Y = ylo; Y <= yhi; ++Y) {
    int y0, y1; float ly;
    bil_src(Y, sy, h, y0, y1, ly);
    float wy = 0.f;
    if (y0 == yi) wy += 1.f - ly;
    if (y1 == yi) wy += ly;
    if (wy == 0.f) continue;
    for (int X = xlo; X <= xhi; ++X) {
      int x0, x1; float lx;
      bil_src(X, sx, w, x0, x1, lx);
      float wx = 0.f;
      if (x0 == xi) wx += 1.f - lx;
      if (x1 == xi) wx += lx;
      if (wx == 0.f) continue;
      const long long o = (p * H + Y) * W + X;
      float g = dy[o];
      if (yact) { const float s = yact[o]; g *= s * (1.f - s); }
      acc = fmaf(wy * wx, g, acc);
    }
  }
  dx[idx] = accumulate ? dx[idx] + acc : acc;
}

int upsample_nchw_bwd(const float* dy, const float* yact, float* dx, long long planes, int h, int w, int H, int W,
                      int accumulate, cudaStream_t stream) {
  const long long total = planes * h * w;
  if (total == 0) return 0;
  if (total < (1LL << 31) && planes * H * W < (1LL << 31))
    launch_k(upsample_nchw_bwd_kernel<unsigned>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, dy, yact, dx, planes, h, w, H, W, accumulate);
  else
    launch_k(upsample_nchw_bwd_kernel<long long>, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, stream, dy, yact, dx, planes, h, w, H, W, accumulate);
  return check_launch("upsample_nchw_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// Dice (LOSS:33-41):  per sample n: tp = sum g*p, fp = sum |p|, fn = sum g ; score = (2|tp|+s)/(fp+fn+s)
// pred stride / target stride let the caller pass channel 0 of an [N,C,H,W] tensor (LOSS:83: only channel 0 is used).
// ------------------------------------------------------------------------------------------------
static constexpr int RED_CHUNKS = 64;  // partial blocks per sample

template <typename TG>
__global__ void __launch_bounds__(256) dice_partial_kernel(const float* __restrict__ p, long long p_stride, const TG* __restrict__ g,
                                                           long long g_stride, float* __restrict__ partial, long long P,
                                                           int hard_threshold) {
  pdl_sync();
  __shared__ float sh[3 * 8];
  const int n = blockIdx.y;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per;
  long long i1 = i0 + per;
  if (i1 > P) i1 = P;
  float v[3] = {0.f, 0.f, 0.f};
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    float pv = p[n * p_stride + i];
    if (hard_threshold) pv = pv >= 0.5f ? 1.f : 0.f;
    const float gv = (float)g[n * g_stride + i];
    v[0] = fmaf(gv, pv, v[0]);
    v[1] += fabsf(pv);
    v[2] += gv;
  }
  block_reduce<3>(v, sh);
  if (threadIdx.x == 0) {
    float* o = partial + ((long long)n * gridDim.x + blockIdx.x) * 3;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2];
  }
}

// sums[n] = {tp, fp, fn};  loss = mean_n w_n * (1 - score_n), w_n = 1 without sample weights   (one block)
__global__ void dice_finish_kernel(const float* __restrict__ partial, int N, int chunks, float smooth, float* __restrict__ sums,
                                   float* __restrict__ loss, const float* __restrict__ sample_weight) {
  pdl_sync();
  __shared__ float sc[256];
  float local = 0.f;
  for (int n = threadIdx.x; n < N; n += 256) {
    float tp = 0.f, fp = 0.f, fn = 0.f;
    for (int c = 0; c < chunks; ++c) {
      const float* o = partial + ((long long)n * chunks + c) * 3;
      tp += o[0]; fp += o[1]; fn += o[2];
    }
    sums[n * 3 + 0] = tp; sums[n * 3 + 1] = fp; sums[n * 3 + 2] = fn;
    const float score = (2.f * fabsf(tp) + smooth) / (fp + fn + smooth);
    local += sample_weight != nullptr ? sample_weight[n] * (1.f - score) : score;
  }
  sc[threadIdx.x] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int i = 0; i < 256; ++i) t += sc[i];
    loss[0] = sample_weight != nullptr ? t / (float)N : 1.f - t / (float)N;
  }
}

size_t dice_workspace_bytes(int N) { return (size_t)N * RED_CHUNKS * 4 * sizeof(float); }

int dice_fwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, int N, long long P,
             int hard_threshold, float smooth, const float* sample_weight, float* workspace, float* sums, float* loss,
             cudaStream_t stream) {
  SFK_REQUIRE(N > 0 && N <= 65535 && P > 0, "dice: bad shape N=%d P=%lld", N, P);
  dim3 grid(RED_CHUNKS, N);
  if (g_is_i64)
    launch_k(dice_partial_kernel<long long>, dim3(grid), dim3(256), 0, stream, p, p_stride, (const long long*)g, g_stride, workspace, P, hard_threshold);
  else
    launch_k(dice_partial_kernel<float>, dim3(grid), dim3(256), 0, stream, p, p_stride, (const float*)g, g_stride, workspace, P, hard_threshold);
  int rc = check_launch("dice_partial_kernel");
  if (rc) return rc;
  launch_k(dice_finish_kernel, dim3(1), dim3(256), 0, stream, workspace, N, RED_CHUNKS, smooth, sums, loss, sample_weight);
  return check_launch("dice_finish_kernel");
}

// dp[n,i] = -dloss/N * (2*sign(tp)*g*F - (2|tp|+s)*sign(p)) / F^2 ,  F = fp+fn+s
template <typename TG>
__global__ void __launch_bounds__(256) dice_bwd_kernel(const float* __restrict__ p, long long p_stride, const TG* __restrict__ g,
                                                       long long g_stride, const float* __restrict__ sums,
                                                       const float* __restrict__ dloss, float smooth,
                                                       const float* __restrict__ sample_weight, float* __restrict__ dp,
                                                       long long dp_stride, int N, long long P, int accumulate) {
  pdl_sync();
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= (long long)N * P) return;
  const int n = (int)(idx / P);
  const long long i = idx - (long long)n * P;
  const float tp = sums[n * 3], F = sums[n * 3 + 1] + sums[n * 3 + 2] + smooth;
  const float num = 2.f * fabsf(tp) + smooth;
  const float stp = tp > 0.f ? 1.f : (tp < 0.f ? -1.f : 0.f);
  const float pv = p[n * p_stride + i];
  const float sp = pv > 0.f ? 1.f : (pv < 0.f ? -1.f : 0.f);
  const float gv = (float)g[n * g_stride + i];
  const float wn = sample_weight != nullptr ? sample_weight[n] : 1.f;
  const float d = -(wn * dloss[0] / (float)N) * (2.f * stp * gv * F - num * sp) / (F * F);
  float* o = dp + n * dp_stride + i;
  *o = accumulate ? *o + d : d;
}

int dice_bwd(const float* p, long long p_stride, const void* g, int g_is_i64, long long g_stride, const float* sums,
             const float* dloss, float smooth, const float* sample_weight, float* dp, long long dp_stride, int N, long long P,
             int accumulate, cudaStream_t stream) {
  const long long total = (long long)N * P;
  if (total == 0) return 0;
  const unsigned grid = (unsigned)((total + 255) / 256);
  if (g_is_i64)
    launch_k(dice_bwd_kernel<long long>, dim3(grid), dim3(256), 0, stream, p, p_stride, (const long long*)g, g_stride, sums, dloss, smooth, sample_weight,
                                                         dp, dp_stride, N, P, accumulate);
  else
    launch_k(dice_bwd_kernel<float>, dim3(grid), dim3(256), 0, stream, p, p_stride, (const float*)g, g_stride, sums, dloss, smooth, sample_weight, dp,
                                                     dp_stride, N, P, accumulate);
  return check_launch("dice_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// Generic deterministic sum of a per-element loss: kind 0 = BCE-with-logits * valid (MS:863-866),
// kind 1 = squared difference (LOSS:52).  partial[blk] = {sum, count}
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bce_partial_kernel(const float* __restrict__ z, const long long* __restrict__ labels,
                                                          long long ignore_index, float* __restrict__ partial, long long n) {
  pdl_sync();
  __shared__ float sh[2 * 8];
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per;
  long long i1 = i0 + per;
  if (i1 > n) i1 = n;
  float v[2] = {0.f, 0.f};
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const long long y = labels[i];
    if (y >= 0 && y != ignore_index) {
      const float zz = z[i];
      v[0] += fmaxf(zz, 0.f) - zz * (float)y + log1pf(expf(-fabsf(zz)));
      v[1] += 1.f;
    }
  }
  block_reduce<2>(v, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x * 2] = v[0]; partial[blockIdx.x * 2 + 1] = v[1]; }
}

// loss = sum / denom ;  denom_mode 0: fixed denominator `denom`, 1: the summed count (CE mean over non-ignored)
__global__ void sum_finish_kernel(const float* __restrict__ partial, int nblocks, int denom_mode, float denom,
                                  float* __restrict__ out /* [2]: loss, count */) {
  pdl_sync();
  if (threadIdx.x != 0) return;
  float s = 0.f, c = 0.f;
  for (int i = 0; i < nblocks; ++i) { s += partial[i * 2]; c += partial[i * 2 + 1]; }
  out[0] = denom_mode ? s / c : s / denom;
  out[1] = c;
}

static constexpr int SUM_BLOCKS = 512;
size_t loss_workspace_bytes() { return (size_t)SUM_BLOCKS * 2 * sizeof(float); }

int bce_valid_fwd(const float* z, const long long* labels, long long ignore_index, long long n, float* workspace, float* out,
                  cudaStream_t stream) {
  SFK_REQUIRE(n > 0, "bce_valid: empty input");
  launch_k(bce_partial_kernel, dim3(SUM_BLOCKS), dim3(256), 0, stream, z, labels, ignore_index, workspace, n);
  int rc = check_launch("bce_partial_kernel");
  if (rc) return rc;
  launch_k(sum_finish_kernel, dim3(1), dim3(32), 0, stream, workspace, SUM_BLOCKS, 0, (float)n, out);
  return check_launch("sum_finish_kernel");
}

__global__ void __launch_bounds__(256) bce_valid_bwd_kernel(const float* __restrict__ z, const long long* __restrict__ labels,
                                                            long long ignore_index, const float* __restrict__ dloss,
                                                            float* __restrict__ dz, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const long long y = labels[i];
  float d = 0.f;
  if (y >= 0 && y != ignore_index) d = (1.f / (1.f + expf(-z[i])) - (float)y) * dloss[0] / (float)n;
  dz[i] = d;
}
int bce_valid_bwd(const float* z, const long long* labels, long long ignore_index, const float* dloss, float* dz, long long n,
                  cudaStream_t stream) {
  if (n == 0) return 0;
  launch_k(bce_valid_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, z, labels, ignore_index, dloss, dz, n);
  return check_launch("bce_valid_bwd_kernel");
}

// CrossEntropy(ignore_index) over z [N, L, P] (NCHW), labels [N, P]
__global__ void __launch_bounds__(256) ce_partial_kernel(const float* __restrict__ z, const long long* __restrict__ labels,
                                                         long long ignore_index, float* __restrict__ partial, int N, int L,
                                                         long long P) {
  pdl_sync();
  __shared__ float sh[2 * 8];
  const long long n = (long long)N * P;
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per;
  long long i1 = i0 + per;
  if (i1 > n) i1 = n;
  float v[2] = {0.f, 0.f};
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const long long y = labels[i];
    if (y == ignore_index || y < 0 || y >= L) continue;
    const long long b = i / P, px = i - b * P;
    const float* zp = z + b * L * P + px;
    float m = -INFINITY;
    for (int l = 0; l < L; ++l) m = fmaxf(m, zp[(long long)l * P]);
    float s = 0.f;
    for (int l = 0; l < L; ++l) s += expf(zp[(long long)l * P] - m);
    v[0] += m + logf(s) - zp[y * P];
    v[1] += 1.f;
  }
  block_reduce<2>(v, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x * 2] = v[0]; partial[blockIdx.x * 2 + 1] = v[1]; }
}

int ce_ignore_fwd(const float* z, const long long* labels, long long ignore_index, int N, int L, long long P, float* workspace,
                  float* out, cudaStream_t stream) {
  SFK_REQUIRE(N > 0 && L > 0 && P > 0, "ce_ignore: empty input");
  launch_k(ce_partial_kernel, dim3(SUM_BLOCKS), dim3(256), 0, stream, z, labels, ignore_index, workspace, N, L, P);
  int rc = check_launch("ce_partial_kernel");
  if (rc) return rc;
  launch_k(sum_finish_kernel, dim3(1), dim3(32), 0, stream, workspace, SUM_BLOCKS, 1, 1.f, out);
  return check_launch("sum_finish_kernel");
}

__global__ void __launch_bounds__(256) ce_ignore_bwd_kernel(const float* __restrict__ z, const long long* __restrict__ labels,
                                                            long long ignore_index, const float* __restrict__ out /* loss,count */,
                                                            const float* __restrict__ dloss, float* __restrict__ dz, int N, int L,
                                                            long long P) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= (long long)N * P) return;
  const long long b = i / P, px = i - b * P;
  const float* zp = z + b * L * P + px;
  float* dp = dz + b * L * P + px;
  const long long y = labels[i];
  if (y == ignore_index || y < 0 || y >= L) {
    for (int l = 0; l < L; ++l) dp[(long long)l * P] = 0.f;
    return;
  }
  float m = -INFINITY;
  for (int l = 0; l < L; ++l) m = fmaxf(m, zp[(long long)l * P]);
  float s = 0.f;
  for (int l = 0; l < L; ++l) s += expf(zp[(long long)l * P] - m);
  const float scale = dloss[0] / out[1];
  for (int l = 0; l < L; ++l) dp[(long long)l * P] = (expf(zp[(long long)l * P] - m) / s - (l == y ? 1.f : 0.f)) * scale;
}
int ce_ignore_bwd(const float* z, const long long* labels, long long ignore_index, const float* out, const float* dloss,
                  float* dz, int N, int L, long long P, cudaStream_t stream) {
  const long long n = (long long)N * P;
  if (n == 0) return 0;
  launch_k(ce_ignore_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, z, labels, ignore_index, out, dloss, dz, N, L, P);
  return check_launch("ce_ignore_bwd_kernel");
}

// MSE of LOSS:44-54:  mean_n( sum((gt-pred)^2) / (N*C) )  ==  sum_all / (N*C*N)
__global__ void __launch_bounds__(256) sqdiff_partial_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                             float* __restrict__ partial, long long n) {
  pdl_sync();
  __shared__ float sh[2 * 8];
  const long long per = (n + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per;
  long long i1 = i0 + per;
  if (i1 > n) i1 = n;
  float v[2] = {0.f, 0.f};
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float d = a[i] - b[i];
    v[0] = fmaf(d, d, v[0]);
  }
  block_reduce<2>(v, sh);
  if (threadIdx.x == 0) { partial[blockIdx.x * 2] = v[0]; partial[blockIdx.x * 2 + 1] = 0.f; }
}
int mse_fwd(const float* a, const float* b, long long n, float denom, float* workspace, float* out, cudaStream_t stream) {
  SFK_REQUIRE(n > 0, "mse: empty input");
  launch_k(sqdiff_partial_kernel, dim3(SUM_BLOCKS), dim3(256), 0, stream, a, b, workspace, n);
  int rc = check_launch("sqdiff_partial_kernel");
  if (rc) return rc;
  launch_k(sum_finish_kernel, dim3(1), dim3(32), 0, stream, workspace, SUM_BLOCKS, 0, denom, out);
  return check_launch("sum_finish_kernel");
}
// da = 2(a-b)/denom * dloss ; db = -da
__global__ void __launch_bounds__(256) mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                      const float* __restrict__ dloss, float denom, float* __restrict__ da,
                                                      float* __restrict__ db, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  const float d = 2.f * (a[i] - b[i]) / denom * dloss[0];
  if (da) da[i] = d;
  if (db) db[i] = -d;
}
int mse_bwd(const float* a, const float* b, const float* dloss, float denom, float* da, float* db, long long n,
            cudaStream_t stream) {
  if (n == 0) return 0;
  launch_k(mse_bwd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, a, b, dloss, denom, da, db, n);
  return check_launch("mse_bwd_kernel");
}

// ------------------------------------------------------------------------------------------------
// threshold_pseudo_masks (TS:24-56) — fp32 comparisons against (float)thr and (float)(1-thr)
//   pseudo[n,i] = p >= thr_hi ? 1 : 0 (int64)
//   partial[n][chunk] = { sum |p|, #confident, #(p >= thr_hi), sum p*[p >= thr_hi] }
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pseudo_partial_kernel(const float* __restrict__ p, long long* __restrict__ pseudo,
                                                             float* __restrict__ partial, long long P, float thr_hi,
                                                             float thr_lo) {
  pdl_sync();
  __shared__ float sh[4 * 8];
  const int n = blockIdx.y;
  const long long per = (P + gridDim.x - 1) / gridDim.x;
  const long long i0 = (long long)blockIdx.x * per;
  long long i1 = i0 + per;
  if (i1 > P) i1 = P;
  float v[4] = {0.f, 0.f, 0.f, 0.f};
  for (long long i = i0 + threadIdx.x; i < i1; i += 256) {
    const float pv = p[(long long)n * P + i];
    const bool pos = pv >= thr_hi;
    const bool conf = pos || (pv <= thr_lo);
    pseudo[(long long)n * P + i] = pos ? 1 : 0;
    v[0] += fabsf(pv);
    v[1] += conf ? 1.f : 0.f;
    v[2] += pos ? 1.f : 0.f;
    v[3] += pos ? pv : 0.f;
  }
  block_reduce<4>(v, sh);
  if (threadIdx.x == 0) {
    float* o = partial + ((long long)n * gridDim.x + blockIdx.x) * 4;
    o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; o[3] = v[3];
  }
}

// stats[n] = { pixel_num, confidence, target_num, dice_loss_n, keep }
__global__ void pseudo_finish_kernel(const float* __restrict__ partial, int N, int chunks, long long P, float conf_thr,
                                     float min_pixels, int allow_throw, float* __restrict__ stats) {
  pdl_sync();
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float px = 0.f, conf = 0.f, pos = 0.f, tp = 0.f;
  for (int c = 0; c < chunks; ++c) {
    const float* o = partial + ((long long)n * chunks + c) * 4;
    px += o[0]; conf += o[1]; pos += o[2]; tp += o[3];
  }
  const float confidence = conf / (float)P;  // int count / numel, evaluated in fp32 like torch (TS:31)
  const float loss = 1.f - (2.f * fabsf(tp) + 1.f) / (px + pos + 1.f);
  const bool keep = !allow_throw || (px > min_pixels && confidence >= conf_thr);
  float* s = stats + n * 5;
  s[0] = px; s[1] = confidence; s[2] = pos; s[3] = loss; s[4] = keep ? 1.f : 0.f;
}

size_t pseudo_workspace_bytes(int N) { return (size_t)N * RED_CHUNKS * 4 * sizeof(float); }

int pseudo_label_threshold(const float* p, long long* pseudo, float* stats, int N, long long P, float thr_hi, float thr_lo,
                           float conf_thr, float min_pixels, int allow_throw, float* workspace, cudaStream_t stream) {
  SFK_REQUIRE(N > 0 && N <= 65535 && P > 0, "pseudo_label_threshold: bad shape N=%d P=%lld", N, P);
  launch_k(pseudo_partial_kernel, dim3(dim3(RED_CHUNKS, N)), dim3(256), 0, stream, p, pseudo, workspace, P, thr_hi, thr_lo);
  int rc = check_launch("pseudo_partial_kernel");
  if (rc) return rc;
  launch_k(pseudo_finish_kernel, dim3((N + 63) / 64), dim3(64), 0, stream, workspace, N, RED_CHUNKS, P, conf_thr, min_pixels, allow_throw, stats);
  return check_launch("pseudo_finish_kernel");
}

// denoise-label trick (TS:111-114): t = clamp(t0 + 0.2*gt - 0.1, 0, 1) with the reference op order
__global__ void __launch_bounds__(256) denoise_kernel(const float* __restrict__ t0, const float* __restrict__ gt,
                                                      float* __restrict__ out, long long n) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= n) return;
  float t = __fadd_rn(t0[i], __fmul_rn(0.2f, gt[i]));  // no FMA contraction: torch rounds the product first
  t = __fadd_rn(t, -0.1f);
  t = t > 1.f ? 1.f : t;
  t = t < 0.f ? 0.f : t;
  out[i] = t;
}
int denoise_labels(const float* t0, const float* gt, float* out, long long n, cudaStream_t stream) {
  if (n == 0) return 0;
  launch_k(denoise_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, stream, t0, gt, out, n);
  return check_launch("denoise_kernel");
}

}  // namespace sfk
