// Optimiser-side bandwidth kernels over flat fp32 buffers:
//   clip_grad_value_(1.2) + torch.optim.Adam(betas=(0.5,0.999), L2 weight decay)  (SFM:32-35,158-162), fused with the
//   bf16 shadow-weight refresh;  teacher <- d*teacher + (1-d)*student  (EMA; d = 0 reproduces the hard reset of TS:202-204).
#include "common.cuh"
#include "sfk_internal.h"

namespace sfk {

__global__ void __launch_bounds__(256) clip_adam_l2_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                           float* __restrict__ v, __nv_bfloat16* __restrict__ w16, long long n,
                                                           float lr, float beta1, float beta2, float eps, float wd, float clip,
                                                           float bc1, float bc2_sqrt) {
  pdl_sync();
  const long long i4 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n) {
    float4 pv = *reinterpret_cast<float4*>(p + i4);
    const float4 gv = *reinterpret_cast<const float4*>(g + i4);
    float4 mv = *reinterpret_cast<float4*>(m + i4);
    float4 vv = *reinterpret_cast<float4*>(v + i4);
    float* pp = reinterpret_cast<float*>(&pv);
    const float* gp = reinterpret_cast<const float*>(&gv);
    float* mp = reinterpret_cast<float*>(&mv);
    float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = fminf(fmaxf(gp[k], -clip), clip);
      gg = fmaf(wd, pp[k], gg);
      mp[k] = beta1 * mp[k] + (1.f - beta1) * gg;
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gg * gg;
      const float denom = sqrtf(vp[k]) / bc2_sqrt + eps;
      pp[k] -= (lr / bc1) * (mp[k] / denom);
    }
    *reinterpret_cast<float4*>(p + i4) = pv;
    *reinterpret_cast<float4*>(m + i4) = mv;
    *reinterpret_cast<float4*>(v + i4) = vv;
    if (w16) {  // refresh the bf16 GEMM shadow of the updated weights in the same pass
      uint2 u;
      u.x = pack_bf16x2(pp[0], pp[1]);
      u.y = pack_bf16x2(pp[2], pp[3]);
      *reinterpret_cast<uint2*>(w16 + i4) = u;
    }
  } else {
    for (long long i = i4; i < n; ++i) {
      float gg = fminf(fmaxf(g[i], -clip), clip);
      gg = fmaf(wd, p[i], gg);
      m[i] = beta1 * m[i] + (1.f - beta1) * gg;
      v[i] = beta2 * v[i] + (1.f - beta2) * gg * gg;
      p[i] -= (lr / bc1) * (m[i] / (sqrtf(v[i]) / bc2_sqrt + eps));
      if (w16) w16[i] = __float2bfloat16(p[i]);
    }
  }
}

int clip_adam_l2(float* p, const float* g, float* m, float* v, void* w16, long long n, float lr, float beta1, float beta2,
                 float eps, float wd, float clip, int step, cudaStream_t stream) {
  SFK_REQUIRE(step >= 1, "clip_adam_l2: step must be >= 1");
  SFK_REQUIRE(((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                reinterpret_cast<uintptr_t>(v)) & 15) == 0, "clip_adam_l2: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step);
  const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
  const long long threads = (n + 3) / 4;
  launch_k(clip_adam_l2_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, stream, p, g, m, v, (__nv_bfloat16*)w16, n, lr, beta1, beta2, eps,
                                                                             wd, clip, bc1, bc2_sqrt);
  return check_launch("clip_adam_l2_kernel");
}

__global__ void __launch_bounds__(256) ema_lerp_kernel(float* __restrict__ t, const float* __restrict__ s,
                                                       __nv_bfloat16* __restrict__ w16, long long n, float decay) {
  pdl_sync();
  const long long i4 = ((long long)blockIdx.x * 256 + threadIdx.x) * 4;
  if (i4 >= n) return;
  if (i4 + 4 <= n) {
    float4 tv = *reinterpret_cast<float4*>(t + i4);
    const float4 sv = *reinterpret_cast<const float4*>(s + i4);
    tv.x = decay * tv.x + (1.f - decay) * sv.x;
    tv.y = decay * tv.y + (1.f - decay) * sv.y;
    tv.z = decay * tv.z + (1.f - decay) * sv.z;
    tv.w = decay * tv.w + (1.f - decay) * sv.w;
    *reinterpret_cast<float4*>(t + i4) = tv;
    if (w16) {
      uint2 u;
      u.x = pack_bf16x2(tv.x, tv.y);
      u.y = pack_bf16x2(tv.z, tv.w);
      *reinterpret_cast<uint2*>(w16 + i4) = u;
    }
  } else {
    for (long long i = i4; i < n; ++i) {
      t[i] = decay * t[i] + (1.f - decay) * s[i];
      if (w16) w16[i] = __float2bfloat16(t[i]);
    }
  }
}

int ema_lerp(float* t, const float* s, void* w16, long long n, float decay, cudaStream_t stream) {
  SFK_REQUIRE(((reinterpret_cast<uintptr_t>(t) | reinterpret_cast<uintptr_t>(s)) & 15) == 0,
              "ema_lerp: buffers must be 16-byte aligned");
  if (n == 0) return 0;
  const long long threads = (n + 3) / 4;
  launch_k(ema_lerp_kernel, dim3((unsigned)((threads + 255) / 256)), dim3(256), 0, stream, t, s, (__nv_bfloat16*)w16, n, decay);
  return check_launch("ema_lerp_kernel");
}

}  // namespace sfk
