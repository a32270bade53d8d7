#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long ra, rb, rc, rd;
  ra = *reinterpret_cast<unsigned long long*>(&a);
  rb = *reinterpret_cast<unsigned long long*>(&b);
  rc = *reinterpret_cast<unsigned long long*>(&c);
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
  return *reinterpret_cast<float2*>(&rd);
}
__global__ void k(const float2* a, const float2* b, float2* c, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { float2 acc = c[i]; for (int j = 0; j < 8; ++j) acc = ffma2(a[i + j], b[i + j], acc); c[i] = acc; }
}
