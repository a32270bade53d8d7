// Host-side plumbing shared by all ops: thread-local error string, launch checks, SM count,
// TMA descriptor encoding through the runtime-resolved driver entry point.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "sfk_internal.h"
#include "../../include/sfk.h"

namespace sfk {

// auxiliary streams created by layer.cu: kernels launched on them use the second column-sum slot bank
static std::mutex g_aux_streams_mu;
static std::vector<cudaStream_t> g_aux_streams;
void register_aux_stream(cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_aux_streams_mu);
  g_aux_streams.push_back(s);
}
int colsum_bank(cudaStream_t s) {
  std::lock_guard<std::mutex> lock(g_aux_streams_mu);
  for (cudaStream_t a : g_aux_streams)
    if (a == s) return 1;
  return 0;
}

static thread_local char g_err[1024] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};

int check_launch(const char* what) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
    return -2;
  }
  return 0;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e == cudaSuccess && qres == cudaDriverEntryPointSuccess) fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t row_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  SFK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  SFK_REQUIRE(box_inner * 2 == 128, "tensor map: inner box must be 128 bytes for SWIZZLE_128B");
  SFK_REQUIRE(box_outer >= 1 && box_outer <= 256, "tensor map: outer box %u out of range", box_outer);
  SFK_REQUIRE(row_stride_bytes % 16 == 0, "tensor map: row stride %llu not a multiple of 16 bytes",
              (unsigned long long)row_stride_bytes);
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu)",
              (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
  return 0;
}

int make_tmap_2d_any(CUtensorMap* out, const void* base, int elem_bytes, uint64_t inner, uint64_t outer,
                     uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  SFK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  SFK_REQUIRE(elem_bytes == 2 || elem_bytes == 4, "tensor map: element size %d unsupported", elem_bytes);
  SFK_REQUIRE((int)box_inner * elem_bytes == swizzle_bytes && (swizzle_bytes == 64 || swizzle_bytes == 128),
              "tensor map: inner box (%u x %d B) must equal the swizzle span (%d B)", box_inner, elem_bytes, swizzle_bytes);
  SFK_REQUIRE(box_outer >= 1 && box_outer <= 256, "tensor map: outer box %u out of range", box_outer);
  SFK_REQUIRE(row_stride_bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(base) & 15) == 0,
              "tensor map: base / row stride must be multiples of 16 bytes");
  cuuint64_t gdim[2] = {inner, outer};
  cuuint64_t gstride[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d any) failed with CUresult %d (inner=%llu outer=%llu stride=%llu)",
              (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes);
  return 0;
}

// rank-5 tiled map (bf16, SWIZZLE_128B, inner box = 64 elements): used for the k == stride patch view of a channels-last map
int make_tmap_5d_bf16(CUtensorMap* out, const void* base, const uint64_t dims[5], const uint64_t strides_bytes[4],
                      const uint32_t box[5]) {
  PFN_encodeTiled enc = get_encode();
  SFK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  SFK_REQUIRE(box[0] * 2 == 128, "tensor map: inner box must be 128 bytes for SWIZZLE_128B");
  SFK_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base must be 16-byte aligned");
  cuuint64_t gdim[5], gstride[4];
  cuuint32_t bx[5], estr[5];
  for (int i = 0; i < 5; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; SFK_REQUIRE(box[i] >= 1 && box[i] <= 256, "tensor map: box out of range"); }
  for (int i = 0; i < 4; ++i) {
    gstride[i] = strides_bytes[i];
    SFK_REQUIRE(strides_bytes[i] % 16 == 0, "tensor map: stride %llu not a multiple of 16 bytes", (unsigned long long)strides_bytes[i]);
  }
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), gdim, gstride, bx, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(5d) failed with CUresult %d", (int)r);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                      uint64_t stride2_bytes, uint32_t box0, uint32_t box1, uint32_t box2, int swizzle_bytes) {
  PFN_encodeTiled enc = get_encode();
  SFK_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available (no CUDA driver?)");
  SFK_REQUIRE((int)box0 * 2 == swizzle_bytes, "tensor map: inner box (%u elems) must equal the swizzle span (%d B)", box0,
              swizzle_bytes);
  SFK_REQUIRE(box1 >= 1 && box1 <= 256 && box2 >= 1 && box2 <= 256, "tensor map: box out of range");
  SFK_REQUIRE(stride1_bytes % 16 == 0 && stride2_bytes % 16 == 0, "tensor map: strides must be multiples of 16 bytes");
  SFK_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map: base must be 16-byte aligned");
  CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                        : swizzle_bytes == 64  ? CU_TENSOR_MAP_SWIZZLE_64B
                                               : CU_TENSOR_MAP_SWIZZLE_32B;
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {stride1_bytes, stride2_bytes};
  cuuint32_t box[3] = {box0, box1, box2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  SFK_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed with CUresult %d (dims %llu,%llu,%llu strides %llu,%llu)",
              (int)r, (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
              (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes);
  return 0;
}

// ---- live per-launch timing of the dominant kernel (bench.py roofline): CUDA events around every GEMM launch ----
struct ProfState {
  bool on = false;
  std::vector<cudaEvent_t> ev;   // pairs
  size_t used = 0;
  double flops = 0.0;
  long long launches = 0;
  std::vector<long long> shapes;  // M, N, K, flags per launch
};
static ProfState g_prof;

void prof_gemm_shape(long long M, long long N, long long K, long long flags) {
  if (!g_prof.on) return;
  g_prof.shapes.push_back(M); g_prof.shapes.push_back(N); g_prof.shapes.push_back(K); g_prof.shapes.push_back(flags);
}

void prof_gemm_begin(double flops, cudaStream_t st) {
  if (!g_prof.on) return;
  if (g_prof.used + 2 > g_prof.ev.size()) {
    for (int i = 0; i < 2; ++i) {
      cudaEvent_t e;
      if (cudaEventCreate(&e) != cudaSuccess) { g_prof.on = false; return; }
      g_prof.ev.push_back(e);
    }
  }
  cudaEventRecord(g_prof.ev[g_prof.used], st);
  g_prof.flops += flops;
  g_prof.launches += 1;
}
void prof_gemm_end(cudaStream_t st) {
  if (!g_prof.on) return;
  cudaEventRecord(g_prof.ev[g_prof.used + 1], st);
  g_prof.used += 2;
}

}  // namespace sfk

extern "C" {

long long sfk_launch_count(void) { return sfk::g_launches.load(); }

int sfk_profile_gemm_begin(void) {
  sfk::g_prof.on = true;
  sfk::g_prof.used = 0;
  sfk::g_prof.flops = 0.0;
  sfk::g_prof.launches = 0;
  sfk::g_prof.shapes.clear();
  return 0;
}

/* After sfk_profile_gemm_end: write one CSV line per profiled GEMM launch (M,N,K,flags,microseconds). */
int sfk_profile_gemm_dump(const char* path) {
  FILE* f = fopen(path, "w");
  if (!f) { sfk::set_error("profile: cannot open %s", path); return -1; }
  fprintf(f, "M,N,K,flags,us\n");
  size_t n = sfk::g_prof.used / 2;
  for (size_t i = 0; i < n && 4 * i + 3 < sfk::g_prof.shapes.size(); ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, sfk::g_prof.ev[2 * i], sfk::g_prof.ev[2 * i + 1]);
    fprintf(f, "%lld,%lld,%lld,%lld,%.3f\n", sfk::g_prof.shapes[4 * i], sfk::g_prof.shapes[4 * i + 1], sfk::g_prof.shapes[4 * i + 2],
            sfk::g_prof.shapes[4 * i + 3], t * 1000.f);
  }
  fclose(f);
  return 0;
}

/* Synchronises the device, sums the per-launch GEMM durations recorded since sfk_profile_gemm_begin. */
int sfk_profile_gemm_end(double* total_ms, double* total_flops, long long* launches) {
  sfk::g_prof.on = false;
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { sfk::set_error("profile: sync failed: %s", cudaGetErrorString(e)); return -1; }
  double ms = 0.0;
  for (size_t i = 0; i + 1 < sfk::g_prof.used; i += 2) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, sfk::g_prof.ev[i], sfk::g_prof.ev[i + 1]) == cudaSuccess) ms += t;
  }
  *total_ms = ms;
  *total_flops = sfk::g_prof.flops;
  *launches = sfk::g_prof.launches;
  return 0;
}


const char* sfk_last_error(void) { return sfk::g_err; }

int sfk_version(void) { return SFK_VERSION; }

int sfk_device_check(void) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) {
    sfk::set_error("no CUDA device: %s", cudaGetErrorString(e));
    return -1;
  }
  int major = 0, minor = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (major != 10) {
    sfk::set_error("device compute capability %d.%d is not sm_100; this library has no other code path", major, minor);
    return -1;
  }
  return 0;
}

}  // extern "C"
