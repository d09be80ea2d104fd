// common.cuh -- shared types for libxrd (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdarg.h>
#include <stdio.h>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <utility>
#include <vector>
#include <stdexcept>
#include <string>

#include "xrd.h"

namespace xrd {

// ---------------------------------------------------------------- errors
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

[[noreturn]] inline void fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  throw Error(code, buf);
}

#define XRD_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t e__ = (expr);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      ::xrd::fail(XRD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                  __LINE__);                                                                   \
  } while (0)

#define XRD_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) ::xrd::fail(XRD_ERR_INVALID, __VA_ARGS__);     \
  } while (0)

extern std::atomic<uint64_t> g_launches;

// In-situ launch profile (xrd_profile_begin / xrd_profile_end): while it is on for the calling thread every XRD_LAUNCH is bracketed
// by two CUDA events on the launching stream, so that a kernel's duration is measured where it runs -- between its real
// neighbours, at the clocks of the real kernel mix -- and not in a loop of its own.  Eager launches only (events recorded into a
// stream capture carry no time).
struct LaunchRec { const char* name; cudaEvent_t e0, e1; };
struct LaunchProf { std::vector<LaunchRec> recs; };
extern thread_local LaunchProf* g_prof;

// Makes `dev` the calling thread's current device for the lifetime of the object and restores the previous one afterwards
// (the C entry points must not leave the caller's thread on another device).
struct DeviceScope {
  int prev = -1;
  explicit DeviceScope(int dev) {
    int cur = -1;
    XRD_CUDA(cudaGetDevice(&cur));
    if (cur != dev) { XRD_CUDA(cudaSetDevice(dev)); prev = cur; }
  }
  ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
  DeviceScope(const DeviceScope&) = delete;
  DeviceScope& operator=(const DeviceScope&) = delete;
};
// device that owns a device pointer (handle-less entry points take the device from their buffers, not from the thread state)
inline int device_of(const void* p) {
  cudaPointerAttributes a;
  XRD_CUDA(cudaPointerGetAttributes(&a, p));
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged)
    fail(XRD_ERR_INVALID, "pointer %p is not device memory (this library has no CPU path)", p);
  return a.device;
}

// Per-device launch state.  cudaFuncSetAttribute and the SM count belong to the CURRENT device, and a process may hold handles on
// several GPUs (xrd_create(device, ...); one serving thread per GPU): a process-wide "done once" flag would leave every device but
// the first at the 48 KB default and fail its first launch of a large-shared-memory kernel.
inline void ensure_dyn_smem(const void* kern, int bytes) {
  static std::mutex mu;
  static std::vector<std::pair<int, const void*>> done;       // (device, kernel) pairs whose limit was raised
  int dev = 0;
  XRD_CUDA(cudaGetDevice(&dev));
  const std::pair<int, const void*> key(dev, kern);
  std::lock_guard<std::mutex> lk(mu);
  if (std::find(done.begin(), done.end(), key) != done.end()) return;
  XRD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done.push_back(key);
}
template <typename F> inline void ensure_dyn_smem(F* kern, int bytes) { ensure_dyn_smem((const void*)kern, bytes); }
// SMs of the current device (grids of the persistent kernels are sized by it)
inline int sm_count() {
  constexpr int kMaxDev = 64;
  static std::atomic<int> cache[kMaxDev];
  int dev = 0;
  XRD_CUDA(cudaGetDevice(&dev));
  if (dev >= 0 && dev < kMaxDev) {
    const int v = cache[dev].load(std::memory_order_relaxed);
    if (v > 0) return v;
  }
  int n = 0;
  XRD_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  if (dev >= 0 && dev < kMaxDev) cache[dev].store(n, std::memory_order_relaxed);
  return n;
}

// ---------------------------------------------------------------- dtypes / tensors
enum DType : int { DT_F32 = 0, DT_BF16 = 1, DT_F16 = 2 };
inline size_t dsize(DType d) { return d == DT_F32 ? 4 : 2; }

// NHWC activation tensor
struct Tens {
  void* p = nullptr;
  int n = 0, h = 0, w = 0, c = 0;
  DType dt = DT_F32;
  size_t numel() const { return (size_t)n * h * w * c; }
  size_t bytes() const { return numel() * dsize(dt); }
  bool valid() const { return p != nullptr; }
};

// Bump allocator over one device block.  In `dry` mode only the peak is tracked
// (planning pass); kernels are not launched.
struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  bool dry = false;
  void* alloc(size_t bytes) {
    off = (off + 255) & ~(size_t)255;
    size_t o = off;
    off += bytes;
    if (off > peak) peak = off;
    if (dry) return (void*)(uintptr_t)(0x10000 + o);
    if (off > cap) fail(XRD_ERR_STATE, "arena overflow: need %zu, have %zu", off, cap);
    return base + o;
  }
  size_t mark() const { return off; }
  void release(size_t m) { off = m; }
};

// Range audit of the 16-bit activation tensors (xrd_set_range_audit): device-side counters, one block per handle.
struct RangeAudit {
  unsigned long long saturated;   // f16 elements stored at +-65504 (the saturating stores clipped them)
  unsigned long long nonfinite;   // NaN / inf elements
  unsigned long long elements;    // elements audited
  unsigned int absmax_bits;       // max |v| over finite elements (float bits; non-negative floats order like unsigned ints)
  unsigned int tensors;           // tensors audited
};

struct Ctx {
  RangeAudit* audit = nullptr;   // non-null: every 16-bit activation tensor the networks produce is scanned after its producer
  cudaStream_t s = nullptr;
  Arena* a = nullptr;
  bool dry = false;     // planning pass: no launches
  DType adt = DT_BF16;  // activation storage dtype of the current mode
  bool tc = true;       // tcgen05 contractions (false in the fp32 check mode)
  Tens alloc(int n, int h, int w, int c, DType dt) {
    Tens t;
    t.n = n; t.h = h; t.w = w; t.c = c; t.dt = dt;
    t.p = a->alloc(t.bytes());
    return t;
  }
  Tens alloc(int n, int h, int w, int c) { return alloc(n, h, w, c, adt); }
  float* allocf(size_t n) { return (float*)a->alloc(n * 4); }
  double* allocd(size_t n) { return (double*)a->alloc(n * 8); }
  // pre-zeroed pool for GroupNorm sums: one memset per network evaluation instead of one per tensor
  double* zpool = nullptr;
  size_t zpool_off = 0, zpool_cap = 0;
  double* alloc_zeroed(size_t n) {          // returns null when the pool is exhausted (caller zeroes its own buffer)
    if (!zpool || zpool_off + n > zpool_cap) return nullptr;
    double* p = zpool + zpool_off;
    zpool_off += n;
    return p;
  }
};

// launch helper: counts launches, surfaces configuration errors immediately
#define XRD_LAUNCH(ctx, kernel, grid, block, smem, ...)                                   \
  do {                                                                                    \
    if (!(ctx).dry) {                                                                     \
      ::xrd::LaunchRec* pr__ = nullptr;                                                   \
      if (::xrd::g_prof) {                                                                \
        const char* nm__ = nullptr;   /* the launched function's own (mangled) name: #kernel is "kern" inside a dispatch lambda */ \
        if (cudaFuncGetName(&nm__, (const void*)(kernel)) != cudaSuccess || !nm__) { (void)cudaGetLastError(); nm__ = #kernel; } \
        ::xrd::g_prof->recs.push_back({nm__, nullptr, nullptr});                          \
        pr__ = &::xrd::g_prof->recs.back();                                               \
        cudaEventCreate(&pr__->e0); cudaEventCreate(&pr__->e1);                           \
        cudaEventRecord(pr__->e0, (ctx).s);                                               \
      }                                                                                   \
      kernel<<<(grid), (block), (smem), (ctx).s>>>(__VA_ARGS__);                          \
      if (pr__) cudaEventRecord(pr__->e1, (ctx).s);                                       \
      ::xrd::g_launches.fetch_add(1, std::memory_order_relaxed);                          \
      cudaError_t e__ = cudaPeekAtLastError();                                            \
      if (e__ != cudaSuccess)                                                             \
        ::xrd::fail(XRD_ERR_CUDA, "launch of %s failed: %s (%s:%d)", #kernel,             \
                    cudaGetErrorString(e__), __FILE__, __LINE__);                         \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------- device load/store helpers
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <> __device__ __forceinline__ float ldf<__half>(const __half* p) { return __half2float(*p); }

template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
// f16 stores saturate (finite overflow -> +-65504) so that one large activation cannot turn into inf/NaN downstream
// torch.clamp semantics: NaN stays NaN (fminf/fmaxf would drop it).  max.NaN / min.NaN are single FMNMX.NAN instructions.
__device__ __forceinline__ float clamp_nan(float v, float lo, float hi) {
  float r;
  asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(lo));
  asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(hi));
  return r;
}
// A NaN is preserved (as the reference propagates it to the final nan_to_num, HYB:615-620); +-inf saturates like any overflow.
__device__ __forceinline__ float sat_h(float v) { return clamp_nan(v, -65504.f, 65504.f); }
// The saturating f16 conversions as ONE instruction each (F2FP.SATFINITE): round to nearest, |v| > 65504 (inf included) -> +-65504,
// NaN -> NaN.  (The first version clamped with two FMNMX per element before the conversion: a quarter of the instructions of the
// tensor-core kernels' epilogues, which ncu showed to be as loaded as the MMA issue thread.)
__device__ __forceinline__ uint32_t pack_h2_sat(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ __half f2h_sat(float v) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(v));
  return __ushort_as_half(r);
}
template <> __device__ __forceinline__ void stf<__half>(__half* p, float v) { *p = f2h_sat(v); }

// 4 consecutive elements (pointer must be aligned to 4 elements)
template <typename T> __device__ __forceinline__ void ld4(const T* p, float (&v)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void ld4<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x), b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
  float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
template <> __device__ __forceinline__ void ld4<__half>(const __half* p, float (&v)[4]) {
  uint2 t = *reinterpret_cast<const uint2*>(p);
  __half2 a = *reinterpret_cast<__half2*>(&t.x), b = *reinterpret_cast<__half2*>(&t.y);
  float2 fa = __half22float2(a), fb = __half22float2(b);
  v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const float (&v)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void st4<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[4]) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  uint2 t;
  t.x = *reinterpret_cast<uint32_t*>(&a);
  t.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = t;
}
template <> __device__ __forceinline__ void st4<__half>(__half* p, const float (&v)[4]) {
  uint2 t;
  t.x = pack_h2_sat(v[0], v[1]);
  t.y = pack_h2_sat(v[2], v[3]);
  *reinterpret_cast<uint2*>(p) = t;
}

// dispatch a lambda on a runtime dtype:  XRD_DISPATCH(dt, T, { kernel<T>... })
#define XRD_DISPATCH(dt, T, ...)                                   \
  do {                                                             \
    switch (dt) {                                                  \
      case ::xrd::DT_F32: { using T = float; __VA_ARGS__; } break; \
      case ::xrd::DT_BF16: { using T = __nv_bfloat16; __VA_ARGS__; } break; \
      case ::xrd::DT_F16: { using T = __half; __VA_ARGS__; } break;  \
    }                                                              \
  } while (0)

enum Act : int { ACT_NONE = 0, ACT_SILU = 1, ACT_GELU = 2, ACT_SIGMOID = 3, ACT_RELU = 4 };

__device__ __forceinline__ float act_apply(float v, int act) {
  switch (act) {
    case ACT_SILU: return v / (1.0f + expf(-v));
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case ACT_RELU: return v < 0.0f ? 0.0f : v;          // NaN stays NaN, like torch.relu
    default: return v;
  }
}

// nan_to_num(nan=0, posinf=1, neginf=0) followed by clamp(0,1)   (HYB:615-616)
__device__ __forceinline__ float sanitize01(float v) {
  if (v != v) return 0.0f;
  return fminf(fmaxf(v, 0.0f), 1.0f);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace xrd
