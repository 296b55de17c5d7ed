// Shared helpers for libgcf (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cmath>
#include <cstdint>
#include <cstddef>
#include "../../include/gcf.h"

namespace gcf {

void set_error(const char* fmt, ...);
int sm_count();

static inline int64_t cdiv(int64_t a, int64_t b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base; size_t cap; size_t off;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), cap(n), off(0) {}
  template <typename T> T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (base == nullptr || off + bytes > cap) { off = cap + 1; return nullptr; }
    T* r = reinterpret_cast<T*>(base + off);
    off += bytes;
    return r;
  }
  bool ok() const { return off <= cap; }
};

}  // namespace gcf

#define GCF_REQUIRE(cond, ...)                                  \
  do {                                                          \
    if (!(cond)) { gcf::set_error(__VA_ARGS__); return GCF_EINVAL; } \
  } while (0)

#define GCF_LAUNCH_CHECK(name)                                                  \
  do {                                                                          \
    cudaError_t e_ = cudaGetLastError();                                        \
    if (e_ != cudaSuccess) {                                                    \
      gcf::set_error("%s: CUDA error: %s", name, cudaGetErrorString(e_));       \
      return GCF_ECUDA;                                                         \
    }                                                                           \
  } while (0)

#define GCF_CUDA(call)                                                          \
  do {                                                                          \
    cudaError_t e_ = (call);                                                    \
    if (e_ != cudaSuccess) {                                                    \
      gcf::set_error("%s failed: %s", #call, cudaGetErrorString(e_));           \
      return GCF_ECUDA;                                                         \
    }                                                                           \
  } while (0)

// ---- device helpers --------------------------------------------------------------------
namespace gcf {

// Streaming loads of read-only index / value arrays: bypass L1 allocation, NOT volatile so that the
// compiler may hoist them and overlap the dependent row gathers of consecutive iterations.
__device__ __forceinline__ int ld_stream_i32(const int32_t* p) {
  int v;
  asm("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_stream_f32(const float* p) {
  float v;
  asm("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ long long ld_stream_i64(const int64_t* p) {
  long long v;
  asm("ld.global.nc.L1::no_allocate.s64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
// 128-bit vector reduction (sm_90+): one L2 atomic transaction per 16 bytes.
__device__ __forceinline__ void red_add_v4(float* addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
// ---- L2 residency control (createpolicy + .L2::cache_hint) -------------------------------------
// Working sets far beyond the 126 MB L2 (cfg5: 3.84 GB of embedding rows) are gather-bound on DRAM; what can be
// saved is the traffic of the hub rows, which the default replacement policy lets the cold stream evict.  Hub
// gathers are tagged evict_last, everything that is touched once (cold rows, CSR arrays, outputs) evict_first.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ float4 ldg_f4_hint(const float4* p, uint64_t pol) {
  float4 v;
  asm("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
  return v;
}
// read-once rows: no L1 allocation, first in line for L2 eviction
__device__ __forceinline__ float4 ldg_f4_stream(const float4* p, uint64_t pol_first) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
      : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol_first));
  return v;
}
// L1 residency by eviction priority (the SM's L1 and shared memory are one SRAM array): rows of hub columns are kept,
// everything else passes through without allocating a line, so the L1 behaves as a demand-filled hub-row cache.
__device__ __forceinline__ float4 ldg_f4_l1_keep(const float4* p) {
  float4 v;
  asm("ld.global.nc.L1::evict_last.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_f4_l1_bypass(const float4* p) {
  float4 v;
  asm("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ int ld_stream_i32_hint(const int32_t* p, uint64_t pol) {
  int v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ float ld_stream_f32_hint(const float* p, uint64_t pol) {
  float v;
  asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
  return v;
}
__device__ __forceinline__ void st_f4_hint(float4* p, const float4& v, uint64_t pol) {
  asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ void f4_fma(float4& a, float s, const float4& x) {
  a.x = fmaf(s, x.x, a.x); a.y = fmaf(s, x.y, a.y); a.z = fmaf(s, x.z, a.z); a.w = fmaf(s, x.w, a.w);
}
__device__ __forceinline__ void f4_add(float4& a, const float4& x) { a.x += x.x; a.y += x.y; a.z += x.z; a.w += x.w; }
__device__ __forceinline__ float f4_dot(const float4& a, const float4& b) {
  return fmaf(a.x, b.x, fmaf(a.y, b.y, fmaf(a.z, b.z, a.w * b.w)));
}

// ---- Adam (torch.optim.Adam single-tensor update order; see optim.cu) --------------------------
struct AdamArgs {
  float one_minus_b1, b2, one_minus_b2, eps, wd, decay_mul, step_size, bc2_sqrt;
  int decoupled;
};

__device__ __forceinline__ void adam_update(float& p, float g, float& m, float& v, const AdamArgs& a) {
  if (a.decoupled) p *= a.decay_mul;
  else if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  m = fmaf(g - m, a.one_minus_b1, m);
  v = fmaf(a.one_minus_b2 * g, g, a.b2 * v);
  const float denom = sqrtf(v) / a.bc2_sqrt + a.eps;
  p = fmaf(-a.step_size, m / denom, p);
}

// host: hyper-parameters + 1-based step count -> the constants of one update
inline AdamArgs make_adam_args(float lr, float beta1, float beta2, float eps, float weight_decay, int decoupled, long long step) {
  AdamArgs a;
  a.one_minus_b1 = 1.f - beta1;
  a.b2 = beta2;
  a.one_minus_b2 = 1.f - beta2;
  a.eps = eps;
  a.wd = weight_decay;
  a.decoupled = decoupled ? 1 : 0;
  a.decay_mul = 1.f - lr * weight_decay;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  a.step_size = (float)((double)lr / bc1);
  a.bc2_sqrt = (float)sqrt(bc2);
  return a;
}

}  // namespace gcf
