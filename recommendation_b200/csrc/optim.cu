// Fused dense optimiser steps over an embedding table: Adam / AdamW (one pass over param, grad, m, v), SGD with momentum
// (selfcf.py:544, directau.py:214) and a row-sparse Adam for mini-batch models.
// Replaces torch.optim.Adam(...).step() (ncl.py:305,329, selfcf.py:542, directau.py:212, lightgcn.py:80),
// same update order as torch's single-tensor implementation:
//   g' = g + wd*p (Adam) | p *= 1 - lr*wd (AdamW);  m.lerp_(g', 1-b1);  v = b2*v + (1-b2)*g'^2;
//   p -= (lr / (1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t) + eps)
#include "common.cuh"
#include <algorithm>
#include <cmath>

namespace gcf {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ m, float* __restrict__ v,
            long long n, AdamArgs a) {
  const long long n4 = n >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(param);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  float4* m4 = reinterpret_cast<float4*>(m);
  float4* v4 = reinterpret_cast<float4*>(v);
  for (long long i = tid; i < n4; i += nth) {
    float4 p = p4[i], g = g4[i], mm = m4[i], vv = v4[i];
    adam_update(p.x, g.x, mm.x, vv.x, a);
    adam_update(p.y, g.y, mm.y, vv.y, a);
    adam_update(p.z, g.z, mm.z, vv.z, a);
    adam_update(p.w, g.w, mm.w, vv.w, a);
    p4[i] = p; m4[i] = mm; v4[i] = vv;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) adam_update(param[i], grad[i], m[i], v[i], a);
}

// torch.optim.SGD single-tensor update (selfcf.py:544, directau.py:214: momentum = 0.9, dampening 0, no Nesterov):
//   g' = g + wd*p;  buf = g' on the first step, momentum*buf + (1-dampening)*g' afterwards;
//   g'' = g' + momentum*buf (Nesterov) | buf;  p -= lr*g''
struct SgdArgs { float lr, momentum, one_minus_damp, wd; int nesterov, first; };

__device__ __forceinline__ void sgd_update(float& p, float g, float& buf, const SgdArgs& a) {
  if (a.wd != 0.f) g = fmaf(a.wd, p, g);
  if (a.momentum != 0.f) {
    buf = a.first ? g : fmaf(a.momentum, buf, a.one_minus_damp * g);
    g = a.nesterov ? fmaf(a.momentum, buf, g) : buf;
  }
  p = fmaf(-a.lr, g, p);
}

__global__ void __launch_bounds__(256)
sgd_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ buf, long long n, SgdArgs a) {
  const long long n4 = n >> 2;
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nth = (long long)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(param);
  const float4* g4 = reinterpret_cast<const float4*>(grad);
  float4* b4 = reinterpret_cast<float4*>(buf);
  const bool mom = a.momentum != 0.f;
  for (long long i = tid; i < n4; i += nth) {
    float4 p = p4[i], g = g4[i], b = f4_zero();
    if (mom && !a.first) b = b4[i];
    sgd_update(p.x, g.x, b.x, a);
    sgd_update(p.y, g.y, b.y, a);
    sgd_update(p.z, g.z, b.z, a);
    sgd_update(p.w, g.w, b.w, a);
    p4[i] = p;
    if (mom) b4[i] = b;
  }
  for (long long i = (n4 << 2) + tid; i < n; i += nth) {
    float b = (mom && !a.first) ? buf[i] : 0.f;
    sgd_update(param[i], grad[i], b, a);
    if (mom) buf[i] = b;
  }
}

// Row-sparse ("lazy") Adam: only the listed rows of the table are touched -- gradient rows are given densely in list
// order.  Rows must be distinct (the caller de-duplicates and sums, e.g. with gcf_scatter_add_rows into a compact
// buffer).  One sub-warp of d/4 lanes per row.
__global__ void __launch_bounds__(256)
adam_rows_kernel(float* __restrict__ param, long long ld, const float* __restrict__ grad_rows, long long ldg,
                 float* __restrict__ m, float* __restrict__ v, const int64_t* __restrict__ rows, long long n_rows, int dvec,
                 AdamArgs a) {
  const long long total = n_rows * dvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / dvec;
    const int c = (int)(i - r * dvec);
    const long long row = ld_stream_i64(rows + r);
    float4* p4 = reinterpret_cast<float4*>(param + row * ld) + c;
    float4* m4 = reinterpret_cast<float4*>(m + row * ld) + c;
    float4* v4 = reinterpret_cast<float4*>(v + row * ld) + c;
    const float4 g = __ldg(reinterpret_cast<const float4*>(grad_rows + r * ldg) + c);
    float4 p = *p4, mm = *m4, vv = *v4;
    adam_update(p.x, g.x, mm.x, vv.x, a);
    adam_update(p.y, g.y, mm.y, vv.y, a);
    adam_update(p.z, g.z, mm.z, vv.z, a);
    adam_update(p.w, g.w, mm.w, vv.w, a);
    *p4 = p; *m4 = mm; *v4 = vv;
  }
}

// x *= *g, skipped entirely (no traffic) when *g == 1 -- the usual `loss.backward()` seed.
__global__ void __launch_bounds__(256)
scale_by_device_scalar_kernel(float4* __restrict__ x, long long n4, float* __restrict__ tail, int n_tail,
                              const float* __restrict__ g) {
  const float s = __ldg(g);
  if (s == 1.f) return;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 v = x[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x[i] = v;
  }
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail[threadIdx.x] *= s;
}

}  // namespace gcf

using namespace gcf;

extern "C" int gcf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int32_t decoupled, int64_t step,
                             gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0 && step >= 1, "gcf_adam_step: bad n / step");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(param && grad && exp_avg && exp_avg_sq, "gcf_adam_step: null pointers");
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  GCF_REQUIRE(a16(param) && a16(grad) && a16(exp_avg) && a16(exp_avg_sq), "gcf_adam_step: pointers must be 16B aligned");
  const AdamArgs a = make_adam_args(lr, beta1, beta2, eps, weight_decay, decoupled, step);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n / 4 + 1, 256), (long long)sm_count() * 16));
  adam_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, exp_avg, exp_avg_sq, n, a);
  GCF_LAUNCH_CHECK("adam_kernel");
  return GCF_OK;
}

extern "C" int gcf_sgd_momentum_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr,
                                     float momentum, float dampening, float weight_decay, int32_t nesterov,
                                     int32_t first_step, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0, "gcf_sgd_momentum_step: negative n");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(param && grad, "gcf_sgd_momentum_step: null pointers");
  GCF_REQUIRE(momentum == 0.f || momentum_buf != nullptr, "gcf_sgd_momentum_step: momentum needs its buffer");
  GCF_REQUIRE(!nesterov || (momentum > 0.f && dampening == 0.f),
              "gcf_sgd_momentum_step: Nesterov momentum requires a momentum and zero dampening (torch.optim.SGD)");
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  GCF_REQUIRE(a16(param) && a16(grad) && a16(momentum_buf), "gcf_sgd_momentum_step: pointers must be 16B aligned");
  const SgdArgs a{lr, momentum, 1.f - dampening, weight_decay, nesterov ? 1 : 0, first_step ? 1 : 0};
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n / 4 + 1, 256), (long long)sm_count() * 16));
  sgd_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, grad, momentum_buf, n, a);
  GCF_LAUNCH_CHECK("sgd_kernel");
  return GCF_OK;
}

extern "C" int gcf_adam_rows_step(float* param, int64_t ld, const float* grad_rows, int64_t ld_grad, float* exp_avg,
                                  float* exp_avg_sq, const int64_t* rows, int64_t n_rows, int32_t d, float lr, float beta1,
                                  float beta2, float eps, float weight_decay, int32_t decoupled, int64_t step,
                                  gcf_stream_t stream) {
  GCF_REQUIRE(n_rows >= 0 && step >= 1, "gcf_adam_rows_step: bad n_rows / step");
  if (n_rows == 0) return GCF_OK;
  GCF_REQUIRE(param && grad_rows && exp_avg && exp_avg_sq && rows, "gcf_adam_rows_step: null pointers");
  GCF_REQUIRE(d > 0 && (d & 3) == 0 && ld >= d && (ld & 3) == 0 && ld_grad >= d && (ld_grad & 3) == 0,
              "gcf_adam_rows_step: d, ld and ld_grad must be multiples of 4 with ld, ld_grad >= d");
  auto a16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  GCF_REQUIRE(a16(param) && a16(grad_rows) && a16(exp_avg) && a16(exp_avg_sq), "gcf_adam_rows_step: pointers must be 16B aligned");
  const AdamArgs a = make_adam_args(lr, beta1, beta2, eps, weight_decay, decoupled, step);
  const long long total = n_rows * (d / 4);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(total, 256), (long long)sm_count() * 16));
  adam_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(param, ld, grad_rows, ld_grad, exp_avg, exp_avg_sq,
                                                                          rows, n_rows, d / 4, a);
  GCF_LAUNCH_CHECK("adam_rows_kernel");
  return GCF_OK;
}

extern "C" int gcf_scale_by_device_scalar(float* x, int64_t n, const float* g, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0, "gcf_scale_by_device_scalar: negative n");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(x != nullptr && g != nullptr, "gcf_scale_by_device_scalar: null pointers");
  GCF_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15u) == 0, "gcf_scale_by_device_scalar: x must be 16B aligned");
  const long long n4 = n / 4;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(gcf::cdiv(std::max<long long>(n4, 1), 256), (long long)gcf::sm_count() * 8));
  gcf::scale_by_device_scalar_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<float4*>(x), n4, x + n4 * 4, (int)(n - n4 * 4), g);
  GCF_LAUNCH_CHECK("scale_by_device_scalar_kernel");
  return GCF_OK;
}
