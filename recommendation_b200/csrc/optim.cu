// Fused dense Adam / AdamW step over an embedding table (one pass over param, grad, m, v).
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
