// Fused gather + BPR loss (+ squared-L2 regulariser of the gathered rows), forward and backward.
//
// Replaces the eager op chains x[idx] -> mul -> sum -> sigmoid -> log -> mean (+ norm(2).pow(2))
// of ncl.py:116-120,314-317, lightgcn.py:95-118, gcl.py:216-223, mhcn.py:35-39, diffnet.py:1110-1115
// and, in the backward, the three index_put_(accumulate=True) scatter-adds.
//
// Layout: a sub-warp of LPR = d/4 lanes owns a run of kRun consecutive triples; every lane keeps
// a float4 slice of the user / positive / negative rows.  Consecutive triples with the same user
// (the full-batch LightGCN case once the triples are in CSR order) reuse the user row and, in the
// backward, accumulate the user gradient in registers: one red.global.add.v4.f32 per run instead of
// one per triple.  The [T, d] gathered tensors of the reference are never materialised.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace gcf {

constexpr int kRun = 8;
constexpr int kBprThreads = 256;

template <int LPR, int VPL, bool GUARD>
__device__ __forceinline__ void load_row(const float* __restrict__ base, long long ld, long long row, int sl, int dvec,
                                         float4 (&r)[VPL]) {
  const float4* p = reinterpret_cast<const float4*>(base + row * ld);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = sl + k * LPR;
    r[k] = (!GUARD || idx < dvec) ? __ldg(p + idx) : f4_zero();
  }
}

// Reduce two per-lane partials (a, b) over a sub-warp of LPR lanes with log2(LPR) shuffles instead of
// 2*log2(LPR): the first butterfly step trades one value for the other, so afterwards the lower half of
// the sub-warp reduces `a` and the upper half reduces `b`.  Returns the lane's total (a for sl < LPR/2,
// b otherwise).
template <int LPR>
__device__ __forceinline__ float pair_reduce(float a, float b, int sl, unsigned mask) {
  const bool hi = (sl & (LPR / 2)) != 0;
  float keep = hi ? b : a;
  const float send = hi ? a : b;
  keep += __shfl_xor_sync(mask, send, LPR / 2);
#pragma unroll
  for (int off = LPR / 4; off > 0; off >>= 1) keep += __shfl_xor_sync(mask, keep, off);
  return keep;
}

// Pointwise loss and derivative with the fast SFU intrinsics (ex2/lg2.approx): absolute error ~1e-7 per
// term, far inside the 1e-3 parity tolerance, and ~5x fewer instructions than expf/logf/log1pf -- the forward
// kernel was issue-bound on the libm sequences (ncu r01: 102 M instructions for 1.03 M triples).
__device__ __forceinline__ void bpr_pointwise(int variant, float eps, float x, float& loss, float& dl) {
  if (variant == GCF_BPR_RAW_SCORE) {  // feature-sharded tables: emit the partial score, the loss is applied after the all-reduce
    loss = 0.f;
    dl = x;
  } else if (variant == GCF_BPR_LOG_EPS_SIGMOID) {
    const float sg = __fdividef(1.f, 1.f + __expf(-x));
    loss = -__logf(eps + sg);
    dl = -__fdividef(sg * (1.f - sg), eps + sg);
  } else {
    // -log(sigmoid(x)) = softplus(-x) = max(-x, 0) + log1p(exp(-|x|)), evaluated without overflow
    const float e = __expf(-fabsf(x));
    const float l1p = (e < 1e-3f) ? e * (1.f - 0.5f * e) : __logf(1.f + e);
    loss = fmaxf(-x, 0.f) + l1p;
    // -sigmoid(-x): for x >= 0 it is -e/(1+e), for x < 0 it is -1/(1+e)
    dl = -__fdividef(x >= 0.f ? e : 1.f, 1.f + e);
  }
}

// BATCH triples are fetched together: all index loads first, then all 3*BATCH row gathers, then the math --
// the dependent idx -> row chains of the batch overlap instead of running back to back.
template <int LPR, int VPL, bool GUARD, int BATCH>
__global__ void __launch_bounds__(kBprThreads)
bpr_fwd_kernel(const float* __restrict__ uemb, long long ldu, const float* __restrict__ iemb, long long ldi, int dvec,
               const int64_t* __restrict__ u_idx, const int64_t* __restrict__ p_idx, const int64_t* __restrict__ n_idx,
               long long n, int n_negs, int variant, float eps, float w_loss, float reg_u, float reg_p, float reg_n,
               float* __restrict__ coef_out, double* __restrict__ block_partials) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const long long group = ((long long)blockIdx.x * (kBprThreads / 32) + (threadIdx.x >> 5)) * RPW + sub;
  const long long t0 = group * kRun;
  const float inv_negs = 1.f / (float)n_negs;
  const bool x_lane = sl == 0, r_lane = sl == LPR / 2;  // pair_reduce leaves x in the low half, the reg term in the high half

  double local = 0.0;
  if (t0 < n) {
    const long long t1 = min(t0 + (long long)kRun, n);
    for (long long tb = t0; tb < t1; tb += BATCH) {
      long long u[BATCH], p[BATCH], q[BATCH];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const long long t = min(tb + b, t1 - 1);  // ragged tail: replay the last triple, result discarded
        u[b] = ld_stream_i64(u_idx + t);
        p[b] = ld_stream_i64(p_idx + t);
        q[b] = ld_stream_i64(n_idx + t * n_negs);
      }
      float4 ur[BATCH][VPL], pr[BATCH][VPL], nr[BATCH][VPL];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        load_row<LPR, VPL, GUARD>(uemb, ldu, u[b], sl, dvec, ur[b]);
        load_row<LPR, VPL, GUARD>(iemb, ldi, p[b], sl, dvec, pr[b]);
        load_row<LPR, VPL, GUARD>(iemb, ldi, q[b], sl, dvec, nr[b]);
      }
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const long long t = tb + b;
        float xs = 0.f, dn = 0.f, rs = 0.f, sn = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          xs += f4_dot(ur[b][k], pr[b][k]);
          dn += f4_dot(ur[b][k], nr[b][k]);
          rs += reg_u * f4_dot(ur[b][k], ur[b][k]) + reg_p * f4_dot(pr[b][k], pr[b][k]);
          sn += f4_dot(nr[b][k], nr[b][k]);
        }
        if (n_negs > 1 && t < t1) {  // rare path (lightgcn.py n_neg in {3,5}): remaining negatives one by one
          for (int j = 1; j < n_negs; ++j) {
            const long long qj = ld_stream_i64(n_idx + t * n_negs + j);
            float4 nj[VPL];
            load_row<LPR, VPL, GUARD>(iemb, ldi, qj, sl, dvec, nj);
#pragma unroll
            for (int k = 0; k < VPL; ++k) { dn += f4_dot(ur[b][k], nj[k]); sn += f4_dot(nj[k], nj[k]); }
          }
        }
        xs -= dn * inv_negs;
        rs += reg_n * sn;
        const float tot = pair_reduce<LPR>(xs, rs, sl, mask);
        if (t < t1) {
          if (x_lane) {
            float loss, dl;
            bpr_pointwise(variant, eps, tot, loss, dl);
            coef_out[t] = dl * w_loss;
            local += (double)(loss * w_loss);
          } else if (r_lane) {
            local += (double)tot;
          }
        }
      }
    }
  }
  // block reduction of the double partials
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  __shared__ double warp_part[kBprThreads / 32];
  if (lane == 0) warp_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kBprThreads / 32; ++w) s += warp_part[w];
    block_partials[blockIdx.x] = s;
  }
}

// deterministic final reduction: one block sums the per-block partials in a fixed order
__global__ void __launch_bounds__(256) bpr_reduce_kernel(const double* __restrict__ partials, long long n_blocks,
                                                       float* __restrict__ loss_out) {
  __shared__ double sh[256];
  double s = 0.0;
  for (long long i = threadIdx.x; i < n_blocks; i += 256) s += partials[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int off = 128; off > 0; off >>= 1) {
    if ((int)threadIdx.x < off) sh[threadIdx.x] += sh[threadIdx.x + off];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = (float)sh[0];
}

// coef[t] = w * l'(x[t]),  block partial of sum_t w * l(x[t])   (x = complete scores, e.g. after an all-reduce of the
// per-feature-slice partial scores)
__global__ void __launch_bounds__(256)
bpr_coef_kernel(const float* __restrict__ x, long long n, int variant, float eps, float w_loss, float* __restrict__ coef,
                double* __restrict__ block_partials) {
  double local = 0.0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    float loss, dl;
    bpr_pointwise(variant, eps, x[t], loss, dl);
    coef[t] = dl * w_loss;
    local += (double)(loss * w_loss);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  __shared__ double warp_part[8];
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < 8; ++w) s += warp_part[w];
    block_partials[blockIdx.x] = s;
  }
}

template <int LPR, int VPL, bool GUARD>
__device__ __forceinline__ void red_row(float* __restrict__ base, long long ld, long long row, int sl, int dvec,
                                        const float4 (&v)[VPL]) {
  float* p = base + row * ld;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = sl + k * LPR;
    if (!GUARD || idx < dvec) {
      // no "memory" clobber on purpose: gradient tables never alias the embedding tables, so
      // the compiler may hoist the next triple's gathers above these reductions.
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p + 4 * idx), "f"(v[k].x), "f"(v[k].y),
                   "f"(v[k].z), "f"(v[k].w));
    }
  }
}

template <int LPR, int VPL, bool GUARD, int BATCH>
__global__ void __launch_bounds__(kBprThreads)
bpr_bwd_kernel(const float* __restrict__ uemb, long long ldu, const float* __restrict__ iemb, long long ldi, int dvec,
               const int64_t* __restrict__ u_idx, const int64_t* __restrict__ p_idx, const int64_t* __restrict__ n_idx,
               long long n, int n_negs, const float* __restrict__ coef, const float* __restrict__ grad_out,
               float reg_u, float reg_p, float reg_n, float* __restrict__ g_user, long long ldgu,
               float* __restrict__ g_item, long long ldgi) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const long long group = ((long long)blockIdx.x * (kBprThreads / 32) + (threadIdx.x >> 5)) * RPW + sub;
  const long long t0 = group * kRun;
  if (t0 >= n) return;
  const long long t1 = min(t0 + (long long)kRun, n);
  const float g = (grad_out != nullptr) ? __ldg(grad_out) : 1.f;
  const float inv_negs = 1.f / (float)n_negs;
  const float ru2 = 2.f * reg_u * g, rp2 = 2.f * reg_p * g, rn2 = 2.f * reg_n * g;

  long long cur_u = -1;
  float4 gu[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) gu[k] = f4_zero();

  for (long long tb = t0; tb < t1; tb += BATCH) {
    long long u[BATCH], p[BATCH], q[BATCH];
    float c[BATCH];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const long long t = min(tb + b, t1 - 1);
      u[b] = ld_stream_i64(u_idx + t);
      p[b] = ld_stream_i64(p_idx + t);
      q[b] = ld_stream_i64(n_idx + t * n_negs);
      c[b] = ld_stream_f32(coef + t) * g;
    }
    float4 ur[BATCH][VPL], pr[BATCH][VPL], nr[BATCH][VPL];
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      load_row<LPR, VPL, GUARD>(uemb, ldu, u[b], sl, dvec, ur[b]);
      load_row<LPR, VPL, GUARD>(iemb, ldi, p[b], sl, dvec, pr[b]);
      load_row<LPR, VPL, GUARD>(iemb, ldi, q[b], sl, dvec, nr[b]);
    }
#pragma unroll
    for (int b = 0; b < BATCH; ++b) {
      const long long t = tb + b;
      if (t >= t1) break;
      if (u[b] != cur_u) {  // user changed: flush the register-accumulated user gradient
        if (cur_u >= 0) red_row<LPR, VPL, GUARD>(g_user, ldgu, cur_u, sl, dvec, gu);
        cur_u = u[b];
#pragma unroll
        for (int k = 0; k < VPL; ++k) gu[k] = f4_zero();
      }
      // d/du: c * (p - mean_j n_j) + 2 reg_u u ;  d/dp: c * u + 2 reg_p p ;  d/dn_j: -c/n_negs * u + 2 reg_n n_j
      const float cn = -c[b] * inv_negs;
      float4 gp[VPL], gn[VPL];
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        gp[k] = make_float4(rp2 * pr[b][k].x, rp2 * pr[b][k].y, rp2 * pr[b][k].z, rp2 * pr[b][k].w);
        f4_fma(gp[k], c[b], ur[b][k]);
        gn[k] = make_float4(rn2 * nr[b][k].x, rn2 * nr[b][k].y, rn2 * nr[b][k].z, rn2 * nr[b][k].w);
        f4_fma(gn[k], cn, ur[b][k]);
        f4_fma(gu[k], c[b], pr[b][k]);
        f4_fma(gu[k], cn, nr[b][k]);
        f4_fma(gu[k], ru2, ur[b][k]);
      }
      red_row<LPR, VPL, GUARD>(g_item, ldgi, p[b], sl, dvec, gp);
      red_row<LPR, VPL, GUARD>(g_item, ldgi, q[b], sl, dvec, gn);
      for (int j = 1; j < n_negs; ++j) {  // rare path: remaining negatives
        const long long qj = ld_stream_i64(n_idx + t * n_negs + j);
        float4 nj[VPL], gj[VPL];
        load_row<LPR, VPL, GUARD>(iemb, ldi, qj, sl, dvec, nj);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          gj[k] = make_float4(rn2 * nj[k].x, rn2 * nj[k].y, rn2 * nj[k].z, rn2 * nj[k].w);
          f4_fma(gj[k], cn, ur[b][k]);
          f4_fma(gu[k], cn, nj[k]);
        }
        red_row<LPR, VPL, GUARD>(g_item, ldgi, qj, sl, dvec, gj);
      }
    }
  }
  if (cur_u >= 0) red_row<LPR, VPL, GUARD>(g_user, ldgu, cur_u, sl, dvec, gu);
}

template <int LPR, int VPL, bool GUARD>
__device__ __forceinline__ void red_row_hint(float* __restrict__ base, long long ld, long long row, int sl, int dvec,
                                             const float4 (&v)[VPL], uint64_t pol) {
  float* p = base + row * ld;
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = sl + k * LPR;
    if (!GUARD || idx < dvec)
      asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p + 4 * idx), "f"(v[k].x),
                   "f"(v[k].y), "f"(v[k].z), "f"(v[k].w), "l"(pol));
  }
}

template <int LPR, int VPL, bool GUARD>
__device__ __forceinline__ void load_row_stream(const float* __restrict__ base, long long ld, long long row, int sl, int dvec,
                                                float4 (&r)[VPL], uint64_t pol) {
  const float4* p = reinterpret_cast<const float4*>(base + row * ld);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int idx = sl + k * LPR;
    r[k] = (!GUARD || idx < dvec) ? ldg_f4_stream(p + idx, pol) : f4_zero();
  }
}

// Forward AND backward in one pass over the triples: the three row gathers are done once, the loss derivative is
// applied on the spot (dL/dloss = grad_scale is known up front: a scalar loss, lightgcn.py:119 `loss.backward()`),
// and the gradients leave as red.global.add.v4.f32 -- the user gradient once per run of equal users.  Halves the
// gather traffic of the separate fwd + bwd kernels, which is what bounds them once the tables outgrow the L2.
// Negative rows are uniformly random (no reuse): gathered and reduced with L2 evict_first so they do not push the
// popular positive rows out of the cache.
template <int LPR, int VPL, bool GUARD, int BATCH, int MINB = 1>
__global__ void __launch_bounds__(kBprThreads, MINB)
bpr_fused_kernel(const float* __restrict__ uemb, long long ldu, const float* __restrict__ iemb, long long ldi, int dvec,
                 const int64_t* __restrict__ u_idx, const int64_t* __restrict__ p_idx, const int64_t* __restrict__ n_idx,
                 long long n, int n_negs, int variant, float eps, float w_loss, float reg_u, float reg_p, float reg_n,
                 float grad_scale, float* __restrict__ coef_out, double* __restrict__ block_partials,
                 float* __restrict__ g_user, long long ldgu, float* __restrict__ g_item, long long ldgi) {
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
  const long long group = ((long long)blockIdx.x * (kBprThreads / 32) + (threadIdx.x >> 5)) * RPW + sub;
  const long long t0 = group * kRun;
  const float inv_negs = 1.f / (float)n_negs;
  const float ru2 = 2.f * reg_u * grad_scale, rp2 = 2.f * reg_p * grad_scale, rn2 = 2.f * reg_n * grad_scale;
  const uint64_t pol_first = l2_policy_evict_first();

  double local = 0.0;
  if (t0 < n) {
    const long long t1 = min(t0 + (long long)kRun, n);
    long long cur_u = -1;
    float4 gu[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) gu[k] = f4_zero();
    float reg_part = 0.f;  // this lane's slice of the squared-norm regulariser
    for (long long tb = t0; tb < t1; tb += BATCH) {
      long long u[BATCH], p[BATCH], q[BATCH];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const long long t = min(tb + b, t1 - 1);
        u[b] = ld_stream_i64(u_idx + t);
        p[b] = ld_stream_i64(p_idx + t);
        q[b] = ld_stream_i64(n_idx + t * n_negs);
      }
      float4 ur[BATCH][VPL], pr[BATCH][VPL], nr[BATCH][VPL];
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        load_row<LPR, VPL, GUARD>(uemb, ldu, u[b], sl, dvec, ur[b]);
        load_row<LPR, VPL, GUARD>(iemb, ldi, p[b], sl, dvec, pr[b]);
        load_row_stream<LPR, VPL, GUARD>(iemb, ldi, q[b], sl, dvec, nr[b], pol_first);
      }
#pragma unroll
      for (int b = 0; b < BATCH; ++b) {
        const long long t = tb + b;
        if (t >= t1) break;
        float xs = 0.f, dn = 0.f, rs = 0.f, sn = 0.f;
        float4 nsum[VPL];  // sum of the negative rows (for the user gradient)
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          xs += f4_dot(ur[b][k], pr[b][k]);
          dn += f4_dot(ur[b][k], nr[b][k]);
          rs += reg_u * f4_dot(ur[b][k], ur[b][k]) + reg_p * f4_dot(pr[b][k], pr[b][k]);
          sn += f4_dot(nr[b][k], nr[b][k]);
          nsum[k] = nr[b][k];
        }
        for (int j = 1; j < n_negs; ++j) {  // rare path (lightgcn.py n_neg in {3,5})
          const long long qj = ld_stream_i64(n_idx + t * n_negs + j);
          float4 nj[VPL];
          load_row_stream<LPR, VPL, GUARD>(iemb, ldi, qj, sl, dvec, nj, pol_first);
#pragma unroll
          for (int k = 0; k < VPL; ++k) { dn += f4_dot(ur[b][k], nj[k]); sn += f4_dot(nj[k], nj[k]); f4_add(nsum[k], nj[k]); }
        }
        float x = xs - dn * inv_negs;
#pragma unroll
        for (int off = LPR / 2; off > 0; off >>= 1) x += __shfl_xor_sync(mask, x, off);  // every lane gets the score
        reg_part += rs + reg_n * sn;
        float loss, dl;
        bpr_pointwise(variant, eps, x, loss, dl);
        const float cf = dl * w_loss;
        if (sl == 0) {
          if (coef_out != nullptr) coef_out[t] = cf;
          local += (double)(loss * w_loss);
        }
        const float c = cf * grad_scale, cn = -c * inv_negs;
        if (u[b] != cur_u) {
          if (cur_u >= 0) red_row<LPR, VPL, GUARD>(g_user, ldgu, cur_u, sl, dvec, gu);
          cur_u = u[b];
#pragma unroll
          for (int k = 0; k < VPL; ++k) gu[k] = f4_zero();
        }
        float4 gp[VPL], gn[VPL];
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          gp[k] = make_float4(rp2 * pr[b][k].x, rp2 * pr[b][k].y, rp2 * pr[b][k].z, rp2 * pr[b][k].w);
          f4_fma(gp[k], c, ur[b][k]);
          gn[k] = make_float4(rn2 * nr[b][k].x, rn2 * nr[b][k].y, rn2 * nr[b][k].z, rn2 * nr[b][k].w);
          f4_fma(gn[k], cn, ur[b][k]);
          f4_fma(gu[k], c, pr[b][k]);
          f4_fma(gu[k], cn, nsum[k]);
          f4_fma(gu[k], ru2, ur[b][k]);
        }
        red_row<LPR, VPL, GUARD>(g_item, ldgi, p[b], sl, dvec, gp);
        red_row_hint<LPR, VPL, GUARD>(g_item, ldgi, q[b], sl, dvec, gn, pol_first);
        for (int j = 1; j < n_negs; ++j) {
          const long long qj = ld_stream_i64(n_idx + t * n_negs + j);
          float4 nj[VPL], gj[VPL];
          load_row_stream<LPR, VPL, GUARD>(iemb, ldi, qj, sl, dvec, nj, pol_first);
#pragma unroll
          for (int k = 0; k < VPL; ++k) {
            gj[k] = make_float4(rn2 * nj[k].x, rn2 * nj[k].y, rn2 * nj[k].z, rn2 * nj[k].w);
            f4_fma(gj[k], cn, ur[b][k]);
          }
          red_row_hint<LPR, VPL, GUARD>(g_item, ldgi, qj, sl, dvec, gj, pol_first);
        }
      }
    }
    if (cur_u >= 0) red_row<LPR, VPL, GUARD>(g_user, ldgu, cur_u, sl, dvec, gu);
    local += (double)reg_part;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  __shared__ double warp_part[kBprThreads / 32];
  if (lane == 0) warp_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < kBprThreads / 32; ++w) s += warp_part[w];
    block_partials[blockIdx.x] = s;
  }
}

static inline long long bpr_blocks(long long n, int lpr) {
  const long long groups_per_block = (long long)(kBprThreads / 32) * (32 / lpr);
  return cdiv(cdiv(n, kRun), groups_per_block);
}

static inline int lpr_for(int d) {
  switch (d) {
    case 8: return 2;
    case 16: return 4;
    case 32: return 8;
    case 64: return 16;
    default: return 32;
  }
}

}  // namespace gcf

using namespace gcf;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" size_t gcf_bpr_workspace_bytes(int64_t n_triples) {
  // one double per block of the forward kernel; LPR = 32 has the fewest triples per block, so it bounds the block count
  const long long blocks = bpr_blocks(n_triples > 0 ? n_triples : 1, 32);
  return align_up((size_t)blocks * sizeof(double));
}

#define GCF_BPR_DISPATCH(KERNEL, ...)                                                            \
  do {                                                                                           \
    switch (d) {                                                                                 \
      case 8:   KERNEL<2, 1, false, 4><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__); break;        \
      case 16:  KERNEL<4, 1, false, 4><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__); break;        \
      case 32:  KERNEL<8, 1, false, 4><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__); break;        \
      case 64:  KERNEL<16, 1, false, 4><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__); break;       \
      case 128: KERNEL<32, 1, false, 4><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__); break;       \
      case 256: KERNEL<32, 2, false, 2><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__); break;       \
      default:                                                                                   \
        if (d <= 128)      KERNEL<32, 1, true, 4><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__);    \
        else if (d <= 256) KERNEL<32, 2, true, 2><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__);    \
        else if (d <= 512) KERNEL<32, 4, true, 1><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__);    \
        else               KERNEL<32, 8, true, 1><<<grid, kBprThreads, 0, st>>>(__VA_ARGS__);    \
    }                                                                                            \
  } while (0)

static int bpr_check(const char* who, const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item,
                     int32_t d, const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n,
                     int32_t n_negs) {
  if (d <= 0 || (d & 3) != 0 || d > 1024) {
    set_error("%s: d=%d unsupported (need d %% 4 == 0 and d <= 1024)", who, d);
    return GCF_EUNSUPPORTED;
  }
  GCF_REQUIRE(n >= 0 && n_negs >= 1, "%s: bad n_triples / n_negs", who);
  GCF_REQUIRE(user_emb && item_emb && aligned16(user_emb) && aligned16(item_emb), "%s: null/misaligned tables", who);
  GCF_REQUIRE(ld_user >= d && ld_item >= d && (ld_user & 3) == 0 && (ld_item & 3) == 0, "%s: bad leading dims", who);
  GCF_REQUIRE(n == 0 || (u_idx && p_idx && n_idx), "%s: null index arrays", who);
  return GCF_OK;
}

extern "C" int gcf_bpr_fwd(const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item, int32_t d,
                           const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n_triples,
                           int32_t n_negs, int32_t variant, float eps, int32_t reduction, float reg_u, float reg_p,
                           float reg_n, float* loss_out, float* coef_out, void* workspace, size_t workspace_bytes,
                           gcf_stream_t stream) {
  int rc = bpr_check("gcf_bpr_fwd", user_emb, ld_user, item_emb, ld_item, d, u_idx, p_idx, n_idx, n_triples, n_negs);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(loss_out != nullptr && (n_triples == 0 || coef_out != nullptr), "gcf_bpr_fwd: null outputs");
  GCF_REQUIRE(variant == GCF_BPR_LOG_EPS_SIGMOID || variant == GCF_BPR_SOFTPLUS || variant == GCF_BPR_RAW_SCORE,
              "gcf_bpr_fwd: bad variant");
  GCF_REQUIRE(variant != GCF_BPR_RAW_SCORE || reduction == GCF_REDUCE_SUM, "gcf_bpr_fwd: raw scores need reduction = sum");
  GCF_REQUIRE(reduction == GCF_REDUCE_MEAN || reduction == GCF_REDUCE_SUM, "gcf_bpr_fwd: bad reduction");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_triples == 0) {
    GCF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    return GCF_OK;
  }
  const long long grid = bpr_blocks(n_triples, lpr_for(d));
  GCF_REQUIRE(grid < 2147483647LL, "gcf_bpr_fwd: too many triples for one launch");
  const size_t need = align_up((size_t)grid * sizeof(double));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_bpr_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  double* partials = static_cast<double*>(workspace);
  const float w_loss = (reduction == GCF_REDUCE_MEAN) ? (1.f / (float)n_triples) : 1.f;
  const int dvec = d / 4;
  GCF_BPR_DISPATCH(bpr_fwd_kernel, user_emb, ld_user, item_emb, ld_item, dvec, u_idx, p_idx, n_idx, n_triples, n_negs,
                   variant, eps, w_loss, reg_u, reg_p, reg_n, coef_out, partials);
  GCF_LAUNCH_CHECK("bpr_fwd_kernel");
  bpr_reduce_kernel<<<1, 256, 0, st>>>(partials, grid, loss_out);
  GCF_LAUNCH_CHECK("bpr_reduce_kernel");
  return GCF_OK;
}

extern "C" int gcf_bpr_bwd(const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item, int32_t d,
                           const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n_triples,
                           int32_t n_negs, const float* coef, const float* grad_out, float reg_u, float reg_p,
                           float reg_n, float* g_user, int64_t ldg_user, float* g_item, int64_t ldg_item,
                           gcf_stream_t stream) {
  int rc = bpr_check("gcf_bpr_bwd", user_emb, ld_user, item_emb, ld_item, d, u_idx, p_idx, n_idx, n_triples, n_negs);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(g_user && g_item && aligned16(g_user) && aligned16(g_item), "gcf_bpr_bwd: null/misaligned gradient tables");
  GCF_REQUIRE(ldg_user >= d && ldg_item >= d && (ldg_user & 3) == 0 && (ldg_item & 3) == 0, "gcf_bpr_bwd: bad gradient leading dims");
  if (n_triples == 0) return GCF_OK;
  GCF_REQUIRE(coef != nullptr, "gcf_bpr_bwd: null coef");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long grid = bpr_blocks(n_triples, lpr_for(d));
  GCF_REQUIRE(grid < 2147483647LL, "gcf_bpr_bwd: too many triples for one launch");
  const int dvec = d / 4;
  // d = 64 / 128: two triples in flight per sub-warp (128 registers with four) -- measured best on cfg1 and cfg5
  if (d == 64)
    bpr_bwd_kernel<16, 1, false, 2><<<grid, kBprThreads, 0, st>>>(user_emb, ld_user, item_emb, ld_item, dvec, u_idx, p_idx, n_idx, n_triples, n_negs, coef, grad_out, reg_u, reg_p, reg_n, g_user, ldg_user, g_item, ldg_item);
  else if (d == 128)
    bpr_bwd_kernel<32, 1, false, 2><<<grid, kBprThreads, 0, st>>>(user_emb, ld_user, item_emb, ld_item, dvec, u_idx, p_idx, n_idx, n_triples, n_negs, coef, grad_out, reg_u, reg_p, reg_n, g_user, ldg_user, g_item, ldg_item);
  else
    GCF_BPR_DISPATCH(bpr_bwd_kernel, user_emb, ld_user, item_emb, ld_item, dvec, u_idx, p_idx, n_idx, n_triples, n_negs,
                     coef, grad_out, reg_u, reg_p, reg_n, g_user, ldg_user, g_item, ldg_item);
  GCF_LAUNCH_CHECK("bpr_bwd_kernel");
  return GCF_OK;
}

extern "C" int gcf_bpr_fwd_bwd(const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item, int32_t d,
                               const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n_triples,
                               int32_t n_negs, int32_t variant, float eps, int32_t reduction, float reg_u, float reg_p,
                               float reg_n, float grad_scale, float* loss_out, float* coef_out, float* g_user,
                               int64_t ldg_user, float* g_item, int64_t ldg_item, void* workspace, size_t workspace_bytes,
                               gcf_stream_t stream) {
  int rc = bpr_check("gcf_bpr_fwd_bwd", user_emb, ld_user, item_emb, ld_item, d, u_idx, p_idx, n_idx, n_triples, n_negs);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(loss_out != nullptr, "gcf_bpr_fwd_bwd: null loss_out");
  GCF_REQUIRE(variant == GCF_BPR_LOG_EPS_SIGMOID || variant == GCF_BPR_SOFTPLUS, "gcf_bpr_fwd_bwd: bad variant");
  GCF_REQUIRE(reduction == GCF_REDUCE_MEAN || reduction == GCF_REDUCE_SUM, "gcf_bpr_fwd_bwd: bad reduction");
  GCF_REQUIRE(g_user && g_item && aligned16(g_user) && aligned16(g_item), "gcf_bpr_fwd_bwd: null/misaligned gradient tables");
  GCF_REQUIRE(ldg_user >= d && ldg_item >= d && (ldg_user & 3) == 0 && (ldg_item & 3) == 0, "gcf_bpr_fwd_bwd: bad gradient leading dims");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_triples == 0) {
    GCF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    return GCF_OK;
  }
  const long long grid = bpr_blocks(n_triples, lpr_for(d));
  GCF_REQUIRE(grid < 2147483647LL, "gcf_bpr_fwd_bwd: too many triples for one launch");
  const size_t need = align_up((size_t)grid * sizeof(double));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_bpr_fwd_bwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  double* partials = static_cast<double*>(workspace);
  const float w_loss = (reduction == GCF_REDUCE_MEAN) ? (1.f / (float)n_triples) : 1.f;
  const int dvec = d / 4;
#define GCF_BPR_FUSED_ARGS user_emb, ld_user, item_emb, ld_item, dvec, u_idx, p_idx, n_idx, n_triples, n_negs, variant, eps, \
                           w_loss, reg_u, reg_p, reg_n, grad_scale, coef_out, partials, g_user, ldg_user, g_item, ldg_item
  switch (d) {
    case 8:   bpr_fused_kernel<2, 1, false, 2><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS); break;
    case 16:  bpr_fused_kernel<4, 1, false, 2><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS); break;
    case 32:  bpr_fused_kernel<8, 1, false, 2><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS); break;
    // d = 64: 4 triples (12 row gathers) in flight per sub-warp at <= 128 registers, 2 CTAs / SM -- measured best of
    // {batch 1, 2, 4, 8} x {1..4 CTAs / SM} on cfg5 (29.4 -> 25.7 ms, profiles/r01_exp_variants_cfg5.log)
    case 64:  bpr_fused_kernel<16, 1, false, 4, 2><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS); break;
    case 128: bpr_fused_kernel<32, 1, false, 2><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS); break;
    case 256: bpr_fused_kernel<32, 2, false, 1><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS); break;
    default:
      if (d <= 128)      bpr_fused_kernel<32, 1, true, 2><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS);
      else if (d <= 256) bpr_fused_kernel<32, 2, true, 1><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS);
      else if (d <= 512) bpr_fused_kernel<32, 4, true, 1><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS);
      else               bpr_fused_kernel<32, 8, true, 1><<<grid, kBprThreads, 0, st>>>(GCF_BPR_FUSED_ARGS);
  }
#undef GCF_BPR_FUSED_ARGS
  GCF_LAUNCH_CHECK("bpr_fused_kernel");
  bpr_reduce_kernel<<<1, 256, 0, st>>>(partials, grid, loss_out);
  GCF_LAUNCH_CHECK("bpr_reduce_kernel");
  return GCF_OK;
}

extern "C" int gcf_bpr_coef_from_scores(const float* x, int64_t n_triples, int32_t variant, float eps, int32_t reduction,
                                        float* loss_out, float* coef_out, void* workspace, size_t workspace_bytes,
                                        gcf_stream_t stream) {
  GCF_REQUIRE(n_triples >= 0, "gcf_bpr_coef_from_scores: negative n_triples");
  GCF_REQUIRE(loss_out != nullptr, "gcf_bpr_coef_from_scores: null loss_out");
  GCF_REQUIRE(variant == GCF_BPR_LOG_EPS_SIGMOID || variant == GCF_BPR_SOFTPLUS, "gcf_bpr_coef_from_scores: bad variant");
  GCF_REQUIRE(reduction == GCF_REDUCE_MEAN || reduction == GCF_REDUCE_SUM, "gcf_bpr_coef_from_scores: bad reduction");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_triples == 0) {
    GCF_CUDA(cudaMemsetAsync(loss_out, 0, sizeof(float), st));
    return GCF_OK;
  }
  GCF_REQUIRE(x != nullptr && coef_out != nullptr, "gcf_bpr_coef_from_scores: null arrays");
  const long long grid = std::max<long long>(1, std::min<long long>(cdiv(n_triples, 256), (long long)sm_count() * 8));
  const size_t need = align_up((size_t)grid * sizeof(double));
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_bpr_coef_from_scores: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  const float w_loss = (reduction == GCF_REDUCE_MEAN) ? (1.f / (float)n_triples) : 1.f;
  bpr_coef_kernel<<<(unsigned)grid, 256, 0, st>>>(x, n_triples, variant, eps, w_loss, coef_out, static_cast<double*>(workspace));
  GCF_LAUNCH_CHECK("bpr_coef_kernel");
  bpr_reduce_kernel<<<1, 256, 0, st>>>(static_cast<double*>(workspace), grid, loss_out);
  GCF_LAUNCH_CHECK("bpr_reduce_kernel");
  return GCF_OK;
}
