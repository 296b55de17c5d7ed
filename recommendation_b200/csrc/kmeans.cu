// Lloyd's k-means on the device for NCL's E-step (ncl.py:339-356: faiss.Kmeans(d, k).train(x) + index.search(x, 1), run by
// the reference inside EVERY batch, ncl.py:313,324).
//
// Assignment = arg-min_c |x - c|^2 = arg-max_c (<x, c> - |c|^2 / 2): the N x k inner products are a dense contraction and
// run on the tcgen05 tensor cores, same pipeline as the InfoNCE kernels (csrc/infonce.cu): a CTA owns 128 points, the
// centroid tiles (256 rows) stream through a TMA ring, one thread issues tcgen05.mma into a double-buffered TMEM
// accumulator and eight epilogue warps read it back with tcgen05.ld and keep a running (best score, best index) per point.
// Nothing of the N x k distance matrix is ever written.
//
// Numerics: bf16 tensor-core products alone (8 mantissa bits) would flip near-ties, so both operands are split into
// bf16 hi + lo parts and the product is formed as  xh.ch + xh.cl + xl.ch  -- three K-slabs of ONE accumulation
// (K = 3 d_pad) -- which carries ~16 mantissa bits (dropped term xl.cl ~ 2^-18 relative); |c|^2 and |x|^2 are exact fp32.
// That is the accuracy class of faiss's fp32 sgemm distance kernel.  Ties go to the lowest centroid index.
//
// Update: stable radix sort of (cluster, point) -> each cluster's members contiguous in ascending point order; one CTA per
// cluster sums its rows in a fixed order (deterministic, no atomics) and divides by the count.  Empty clusters are
// re-seeded on the device (faiss: a slightly perturbed copy of a large cluster's centroid): no host synchronisation
// anywhere in the loop.
#include "common.cuh"
#include "radix.cuh"
#include "tc05.cuh"
#include <cuda_bf16.h>
#include <algorithm>

namespace gcf {

using namespace tc;

constexpr int kKmTileM = 128;     // points per CTA (= TMEM lanes)
constexpr int kKmTileN = 256;     // centroids per MMA tile (= TMEM columns per accumulator stage)
constexpr int kKmChunk = 64;      // bf16 elements per 128-byte swizzled row
constexpr int kKmStages = 4;      // TMA ring depth
constexpr int kKmThreads = 320;   // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue
constexpr int kKmABytes = kKmTileM * 128;   // one [128 x 64] bf16 chunk
constexpr int kKmBBytes = kKmTileN * 128;   // one [256 x 64] bf16 chunk
constexpr float kKmSplitEps = 1.f / 1024.f; // faiss ClusteringParameters: EPS of split_clusters

// rows -> K-concatenated bf16 split [P0 | P1 | P2]: points (hi, hi, lo), centroids (hi, lo, hi); sq[row] = |row|^2 (fp32).
// Rows beyond n are zero; on the centroid side their |c|^2 is +inf so that a padded column can never win the arg-max.
__global__ void __launch_bounds__(256)
km_prep_kernel(const float* __restrict__ x, long long ld, long long n, int d, int dp, long long n_pad, int centroid_side,
               __nv_bfloat16* __restrict__ out, float* __restrict__ sq) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_pad) return;
  __nv_bfloat16* o = out + row * (3LL * dp);
  const bool live = row < n;
  float ss = 0.f;
  for (int c = lane; c < dp; c += 32) {
    const float v = (live && c < d) ? x[row * ld + c] : 0.f;
    const __nv_bfloat16 hi = __float2bfloat16(v);
    const __nv_bfloat16 lo = __float2bfloat16(v - __bfloat162float(hi));
    o[c] = hi;
    o[dp + c] = centroid_side ? lo : hi;
    o[2 * dp + c] = centroid_side ? hi : lo;
    ss = fmaf(v, v, ss);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
  if (lane == 0) sq[row] = live ? ss : (centroid_side ? INFINITY : 0.f);
}

// KC > 0: the points' KC chunks stay resident in shared memory; KC == 0: they ride in the ring with the centroid chunks.
template <int KC>
__global__ void __launch_bounds__(kKmThreads, 1)
km_assign_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, int n_tiles,
                 int tiles_per_split, int kc_rt, const float* __restrict__ c_sq, float* __restrict__ part_v,
                 int* __restrict__ part_i, long long m_pad) {
  constexpr bool kStreamA = KC == 0;
  const int n_kc = kStreamA ? kc_rt : KC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + (kStreamA ? kKmStages : KC) * kKmABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kKmStages * kKmBBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kKmStages;
  uint64_t* a_bar = bars + 2 * kKmStages;
  uint64_t* acc_full = a_bar + 1;
  uint64_t* acc_empty = acc_full + 2;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, split = blockIdx.y;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(t_begin + tiles_per_split, n_tiles);
  const int my_tiles = max(t_end - t_begin, 0);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_a);
    tma_prefetch_desc(&tm_b);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < kKmStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
      mbar_init(a_bar, 1);
      for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 2 * kKmTileN);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    if (lane == 0 && my_tiles > 0) {
      if (!kStreamA) {
        mbar_expect_tx(a_bar, KC * kKmABytes);
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(&tm_a, a_bar, smem_a + kc * kKmABytes, kc * kKmChunk, m_tile * kKmTileM);
      }
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_expect_tx(full_bar + stage, kKmBBytes + (kStreamA ? kKmABytes : 0));
          if (kStreamA) tma_load_2d(&tm_a, full_bar + stage, smem_a + stage * kKmABytes, kc * kKmChunk, m_tile * kKmTileM);
          tma_load_2d(&tm_b, full_bar + stage, smem_b + stage * kKmBBytes, kc * kKmChunk, t * kKmTileN);
          if (++stage == kKmStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(kKmTileM, kKmTileN, 0, 0);
      if (!kStreamA) {
        mbar_wait(a_bar, 0);
        fence_after_sync();
      }
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        mbar_wait(acc_empty + acc, ((it >> 1) & 1) ^ 1);
        fence_after_sync();
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait(full_bar + stage, phase);
          fence_after_sync();
          const uint32_t a_addr = smem_u32(smem_a + (kStreamA ? stage : kc) * kKmABytes);
          const uint32_t b_addr = smem_u32(smem_b + stage * kKmBBytes);
#pragma unroll
          for (int kk = 0; kk < kKmChunk / 16; ++kk) {
            const uint64_t da = smem_desc_sw128(a_addr + kk * 32, 0, 1024);
            const uint64_t db = smem_desc_sw128(b_addr + kk * 32, 0, 1024);
            umma_bf16(tmem_base + acc * kKmTileN, da, db, idesc, (kc | kk) != 0);
          }
          umma_commit(empty_bar + stage);
          if (++stage == kKmStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full + acc);
      }
    }
  } else {
    // ===== epilogue: thread = one point x one half of the tile's centroids; running arg-max of <x, c> - |c|^2 / 2 =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    float best = -INFINITY;
    int best_i = 0x7fffffff;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const int t = t_begin + it;
      mbar_wait(acc_full + acc, (it >> 1) & 1);
      fence_after_sync();
      const int col0 = t * kKmTileN + half * (kKmTileN / 2);
#pragma unroll 1
      for (int c = 0; c < kKmTileN / 64; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kKmTileN + half * (kKmTileN / 2) + c * 32), v);
        const float4* cs4 = reinterpret_cast<const float4*>(c_sq + col0 + c * 32);   // same address in every lane: broadcast
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 q = __ldg(cs4 + j4);
          const float qq[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const float s = fmaf(-0.5f, qq[u], v[4 * j4 + u]);
            if (s > best) { best = s; best_i = col0 + c * 32 + 4 * j4 + u; }   // ascending columns + strict '>': lowest index wins ties
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
    }
    const long long out = ((long long)split * 2 + half) * m_pad + (long long)m_tile * kKmTileM + row;
    part_v[out] = best;
    part_i[out] = best_i;
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 2 * kKmTileN);
  }
}

// partial (score, index) pairs -> assignment + squared distance (exact ties between partials: lowest centroid index)
__global__ void __launch_bounds__(256)
km_combine_kernel(const float* __restrict__ part_v, const int* __restrict__ part_i, int n_parts, long long m_pad, long long n,
                  const float* __restrict__ x_sq, uint32_t* __restrict__ assign_u32, int64_t* __restrict__ assign_i64,
                  float* __restrict__ dist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int p = 0; p < n_parts; ++p) {
    const float v = part_v[p * m_pad + i];
    const int vi = part_i[p * m_pad + i];
    if (v > best || (v == best && vi < bi)) { best = v; bi = vi; }
  }
  if (assign_u32 != nullptr) assign_u32[i] = (uint32_t)bi;
  if (assign_i64 != nullptr) assign_i64[i] = (int64_t)bi;
  if (dist != nullptr) dist[i] = fmaxf(fmaf(-2.f, best, x_sq[i]), 0.f);
}

// One CTA per cluster: its members are sorted_pos[s .. e) (ascending point id); groups of d/4 lanes sum every G-th member,
// the G partial sums are added in group order.  centroid = sum / count; an empty cluster keeps its centroid (fixed below).
__global__ void __launch_bounds__(256)
km_centroid_kernel(const float* __restrict__ x, long long ld, int dvec, const uint32_t* __restrict__ sorted_keys,
                   const uint32_t* __restrict__ sorted_pos, long long n, float* __restrict__ centroids,
                   int* __restrict__ counts, int* __restrict__ n_empty) {
  __shared__ long long seg[2];
  __shared__ float4 partial[256];
  const int c = blockIdx.x;
  if (threadIdx.x < 2) {
    const uint32_t key = (uint32_t)c + threadIdx.x;   // lower_bound(c), lower_bound(c + 1)
    long long lo = 0, hi = n;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (sorted_keys[mid] < key) lo = mid + 1; else hi = mid;
    }
    seg[threadIdx.x] = lo;
  }
  __syncthreads();
  const long long s = seg[0], e = seg[1];
  const int cnt = (int)(e - s);
  if (threadIdx.x == 0) {
    counts[c] = cnt;
    if (cnt == 0) atomicAdd(n_empty, 1);
  }
  if (cnt == 0) return;
  const int groups = 256 / dvec;
  const int g = threadIdx.x / dvec, v = threadIdx.x - g * dvec;
  float4 acc = f4_zero();
  if (g < groups) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    const long long ld4 = ld >> 2;
    for (long long j = s + g; j < e; j += groups) f4_add(acc, __ldg(x4 + (long long)sorted_pos[j] * ld4 + v));
    partial[threadIdx.x] = acc;
  }
  __syncthreads();
  if (g == 0) {
    float4 tot = partial[v];
    for (int q = 1; q < groups; ++q) f4_add(tot, partial[q * dvec + v]);
    const float inv = 1.f / (float)cnt;
    reinterpret_cast<float4*>(centroids + (long long)c * dvec * 4)[v] = make_float4(tot.x * inv, tot.y * inv, tot.z * inv, tot.w * inv);
  }
}

// faiss split_clusters, deterministic flavour (the reference's own draw cannot be reproduced): the j-th empty cluster (by
// index) takes centroid * (1 + eps) of the j-th largest cluster (ties: lower index), which keeps centroid * (1 - eps).
// Single CTA; returns at once when no cluster is empty.
__global__ void __launch_bounds__(1024)
km_fix_empty_kernel(float* __restrict__ centroids, int d, int k, const int* __restrict__ counts, int* __restrict__ n_empty) {
  extern __shared__ int cnt_s[];            // k counts (consumed: a chosen donor is marked -1)
  __shared__ int red_v[32], red_i[32];
  __shared__ int donor_s, empty_s;
  const int ne = *n_empty;
  if (ne == 0) return;
  for (int i = threadIdx.x; i < k; i += blockDim.x) cnt_s[i] = counts[i];
  __syncthreads();
  int next_empty = 0;                       // scan position for the j-th empty cluster (all threads agree)
  for (int j = 0; j < ne; ++j) {
    // j-th empty cluster by index
    if (threadIdx.x == 0) {
      int q = next_empty;
      while (q < k && counts[q] != 0) ++q;
      empty_s = q;
    }
    // largest remaining cluster, ties to the lower index
    int bv = -1, bi = 0x7fffffff;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
      const int cv = cnt_s[i];
      if (cv > bv || (cv == bv && i < bi)) { bv = cv; bi = i; }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const int ov = __shfl_xor_sync(0xffffffffu, bv, off), oi = __shfl_xor_sync(0xffffffffu, bi, off);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { red_v[threadIdx.x >> 5] = bv; red_i[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      int v0 = red_v[0], i0 = red_i[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (red_v[w] > v0 || (red_v[w] == v0 && red_i[w] < i0)) { v0 = red_v[w]; i0 = red_i[w]; }
      donor_s = i0;
      cnt_s[i0] = -1;
    }
    __syncthreads();
    const int donor = donor_s, em = empty_s;
    next_empty = em + 1;
    if (em < k && donor != em) {
      for (int c = threadIdx.x; c < d; c += blockDim.x) {
        const float cv = centroids[(long long)donor * d + c];
        centroids[(long long)em * d + c] = cv * (1.f + kKmSplitEps);
        centroids[(long long)donor * d + c] = cv * (1.f - kKmSplitEps);
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) *n_empty = 0;
}

struct KmPlan {
  long long a_pad, b_pad;
  int dp, n_kc, m_tiles, n_tiles, n_splits, tiles_per_split;
};

static KmPlan km_plan(long long n, int k, int d) {
  KmPlan p;
  p.dp = (int)((d + kKmChunk - 1) / kKmChunk * kKmChunk);
  p.n_kc = 3 * p.dp / kKmChunk;
  p.a_pad = (std::max<long long>(n, 1) + kKmTileM - 1) / kKmTileM * kKmTileM;
  p.b_pad = ((long long)std::max(k, 1) + kKmTileN - 1) / kKmTileN * kKmTileN;
  p.m_tiles = (int)(p.a_pad / kKmTileM);
  p.n_tiles = (int)(p.b_pad / kKmTileN);
  const int splits = std::max(1, std::min(p.n_tiles, sm_count() / std::max(p.m_tiles, 1)));
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

struct KmWs {
  __nv_bfloat16 *ab, *bb;
  float *x_sq, *c_sq, *part_v;
  int* part_i;
  uint32_t *assign, *sorted_keys, *sorted_pos;
  int *counts, *n_empty;
  void* sort_ws;
  size_t sort_ws_bytes;
  bool ok;
};

static size_t km_ws_bytes(long long n, int k, int d) {
  const KmPlan p = km_plan(n, k, d);
  size_t b = 1024;
  b += align_up((size_t)p.a_pad * 3 * p.dp * sizeof(__nv_bfloat16)) + align_up((size_t)p.b_pad * 3 * p.dp * sizeof(__nv_bfloat16));
  b += align_up((size_t)p.a_pad * 4) + align_up((size_t)p.b_pad * 4);
  b += 2 * align_up((size_t)2 * p.n_splits * p.a_pad * 4);
  b += 3 * align_up((size_t)std::max<long long>(n, 1) * 4);
  b += align_up((size_t)std::max(k, 1) * 4) + align_up(4);
  b += align_up(radix_sort_workspace_bytes(std::max<long long>(n, 1), 4, true));
  return b;
}

static KmWs km_carve(void* ws, size_t ws_bytes, long long n, int k, int d) {
  const KmPlan p = km_plan(n, k, d);
  char* base = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(ws) + 1023) & ~uintptr_t(1023));
  const size_t lost = (size_t)(base - static_cast<char*>(ws));
  Arena ar(base, ws_bytes > lost ? ws_bytes - lost : 0);
  KmWs w;
  w.ab = ar.take<__nv_bfloat16>((size_t)p.a_pad * 3 * p.dp);
  w.bb = ar.take<__nv_bfloat16>((size_t)p.b_pad * 3 * p.dp);
  w.x_sq = ar.take<float>(p.a_pad);
  w.c_sq = ar.take<float>(p.b_pad);
  w.part_v = ar.take<float>((size_t)2 * p.n_splits * p.a_pad);
  w.part_i = ar.take<int>((size_t)2 * p.n_splits * p.a_pad);
  const size_t nn = (size_t)std::max<long long>(n, 1);
  w.assign = ar.take<uint32_t>(nn);
  w.sorted_keys = ar.take<uint32_t>(nn);
  w.sorted_pos = ar.take<uint32_t>(nn);
  w.counts = ar.take<int>(std::max(k, 1));
  w.n_empty = ar.take<int>(1);
  w.sort_ws_bytes = radix_sort_workspace_bytes((long long)nn, 4, true);
  w.sort_ws = ar.take<char>(w.sort_ws_bytes);
  w.ok = ar.ok();
  return w;
}

static size_t km_smem_bytes(int n_kc) {
  const int a_chunks = n_kc > 4 ? kKmStages : n_kc;
  return 1024 + (size_t)a_chunks * kKmABytes + (size_t)kKmStages * kKmBBytes + 256;
}

static int km_bits_for(uint32_t max_value) {
  int b = 1;
  while (b < 32 && (max_value >> b) != 0) ++b;
  return b;
}

// centroids [k, d] (fp32, contiguous) -> assignment of the n prepared points
static int km_assign(const KmPlan& p, const KmWs& w, const CUtensorMap& tm_a, const CUtensorMap& tm_b, const float* centroids,
                     long long n, int k, int d, int64_t* assign_i64, float* dist, cudaStream_t st) {
  km_prep_kernel<<<(unsigned)cdiv(p.b_pad * 32, 256), 256, 0, st>>>(centroids, d, k, d, p.dp, p.b_pad, 1, w.bb, w.c_sq);
  GCF_LAUNCH_CHECK("km_prep_kernel");
  const size_t smem = km_smem_bytes(p.n_kc);
  dim3 grid(p.m_tiles, p.n_splits);
#define GCF_KM_LAUNCH(KC)                                                                                                  \
  do {                                                                                                                     \
    GCF_CUDA(cudaFuncSetAttribute(km_assign_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    km_assign_kernel<KC><<<grid, kKmThreads, smem, st>>>(tm_a, tm_b, p.n_tiles, p.tiles_per_split, p.n_kc, w.c_sq, w.part_v, \
                                                         w.part_i, p.a_pad);                                               \
  } while (0)
  if (p.n_kc == 3) GCF_KM_LAUNCH(3);
  else GCF_KM_LAUNCH(0);
#undef GCF_KM_LAUNCH
  GCF_LAUNCH_CHECK("km_assign_kernel");
  km_combine_kernel<<<(unsigned)cdiv(n, 256), 256, 0, st>>>(w.part_v, w.part_i, 2 * p.n_splits, p.a_pad, n, w.x_sq, w.assign,
                                                            assign_i64, dist);
  GCF_LAUNCH_CHECK("km_combine_kernel");
  return GCF_OK;
}

}  // namespace gcf

using namespace gcf;

extern "C" size_t gcf_kmeans_workspace_bytes(int64_t n, int32_t k, int32_t d) {
  if (n <= 0 || k <= 0 || d <= 0) return 0;
  return km_ws_bytes(n, k, d);
}

extern "C" int gcf_kmeans_lloyd(const float* x, int64_t ldx, int64_t n, int32_t d, int32_t k, int32_t n_iter, float* centroids,
                                int64_t* assign, float* dist, void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 1 && k >= 1 && k <= n && n < 2147483647LL, "gcf_kmeans_lloyd: need 1 <= k <= n < 2^31");
  GCF_REQUIRE(n_iter >= 0, "gcf_kmeans_lloyd: negative n_iter");
  if (d <= 0 || (d & 3) != 0 || d > 1024) {
    set_error("gcf_kmeans_lloyd: d=%d unsupported (need d %% 4 == 0 and d <= 1024)", d);
    return GCF_EUNSUPPORTED;
  }
  GCF_REQUIRE(x && centroids, "gcf_kmeans_lloyd: null points / centroids");
  GCF_REQUIRE(ldx >= d && (ldx & 3) == 0 && (reinterpret_cast<uintptr_t>(x) & 15u) == 0 &&
                  (reinterpret_cast<uintptr_t>(centroids) & 15u) == 0,
              "gcf_kmeans_lloyd: x / centroids must be 16B aligned with ld %% 4 == 0");
  const size_t need = km_ws_bytes(n, k, d);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_kmeans_lloyd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  GCF_REQUIRE((size_t)k * sizeof(int) <= 200 * 1024, "gcf_kmeans_lloyd: k too large for the empty-cluster pass (k <= 51200)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const KmPlan p = km_plan(n, k, d);
  const KmWs w = km_carve(workspace, workspace_bytes, n, k, d);
  GCF_REQUIRE(w.ok, "gcf_kmeans_lloyd: workspace carve-up failed");
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap(&tm_a, w.ab, p.a_pad, 3 * p.dp, kKmTileM);
  if (rc != GCF_OK) return rc;
  rc = make_tmap(&tm_b, w.bb, p.b_pad, 3 * p.dp, kKmTileN);
  if (rc != GCF_OK) return rc;
  km_prep_kernel<<<(unsigned)cdiv(p.a_pad * 32, 256), 256, 0, st>>>(x, ldx, n, d, p.dp, p.a_pad, 0, w.ab, w.x_sq);
  GCF_LAUNCH_CHECK("km_prep_kernel");
  GCF_CUDA(cudaMemsetAsync(w.n_empty, 0, sizeof(int), st));
  GCF_CUDA(cudaFuncSetAttribute(km_fix_empty_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)k * sizeof(int))));
  const int key_bits = km_bits_for((uint32_t)(k - 1));
  for (int it = 0; it < n_iter; ++it) {
    rc = km_assign(p, w, tm_a, tm_b, centroids, n, k, d, nullptr, nullptr, st);
    if (rc != GCF_OK) return rc;
    rc = radix_sort_u32(w.assign, nullptr, w.sorted_keys, w.sorted_pos, n, key_bits, w.sort_ws, w.sort_ws_bytes, st);
    if (rc != GCF_OK) return rc;
    km_centroid_kernel<<<k, 256, 0, st>>>(x, ldx, d / 4, w.sorted_keys, w.sorted_pos, n, centroids, w.counts, w.n_empty);
    GCF_LAUNCH_CHECK("km_centroid_kernel");
    km_fix_empty_kernel<<<1, 1024, (size_t)k * sizeof(int), st>>>(centroids, d, k, w.counts, w.n_empty);
    GCF_LAUNCH_CHECK("km_fix_empty_kernel");
  }
  if (assign != nullptr || dist != nullptr) {
    rc = km_assign(p, w, tm_a, tm_b, centroids, n, k, d, assign, dist, st);   // faiss: index.search(x, 1) on the trained centroids
    if (rc != GCF_OK) return rc;
  }
  return GCF_OK;
}
