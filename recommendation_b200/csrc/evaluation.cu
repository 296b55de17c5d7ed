// Batched top-N recommendation and ranking metrics.
//
// Replaces the reference's per-user evaluation loop -- predict (one [I, d] x [d] matmul per user), Python masking of the
// rated items with -1e8, torch.topk, dict building, and the set-intersection metrics of ranking_evaluation
// (ncl.py:133-178,253-277; lightgcn.py:48-74).  The score block itself is a plain dense GEMM (left to cuBLAS through
// torch.matmul, as the reference does); this file owns what follows it:
//   masked_topn_kernel : one CTA per query row: mask the user's training items, exact radix select of the n-th largest
//                        score (3 passes over the row, order-preserving float -> uint keys), ordered collection
//                        (ties broken by the lower item id) and an in-smem bitonic sort.  Deterministic.
//   ranking_hits_kernel: per query, hit count / DCG at every cut-off against the user's sorted test items.
#include "common.cuh"
#include <algorithm>

namespace gcf {

constexpr int kTopnThreads = 256;
constexpr int kTopnMax = 128;

__device__ __forceinline__ uint32_t float_key(float f) {  // larger float -> larger key (total order, -0 < +0)
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// block-wide exclusive scan of one int per thread (256 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_sums, int* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, off);
    if (lane >= off) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < kTopnThreads / 32; ++w) {
    const int s = warp_sums[w];
    if (w < warp) base += s;
    tot += s;
  }
  __syncthreads();
  *total = tot;
  return base + inc - v;
}

__global__ void __launch_bounds__(kTopnThreads)
masked_topn_kernel(float* __restrict__ scores, long long ld, long long n_items, const int64_t* __restrict__ users,
                   const int* __restrict__ pos_row_ptr, const int* __restrict__ pos_col_idx, float mask_value, int n_top,
                   int64_t* __restrict__ out_idx, float* __restrict__ out_val) {
  __shared__ unsigned int hist[2048];
  __shared__ int warp_sums[kTopnThreads / 32];
  __shared__ unsigned long long cand[kTopnMax];
  __shared__ unsigned int s_prefix, s_remaining;
  const long long q = blockIdx.x;
  float* row = scores + q * ld;
  const int tid = threadIdx.x;

  // 0. the user's training items can never be recommended (ncl.py:257-259: candidates[item] = -1e8)
  if (pos_row_ptr != nullptr) {
    const long long u = users != nullptr ? users[q] : q;
    for (int j = pos_row_ptr[u] + tid; j < pos_row_ptr[u + 1]; j += kTopnThreads) {
      const int it = pos_col_idx[j];
      if (it >= 0 && it < n_items) row[it] = mask_value;
    }
  }
  __syncthreads();

  // 1. radix select: the key of the n_top-th largest element.  prefix = the already fixed high bits.
  if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)n_top; }
  const int shifts[3] = {21, 10, 0};
  const int widths[3] = {11, 11, 10};
  unsigned int mask_hi = 0u;
  for (int pass = 0; pass < 3; ++pass) {
    for (int b = tid; b < 2048; b += kTopnThreads) hist[b] = 0u;
    __syncthreads();
    const unsigned int prefix = s_prefix;
    const int sh = shifts[pass], nb = 1 << widths[pass];
    for (long long i = tid; i < n_items; i += kTopnThreads) {
      const uint32_t k = float_key(row[i]);
      if ((k & mask_hi) == prefix) atomicAdd(&hist[(k >> sh) & (nb - 1)], 1u);
    }
    __syncthreads();
    if (tid == 0) {  // walk the bins from the top until the n-th element is inside one
      unsigned int rem = s_remaining;
      int b = nb - 1;
      for (; b > 0; --b) {
        if (hist[b] >= rem) break;
        rem -= hist[b];
      }
      s_prefix = prefix | ((unsigned)b << sh);
      s_remaining = rem;  // how many elements of this bin (and, at the end, of the exact key) are still needed
    }
    mask_hi |= (unsigned)(nb - 1) << sh;
    __syncthreads();
  }
  const unsigned int kth = s_prefix;       // exact key of the n_top-th largest score
  const int need_ties = (int)s_remaining;  // how many elements equal to it belong to the result
  __syncthreads();

  // 2. ordered collection: everything above the threshold, then the first `need_ties` equal elements (lowest ids)
  int n_gt = 0, n_eq = 0;
  for (long long base = 0; base < n_items; base += kTopnThreads) {
    const long long i = base + tid;
    uint32_t k = 0u;
    int gt = 0, eq = 0;
    if (i < n_items) { k = float_key(row[i]); gt = k > kth; eq = k == kth; }
    int tot_gt, tot_eq;
    const int p_gt = block_exclusive_scan(gt, warp_sums, &tot_gt);
    const int p_eq = block_exclusive_scan(eq, warp_sums, &tot_eq);
    if (gt) cand[n_gt + p_gt] = ((unsigned long long)(~k) << 32) | (unsigned long long)(uint32_t)i;
    if (eq && n_eq + p_eq < need_ties)
      cand[(n_top - need_ties) + n_eq + p_eq] = ((unsigned long long)(~k) << 32) | (unsigned long long)(uint32_t)i;
    n_gt += tot_gt;
    n_eq += tot_eq;
  }
  __syncthreads();
  // 3. sort by (score descending, item id ascending): ascending on (~key, id); pad to a power of two
  int m = 1;
  while (m < n_top) m <<= 1;
  for (int i = n_top + tid; i < m; i += kTopnThreads) cand[i] = ~0ull;
  __syncthreads();
  for (int size = 2; size <= m; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < m; i += kTopnThreads) {
        const int j = i ^ stride;
        if (j > i) {
          const bool up = (i & size) == 0;
          const unsigned long long a = cand[i], b = cand[j];
          if ((a > b) == up) { cand[i] = b; cand[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int i = tid; i < n_top; i += kTopnThreads) {
    const unsigned long long c = cand[i];
    const uint32_t id = (uint32_t)(c & 0xffffffffull);
    out_idx[q * n_top + i] = (int64_t)id;
    out_val[q * n_top + i] = row[id];
  }
}

// hits[q, c] = |top-N_c(q) ∩ test(u_q)|,  dcg[q, c] = sum over hit positions i < N_c of 1 / log2(i + 2)
__global__ void __launch_bounds__(128)
ranking_hits_kernel(const int64_t* __restrict__ topn, long long n_queries, int n_top, const int64_t* __restrict__ users,
                    const int* __restrict__ test_row_ptr, const int* __restrict__ test_col_idx, const int* __restrict__ cutoffs,
                    int n_cut, int* __restrict__ hits, float* __restrict__ dcg) {
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n_queries) return;
  const long long u = users != nullptr ? users[q] : q;
  const int s = test_row_ptr[u], e = test_row_ptr[u + 1];
  int h = 0;
  float g = 0.f;
  int c = 0;
  for (int i = 0; i < n_top && c < n_cut; ++i) {
    const int item = (int)topn[q * n_top + i];
    int lo = s, hi = e;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (test_col_idx[mid] < item) lo = mid + 1; else hi = mid;
    }
    if (lo < e && test_col_idx[lo] == item) { ++h; g += 1.f / log2f((float)(i + 2)); }
    while (c < n_cut && cutoffs[c] == i + 1) { hits[q * n_cut + c] = h; dcg[q * n_cut + c] = g; ++c; }
  }
  for (; c < n_cut; ++c) { hits[q * n_cut + c] = h; dcg[q * n_cut + c] = g; }  // cut-offs beyond n_top
}

}  // namespace gcf

using namespace gcf;

extern "C" int gcf_masked_topn(float* scores, int64_t ld, int64_t n_queries, int64_t n_items, const int64_t* users,
                               const int32_t* pos_row_ptr, const int32_t* pos_col_idx, float mask_value, int32_t n_top,
                               int64_t* out_idx, float* out_val, gcf_stream_t stream) {
  GCF_REQUIRE(n_queries >= 0 && n_items >= 1, "gcf_masked_topn: bad sizes");
  GCF_REQUIRE(n_top >= 1 && n_top <= kTopnMax && n_top <= n_items, "gcf_masked_topn: n_top must be in [1, min(%d, n_items)]", kTopnMax);
  GCF_REQUIRE(n_items < (1LL << 31) && n_queries < (1LL << 31), "gcf_masked_topn: sizes must fit int32");
  if (n_queries == 0) return GCF_OK;
  GCF_REQUIRE(scores && out_idx && out_val && ld >= n_items, "gcf_masked_topn: null pointers / bad leading dim");
  GCF_REQUIRE((pos_row_ptr == nullptr) == (pos_col_idx == nullptr), "gcf_masked_topn: give both mask arrays or none");
  masked_topn_kernel<<<(unsigned)n_queries, kTopnThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      scores, ld, n_items, users, pos_row_ptr, pos_col_idx, mask_value, n_top, out_idx, out_val);
  GCF_LAUNCH_CHECK("masked_topn_kernel");
  return GCF_OK;
}

extern "C" int gcf_ranking_hits(const int64_t* topn, int64_t n_queries, int32_t n_top, const int64_t* users,
                                const int32_t* test_row_ptr, const int32_t* test_col_idx, const int32_t* cutoffs,
                                int32_t n_cutoffs, int32_t* hits, float* dcg, gcf_stream_t stream) {
  GCF_REQUIRE(n_queries >= 0 && n_top >= 1 && n_cutoffs >= 1, "gcf_ranking_hits: bad sizes");
  if (n_queries == 0) return GCF_OK;
  GCF_REQUIRE(topn && test_row_ptr && test_col_idx && cutoffs && hits && dcg, "gcf_ranking_hits: null pointers");
  ranking_hits_kernel<<<(unsigned)cdiv(n_queries, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      topn, n_queries, n_top, users, test_row_ptr, test_col_idx, cutoffs, n_cutoffs, hits, dcg);
  GCF_LAUNCH_CHECK("ranking_hits_kernel");
  return GCF_OK;
}
