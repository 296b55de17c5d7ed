// Sparse x sparse products behind the motif-induced adjacency matrices of MHCN (univariate/mhcn.py:340-368,
// build_hyper_adj_mats) -- SURVEY.md 8f row 4.  The reference evaluates sixteen expressions of the form
// (P.dot(Q)).multiply(M) plus the full Y.dot(Y.T) with scipy on the host.  Here:
//
//   * gcf_csr_sample      out[e] = X[i_e, j_e] for every stored entry e of a pattern P (0 where X has no entry): the
//                         building block of S.multiply(S.T), S - B, and of aligning two matrices on one pattern;
//   * gcf_spgemm_masked   out[e] = M[e] * sum_k A[i_e, k] * B[k, j_e] on M's pattern only -- the product is never
//                         materialised; one thread per mask entry intersects two sorted index lists (the shorter one is
//                         walked, the longer one binary-searched);
//   * gcf_spgemm_count / gcf_spgemm_expand   full product by expand-sort-compress: every A entry (i, k) emits one COO
//                         product per entry of B's row k at an offset given by an exclusive scan; the stable COO -> CSR
//                         build (gcf_coo_to_csr_stable) then sorts and sums them in emission order.
//
// All index work is integer-exact; values are fp32 sums of products (integer-valued for the 0/1 matrices of the
// reference, hence exact as well).
#include "common.cuh"
#include "radix.cuh"
#include <algorithm>

namespace gcf {

// row of entry e: largest r with row_ptr[r] <= e
__device__ __forceinline__ int row_of_entry(const int* __restrict__ row_ptr, int n_rows, int e) {
  int lo = 0, hi = n_rows;  // invariant: row_ptr[lo] <= e < row_ptr[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (__ldg(row_ptr + mid) <= e) lo = mid; else hi = mid;
  }
  return lo;
}

// position of `key` in the ascending list idx[b, e), or -1
__device__ __forceinline__ int find_col(const int* __restrict__ idx, int b, int e, int key) {
  while (b < e) {
    const int mid = (b + e) >> 1;
    const int c = __ldg(idx + mid);
    if (c == key) return mid;
    if (c < key) b = mid + 1; else e = mid;
  }
  return -1;
}

__global__ void __launch_bounds__(256)
csr_sample_kernel(const int* __restrict__ x_row_ptr, const int* __restrict__ x_col, const float* __restrict__ x_val,
                  const int* __restrict__ p_row_ptr, const int* __restrict__ p_col, int p_rows, long long p_nnz,
                  float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < p_nnz; e += (long long)gridDim.x * blockDim.x) {
    const int i = row_of_entry(p_row_ptr, p_rows, (int)e);
    const int j = __ldg(p_col + e);
    const int pos = find_col(x_col, __ldg(x_row_ptr + i), __ldg(x_row_ptr + i + 1), j);
    out[e] = pos >= 0 ? __ldg(x_val + pos) : 0.f;
  }
}

__global__ void __launch_bounds__(256)
spgemm_masked_kernel(const int* __restrict__ a_row_ptr, const int* __restrict__ a_col, const float* __restrict__ a_val,
                     const int* __restrict__ bt_row_ptr, const int* __restrict__ bt_col, const float* __restrict__ bt_val,
                     const int* __restrict__ m_row_ptr, const int* __restrict__ m_col, const float* __restrict__ m_val,
                     int m_rows, long long m_nnz, float* __restrict__ out) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < m_nnz; e += (long long)gridDim.x * blockDim.x) {
    const int i = row_of_entry(m_row_ptr, m_rows, (int)e);
    const int j = __ldg(m_col + e);
    int sb = __ldg(a_row_ptr + i), se = __ldg(a_row_ptr + i + 1);        // row i of A
    int lb = __ldg(bt_row_ptr + j), le = __ldg(bt_row_ptr + j + 1);      // row j of B^T = column j of B
    const int* s_col = a_col; const float* s_val = a_val;
    const int* l_col = bt_col; const float* l_val = bt_val;
    if (se - sb > le - lb) {  // walk the shorter list
      int t = sb; sb = lb; lb = t;
      t = se; se = le; le = t;
      s_col = bt_col; s_val = bt_val; l_col = a_col; l_val = a_val;
    }
    float acc = 0.f;
    for (int p = sb; p < se && lb < le; ++p) {
      const int k = __ldg(s_col + p);
      // both lists ascend: everything before the hit (or the insertion point) can be skipped from now on
      int b = lb, en = le;
      while (b < en) {
        const int mid = (b + en) >> 1;
        if (__ldg(l_col + mid) < k) b = mid + 1; else en = mid;
      }
      lb = b;
      if (b < le && __ldg(l_col + b) == k) acc = fmaf(__ldg(s_val + p), __ldg(l_val + b), acc);
    }
    out[e] = __ldg(m_val + e) * acc;
  }
}

// cnt[t] = length of B's row k_t; the 64-bit total goes to *total64 (zeroed by the caller) so that the host can tell
// when a range exceeds the 32-bit offsets of the scan and must be split
__global__ void __launch_bounds__(256)
spgemm_count_kernel(const int* __restrict__ a_col, const int* __restrict__ b_row_ptr, long long e0, long long n,
                    uint32_t* __restrict__ cnt, unsigned long long* __restrict__ total64) {
  unsigned long long local = 0;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const int k = __ldg(a_col + e0 + t);
    const uint32_t c = (uint32_t)(__ldg(b_row_ptr + k + 1) - __ldg(b_row_ptr + k));
    cnt[t] = c;
    local += c;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) local += __shfl_xor_sync(0xffffffffu, local, off);
  if ((threadIdx.x & 31) == 0 && local != 0) atomicAdd(total64, local);
}

// one warp per A entry: products of A[i, k] with row k of B, written at offs[entry]
__global__ void __launch_bounds__(256)
spgemm_expand_kernel(const int* __restrict__ a_row_ptr, const int* __restrict__ a_col, const float* __restrict__ a_val,
                     int a_rows, const int* __restrict__ b_row_ptr, const int* __restrict__ b_col,
                     const float* __restrict__ b_val, long long e0, long long n, const uint32_t* __restrict__ offs,
                     int64_t* __restrict__ rows_out, int64_t* __restrict__ cols_out, float* __restrict__ vals_out) {
  const int lane = threadIdx.x & 31;
  const long long warps = (long long)gridDim.x * (blockDim.x >> 5);
  for (long long t = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); t < n; t += warps) {
    const long long e = e0 + t;
    const int i = row_of_entry(a_row_ptr, a_rows, (int)e);
    const int k = __ldg(a_col + e);
    const float a = __ldg(a_val + e);
    const int b0 = __ldg(b_row_ptr + k), b1 = __ldg(b_row_ptr + k + 1);
    const long long o = offs[t];
    for (int q = b0 + lane; q < b1; q += 32) {
      rows_out[o + (q - b0)] = i;
      cols_out[o + (q - b0)] = __ldg(b_col + q);
      vals_out[o + (q - b0)] = a * __ldg(b_val + q);
    }
  }
}

static bool csr_ok(const gcf_csr_t* m) {
  return m != nullptr && m->n_rows >= 0 && m->n_cols >= 0 && m->nnz >= 0 && m->nnz < 2147483647LL &&
         (m->n_rows == 0 || m->row_ptr != nullptr) && (m->nnz == 0 || (m->col_idx != nullptr && m->vals != nullptr));
}

static int grid_for(long long n, int per_block = 256) {
  return (int)std::max<long long>(1, std::min<long long>(cdiv(n, per_block), (long long)sm_count() * 16));
}

}  // namespace gcf

using namespace gcf;

extern "C" int gcf_csr_sample(const gcf_csr_t* X, const gcf_csr_t* P, float* out, gcf_stream_t stream) {
  GCF_REQUIRE(csr_ok(X) && csr_ok(P), "gcf_csr_sample: malformed operand");
  GCF_REQUIRE(X->n_rows == P->n_rows && X->n_cols == P->n_cols, "gcf_csr_sample: shapes differ (%lld x %lld vs %lld x %lld)",
              (long long)X->n_rows, (long long)X->n_cols, (long long)P->n_rows, (long long)P->n_cols);
  if (P->nnz == 0) return GCF_OK;
  GCF_REQUIRE(out != nullptr, "gcf_csr_sample: null output");
  csr_sample_kernel<<<grid_for(P->nnz), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X->row_ptr, X->col_idx, X->vals, P->row_ptr, P->col_idx, (int)P->n_rows, P->nnz, out);
  GCF_LAUNCH_CHECK("csr_sample_kernel");
  return GCF_OK;
}

extern "C" int gcf_spgemm_masked(const gcf_csr_t* A, const gcf_csr_t* Bt, const gcf_csr_t* mask, float* out,
                                 gcf_stream_t stream) {
  GCF_REQUIRE(csr_ok(A) && csr_ok(Bt) && csr_ok(mask), "gcf_spgemm_masked: malformed operand");
  GCF_REQUIRE(A->n_cols == Bt->n_cols, "gcf_spgemm_masked: inner dimensions differ (A is %lld wide, B^T is %lld wide)",
              (long long)A->n_cols, (long long)Bt->n_cols);
  GCF_REQUIRE(mask->n_rows == A->n_rows && mask->n_cols == Bt->n_rows, "gcf_spgemm_masked: mask must be %lld x %lld",
              (long long)A->n_rows, (long long)Bt->n_rows);
  if (mask->nnz == 0) return GCF_OK;
  GCF_REQUIRE(out != nullptr, "gcf_spgemm_masked: null output");
  spgemm_masked_kernel<<<grid_for(mask->nnz), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      A->row_ptr, A->col_idx, A->vals, Bt->row_ptr, Bt->col_idx, Bt->vals, mask->row_ptr, mask->col_idx, mask->vals,
      (int)mask->n_rows, mask->nnz, out);
  GCF_LAUNCH_CHECK("spgemm_masked_kernel");
  return GCF_OK;
}

extern "C" size_t gcf_spgemm_workspace_bytes(int64_t n_entries) {
  if (n_entries <= 0) return 256;
  return align_up((size_t)n_entries * sizeof(uint32_t)) + align_up(sizeof(uint32_t)) + align_up(scan_workspace_bytes(n_entries));
}

extern "C" int gcf_spgemm_count(const gcf_csr_t* A, const gcf_csr_t* B, int64_t entry_begin, int64_t entry_end,
                                int64_t* n_products, void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(csr_ok(A) && csr_ok(B), "gcf_spgemm_count: malformed operand");
  GCF_REQUIRE(A->n_cols == B->n_rows, "gcf_spgemm_count: inner dimensions differ");
  GCF_REQUIRE(0 <= entry_begin && entry_begin <= entry_end && entry_end <= A->nnz, "gcf_spgemm_count: bad entry range");
  GCF_REQUIRE(n_products != nullptr, "gcf_spgemm_count: null output");
  const long long n = entry_end - entry_begin;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) { GCF_CUDA(cudaMemsetAsync(n_products, 0, sizeof(int64_t), st)); return GCF_OK; }
  if (workspace == nullptr || workspace_bytes < gcf_spgemm_workspace_bytes(n)) {
    set_error("gcf_spgemm_count: workspace too small (%zu < %zu)", workspace_bytes, gcf_spgemm_workspace_bytes(n));
    return GCF_EWORKSPACE;
  }
  Arena ar(workspace, workspace_bytes);
  uint32_t* offs = ar.take<uint32_t>(n);
  uint32_t* total = ar.take<uint32_t>(1);
  const size_t sb = scan_workspace_bytes(n);
  void* scan_ws = ar.take<char>(sb);
  GCF_REQUIRE(ar.ok(), "gcf_spgemm_count: workspace carve-up failed");
  GCF_CUDA(cudaMemsetAsync(n_products, 0, sizeof(int64_t), st));
  spgemm_count_kernel<<<grid_for(n), 256, 0, st>>>(A->col_idx, B->row_ptr, entry_begin, n, offs,
                                                   reinterpret_cast<unsigned long long*>(n_products));
  GCF_LAUNCH_CHECK("spgemm_count_kernel");
  // offsets stay in the workspace for gcf_spgemm_expand (valid only while *n_products < 2^32: the caller splits otherwise)
  return exclusive_scan_u32(offs, offs, n, total, scan_ws, sb, st);
}

extern "C" int gcf_spgemm_expand(const gcf_csr_t* A, const gcf_csr_t* B, int64_t entry_begin, int64_t entry_end,
                                 int64_t* rows_out, int64_t* cols_out, float* vals_out, void* workspace,
                                 size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(csr_ok(A) && csr_ok(B), "gcf_spgemm_expand: malformed operand");
  GCF_REQUIRE(A->n_cols == B->n_rows, "gcf_spgemm_expand: inner dimensions differ");
  GCF_REQUIRE(0 <= entry_begin && entry_begin <= entry_end && entry_end <= A->nnz, "gcf_spgemm_expand: bad entry range");
  const long long n = entry_end - entry_begin;
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(rows_out && cols_out && vals_out, "gcf_spgemm_expand: null output");
  GCF_REQUIRE(workspace != nullptr && workspace_bytes >= gcf_spgemm_workspace_bytes(n),
              "gcf_spgemm_expand: pass the workspace gcf_spgemm_count filled for the same entry range");
  const uint32_t* offs = reinterpret_cast<const uint32_t*>(workspace);
  spgemm_expand_kernel<<<grid_for(n, 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      A->row_ptr, A->col_idx, A->vals, (int)A->n_rows, B->row_ptr, B->col_idx, B->vals, entry_begin, n, offs, rows_out,
      cols_out, vals_out);
  GCF_LAUNCH_CHECK("spgemm_expand_kernel");
  return GCF_OK;
}
