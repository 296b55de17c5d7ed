// Adjacency construction as integer CUDA kernels: degree count, bidirectional edge index,
// COO -> canonical CSR (stable radix sort + duplicate merge), value normalisation and CSR transpose.
//
// Replaces the scipy / pandas / PyG host code of the reference:
//   lightgcn.py:36-39 (edge_index), lightgcn.py:17,25 (PyG gcn_norm, recomputed K times per forward),
//   selfcf.py:240-255,297-306 and ssl4rec.py:79-88 (csr_matrix + A + A^T + D^-1/2 A D^-1/2),
//   ncl.py:76-85 / directau.py:132-141 (raw COO with duplicates, summed later by the SpMM).
// Index results are bit-exact with scipy's canonical CSR; values follow (dinv[r]*a)*dinv[c].
#include "common.cuh"
#include "radix.cuh"
#include <algorithm>

namespace gcf {

static inline int launch_blocks(long long n, int threads = 256, int waves = 8) {
  return (int)std::max<long long>(1, std::min<long long>(cdiv(n, threads), (long long)sm_count() * waves));
}

static int bits_for_count(int64_t n_values) {  // bits needed to store values in [0, n_values)
  int b = 1;
  while (b < 63 && ((int64_t)1 << b) < n_values) ++b;
  return b;
}

__global__ void __launch_bounds__(256)
degree_count_kernel(const int64_t* __restrict__ idx, long long n, int* __restrict__ deg, long long n_nodes) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long v = idx[i];
    if (v >= 0 && v < n_nodes) atomicAdd(deg + v, 1);
  }
}

__global__ void __launch_bounds__(256)
bipartite_edge_index_kernel(const int64_t* __restrict__ users, const int64_t* __restrict__ items, long long n_edges,
                            long long n_users, int64_t* __restrict__ rows, int64_t* __restrict__ cols) {
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n_edges;
       e += (long long)gridDim.x * blockDim.x) {
    const long long u = users[e], i = items[e] + n_users;
    rows[e] = u;            cols[e] = i;
    rows[n_edges + e] = i;  cols[n_edges + e] = u;
  }
}

__global__ void __launch_bounds__(256)
pack_keys_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ cols, long long n, int col_bits,
                 long long n_rows, long long n_cols, uint64_t* __restrict__ keys, int* __restrict__ bad_flag) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = rows[i], c = cols[i];
    if (r < 0 || r >= n_rows || c < 0 || c >= n_cols) {
      *bad_flag = 1;
      keys[i] = ~0ull;  // sorts last; dropped by the compaction
    } else {
      keys[i] = ((uint64_t)r << col_bits) | (uint64_t)c;
    }
  }
}

__global__ void __launch_bounds__(256)
head_flags_kernel(const uint64_t* __restrict__ keys, long long n, uint32_t* __restrict__ flags) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    flags[i] = (k != ~0ull && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
  }
}

// One thread per run head: writes the distinct entry, sums duplicates in their original order.
__global__ void __launch_bounds__(256)
compact_runs_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ payload,
                    const float* __restrict__ vals, const uint32_t* __restrict__ pos, long long n, int col_bits,
                    int* __restrict__ col_idx, float* __restrict__ out_vals, int* __restrict__ row_cnt) {
  const uint64_t cmask = (col_bits >= 64) ? ~0ull : ((1ull << col_bits) - 1ull);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    if (k == ~0ull) continue;
    if (i > 0 && keys[i - 1] == k) continue;
    float s = 0.f;
    for (long long j = i; j < n && keys[j] == k; ++j) s += (vals != nullptr) ? vals[payload[j]] : 1.f;
    const uint32_t dst = pos[i];
    col_idx[dst] = (int)(k & cmask);
    out_vals[dst] = s;
    atomicAdd(row_cnt + (long long)(k >> col_bits), 1);
  }
}

// number of distinct entries, or -1 when a row / column index was out of range (scipy and torch raise on such input,
// selfcf.py:297-306; the caller turns the -1 into its own error instead of building a graph with entries missing)
__global__ void write_nnz_kernel(const uint32_t* __restrict__ total, const int* __restrict__ bad, int64_t* __restrict__ nnz_out) {
  *nnz_out = (*bad != 0) ? -1 : (int64_t)(*total);
}

// warp per row: rowsum + scaling vector
__global__ void __launch_bounds__(256)
rowsum_kernel(const int* __restrict__ row_ptr, const float* __restrict__ vals, long long n_rows, int mode,
              float* __restrict__ rowsum, float* __restrict__ dinv) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int s = row_ptr[row], e = row_ptr[row + 1];
  float acc = 0.f;
  for (int j = s + lane; j < e; j += 32) acc += vals[j];
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) {
    if (rowsum != nullptr) rowsum[row] = acc;
    float di = 1.f;
    if (mode == 1) di = 1.0f / sqrtf(acc);  // IEEE div + sqrt == torch pow(-0.5) (SURVEY 8a numerics note)
    else if (mode == 2) di = 1.0f / acc;
    if (isinf(di)) di = 0.f;                // d_inv[np.isinf(d_inv)] = 0  (selfcf.py:246)
    dinv[row] = di;
  }
}

__global__ void __launch_bounds__(256)
scale_values_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx, const float* __restrict__ vals_in,
                    long long n_rows, int mode, const float* __restrict__ dinv, float* __restrict__ vals_out) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int s = row_ptr[row], e = row_ptr[row + 1];
  const float dr = dinv[row];
  for (int j = s + lane; j < e; j += 32) {
    float v = dr * vals_in[j];                     // d_mat_inv.dot(adj)
    if (mode == 1) v = v * dinv[col_idx[j]];       // .dot(d_mat_inv)
    vals_out[j] = v;
  }
}

__global__ void __launch_bounds__(256)
scale_given_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx, const float* __restrict__ vals_in,
                   long long n_rows, const float* __restrict__ row_scale, const float* __restrict__ col_scale,
                   float* __restrict__ vals_out) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int s = row_ptr[row], e = row_ptr[row + 1];
  const float dr = row_scale != nullptr ? row_scale[row] : 1.f;
  for (int j = s + lane; j < e; j += 32) {
    float v = dr * vals_in[j];
    if (col_scale != nullptr) v = v * col_scale[col_idx[j]];
    vals_out[j] = v;
  }
}

__global__ void __launch_bounds__(256)
col_hist_kernel(const int* __restrict__ col_idx, long long nnz, int* __restrict__ cnt, uint32_t* __restrict__ keys) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < nnz; j += (long long)gridDim.x * blockDim.x) {
    const int c = col_idx[j];
    atomicAdd(cnt + c, 1);
    keys[j] = (uint32_t)c;
  }
}

__global__ void __launch_bounds__(256)
transpose_fill_kernel(const int* __restrict__ row_ptr, long long n_rows, const float* __restrict__ vals,
                      const uint32_t* __restrict__ sorted_pos, long long nnz, int* __restrict__ t_col_idx,
                      float* __restrict__ t_vals) {
  for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < nnz; k += (long long)gridDim.x * blockDim.x) {
    const uint32_t j = sorted_pos[k];
    // row of entry j: last r with row_ptr[r] <= j
    long long lo = 0, hi = n_rows;  // invariant: row_ptr[lo] <= j < row_ptr[hi]
    while (hi - lo > 1) {
      const long long mid = (lo + hi) >> 1;
      if ((uint32_t)row_ptr[mid] <= j) lo = mid; else hi = mid;
    }
    t_col_idx[k] = (int)lo;
    t_vals[k] = vals[j];
  }
}

}  // namespace gcf

using namespace gcf;

extern "C" int gcf_degree_count(const int64_t* idx, int64_t n, int32_t* deg, int64_t n_nodes, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0 && n_nodes >= 0, "gcf_degree_count: negative sizes");
  GCF_REQUIRE(n_nodes == 0 || deg != nullptr, "gcf_degree_count: null deg");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_nodes > 0) GCF_CUDA(cudaMemsetAsync(deg, 0, (size_t)n_nodes * sizeof(int32_t), st));
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(idx != nullptr, "gcf_degree_count: null idx");
  degree_count_kernel<<<launch_blocks(n), 256, 0, st>>>(idx, n, deg, n_nodes);
  GCF_LAUNCH_CHECK("degree_count_kernel");
  return GCF_OK;
}

extern "C" int gcf_bipartite_edge_index(const int64_t* users, const int64_t* items, int64_t n_edges, int64_t n_users,
                                        int64_t* rows, int64_t* cols, gcf_stream_t stream) {
  GCF_REQUIRE(n_edges >= 0, "gcf_bipartite_edge_index: negative n_edges");
  if (n_edges == 0) return GCF_OK;
  GCF_REQUIRE(users && items && rows && cols, "gcf_bipartite_edge_index: null pointers");
  bipartite_edge_index_kernel<<<launch_blocks(n_edges), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      users, items, n_edges, n_users, rows, cols);
  GCF_LAUNCH_CHECK("bipartite_edge_index_kernel");
  return GCF_OK;
}

namespace {
struct CooLayout {
  size_t keys_in, keys_sorted, payload, flags, total, bad, sort_ws, scan_ws, end;
  size_t sort_bytes, scan_bytes;
};
CooLayout coo_layout(int64_t nnz, int64_t n_rows, bool with_vals) {
  CooLayout L;
  const size_t n = (size_t)std::max<int64_t>(nnz, 1);
  size_t off = 0;
  L.keys_in = off;     off += align_up(n * sizeof(uint64_t));
  L.keys_sorted = off; off += align_up(n * sizeof(uint64_t));
  L.payload = off;     off += with_vals ? align_up(n * sizeof(uint32_t)) : 0;
  L.flags = off;       off += align_up(n * sizeof(uint32_t));
  L.total = off;       off += 256;
  L.bad = off;         off += 256;
  L.sort_bytes = radix_sort_workspace_bytes(nnz, 8, with_vals);
  L.sort_ws = off;     off += align_up(L.sort_bytes);
  L.scan_bytes = std::max(scan_workspace_bytes(nnz), scan_workspace_bytes(n_rows + 1));
  L.scan_ws = off;     off += align_up(L.scan_bytes);
  L.end = off;
  return L;
}
}  // namespace

extern "C" size_t gcf_coo_to_csr_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols) {
  (void)n_cols;
  return coo_layout(nnz, n_rows, true).end;
}

extern "C" int gcf_coo_to_csr_stable(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz,
                                     int64_t n_rows, int64_t n_cols, int32_t* row_ptr, int32_t* col_idx,
                                     float* out_vals, int64_t* nnz_out, void* workspace, size_t workspace_bytes,
                                     gcf_stream_t stream) {
  GCF_REQUIRE(nnz >= 0 && n_rows >= 0 && n_cols >= 0, "gcf_coo_to_csr_stable: negative sizes");
  GCF_REQUIRE(nnz < 2147483647LL && n_rows < 2147483647LL && n_cols < 2147483647LL,
              "gcf_coo_to_csr_stable: sizes must fit int32 (nnz=%lld)", (long long)nnz);
  GCF_REQUIRE(row_ptr != nullptr && nnz_out != nullptr, "gcf_coo_to_csr_stable: null row_ptr / nnz_out");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GCF_CUDA(cudaMemsetAsync(row_ptr, 0, (size_t)(n_rows + 1) * sizeof(int32_t), st));
  if (nnz == 0) {
    GCF_CUDA(cudaMemsetAsync(nnz_out, 0, sizeof(int64_t), st));
    return GCF_OK;
  }
  GCF_REQUIRE(rows && cols && col_idx && out_vals, "gcf_coo_to_csr_stable: null pointers");
  const bool with_vals = vals != nullptr;
  const CooLayout L = coo_layout(nnz, n_rows, with_vals);
  if (workspace == nullptr || workspace_bytes < L.end) {
    set_error("gcf_coo_to_csr_stable: workspace too small (%zu < %zu)", workspace_bytes, L.end);
    return GCF_EWORKSPACE;
  }
  char* ws = static_cast<char*>(workspace);
  uint64_t* keys_in = reinterpret_cast<uint64_t*>(ws + L.keys_in);
  uint64_t* keys_sorted = reinterpret_cast<uint64_t*>(ws + L.keys_sorted);
  uint32_t* payload = with_vals ? reinterpret_cast<uint32_t*>(ws + L.payload) : nullptr;
  uint32_t* flags = reinterpret_cast<uint32_t*>(ws + L.flags);
  uint32_t* total = reinterpret_cast<uint32_t*>(ws + L.total);
  int* bad = reinterpret_cast<int*>(ws + L.bad);

  const int col_bits = bits_for_count(n_cols);
  const int row_bits = bits_for_count(n_rows);
  GCF_REQUIRE(col_bits + row_bits <= 63, "gcf_coo_to_csr_stable: key does not fit 63 bits");
  GCF_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
  pack_keys_kernel<<<launch_blocks(nnz), 256, 0, st>>>(rows, cols, nnz, col_bits, n_rows, n_cols, keys_in, bad);
  GCF_LAUNCH_CHECK("pack_keys_kernel");
  // sort on all 64 bits only when invalid entries may exist; the sentinel has every bit set, so sorting the
  // low (row_bits + col_bits + 1) bits still moves it behind every valid key.
  int rc = radix_sort_u64(keys_in, nullptr, keys_sorted, payload, nnz, std::min(64, row_bits + col_bits + 1),
                          ws + L.sort_ws, L.sort_bytes, st);
  if (rc != GCF_OK) return rc;
  head_flags_kernel<<<launch_blocks(nnz), 256, 0, st>>>(keys_sorted, nnz, flags);
  GCF_LAUNCH_CHECK("head_flags_kernel");
  rc = exclusive_scan_u32(flags, flags, nnz, total, ws + L.scan_ws, L.scan_bytes, st);
  if (rc != GCF_OK) return rc;
  compact_runs_kernel<<<launch_blocks(nnz), 256, 0, st>>>(keys_sorted, payload, vals, flags, nnz, col_bits, col_idx,
                                                          out_vals, row_ptr);
  GCF_LAUNCH_CHECK("compact_runs_kernel");
  rc = exclusive_scan_u32(reinterpret_cast<uint32_t*>(row_ptr), reinterpret_cast<uint32_t*>(row_ptr), n_rows + 1,
                          nullptr, ws + L.scan_ws, L.scan_bytes, st);
  if (rc != GCF_OK) return rc;
  write_nnz_kernel<<<1, 1, 0, st>>>(total, bad, nnz_out);
  GCF_LAUNCH_CHECK("write_nnz_kernel");
  return GCF_OK;
}

extern "C" int gcf_norm_values(int32_t mode, const int32_t* row_ptr, const int32_t* col_idx, const float* vals_in,
                               int64_t n_rows, int64_t n_cols, float* vals_out, float* rowsum_out, float* dinv_out,
                               gcf_stream_t stream) {
  GCF_REQUIRE(mode >= 0 && mode <= 2, "gcf_norm_values: mode must be 0 (none), 1 (sym) or 2 (row)");
  GCF_REQUIRE(mode != 1 || n_rows == n_cols, "gcf_norm_values: sym normalisation needs a square matrix");
  if (n_rows == 0) return GCF_OK;
  GCF_REQUIRE(row_ptr && dinv_out, "gcf_norm_values: null row_ptr / dinv_out");  // col/val may be null when nnz == 0
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long blocks = cdiv(n_rows * 32, 256);
  GCF_REQUIRE(blocks < 2147483647LL, "gcf_norm_values: too many rows");
  rowsum_kernel<<<(unsigned)blocks, 256, 0, st>>>(row_ptr, vals_in, n_rows, mode, rowsum_out, dinv_out);
  GCF_LAUNCH_CHECK("rowsum_kernel");
  scale_values_kernel<<<(unsigned)blocks, 256, 0, st>>>(row_ptr, col_idx, vals_in, n_rows, mode, dinv_out, vals_out);
  GCF_LAUNCH_CHECK("scale_values_kernel");
  return GCF_OK;
}

extern "C" int gcf_scale_csr_values(const int32_t* row_ptr, const int32_t* col_idx, const float* vals_in, int64_t n_rows,
                                    const float* row_scale, const float* col_scale, float* vals_out,
                                    gcf_stream_t stream) {
  GCF_REQUIRE(n_rows >= 0, "gcf_scale_csr_values: negative n_rows");
  if (n_rows == 0) return GCF_OK;
  GCF_REQUIRE(row_ptr != nullptr, "gcf_scale_csr_values: null row_ptr");
  const long long blocks = cdiv(n_rows * 32, 256);
  GCF_REQUIRE(blocks < 2147483647LL, "gcf_scale_csr_values: too many rows");
  scale_given_kernel<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(row_ptr, col_idx, vals_in, n_rows,
                                                                                    row_scale, col_scale, vals_out);
  GCF_LAUNCH_CHECK("scale_given_kernel");
  return GCF_OK;
}

extern "C" size_t gcf_csr_transpose_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols) {
  (void)n_rows;
  const size_t n = (size_t)std::max<int64_t>(nnz, 1);
  return 3 * align_up(n * sizeof(uint32_t)) + align_up(radix_sort_workspace_bytes(nnz, 4, true)) +
         align_up(scan_workspace_bytes(n_cols + 1)) + 256;
}

extern "C" int gcf_csr_transpose(const int32_t* row_ptr, const int32_t* col_idx, const float* vals, int64_t n_rows,
                                 int64_t n_cols, int64_t nnz, int32_t* t_row_ptr, int32_t* t_col_idx, float* t_vals,
                                 void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(nnz >= 0 && n_rows >= 0 && n_cols >= 0 && nnz < 2147483647LL, "gcf_csr_transpose: sizes out of range");
  GCF_REQUIRE(t_row_ptr != nullptr, "gcf_csr_transpose: null t_row_ptr");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GCF_CUDA(cudaMemsetAsync(t_row_ptr, 0, (size_t)(n_cols + 1) * sizeof(int32_t), st));
  if (nnz == 0) return GCF_OK;
  GCF_REQUIRE(row_ptr && col_idx && vals && t_col_idx && t_vals, "gcf_csr_transpose: null pointers");
  const size_t need = gcf_csr_transpose_workspace_bytes(nnz, n_rows, n_cols);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_csr_transpose: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  Arena ar(workspace, workspace_bytes);
  uint32_t* keys = ar.take<uint32_t>(nnz);
  uint32_t* keys_sorted = ar.take<uint32_t>(nnz);
  uint32_t* pos_sorted = ar.take<uint32_t>(nnz);
  const size_t sort_bytes = radix_sort_workspace_bytes(nnz, 4, true);
  void* sort_ws = ar.take<char>(sort_bytes);
  const size_t scan_bytes = scan_workspace_bytes(n_cols + 1);
  void* scan_ws = ar.take<char>(scan_bytes);
  GCF_REQUIRE(ar.ok(), "gcf_csr_transpose: workspace carve-up failed");

  col_hist_kernel<<<launch_blocks(nnz), 256, 0, st>>>(col_idx, nnz, t_row_ptr, keys);
  GCF_LAUNCH_CHECK("col_hist_kernel");
  int rc = exclusive_scan_u32(reinterpret_cast<uint32_t*>(t_row_ptr), reinterpret_cast<uint32_t*>(t_row_ptr),
                              n_cols + 1, nullptr, scan_ws, scan_bytes, st);
  if (rc != GCF_OK) return rc;
  rc = radix_sort_u32(keys, nullptr, keys_sorted, pos_sorted, nnz, bits_for_count(n_cols), sort_ws, sort_bytes, st);
  if (rc != GCF_OK) return rc;
  transpose_fill_kernel<<<launch_blocks(nnz), 256, 0, st>>>(row_ptr, n_rows, vals, pos_sorted, nnz, t_col_idx, t_vals);
  GCF_LAUNCH_CHECK("transpose_fill_kernel");
  return GCF_OK;
}
