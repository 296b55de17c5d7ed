// Batch row gather and its backward, the scatter-add into the embedding tables.
//
// Replaces aten::index (x[idx]) and index_put_(accumulate=True) behind ncl.py:314-316,
// selfcf.py:504-511, directau.py:222, ssl4rec.py:190, gcl.py:216-218.
//
// scatter-add, mode 0 ("warp-aggregated"): a warp takes 32 source rows, groups equal destination
// indices with match.any, sums each group's rows in registers and issues ONE 128-bit
// red.global.add.v4.f32 per distinct destination row and float4 column -- hub items that repeat
// inside a batch cost one L2 atomic instead of many.
// mode 1 (deterministic): destination indices are radix-sorted (stable) together with their
// source positions; one warp per destination row sums its sources in source order and does a
// plain read-modify-write (each destination row is owned by exactly one warp).
#include "common.cuh"
#include "radix.cuh"
#include <algorithm>

namespace gcf {

__global__ void __launch_bounds__(256)
gather_rows_kernel(const float4* __restrict__ table, long long ld4, long long n_table_rows, int dvec,
                   const int64_t* __restrict__ idx, long long n, float4* __restrict__ out, long long ldo4) {
  const long long total = n * dvec;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long t = i / dvec;
    const int v = (int)(i - t * dvec);
    const long long r = idx[t];
    float4 x = f4_zero();
    if (r >= 0 && r < n_table_rows) x = __ldg(table + r * ld4 + v);
    out[t * ldo4 + v] = x;
  }
}

__global__ void __launch_bounds__(256)
scatter_add_agg_kernel(const float4* __restrict__ src, long long lds4, int dvec, const int64_t* __restrict__ idx,
                       long long n, float* __restrict__ table, long long ld, long long n_table_rows) {
  const int lane = threadIdx.x & 31;
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const long long n_warps = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long base = warp * 32; base < n; base += n_warps * 32) {
    const long long t = base + lane;
    long long my = -(long long)(lane + 1);  // distinct sentinels: never grouped, never written
    if (t < n) {
      const long long r = idx[t];
      if (r >= 0 && r < n_table_rows) my = r;
    }
    const unsigned grp = __match_any_sync(0xffffffffu, my);
    const bool leader = (my >= 0) && ((__ffs(grp) - 1) == lane);
    unsigned todo = __ballot_sync(0xffffffffu, leader);
    while (todo) {
      const int L = __ffs(todo) - 1;
      todo &= todo - 1;
      const unsigned members = __shfl_sync(0xffffffffu, grp, L);
      const long long row = __shfl_sync(0xffffffffu, my, L);
      for (int v = lane; v < dvec; v += 32) {
        float4 acc = f4_zero();
        unsigned mm = members;
        while (mm) {
          const int j = __ffs(mm) - 1;
          mm &= mm - 1;
          f4_add(acc, __ldg(src + (base + j) * lds4 + v));
        }
        red_add_v4(table + row * ld + 4 * v, acc);
      }
    }
  }
}

// sorted_idx / sorted_pos: destination rows ascending, ties in source order.
__global__ void __launch_bounds__(256)
scatter_add_sorted_kernel(const float4* __restrict__ src, long long lds4, int dvec,
                          const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ sorted_pos,
                          long long n, float4* __restrict__ table, long long ld4, long long n_table_rows) {
  const int lane = threadIdx.x & 31;
  const long long q = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n) return;
  const uint32_t row = sorted_idx[q];
  if (q > 0 && sorted_idx[q - 1] == row) return;  // not a segment head
  if ((long long)row >= n_table_rows) return;      // invalid indices were mapped past the table
  for (int v = lane; v < dvec; v += 32) {
    float4 acc = table[(long long)row * ld4 + v];
    for (long long j = q; j < n && sorted_idx[j] == row; ++j) f4_add(acc, __ldg(src + (long long)sorted_pos[j] * lds4 + v));
    table[(long long)row * ld4 + v] = acc;
  }
}

__global__ void __launch_bounds__(256)
idx_to_u32_kernel(const int64_t* __restrict__ idx, long long n, long long n_table_rows, uint32_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const long long r = idx[i];
    out[i] = (r >= 0 && r < n_table_rows) ? (uint32_t)r : (uint32_t)n_table_rows;  // invalid -> one past the end
  }
}

static int bits_for(uint64_t max_value) {
  int b = 1;
  while (b < 64 && (max_value >> b) != 0) ++b;
  return b;
}

}  // namespace gcf

using namespace gcf;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" int gcf_gather_rows(const float* table, int64_t ld, int64_t n_table_rows, int32_t d, const int64_t* idx,
                               int64_t n, float* out, int64_t ld_out, gcf_stream_t stream) {
  if (d <= 0 || (d & 3) != 0) {
    set_error("gcf_gather_rows: d=%d unsupported (need d %% 4 == 0)", d);
    return GCF_EUNSUPPORTED;
  }
  GCF_REQUIRE(n >= 0 && n_table_rows >= 0, "gcf_gather_rows: negative sizes");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(table && idx && out && aligned16(table) && aligned16(out), "gcf_gather_rows: null/misaligned pointers");
  GCF_REQUIRE(ld >= d && ld_out >= d && (ld & 3) == 0 && (ld_out & 3) == 0, "gcf_gather_rows: bad leading dims");
  const int dvec = d / 4;
  const long long total = (long long)n * dvec;
  const int blocks = (int)std::min<long long>(cdiv(total, 256), (long long)sm_count() * 16);
  gather_rows_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const float4*>(table), ld / 4, n_table_rows, dvec, idx, n, reinterpret_cast<float4*>(out),
      ld_out / 4);
  GCF_LAUNCH_CHECK("gather_rows_kernel");
  return GCF_OK;
}

extern "C" size_t gcf_scatter_add_workspace_bytes(int64_t n, int64_t n_table_rows, int32_t mode) {
  (void)n_table_rows;
  if (mode != 1 || n <= 0) return 0;
  // keys in, keys out, payload out + the sort's own scratch
  return 3 * align_up((size_t)n * sizeof(uint32_t)) + radix_sort_workspace_bytes(n, 4, true);
}

extern "C" int gcf_scatter_add_rows(const float* src, int64_t ld_src, int32_t d, const int64_t* idx, int64_t n,
                                    float* table_grad, int64_t ld, int64_t n_table_rows, int32_t mode,
                                    void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  if (d <= 0 || (d & 3) != 0) {
    set_error("gcf_scatter_add_rows: d=%d unsupported (need d %% 4 == 0)", d);
    return GCF_EUNSUPPORTED;
  }
  GCF_REQUIRE(mode == 0 || mode == 1, "gcf_scatter_add_rows: mode must be 0 (atomic) or 1 (deterministic)");
  GCF_REQUIRE(n >= 0 && n_table_rows >= 0 && n < 2147483647LL && n_table_rows < 4294967295LL,
              "gcf_scatter_add_rows: sizes out of range");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(src && idx && table_grad && aligned16(src) && aligned16(table_grad), "gcf_scatter_add_rows: null/misaligned pointers");
  GCF_REQUIRE(ld_src >= d && ld >= d && (ld_src & 3) == 0 && (ld & 3) == 0, "gcf_scatter_add_rows: bad leading dims");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dvec = d / 4;
  if (mode == 0) {
    const int blocks = (int)std::min<long long>(cdiv(n, 32 * 8), (long long)sm_count() * 16);
    scatter_add_agg_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), ld_src / 4, dvec, idx, n,
                                                   table_grad, ld, n_table_rows);
    GCF_LAUNCH_CHECK("scatter_add_agg_kernel");
    return GCF_OK;
  }
  const size_t need = gcf_scatter_add_workspace_bytes(n, n_table_rows, 1);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_scatter_add_rows: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  Arena ar(workspace, workspace_bytes);
  uint32_t* keys_in = ar.take<uint32_t>(n);
  uint32_t* keys_out = ar.take<uint32_t>(n);
  uint32_t* pos_out = ar.take<uint32_t>(n);
  const size_t sort_ws = radix_sort_workspace_bytes(n, 4, true);
  void* sort_buf = ar.take<char>(sort_ws);
  GCF_REQUIRE(ar.ok(), "gcf_scatter_add_rows: workspace carve-up failed");
  {
    const int blocks = (int)std::min<long long>(cdiv(n, 256), (long long)sm_count() * 8);
    idx_to_u32_kernel<<<blocks, 256, 0, st>>>(idx, n, n_table_rows, keys_in);
    GCF_LAUNCH_CHECK("idx_to_u32_kernel");
  }
  int rc = radix_sort_u32(keys_in, nullptr, keys_out, pos_out, n, bits_for((uint64_t)n_table_rows), sort_buf, sort_ws, st);
  if (rc != GCF_OK) return rc;
  {
    const long long blocks = cdiv(n, 8);  // one warp per sorted position
    scatter_add_sorted_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(src), ld_src / 4, dvec,
                                                               keys_out, pos_out, n,
                                                               reinterpret_cast<float4*>(table_grad), ld / 4,
                                                               n_table_rows);
    GCF_LAUNCH_CHECK("scatter_add_sorted_kernel");
  }
  return GCF_OK;
}

// ---- column-slice <-> row-major layout conversion around the collectives of the feature-sharded trainers ----
// `blocked` is [n_slices][n_rows][w] (what an all-gather / all-to-all of per-rank [n_rows, w] column slices delivers),
// `rows` is the row-major [n_rows, n_slices * w] table the fused loss kernels read.  One pass, 128-bit accesses on both
// sides (w % 4 == 0), grid-stride over the float4 elements of the row-major side.
namespace gcf {

template <bool TO_ROWS>
__global__ void __launch_bounds__(256)
slices_rows_kernel(float4* __restrict__ blocked, float4* __restrict__ rows, long long ld4, long long n_rows, int n_slices,
                   int w4) {
  const long long d4 = (long long)n_slices * w4;
  const long long total = n_rows * d4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / d4;
    const int q = (int)(i - r * d4);
    const int g = q / w4, qq = q - g * w4;
    float4* b = blocked + ((long long)g * n_rows + r) * w4 + qq;
    float4* x = rows + r * ld4 + q;
    if (TO_ROWS) *x = __ldcs(b); else *b = __ldcs(x);
  }
}

static int slices_rows(bool to_rows, const float* blocked, const float* rows, int64_t ld_rows, int64_t n_rows, int32_t n_slices,
                       int32_t w, gcf_stream_t stream, const char* who) {
  GCF_REQUIRE(blocked != nullptr && rows != nullptr, "%s: null pointer", who);
  GCF_REQUIRE(n_rows >= 0 && n_slices >= 1 && w >= 4 && (w & 3) == 0, "%s: need n_slices >= 1 and w %% 4 == 0 (w=%d)", who, w);
  GCF_REQUIRE(ld_rows >= (int64_t)n_slices * w && (ld_rows & 3) == 0, "%s: ld_rows too small or not a multiple of 4", who);
  GCF_REQUIRE((reinterpret_cast<uintptr_t>(blocked) & 15u) == 0 && (reinterpret_cast<uintptr_t>(rows) & 15u) == 0,
              "%s: buffers must be 16-byte aligned", who);
  if (n_rows == 0) return GCF_OK;
  const long long total = (long long)n_rows * n_slices * (w / 4);
  const int blocks = (int)std::min<long long>(cdiv(total, 256), (long long)sm_count() * 16);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float4* b4 = reinterpret_cast<float4*>(const_cast<float*>(blocked));
  float4* r4 = reinterpret_cast<float4*>(const_cast<float*>(rows));
  if (to_rows) slices_rows_kernel<true><<<blocks, 256, 0, st>>>(b4, r4, ld_rows / 4, n_rows, n_slices, w / 4);
  else slices_rows_kernel<false><<<blocks, 256, 0, st>>>(b4, r4, ld_rows / 4, n_rows, n_slices, w / 4);
  GCF_LAUNCH_CHECK("slices_rows_kernel");
  return GCF_OK;
}

}  // namespace gcf

extern "C" int gcf_slices_to_rows(const float* blocked, float* rows, int64_t ld_rows, int64_t n_rows, int32_t n_slices,
                                  int32_t w, gcf_stream_t stream) {
  return gcf::slices_rows(true, blocked, rows, ld_rows, n_rows, n_slices, w, stream, "gcf_slices_to_rows");
}

extern "C" int gcf_rows_to_slices(const float* rows, int64_t ld_rows, float* blocked, int64_t n_rows, int32_t n_slices,
                                  int32_t w, gcf_stream_t stream) {
  return gcf::slices_rows(false, blocked, rows, ld_rows, n_rows, n_slices, w, stream, "gcf_rows_to_slices");
}
