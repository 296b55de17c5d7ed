// Exclusive scan and stable LSD radix sort, written for this library (no CUB/Thrust).
//
// Radix sort, one pass per 8-bit digit:
//   1. histogram: every block counts the digits of its tile -> hist[digit][block]
//   2. exclusive scan over the digit-major table (all blocks of digit 0, then digit 1, ...)
//   3. scatter: every block re-reads its tile in (warp, round, lane) order; match.any gives each
//      element its rank among equal digits of the same warp-round, a per-warp running count and a
//      cross-warp prefix turn that into a stable rank inside the tile, and the scanned table
//      supplies the tile's base offset per digit.
// Stability of every pass makes the multi-pass sort stable, which the CSR build relies on
// (duplicate (row, col) entries are summed in their original order).
#include "radix.cuh"
#include <algorithm>

namespace gcf {

// ========================================================================================
// exclusive scan
// ========================================================================================
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;  // 2048 elements per block

// Scans one tile per block; writes per-tile totals when tile_sums != NULL.
__global__ void __launch_bounds__(kScanThreads)
scan_tiles_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, long long n,
                  uint32_t* __restrict__ tile_sums) {
  __shared__ uint32_t warp_tot[kScanThreads / 32];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
  uint32_t v[kScanItems];
  uint32_t sum = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    v[k] = (base + k < n) ? in[base + k] : 0u;
    sum += v[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = sum;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += y;
  }
  if (lane == 31) warp_tot[warp] = incl;
  __syncthreads();
  uint32_t warp_off = 0;
#pragma unroll
  for (int w = 0; w < kScanThreads / 32; ++w)
    if (w < warp) warp_off += warp_tot[w];
  uint32_t run = warp_off + incl - sum;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) out[base + k] = run;
    run += v[k];
  }
  if (tile_sums != nullptr && threadIdx.x == kScanThreads - 1) tile_sums[blockIdx.x] = run;
}

__global__ void __launch_bounds__(kScanThreads)
scan_add_offsets_kernel(uint32_t* __restrict__ out, long long n, const uint32_t* __restrict__ tile_offsets) {
  const uint32_t off = tile_offsets[blockIdx.x];
  const long long base = (long long)blockIdx.x * kScanTile + (long long)threadIdx.x * kScanItems;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k)
    if (base + k < n) out[base + k] += off;
}

__global__ void scan_total_kernel(const uint32_t* __restrict__ scanned_last, const uint32_t* __restrict__ in_last,
                                  uint32_t* __restrict__ total_out) {
  *total_out = *scanned_last + *in_last;
}

size_t scan_workspace_bytes(int64_t n) {
  // per level: tile sums (scanned in place) + one saved copy of the last input element
  size_t bytes = align_up(sizeof(uint32_t));
  long long m = n;
  while (m > kScanTile) {
    m = cdiv(m, kScanTile);
    bytes += align_up((size_t)m * sizeof(uint32_t));
  }
  return bytes + 256;
}

static int scan_rec(const uint32_t* in, uint32_t* out, long long n, Arena& ar, cudaStream_t st) {
  const long long tiles = cdiv(n, kScanTile);
  if (tiles <= 1) {
    scan_tiles_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
    GCF_LAUNCH_CHECK("scan_tiles_kernel");
    return GCF_OK;
  }
  uint32_t* sums = ar.take<uint32_t>(tiles);
  GCF_REQUIRE(sums != nullptr, "exclusive_scan_u32: workspace too small");
  scan_tiles_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(in, out, n, sums);
  GCF_LAUNCH_CHECK("scan_tiles_kernel");
  int rc = scan_rec(sums, sums, tiles, ar, st);
  if (rc != GCF_OK) return rc;
  scan_add_offsets_kernel<<<(unsigned)tiles, kScanThreads, 0, st>>>(out, n, sums);
  GCF_LAUNCH_CHECK("scan_add_offsets_kernel");
  return GCF_OK;
}

int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_out, void* ws, size_t ws_bytes,
                       cudaStream_t st) {
  if (n <= 0) {
    if (total_out != nullptr) GCF_CUDA(cudaMemsetAsync(total_out, 0, sizeof(uint32_t), st));
    return GCF_OK;
  }
  GCF_REQUIRE(n < (1LL << 40), "exclusive_scan_u32: n too large");
  Arena ar(ws, ws_bytes);
  uint32_t* last_in = nullptr;
  if (total_out != nullptr) {
    // `out` may alias `in`: keep the last input element before it is overwritten
    last_in = ar.take<uint32_t>(1);
    GCF_REQUIRE(last_in != nullptr, "exclusive_scan_u32: workspace too small");
    GCF_CUDA(cudaMemcpyAsync(last_in, in + (n - 1), sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
  }
  int rc = scan_rec(in, out, n, ar, st);
  if (rc != GCF_OK) return rc;
  if (total_out != nullptr) {
    scan_total_kernel<<<1, 1, 0, st>>>(out + (n - 1), last_in, total_out);
    GCF_LAUNCH_CHECK("scan_total_kernel");
  }
  return GCF_OK;
}

// ========================================================================================
// radix sort
// ========================================================================================
constexpr int kSortThreads = 256;
constexpr int kSortWarps = kSortThreads / 32;
constexpr int kSortRounds = 8;                                  // elements per thread
constexpr int kSortTile = kSortThreads * kSortRounds;           // 2048 keys per block
constexpr int kRadix = 256;

template <typename K>
__device__ __forceinline__ unsigned digit_of(K key, int shift) {
  return (unsigned)((key >> shift) & (K)0xff);
}

template <typename K>
__global__ void __launch_bounds__(kSortThreads)
radix_hist_kernel(const K* __restrict__ keys, long long n, int shift, uint32_t* __restrict__ hist, long long n_blocks) {
  __shared__ uint32_t sh[kRadix];
  sh[threadIdx.x] = 0;  // kSortThreads == kRadix
  __syncthreads();
  const long long base = (long long)blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const long long i = base + (long long)r * kSortThreads + threadIdx.x;
    if (i < n) atomicAdd(&sh[digit_of<K>(keys[i], shift)], 1u);
  }
  __syncthreads();
  hist[(long long)threadIdx.x * n_blocks + blockIdx.x] = sh[threadIdx.x];
}

template <typename K, bool HAS_PAY, bool IOTA>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const K* __restrict__ keys_in, const uint32_t* __restrict__ pay_in, K* __restrict__ keys_out,
                     uint32_t* __restrict__ pay_out, long long n, int shift, const uint32_t* __restrict__ hist_scanned,
                     long long n_blocks) {
  __shared__ uint32_t warp_cnt[kSortWarps][kRadix];  // running per-warp digit counts, then bases
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < kSortWarps * kRadix; i += kSortThreads) (&warp_cnt[0][0])[i] = 0;
  __syncthreads();

  // element order inside the tile: (warp, round, lane) -- each warp owns a contiguous slice
  const long long warp_base = (long long)blockIdx.x * kSortTile + (long long)warp * (32 * kSortRounds);
  K key[kSortRounds];
  uint32_t rank[kSortRounds];
  unsigned dig[kSortRounds];
  const unsigned lt_mask = (1u << lane) - 1u;
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const long long i = warp_base + r * 32 + lane;
    const bool valid = i < n;
    key[r] = valid ? keys_in[i] : (K)0;
    dig[r] = valid ? digit_of<K>(key[r], shift) : (0x100u + (unsigned)lane);  // invalid lanes never match
    const unsigned peers = __match_any_sync(0xffffffffu, dig[r]);
    const int leader = __ffs(peers) - 1;
    uint32_t before = 0;
    if (valid && lane == leader) {
      before = warp_cnt[warp][dig[r]];
      warp_cnt[warp][dig[r]] = before + __popc(peers);
    }
    before = __shfl_sync(0xffffffffu, before, leader);
    rank[r] = before + __popc(peers & lt_mask);
    __syncwarp();
  }
  __syncthreads();
  // cross-warp exclusive prefix per digit + the tile's global base for that digit
  {
    const int dg = threadIdx.x;  // one digit per thread
    uint32_t run = hist_scanned[(long long)dg * n_blocks + blockIdx.x];
#pragma unroll
    for (int w = 0; w < kSortWarps; ++w) {
      const uint32_t c = warp_cnt[w][dg];
      warp_cnt[w][dg] = run;
      run += c;
    }
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < kSortRounds; ++r) {
    const long long i = warp_base + r * 32 + lane;
    if (i < n) {
      const uint32_t dst = warp_cnt[warp][dig[r]] + rank[r];
      keys_out[dst] = key[r];
      if (HAS_PAY) pay_out[dst] = IOTA ? (uint32_t)i : pay_in[i];
    }
  }
}

size_t radix_sort_workspace_bytes(int64_t n, int key_bytes, bool with_payload) {
  if (n <= 0) return 256;
  const long long blocks = cdiv(n, kSortTile);
  size_t bytes = align_up((size_t)n * key_bytes);                   // alternate key buffer
  if (with_payload) bytes += align_up((size_t)n * sizeof(uint32_t));  // alternate payload buffer
  bytes += align_up((size_t)blocks * kRadix * sizeof(uint32_t));    // digit-major histogram
  bytes += scan_workspace_bytes(blocks * kRadix);
  return bytes + 256;
}

template <typename K>
static int radix_sort_impl(const K* keys_in, const uint32_t* pay_in, K* keys_out, uint32_t* pay_out, int64_t n,
                           int end_bit, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (n <= 0) return GCF_OK;
  GCF_REQUIRE(n < 4294967295LL, "radix_sort: n must fit in 32 bits");
  GCF_REQUIRE(end_bit >= 1 && end_bit <= (int)(8 * sizeof(K)), "radix_sort: bad end_bit");
  GCF_REQUIRE(keys_in != nullptr && keys_out != nullptr && (const void*)keys_in != (const void*)keys_out,
              "radix_sort: keys_in/keys_out must be distinct non-null buffers");
  const bool has_pay = pay_out != nullptr;
  const int passes = (end_bit + 7) / 8;
  const long long blocks = cdiv(n, kSortTile);
  Arena ar(ws, ws_bytes);
  K* keys_tmp = ar.take<K>(n);
  uint32_t* pay_tmp = has_pay ? ar.take<uint32_t>(n) : nullptr;
  uint32_t* hist = ar.take<uint32_t>(blocks * kRadix);
  const size_t scan_ws = scan_workspace_bytes(blocks * kRadix);
  void* scan_buf = ar.take<char>(scan_ws);
  GCF_REQUIRE(ar.ok() && keys_tmp && hist && scan_buf, "radix_sort: workspace too small");

  const K* src_k = keys_in;
  const uint32_t* src_p = pay_in;
  for (int p = 0; p < passes; ++p) {
    const bool to_out = ((passes - 1 - p) % 2) == 0;
    K* dst_k = to_out ? keys_out : keys_tmp;
    uint32_t* dst_p = to_out ? pay_out : pay_tmp;
    const int shift = 8 * p;
    radix_hist_kernel<K><<<(unsigned)blocks, kSortThreads, 0, st>>>(src_k, n, shift, hist, blocks);
    GCF_LAUNCH_CHECK("radix_hist_kernel");
    int rc = exclusive_scan_u32(hist, hist, blocks * kRadix, nullptr, scan_buf, scan_ws, st);
    if (rc != GCF_OK) return rc;
    if (!has_pay)
      radix_scatter_kernel<K, false, false><<<(unsigned)blocks, kSortThreads, 0, st>>>(src_k, nullptr, dst_k, nullptr, n, shift, hist, blocks);
    else if (p == 0 && pay_in == nullptr)
      radix_scatter_kernel<K, true, true><<<(unsigned)blocks, kSortThreads, 0, st>>>(src_k, nullptr, dst_k, dst_p, n, shift, hist, blocks);
    else
      radix_scatter_kernel<K, true, false><<<(unsigned)blocks, kSortThreads, 0, st>>>(src_k, src_p, dst_k, dst_p, n, shift, hist, blocks);
    GCF_LAUNCH_CHECK("radix_scatter_kernel");
    src_k = dst_k;
    src_p = dst_p;
  }
  return GCF_OK;
}

int radix_sort_u32(const uint32_t* keys_in, const uint32_t* pay_in, uint32_t* keys_out, uint32_t* pay_out, int64_t n,
                   int end_bit, void* ws, size_t ws_bytes, cudaStream_t st) {
  return radix_sort_impl<uint32_t>(keys_in, pay_in, keys_out, pay_out, n, end_bit, ws, ws_bytes, st);
}
int radix_sort_u64(const uint64_t* keys_in, const uint32_t* pay_in, uint64_t* keys_out, uint32_t* pay_out, int64_t n,
                   int end_bit, void* ws, size_t ws_bytes, cudaStream_t st) {
  return radix_sort_impl<uint64_t>(keys_in, pay_in, keys_out, pay_out, n, end_bit, ws, ws_bytes, st);
}

}  // namespace gcf
