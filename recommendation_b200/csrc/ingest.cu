// Text ingest on the GPU: `user item rating` lines -> id arrays -> dense indices.
//
// Replaces, for the file formats the reference reads,
//   load_data          ncl.py:542-543 (= directau.py, mhcn.py:624-625, ...): [line.strip().split()[:2] + [1.0] ...], blank lines skipped
//   Interaction._build ncl.py:55-70: user / item dictionaries = enumerate(sorted(set(ids)))   (STRING sort: '10' < '2')
//                      selfcf.py:281-288: ids numbered by first appearance
//   load_data          lightgcn.py:29-33: pandas read_csv(sep=' ') of integer ids
// Byte and integer work only:
//   gcf_text_count_records / gcf_text_parse_pairs   every thread owns a 256-byte segment of the text, counts the records
//       (non-blank lines) starting in it, an exclusive scan turns the counts into record numbers, and the second pass parses
//       the first two whitespace-separated tokens of each record into either left-aligned big-endian 8-byte string keys
//       (whose unsigned order IS the byte-wise lexicographic order Python's sorted() uses for ASCII ids) or decimal integers;
//   gcf_sort_unique_u64     stable radix sort + run heads: the distinct keys in ascending order and, for each, the position
//       of its first occurrence (the first-appearance numbering of selfcf.py);
//   gcf_lookup_sorted_u64   binary search of every key in the distinct-key table: the dense index of each token.
#include "common.cuh"
#include "radix.cuh"
#include <algorithm>

namespace gcf {

constexpr int kSeg = 256;  // bytes of text per thread

__device__ __forceinline__ bool is_space(uint8_t c) { return c == ' ' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// A record starts at byte p when p begins a line (p == 0 or text[p-1] == '\n') that holds at least one non-blank byte.
__device__ __forceinline__ bool record_starts_at(const uint8_t* __restrict__ text, long long n, long long p) {
  if (p > 0 && text[p - 1] != '\n') return false;
  for (long long q = p; q < n; ++q) {
    const uint8_t c = text[q];
    if (c == '\n') return false;
    if (!is_space(c)) return true;
  }
  return false;
}

__global__ void __launch_bounds__(256)
count_records_kernel(const uint8_t* __restrict__ text, long long n, long long n_segs, uint32_t* __restrict__ cnt) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += (long long)gridDim.x * blockDim.x) {
    const long long b = s * kSeg, e = min(b + (long long)kSeg, n);
    uint32_t c = 0;
    for (long long p = b; p < e; ++p)
      if ((p == 0 || text[p - 1] == '\n') && record_starts_at(text, n, p)) ++c;
    cnt[s] = c;
  }
}

__global__ void total_to_i64_kernel(const uint32_t* __restrict__ total, int64_t* __restrict__ out) { *out = (int64_t)*total; }

// token [b, e) -> key.  mode 0: up to 8 bytes, left-aligned, big-endian, zero padded.  mode 1: unsigned decimal integer.
__device__ __forceinline__ bool token_to_key(const uint8_t* __restrict__ text, long long b, long long e, int mode, uint64_t& key) {
  key = 0;
  if (mode == 0) {
    if (e - b > 8) return false;
    for (int k = 0; k < 8; ++k) key = (key << 8) | (uint64_t)(b + k < e ? text[b + k] : 0);
    return true;
  }
  if (e - b > 18) return false;
  for (long long q = b; q < e; ++q) {
    const uint8_t c = text[q];
    if (c < '0' || c > '9') return false;
    key = key * 10 + (uint64_t)(c - '0');
  }
  return true;
}

// status bits: 1 = a record with fewer than two tokens, 2 = a token that does not fit the key format
__global__ void __launch_bounds__(256)
parse_pairs_kernel(const uint8_t* __restrict__ text, long long n, long long n_segs, const uint32_t* __restrict__ first_record,
                   int mode, uint64_t* __restrict__ key_a, uint64_t* __restrict__ key_b, int* __restrict__ status) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += (long long)gridDim.x * blockDim.x) {
    const long long b = s * kSeg, e = min(b + (long long)kSeg, n);
    long long rec = first_record[s];
    for (long long p = b; p < e; ++p) {
      if (!((p == 0 || text[p - 1] == '\n') && record_starts_at(text, n, p))) continue;
      long long q = p;
      uint64_t keys[2] = {0, 0};
      int found = 0, bad = 0;
      while (found < 2) {
        while (q < n && text[q] != '\n' && is_space(text[q])) ++q;
        if (q >= n || text[q] == '\n') break;
        long long t0 = q;
        while (q < n && text[q] != '\n' && !is_space(text[q])) ++q;
        if (!token_to_key(text, t0, q, mode, keys[found])) bad = 1;
        ++found;
      }
      if (found < 2) atomicOr(status, 1);
      if (bad) atomicOr(status, 2);
      key_a[rec] = keys[0];
      key_b[rec] = keys[1];
      ++rec;
    }
  }
}

__global__ void __launch_bounds__(256)
run_heads_kernel(const uint64_t* __restrict__ sorted, long long n, uint32_t* __restrict__ flag) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    flag[i] = (i == 0 || sorted[i - 1] != sorted[i]) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
write_heads_kernel(const uint64_t* __restrict__ sorted, const uint32_t* __restrict__ pos_sorted, const uint32_t* __restrict__ slot,
                   long long n, uint64_t* __restrict__ uniq, int64_t* __restrict__ first_pos) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    if (i > 0 && sorted[i - 1] == sorted[i]) continue;
    const uint32_t d = slot[i];
    uniq[d] = sorted[i];
    if (first_pos != nullptr) first_pos[d] = (int64_t)pos_sorted[i];   // stable sort: the head of a run is its first occurrence
  }
}

__global__ void __launch_bounds__(256)
lookup_sorted_kernel(const uint64_t* __restrict__ table, long long n_table, const uint64_t* __restrict__ keys, long long n,
                     int64_t* __restrict__ idx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    long long lo = 0, hi = n_table;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      if (table[mid] < k) lo = mid + 1; else hi = mid;
    }
    idx[i] = (lo < n_table && table[lo] == k) ? lo : -1;
  }
}

// ---- ids longer than 8 bytes: keys of n_words 64-bit words (left-aligned, big-endian, zero padded), word-major [W][R] ----
// Lexicographic order of the word tuples == byte-wise order of the id strings (ncl.py:60-61 sorts arbitrary strings).
__device__ __forceinline__ void token_to_words(const uint8_t* __restrict__ text, long long b, long long e, int n_words,
                                               uint64_t* __restrict__ out, long long stride, long long rec) {
  for (int w = 0; w < n_words; ++w) {
    uint64_t key = 0;
    const long long o = b + 8LL * w;
    for (int k = 0; k < 8; ++k) key = (key << 8) | (uint64_t)(o + k < e ? text[o + k] : 0);
    out[(long long)w * stride + rec] = key;
  }
}

// status[0] bits: 1 = a record with fewer than two tokens, 2 = a token longer than 8 * n_words bytes; status[1] = longest token
__global__ void __launch_bounds__(256)
parse_pairs_words_kernel(const uint8_t* __restrict__ text, long long n, long long n_segs, const uint32_t* __restrict__ first_record,
                         int n_words, long long n_records, uint64_t* __restrict__ key_a, uint64_t* __restrict__ key_b,
                         int* __restrict__ status) {
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < n_segs; s += (long long)gridDim.x * blockDim.x) {
    const long long b = s * kSeg, e = min(b + (long long)kSeg, n);
    long long rec = first_record[s];
    for (long long p = b; p < e; ++p) {
      if (!((p == 0 || text[p - 1] == '\n') && record_starts_at(text, n, p))) continue;
      long long q = p;
      int found = 0, longest = 0;
      while (found < 2) {
        while (q < n && text[q] != '\n' && is_space(text[q])) ++q;
        if (q >= n || text[q] == '\n') break;
        const long long t0 = q;
        while (q < n && text[q] != '\n' && !is_space(text[q])) ++q;
        longest = max(longest, (int)min(q - t0, (long long)0x7fffffff));
        token_to_words(text, t0, q, n_words, found == 0 ? key_a : key_b, n_records, rec);
        ++found;
      }
      if (found < 2) {
        atomicOr(status, 1);
        for (int f = found; f < 2; ++f)
          for (int w = 0; w < n_words; ++w) (f == 0 ? key_a : key_b)[(long long)w * n_records + rec] = 0;
      }
      if (longest > 8 * n_words) atomicOr(status, 2);
      atomicMax(status + 1, longest);
      ++rec;
    }
  }
}

__global__ void __launch_bounds__(256)
gather_u64_kernel(const uint64_t* __restrict__ src, const uint32_t* __restrict__ perm, long long n, uint64_t* __restrict__ out) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = src[perm != nullptr ? perm[i] : i];
}

__global__ void __launch_bounds__(256)
run_heads_words_kernel(const uint64_t* __restrict__ keys, int n_words, long long n, const uint32_t* __restrict__ perm,
                       uint32_t* __restrict__ flag) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    uint32_t head = i == 0;
    if (i > 0) {
      const uint32_t a = perm[i - 1], b = perm[i];
      for (int w = 0; w < n_words && !head; ++w) head = keys[(long long)w * n + a] != keys[(long long)w * n + b];
    }
    flag[i] = head;
  }
}

// flag = run heads, slot = their exclusive scan: the run of sorted position i has number slot[i] - (flag[i] ? 0 : 1)
__global__ void __launch_bounds__(256)
write_heads_words_kernel(const uint64_t* __restrict__ keys, int n_words, long long n, const uint32_t* __restrict__ perm,
                         const uint32_t* __restrict__ flag, const uint32_t* __restrict__ slot, uint64_t* __restrict__ uniq,
                         long long uniq_stride, int64_t* __restrict__ first_pos, int64_t* __restrict__ rank) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint32_t src = perm[i];
    const uint32_t id = slot[i] - (flag[i] ? 0u : 1u);
    if (rank != nullptr) rank[src] = (int64_t)id;
    if (flag[i]) {
      for (int w = 0; w < n_words; ++w) uniq[(long long)w * uniq_stride + id] = keys[(long long)w * n + src];
      if (first_pos != nullptr) first_pos[id] = (int64_t)src;          // stable sort: the head of a run is its first occurrence
    }
  }
}

__global__ void __launch_bounds__(256)
lookup_sorted_words_kernel(const uint64_t* __restrict__ table, int n_words, long long n_table, long long table_stride,
                           const uint64_t* __restrict__ keys, long long n, int64_t* __restrict__ idx) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    long long lo = 0, hi = n_table;
    while (lo < hi) {
      const long long mid = (lo + hi) >> 1;
      int cmp = 0;   // table[mid] <=> key
      for (int w = 0; w < n_words && cmp == 0; ++w) {
        const uint64_t t = table[(long long)w * table_stride + mid], k = keys[(long long)w * n + i];
        cmp = t < k ? -1 : (t > k ? 1 : 0);
      }
      if (cmp < 0) lo = mid + 1; else hi = mid;
    }
    bool eq = lo < n_table;
    for (int w = 0; w < n_words && eq; ++w) eq = table[(long long)w * table_stride + lo] == keys[(long long)w * n + i];
    idx[i] = eq ? lo : -1;
  }
}

static int grid_of(long long n) { return (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256), (long long)sm_count() * 16)); }

}  // namespace gcf

using namespace gcf;

extern "C" size_t gcf_text_workspace_bytes(int64_t n_bytes) {
  const int64_t segs = std::max<int64_t>(1, cdiv(n_bytes, kSeg));
  return align_up((size_t)segs * sizeof(uint32_t)) + align_up(sizeof(uint32_t)) + align_up(scan_workspace_bytes(segs));
}

extern "C" int gcf_text_count_records(const uint8_t* text, int64_t n_bytes, int64_t* n_records, void* workspace,
                                      size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(n_bytes >= 0 && n_records != nullptr, "gcf_text_count_records: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n_bytes == 0) { GCF_CUDA(cudaMemsetAsync(n_records, 0, sizeof(int64_t), st)); return GCF_OK; }
  GCF_REQUIRE(text != nullptr, "gcf_text_count_records: null text");
  if (workspace == nullptr || workspace_bytes < gcf_text_workspace_bytes(n_bytes)) {
    set_error("gcf_text_count_records: workspace too small (%zu < %zu)", workspace_bytes, gcf_text_workspace_bytes(n_bytes));
    return GCF_EWORKSPACE;
  }
  const long long segs = cdiv(n_bytes, kSeg);
  Arena ar(workspace, workspace_bytes);
  uint32_t* cnt = ar.take<uint32_t>(segs);
  uint32_t* total = ar.take<uint32_t>(1);
  const size_t sb = scan_workspace_bytes(segs);
  void* scan_ws = ar.take<char>(sb);
  GCF_REQUIRE(ar.ok(), "gcf_text_count_records: workspace carve-up failed");
  count_records_kernel<<<grid_of(segs), 256, 0, st>>>(text, n_bytes, segs, cnt);
  GCF_LAUNCH_CHECK("count_records_kernel");
  int rc = exclusive_scan_u32(cnt, cnt, segs, total, scan_ws, sb, st);   // per-segment first record number stays in the workspace
  if (rc != GCF_OK) return rc;
  total_to_i64_kernel<<<1, 1, 0, st>>>(total, n_records);
  GCF_LAUNCH_CHECK("total_to_i64_kernel");
  return GCF_OK;
}

extern "C" int gcf_text_parse_pairs(const uint8_t* text, int64_t n_bytes, int32_t mode, uint64_t* first, uint64_t* second,
                                    int32_t* status, void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(n_bytes >= 0 && (mode == 0 || mode == 1) && status != nullptr, "gcf_text_parse_pairs: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GCF_CUDA(cudaMemsetAsync(status, 0, sizeof(int32_t), st));
  if (n_bytes == 0) return GCF_OK;
  GCF_REQUIRE(text && first && second, "gcf_text_parse_pairs: null buffer");
  GCF_REQUIRE(workspace != nullptr && workspace_bytes >= gcf_text_workspace_bytes(n_bytes),
              "gcf_text_parse_pairs: pass the workspace gcf_text_count_records filled for the same text");
  const long long segs = cdiv(n_bytes, kSeg);
  parse_pairs_kernel<<<grid_of(segs), 256, 0, st>>>(text, n_bytes, segs, reinterpret_cast<const uint32_t*>(workspace), mode,
                                                    first, second, status);
  GCF_LAUNCH_CHECK("parse_pairs_kernel");
  return GCF_OK;
}

extern "C" size_t gcf_sort_unique_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return align_up((size_t)n * sizeof(uint64_t)) + 2 * align_up((size_t)n * sizeof(uint32_t)) + align_up(sizeof(uint32_t)) +
         align_up(radix_sort_workspace_bytes(n, 8, true)) + align_up(scan_workspace_bytes(n));
}

extern "C" int gcf_sort_unique_u64(const uint64_t* keys, int64_t n, uint64_t* uniq, int64_t* first_pos, int64_t* n_uniq,
                                   void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0 && n_uniq != nullptr, "gcf_sort_unique_u64: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) { GCF_CUDA(cudaMemsetAsync(n_uniq, 0, sizeof(int64_t), st)); return GCF_OK; }
  GCF_REQUIRE(keys && uniq, "gcf_sort_unique_u64: null buffer");
  if (workspace == nullptr || workspace_bytes < gcf_sort_unique_workspace_bytes(n)) {
    set_error("gcf_sort_unique_u64: workspace too small (%zu < %zu)", workspace_bytes, gcf_sort_unique_workspace_bytes(n));
    return GCF_EWORKSPACE;
  }
  Arena ar(workspace, workspace_bytes);
  uint64_t* sorted = ar.take<uint64_t>(n);
  uint32_t* pos = ar.take<uint32_t>(n);
  uint32_t* flag = ar.take<uint32_t>(n);
  uint32_t* total = ar.take<uint32_t>(1);
  const size_t sort_b = radix_sort_workspace_bytes(n, 8, true);
  void* sort_ws = ar.take<char>(sort_b);
  const size_t scan_b = scan_workspace_bytes(n);
  void* scan_ws = ar.take<char>(scan_b);
  GCF_REQUIRE(ar.ok(), "gcf_sort_unique_u64: workspace carve-up failed");
  int rc = radix_sort_u64(keys, nullptr, sorted, pos, n, 64, sort_ws, sort_b, st);
  if (rc != GCF_OK) return rc;
  run_heads_kernel<<<grid_of(n), 256, 0, st>>>(sorted, n, flag);
  GCF_LAUNCH_CHECK("run_heads_kernel");
  rc = exclusive_scan_u32(flag, flag, n, total, scan_ws, scan_b, st);
  if (rc != GCF_OK) return rc;
  write_heads_kernel<<<grid_of(n), 256, 0, st>>>(sorted, pos, flag, n, uniq, first_pos);
  GCF_LAUNCH_CHECK("write_heads_kernel");
  total_to_i64_kernel<<<1, 1, 0, st>>>(total, n_uniq);
  GCF_LAUNCH_CHECK("total_to_i64_kernel");
  return GCF_OK;
}

extern "C" int gcf_lookup_sorted_u64(const uint64_t* table, int64_t n_table, const uint64_t* keys, int64_t n, int64_t* idx,
                                     gcf_stream_t stream) {
  GCF_REQUIRE(n_table >= 0 && n >= 0, "gcf_lookup_sorted_u64: negative size");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(keys && idx && (n_table == 0 || table), "gcf_lookup_sorted_u64: null buffer");
  lookup_sorted_kernel<<<grid_of(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(table, n_table, keys, n, idx);
  GCF_LAUNCH_CHECK("lookup_sorted_kernel");
  return GCF_OK;
}


// ---- ids of up to 8 * n_words bytes -------------------------------------------------------------------------------------
extern "C" int gcf_text_parse_pairs_words(const uint8_t* text, int64_t n_bytes, int32_t n_words, int64_t n_records,
                                          uint64_t* first, uint64_t* second, int32_t* status, void* workspace,
                                          size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(n_bytes >= 0 && n_words >= 1 && n_words <= 32 && n_records >= 0 && status != nullptr,
              "gcf_text_parse_pairs_words: bad arguments (1 <= n_words <= 32)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GCF_CUDA(cudaMemsetAsync(status, 0, 2 * sizeof(int32_t), st));
  if (n_bytes == 0 || n_records == 0) return GCF_OK;
  GCF_REQUIRE(text && first && second, "gcf_text_parse_pairs_words: null buffer");
  GCF_REQUIRE(workspace != nullptr && workspace_bytes >= gcf_text_workspace_bytes(n_bytes),
              "gcf_text_parse_pairs_words: pass the workspace gcf_text_count_records filled for the same text");
  const long long segs = cdiv(n_bytes, kSeg);
  parse_pairs_words_kernel<<<grid_of(segs), 256, 0, st>>>(text, n_bytes, segs, reinterpret_cast<const uint32_t*>(workspace), n_words,
                                                          n_records, first, second, status);
  GCF_LAUNCH_CHECK("parse_pairs_words_kernel");
  return GCF_OK;
}

extern "C" size_t gcf_sort_unique_words_workspace_bytes(int64_t n) {
  if (n <= 0) return 256;
  return 2 * align_up((size_t)n * sizeof(uint64_t)) + 4 * align_up((size_t)n * sizeof(uint32_t)) + align_up(sizeof(uint32_t)) +
         align_up(radix_sort_workspace_bytes(n, 8, true)) + align_up(scan_workspace_bytes(n));
}

extern "C" int gcf_sort_unique_words(const uint64_t* keys, int32_t n_words, int64_t n, uint64_t* uniq, int64_t* first_pos,
                                     int64_t* n_uniq, int64_t* rank, void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0 && n < 4294967295LL && n_words >= 1 && n_words <= 32 && n_uniq != nullptr, "gcf_sort_unique_words: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n == 0) { GCF_CUDA(cudaMemsetAsync(n_uniq, 0, sizeof(int64_t), st)); return GCF_OK; }
  GCF_REQUIRE(keys && uniq, "gcf_sort_unique_words: null buffer");
  if (workspace == nullptr || workspace_bytes < gcf_sort_unique_words_workspace_bytes(n)) {
    set_error("gcf_sort_unique_words: workspace too small (%zu < %zu)", workspace_bytes, gcf_sort_unique_words_workspace_bytes(n));
    return GCF_EWORKSPACE;
  }
  Arena ar(workspace, workspace_bytes);
  uint64_t* gathered = ar.take<uint64_t>(n);
  uint64_t* sorted = ar.take<uint64_t>(n);
  uint32_t* perm_a = ar.take<uint32_t>(n);
  uint32_t* perm_b = ar.take<uint32_t>(n);
  uint32_t* flag = ar.take<uint32_t>(n);
  uint32_t* slot = ar.take<uint32_t>(n);
  uint32_t* total = ar.take<uint32_t>(1);
  const size_t sort_b = radix_sort_workspace_bytes(n, 8, true);
  void* sort_ws = ar.take<char>(sort_b);
  const size_t scan_b = scan_workspace_bytes(n);
  void* scan_ws = ar.take<char>(scan_b);
  GCF_REQUIRE(ar.ok(), "gcf_sort_unique_words: workspace carve-up failed");
  // LSD over the words, least significant (last) word first; every pass is a stable sort, the payload is the permutation
  uint32_t* perm = nullptr;      // identity
  uint32_t* next = perm_a;
  for (int w = n_words - 1; w >= 0; --w) {
    const uint64_t* word = keys + (long long)w * n;
    const uint64_t* src = word;
    if (perm != nullptr) {
      gather_u64_kernel<<<grid_of(n), 256, 0, st>>>(word, perm, n, gathered);
      GCF_LAUNCH_CHECK("gather_u64_kernel");
      src = gathered;
    }
    int rc = radix_sort_u64(src, perm, sorted, next, n, 64, sort_ws, sort_b, st);
    if (rc != GCF_OK) return rc;
    perm = next;
    next = (next == perm_a) ? perm_b : perm_a;
  }
  run_heads_words_kernel<<<grid_of(n), 256, 0, st>>>(keys, n_words, n, perm, flag);
  GCF_LAUNCH_CHECK("run_heads_words_kernel");
  int rc = exclusive_scan_u32(flag, slot, n, total, scan_ws, scan_b, st);
  if (rc != GCF_OK) return rc;
  write_heads_words_kernel<<<grid_of(n), 256, 0, st>>>(keys, n_words, n, perm, flag, slot, uniq, n, first_pos, rank);
  GCF_LAUNCH_CHECK("write_heads_words_kernel");
  total_to_i64_kernel<<<1, 1, 0, st>>>(total, n_uniq);
  GCF_LAUNCH_CHECK("total_to_i64_kernel");
  return GCF_OK;
}

extern "C" int gcf_lookup_sorted_words(const uint64_t* table, int32_t n_words, int64_t n_table, int64_t table_stride,
                                       const uint64_t* keys, int64_t n, int64_t* idx, gcf_stream_t stream) {
  GCF_REQUIRE(n_table >= 0 && n >= 0 && n_words >= 1 && table_stride >= n_table, "gcf_lookup_sorted_words: bad sizes");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(keys && idx && (n_table == 0 || table), "gcf_lookup_sorted_words: null buffer");
  lookup_sorted_words_kernel<<<grid_of(n), 256, 0, static_cast<cudaStream_t>(stream)>>>(table, n_words, n_table, table_stride, keys, n, idx);
  GCF_LAUNCH_CHECK("lookup_sorted_words_kernel");
  return GCF_OK;
}
