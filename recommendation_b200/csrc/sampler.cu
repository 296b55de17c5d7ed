// Philox4x32-10 counter-based negative sampler.
//
// Replaces the Python samplers of the reference: torch.randint without rejection (lightgcn.py:91-94)
// and random.choice / np.random.randint with rejection against the user's training items
// (ncl.py:91-114, selfcf.py:188-211, directau.py:14-32, ssl4rec.py:33-50, gcl.py:111-125).
// The reference never seeds its RNGs, so only the *distribution* (uniform over items, positives
// rejected) is part of the contract; the bit-exact stream is pinned by oracle/philox.py instead.
//
// Stream definition (also restated in oracle/philox.py):
//   key     = (seed_lo, seed_hi ^ offset_hi)
//   counter = (slot_lo, slot_hi, offset_lo, trial / 4)      slot = t * n_negs + j
//   word    = Philox4x32-10(counter, key)[trial % 4]
//   item    = mulhi32(word, n_items)
#include "common.cuh"
#include <algorithm>

namespace gcf {

__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    const uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__global__ void __launch_bounds__(256)
sample_negatives_kernel(uint32_t seed_lo, uint32_t key_hi, uint32_t offset_lo, const int64_t* __restrict__ users,
                        long long n_slots, int n_negs, uint32_t n_items, const int* __restrict__ pos_row_ptr,
                        const int* __restrict__ pos_col_idx, int max_trials, unsigned long long slot_base,
                        const int64_t* __restrict__ slot_pos, int64_t* __restrict__ out) {
  for (long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x; slot < n_slots;
       slot += (long long)gridDim.x * blockDim.x) {
    int ps = 0, pe = 0;
    if (pos_row_ptr != nullptr) {
      const long long u = users[slot / n_negs];
      ps = pos_row_ptr[u];
      pe = pos_row_ptr[u + 1];
    }
    uint32_t cand = 0;
    uint32_t w[4];
    for (int trial = 0; trial < max_trials; ++trial) {
      if ((trial & 3) == 0) {
        // position in the GLOBAL slot numbering: a window of it (slot_base) or an explicit list of triple positions
        const unsigned long long gslot = slot_pos == nullptr ? slot_base + (unsigned long long)slot
            : (unsigned long long)slot_pos[slot / n_negs] * (unsigned long long)n_negs + (unsigned long long)(slot % n_negs);
        philox4x32_10((uint32_t)gslot, (uint32_t)(gslot >> 32), offset_lo, (uint32_t)(trial >> 2), seed_lo, key_hi, w);
      }
      cand = __umulhi(w[trial & 3], n_items);
      // binary search in the user's sorted positives
      int lo = ps, hi = pe;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if ((uint32_t)pos_col_idx[mid] < cand) lo = mid + 1; else hi = mid;
      }
      if (!(lo < pe && (uint32_t)pos_col_idx[lo] == cand)) break;  // accepted
    }
    out[slot] = (int64_t)cand;
  }
}

// out[j] = keep(e) ? vals[e] / (1 - rate) : 0,  e = index ? index[j] : j,  keep(e) <=> Philox(e; seed, offset)[0] < (1 - rate) 2^32
__global__ void __launch_bounds__(256)
dropout_values_kernel(const float* __restrict__ vals, long long n, const int* __restrict__ index, uint32_t thresh, float scale,
                      uint32_t seed_lo, uint32_t seed_hi, uint32_t off_lo, uint32_t off_hi, float* __restrict__ out) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    const long long e = index != nullptr ? (long long)index[j] : j;
    uint32_t w[4];
    philox4x32_10((uint32_t)e, (uint32_t)((unsigned long long)e >> 32), off_lo, off_hi, seed_lo, seed_hi, w);
    out[j] = (w[0] < thresh) ? vals[e] * scale : 0.f;
  }
}

// out[j] = Philox4x32-10(counter = (j, offset), key = seed)[0]: one uniform 32-bit key per entry (random subsets of exact size)
__global__ void __launch_bounds__(256)
philox_keys_kernel(long long n, uint32_t seed_lo, uint32_t seed_hi, uint32_t off_lo, uint32_t off_hi, int64_t* __restrict__ out) {
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (long long)gridDim.x * blockDim.x) {
    uint32_t w[4];
    philox4x32_10((uint32_t)j, (uint32_t)((unsigned long long)j >> 32), off_lo, off_hi, seed_lo, seed_hi, w);
    out[j] = (int64_t)w[0];
  }
}

}  // namespace gcf

using namespace gcf;

extern "C" int gcf_philox_keys(int64_t n, uint64_t seed, uint64_t offset, int64_t* out, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0, "gcf_philox_keys: negative n");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(out != nullptr, "gcf_philox_keys: null output");
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256), (long long)sm_count() * 16));
  philox_keys_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(n, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                                          (uint32_t)offset, (uint32_t)(offset >> 32), out);
  GCF_LAUNCH_CHECK("philox_keys_kernel");
  return GCF_OK;
}

extern "C" int gcf_csr_dropout_values(const float* vals, int64_t n, const int32_t* index, float rate, uint64_t seed,
                                      uint64_t offset, float* out, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0, "gcf_csr_dropout_values: negative n");
  GCF_REQUIRE(rate >= 0.f && rate < 1.f, "gcf_csr_dropout_values: rate must be in [0, 1)");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(vals != nullptr && out != nullptr, "gcf_csr_dropout_values: null pointers");
  const double keep = 1.0 - (double)rate;
  const uint32_t thresh = keep >= 1.0 ? 0xffffffffu : (uint32_t)(keep * 4294967296.0);
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(n, 256), (long long)sm_count() * 16));
  dropout_values_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      vals, n, index, thresh, (float)(1.0 / keep), (uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)offset,
      (uint32_t)(offset >> 32), out);
  GCF_LAUNCH_CHECK("dropout_values_kernel");
  return GCF_OK;
}


extern "C" int gcf_sample_negatives_at(uint64_t seed, uint64_t offset, int64_t slot_base, const int64_t* users, int64_t n,
                                       int32_t n_negs, int64_t n_items, const int32_t* pos_row_ptr, const int32_t* pos_col_idx,
                                       int32_t max_trials, int64_t* out, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0 && n_negs >= 1 && slot_base >= 0, "gcf_sample_negatives: bad n / n_negs / slot_base");
  GCF_REQUIRE(n_items >= 1 && n_items < 4294967296LL, "gcf_sample_negatives: n_items must be in [1, 2^32)");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(out != nullptr, "gcf_sample_negatives: null out");
  GCF_REQUIRE((pos_row_ptr == nullptr) == (pos_col_idx == nullptr), "gcf_sample_negatives: give both CSR arrays or none");
  GCF_REQUIRE(pos_row_ptr == nullptr || users != nullptr, "gcf_sample_negatives: rejection needs the user of each triple");
  if (pos_row_ptr == nullptr || max_trials < 1) max_trials = 1;
  const long long slots = (long long)n * n_negs;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(slots, 256), (long long)sm_count() * 16));
  sample_negatives_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32), (uint32_t)offset, users, slots, n_negs,
      (uint32_t)n_items, pos_row_ptr, pos_col_idx, max_trials, (unsigned long long)slot_base, nullptr, out);
  GCF_LAUNCH_CHECK("sample_negatives_kernel");
  return GCF_OK;
}

extern "C" int gcf_sample_negatives_pos(uint64_t seed, uint64_t offset, const int64_t* slot_pos, int64_t n, int32_t n_negs,
                                        int64_t n_items, int64_t* out, gcf_stream_t stream) {
  GCF_REQUIRE(n >= 0 && n_negs >= 1, "gcf_sample_negatives_pos: bad n / n_negs");
  GCF_REQUIRE(n_items >= 1 && n_items < 4294967296LL, "gcf_sample_negatives_pos: n_items must be in [1, 2^32)");
  if (n == 0) return GCF_OK;
  GCF_REQUIRE(out != nullptr && slot_pos != nullptr, "gcf_sample_negatives_pos: null positions or output");
  const long long slots = (long long)n * n_negs;
  const int blocks = (int)std::max<long long>(1, std::min<long long>(cdiv(slots, 256), (long long)sm_count() * 16));
  sample_negatives_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(offset >> 32), (uint32_t)offset, nullptr, slots, n_negs,
      (uint32_t)n_items, nullptr, nullptr, 1, 0ULL, slot_pos, out);
  GCF_LAUNCH_CHECK("sample_negatives_kernel");
  return GCF_OK;
}

extern "C" int gcf_sample_negatives(uint64_t seed, uint64_t offset, const int64_t* users, int64_t n, int32_t n_negs,
                                    int64_t n_items, const int32_t* pos_row_ptr, const int32_t* pos_col_idx,
                                    int32_t max_trials, int64_t* out, gcf_stream_t stream) {
  return gcf_sample_negatives_at(seed, offset, 0, users, n, n_negs, n_items, pos_row_ptr, pos_col_idx, max_trials, out, stream);
}
