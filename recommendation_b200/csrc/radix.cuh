// Hand-written device-wide primitives used by the graph build and the deterministic
// scatter-add: exclusive prefix sum and a stable LSD radix sort (8 bits per pass).
#pragma once
#include "common.cuh"

namespace gcf {

// ---- exclusive scan over uint32 (out may alias in) -------------------------------------
size_t scan_workspace_bytes(int64_t n);
// total_out (nullable): device uint32 receiving the sum of all n inputs.
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_out, void* ws, size_t ws_bytes,
                       cudaStream_t st);

// ---- stable LSD radix sort -------------------------------------------------------------
// Sorts by key bits [0, end_bit).  pay_in == NULL with pay_out != NULL: payload = original index.
// pay_out == NULL: keys only.  The result always lands in keys_out / pay_out; inputs are not modified.
size_t radix_sort_workspace_bytes(int64_t n, int key_bytes, bool with_payload);
int radix_sort_u32(const uint32_t* keys_in, const uint32_t* pay_in, uint32_t* keys_out, uint32_t* pay_out, int64_t n,
                   int end_bit, void* ws, size_t ws_bytes, cudaStream_t st);
int radix_sort_u64(const uint64_t* keys_in, const uint32_t* pay_in, uint64_t* keys_out, uint32_t* pay_out, int64_t n,
                   int end_bit, void* ws, size_t ws_bytes, cudaStream_t st);

}  // namespace gcf
