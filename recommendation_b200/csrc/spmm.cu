// CSR SpMM and K-layer LightGCN propagation for sm_100a.
//
// Replaces torch.sparse.mm (ncl.py:419, selfcf.py:479, directau.py:290) / PyG LGConv.propagate
// (lightgcn.py:25) and the layer-combination epilogues (ncl.py:421, selfcf.py:481-482,
// lightgcn.py:26).  HBM/L2-bound gather kernel -- no tensor cores on purpose.
//
// Work decomposition
//   * a sub-warp of LPR = d/4 lanes owns one output row; each lane keeps a float4 slice of
//     the row in registers and streams the row's (col, val) pairs: the sub-warp loads LPR
//     consecutive pairs with one coalesced request, then broadcasts them lane by lane
//     (shfl) and issues UNR independent 128-bit gathers of X rows before consuming them.
//   * rows longer than `chunk` entries (power-law hubs) are cut into chunks; one full warp
//     per chunk, partial sums go to scratch and the last-arriving chunk of a row (atomic
//     ticket, no spinning) reduces them in chunk order and runs the epilogue.  Long chunks
//     occupy the lowest block ids so they are scheduled first.
//   * epilogue fused: optional raw store, optional row-L2-normalise, linear combination
//     with up to 8 addend matrices (layer mean / sum, backward residual terms).
//
// r02 -- flat-stream kernel (spmm_flat_kernel, the default whenever the operator carries a tile schedule):
//   The r01 kernel walked row by row: row_ptr -> (col, val) -> <= 8 gathers, three dependent round trips for a
//   ~13-entry row, and it ran at 22 G gathers/s while a bare gather of the same index stream sustains 25-38 G/s
//   (tools/lab/gather_lab.cu: random 256-byte rows stream from HBM at the full copy rate, 6.5 TB/s, through plain
//   LDG.128; cp.async.bulk / TMA gather4 rings are no faster and lose the L1 hits on hub rows).  What was missing is
//   issue continuity, not a different data path.  So the short rows are cut into TILES of consecutive rows holding
//   ~tile_nnz entries; one sub-warp streams a tile's entries in fixed batches of LPR, always UNR gathers in flight,
//   the (col, val) pairs of batch b+1 requested before the gathers of batch b are issued, and row ends are detected
//   by comparing the running entry index with a register-cached window of row_ptr (segmented sum; empty rows fall
//   out of the same comparison).  Columns and values are broadcast through two 128-byte shared-memory slabs per warp,
//   four per LDS.128, instead of two SHFL per entry.  Hub rows (degree > chunk) keep the chunked path above.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>

namespace gcf {

struct Epi {
  float* Y; long long ldy;
  float* O; long long ldo;
  int epilogue; float alpha, post;
  int n_add;
  const float* add[GCF_MAX_ADDENDS];
  float beta[GCF_MAX_ADDENDS];
  // optional fused optimiser: the epilogue value IS the gradient of row `row` of a dense [n_rows, d] table (ld = ldo);
  // Adam is applied in place and the gradient is not stored unless O is given (gcf_propagate_bwd_adam)
  float* ad_p; float* ad_m; float* ad_v;
  AdamArgs adam;
};

struct LongPlan {
  int chunk, n_long, n_chunks;
  const int* long_rows;
  const int* long_chunk_ptr;
  const int* chunk_long;
  float* partial;
  int* counters;
};

constexpr int kWarpsPerBlock = 8;

// One batch of up to LPR (col, val) pairs -- lane t of the sub-warp holds pair t -- is broadcast lane by lane and
// consumed UNR independent 128-bit row gathers at a time.
template <int LPR, int VPL, int UNR, bool GUARD>
__device__ __forceinline__ void consume_batch(int c, float v, int cnt, const float* __restrict__ X, long long ldx, int sl,
                                              unsigned mask, int dvec, float4 (&acc)[VPL]) {
  if (cnt == LPR) {
#pragma unroll
    for (int t0 = 0; t0 < LPR; t0 += UNR) {
      float4 x[UNR][VPL];
      float w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int ct = __shfl_sync(mask, c, t0 + u, LPR);
        w[u] = __shfl_sync(mask, v, t0 + u, LPR);
        const float4* xr = reinterpret_cast<const float4*>(X + (long long)ct * ldx);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int idx = sl + k * LPR;
          x[u][k] = (!GUARD || idx < dvec) ? __ldg(xr + idx) : f4_zero();
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int k = 0; k < VPL; ++k) f4_fma(acc[k], w[u], x[u][k]);
    }
  } else {
    // ragged tail: replay the last valid column with weight 0 to keep UNR loads in flight
    for (int t0 = 0; t0 < cnt; t0 += UNR) {
      float4 x[UNR][VPL];
      float w[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        const int t = t0 + u;
        const int tt = min(t, cnt - 1);
        const int ct = __shfl_sync(mask, c, tt, LPR);
        const float wt = __shfl_sync(mask, v, tt, LPR);
        w[u] = (t < cnt) ? wt : 0.f;
        const float4* xr = reinterpret_cast<const float4*>(X + (long long)ct * ldx);
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
          const int idx = sl + k * LPR;
          x[u][k] = (!GUARD || idx < dvec) ? __ldg(xr + idx) : f4_zero();
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u)
#pragma unroll
        for (int k = 0; k < VPL; ++k) f4_fma(acc[k], w[u], x[u][k]);
    }
  }
}

template <int LPR, int VPL, int UNR, bool GUARD>
__device__ __forceinline__ void accumulate(const int* __restrict__ col_idx, const float* __restrict__ vals,
                                           int begin, int end, int stride, const float* __restrict__ X,
                                           long long ldx, int sl, unsigned mask, int dvec, float4 (&acc)[VPL]) {
  for (int base = begin; base < end; base += stride) {
    const int j = base + sl;
    int c = 0;
    float v = 0.f;
    if (j < end) {
      c = ld_stream_i32(col_idx + j);
      v = ld_stream_f32(vals + j);
    }
    consume_batch<LPR, VPL, UNR, GUARD>(c, v, min(LPR, end - base), X, ldx, sl, mask, dvec, acc);
  }
}

// Narrow rows (d <= 32 floats: feature-sharded slices): a sub-warp has only LPR = d/4 lanes, so one coalesced
// (col, val) request per iteration would leave just LPR row gathers in flight.  Each lane loads PPL pairs
// (stride LPR) instead and the sub-warp keeps LPR * PPL predicated 128-bit gathers in flight; the FMA order is
// the entry order of accumulate<>, so the result is bit-identical to it.
// NA (tuning variants 5 / 6 only): gathers with L1::no_allocate.  A default (L1-allocating) load of a 32 / 64-byte row fills
// the whole 128-byte L1 line: ncu of the d = 16 kernel shows 3.9 sectors requested from the L2 per 2-sector gather
// (profiles/r02_ncu_narrow16.md).  Bypassing the L1 is nevertheless SLOWER (cfg5: d = 8 2.54 -> 3.39 ms, d = 16 3.70 -> 4.19 ms,
// gpurun_out/r02/exp_narrow_na_d*.log): the neighbour row that rides along is often wanted too, so the default stays.
template <int LPR, int PPL, bool NA = false>
__device__ __forceinline__ void accumulate_multi(const int* __restrict__ col_idx, const float* __restrict__ vals,
                                                 int begin, int end, int stride, const float* __restrict__ X,
                                                 long long ldx, int sl, unsigned mask, float4& acc) {
  // `stride` has accumulate<>'s meaning (distance between a sub-warp's consecutive groups of LPR entries), so the
  // entries a sub-warp sums, and their order, do not depend on PPL
  for (int base = begin; base < end; base += PPL * stride) {
    int c[PPL];
    float v[PPL];
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
      const int j = base + p * stride + sl;
      c[p] = -1;
      v[p] = 0.f;
      if (j < end) {
        c[p] = ld_stream_i32(col_idx + j);
        v[p] = ld_stream_f32(vals + j);
      }
    }
    float4 x[PPL][LPR];
    float w[PPL][LPR];
#pragma unroll
    for (int p = 0; p < PPL; ++p)
#pragma unroll
      for (int t = 0; t < LPR; ++t) {
        const int ct = __shfl_sync(mask, c[p], t, LPR);
        w[p][t] = __shfl_sync(mask, v[p], t, LPR);
        x[p][t] = f4_zero();
        if (ct >= 0) {
          const float4* xp = reinterpret_cast<const float4*>(X + (long long)ct * ldx) + sl;
          x[p][t] = NA ? ldg_f4_l1_bypass(xp) : __ldg(xp);
        }
      }
#pragma unroll
    for (int p = 0; p < PPL; ++p)
#pragma unroll
      for (int t = 0; t < LPR; ++t) f4_fma(acc, w[p][t], x[p][t]);
  }
}

template <int LPR, int VPL, bool GUARD>
__device__ __forceinline__ void finish_row(const Epi& ep, long long row, int sl, unsigned mask, int dvec,
                                           float4 (&acc)[VPL], bool active = true) {
  // `active` = false: the lane only takes part in the shuffles (warp-synchronous callers whose sub-warps hold
  // different numbers of finished rows); `row` may then be anything
  if (ep.Y != nullptr) {
    float4* y = reinterpret_cast<float4*>(ep.Y + row * ep.ldy);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int idx = sl + k * LPR;
      if ((!GUARD || idx < dvec) && active) y[idx] = acc[k];
    }
  }
  if (ep.O != nullptr || ep.ad_p != nullptr) {
    float inv = 1.f;
    if (ep.epilogue == GCF_EPILOGUE_L2NORM) {
      float ss = 0.f;
#pragma unroll
      for (int k = 0; k < VPL; ++k) ss += f4_dot(acc[k], acc[k]);
#pragma unroll
      for (int off = LPR / 2; off > 0; off >>= 1) ss += __shfl_xor_sync(mask, ss, off);
      inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(|x|, eps)
    }
    const float a = ep.alpha * inv;
    float4* o = ep.O != nullptr ? reinterpret_cast<float4*>(ep.O + row * ep.ldo) : nullptr;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int idx = sl + k * LPR;
      if ((!GUARD || idx < dvec) && active) {
        float4 r = f4_zero();
#pragma unroll 1
        for (int q = 0; q < ep.n_add; ++q) {
          const float4 z = __ldg(reinterpret_cast<const float4*>(ep.add[q] + row * ep.ldo) + idx);
          f4_fma(r, ep.beta[q], z);
        }
        f4_fma(r, a, acc[k]);
        r.x *= ep.post; r.y *= ep.post; r.z *= ep.post; r.w *= ep.post;
        if (o != nullptr) o[idx] = r;
        if (ep.ad_p != nullptr) {
          float4* p4 = reinterpret_cast<float4*>(ep.ad_p + row * ep.ldo) + idx;
          float4* m4 = reinterpret_cast<float4*>(ep.ad_m + row * ep.ldo) + idx;
          float4* v4 = reinterpret_cast<float4*>(ep.ad_v + row * ep.ldo) + idx;
          float4 p = *p4, mm = *m4, vv = *v4;
          adam_update(p.x, r.x, mm.x, vv.x, ep.adam);
          adam_update(p.y, r.y, mm.y, vv.y, ep.adam);
          adam_update(p.z, r.z, mm.z, vv.z, ep.adam);
          adam_update(p.w, r.w, mm.w, vv.w, ep.adam);
          *p4 = p; *m4 = mm; *v4 = vv;
        }
      }
    }
  }
}

// Long rows (power-law hubs): one warp per chunk of `lp.chunk` entries; the last-arriving chunk of a row reduces the
// partial sums in chunk order (deterministic) and runs the epilogue.
template <int LPR, int VPL, int UNR, bool GUARD, int PPL, bool NA = false>
__device__ __forceinline__ void long_chunk_path(const int* __restrict__ row_ptr, const int* __restrict__ col_idx,
                                                const float* __restrict__ vals, const float* __restrict__ X,
                                                long long ldx, int dvec, const Epi& ep, const LongPlan& lp, int warp,
                                                int lane, int sub, int sl, unsigned mask, float4 (&acc)[VPL]) {
  constexpr int RPW = 32 / LPR;
  // ---- long-row chunk: one warp per chunk ----
  const int chunk_id = blockIdx.x * kWarpsPerBlock + warp;
  if (chunk_id >= lp.n_chunks) return;
  const int L = lp.chunk_long[chunk_id];
  const int row = lp.long_rows[L];
  const int c0 = lp.long_chunk_ptr[L];
  const int nck = lp.long_chunk_ptr[L + 1] - c0;
  const int rs = row_ptr[row], re = row_ptr[row + 1];
  const int s = rs + (chunk_id - c0) * lp.chunk;
  const int e = min(s + lp.chunk, re);
  if constexpr (PPL > 1)
    accumulate_multi<LPR, PPL, NA>(col_idx, vals, s + sub * LPR, e, LPR * RPW, X, ldx, sl, mask, acc[0]);
  else
    accumulate<LPR, VPL, UNR, GUARD>(col_idx, vals, s + sub * LPR, e, LPR * RPW, X, ldx, sl, mask, dvec, acc);
  __syncwarp();
#pragma unroll
  for (int off = LPR; off < 32; off <<= 1)
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      acc[k].x += __shfl_xor_sync(0xffffffffu, acc[k].x, off);
      acc[k].y += __shfl_xor_sync(0xffffffffu, acc[k].y, off);
      acc[k].z += __shfl_xor_sync(0xffffffffu, acc[k].z, off);
      acc[k].w += __shfl_xor_sync(0xffffffffu, acc[k].w, off);
    }
  const long long dpad = (long long)dvec * 4;
  if (sub == 0) {
    float4* part = reinterpret_cast<float4*>(lp.partial + (long long)chunk_id * dpad);
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int idx = sl + k * LPR;
      if (!GUARD || idx < dvec) part[idx] = acc[k];
    }
  }
  __threadfence();
  __syncwarp();
  int last = 0;
  if (lane == 0) last = (atomicAdd(lp.counters + L, 1) == nck - 1);
  last = __shfl_sync(0xffffffffu, last, 0);
  if (!last) return;
  __threadfence();
  if (sub == 0) {
    float4 tot[VPL];
#pragma unroll
    for (int k = 0; k < VPL; ++k) tot[k] = f4_zero();
    for (int q = 0; q < nck; ++q) {
      const float4* p = reinterpret_cast<const float4*>(lp.partial + (long long)(c0 + q) * dpad);
#pragma unroll
      for (int k = 0; k < VPL; ++k) {
        const int idx = sl + k * LPR;
        if (!GUARD || idx < dvec) f4_add(tot[k], __ldcg(p + idx));
      }
    }
    finish_row<LPR, VPL, GUARD>(ep, row, sl, mask, dvec, tot);
  }
  if (lane == 0) lp.counters[L] = 0;  // self-resetting ticket for the next launch
  return;
}

template <int LPR, int VPL, int UNR, bool GUARD, int MINB, int PPL = 1, bool NA = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
spmm_csr_kernel(const int* __restrict__ row_ptr, const int* __restrict__ col_idx, const float* __restrict__ vals,
                long long n_rows, const float* __restrict__ X, long long ldx, int dvec, Epi ep, LongPlan lp,
                int long_blocks) {
  constexpr int RPW = 32 / LPR;  // rows per warp
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));

  float4 acc[VPL];
#pragma unroll
  for (int k = 0; k < VPL; ++k) acc[k] = f4_zero();

  if ((int)blockIdx.x < long_blocks) {
    long_chunk_path<LPR, VPL, UNR, GUARD, PPL, NA>(row_ptr, col_idx, vals, X, ldx, dvec, ep, lp, warp, lane, sub, sl, mask, acc);
    return;
  }

  // ---- regular rows: one sub-warp per row ----
  const long long row = ((long long)(blockIdx.x - long_blocks) * kWarpsPerBlock + warp) * RPW + sub;
  if (row >= n_rows) return;
  const int s = row_ptr[row], e = row_ptr[row + 1];
  if (lp.n_long > 0 && e - s > lp.chunk) return;  // owned by the long path
  if constexpr (PPL > 1)
    accumulate_multi<LPR, PPL, NA>(col_idx, vals, s, e, LPR, X, ldx, sl, mask, acc[0]);
  else
    accumulate<LPR, VPL, UNR, GUARD>(col_idx, vals, s, e, LPR, X, ldx, sl, mask, dvec, acc);
  finish_row<LPR, VPL, GUARD>(ep, row, sl, mask, dvec, acc);
}

template <int LPR, int VPL, int UNR, bool GUARD, int MINB, int PPL = 1, bool NA = false>
static int launch(const gcf_csr_t* A, const float* X, long long ldx, int dvec, const Epi& ep, const LongPlan& lp,
                  cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int long_blocks = (int)cdiv(lp.n_chunks, kWarpsPerBlock);
  const long long short_blocks = cdiv(A->n_rows, (long long)kWarpsPerBlock * RPW);
  const long long grid = long_blocks + short_blocks;
  if (grid <= 0) return GCF_OK;
  GCF_REQUIRE(grid < 2147483647LL, "gcf_spmm_csr_f32: grid too large");
  static_assert(PPL == 1 || (VPL == 1 && !GUARD), "multi-pair batches are for exact narrow rows");
  spmm_csr_kernel<LPR, VPL, UNR, GUARD, MINB, PPL, NA><<<(unsigned)grid, kWarpsPerBlock * 32, 0, st>>>(
      A->row_ptr, A->col_idx, A->vals, A->n_rows, X, ldx, dvec, ep, lp, long_blocks);
  GCF_LAUNCH_CHECK("spmm_csr_kernel");
  return GCF_OK;
}

// ---- flat-stream kernel ------------------------------------------------------------------------------------------
struct TilePlan { const int2* tiles; int n_tiles; const int* rp; const int* ids; const int* empty_rows; int n_empty; };

__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }

// The epilogue of a row reads rows of other matrices (addends; parameter and moments for the fused Adam).  Waiting for
// them would stall the warp -- both of its sub-warps -- for a DRAM round trip per output row, so they are requested
// into L1 when the row is stashed, one group of gathers before the epilogue runs.
__device__ __forceinline__ void prefetch_epilogue(const Epi& ep, long long row, int sl) {
  const long long off = row * ep.ldo + sl * 4;
#pragma unroll 1
  for (int q = 0; q < ep.n_add; ++q) prefetch_l1(ep.add[q] + off);
  if (ep.ad_p != nullptr) { prefetch_l1(ep.ad_p + off); prefetch_l1(ep.ad_m + off); prefetch_l1(ep.ad_v + off); }
}

// Rows are addressed in the COMPACT numbering of the non-empty rows (tp.rp = their row pointer, tp.ids = their row
// ids, NULL when no row is empty): every row of a tile then has at least one entry, so at most LPR rows end inside a
// batch of LPR entries and they all sit in one LPR-wide window of tp.rp: lane t of a sub-warp looks at row k + t and,
// if that row ends inside the batch, writes "row id + 1" at the batch position of its last entry (mark slab).  The
// window of the NEXT batch is requested as soon as the number of rows ending in this one is known (a ballot), i.e.
// a whole batch before it is needed; the same holds for the (col, val) pairs.  Nothing the gathers depend on is ever
// waited for: per batch the sub-warp exposes one memory round trip per group of UNR gathers and nothing else.
//
// The kernel is warp-synchronous: the RPW sub-warps of a warp stream different tiles but run the same number of
// batches (the longer tile's; tiles hold tile_nnz .. tile_nnz + chunk entries), so every barrier and shuffle uses the
// full mask.
//
// STASH = false: the epilogue is a plain store of the row (Y only) and is issued on the spot.
// STASH = true : completed rows are parked in shared memory (each lane keeps its own float4 slice) and the full
//                epilogue -- linear combination with addends, row-L2-normalise, fused Adam -- runs between two groups
//                of gathers, where no gathered row is live in registers; at most UNR rows complete per group.
// Narrow rows (d/G = 8, 16, 32 floats in the multi-GPU layouts) have LPR = 2, 4, 8 lanes per row: a batch then holds
// BS = LPR * PPL = 16 entries, every lane carrying PPL (col, val) pairs and PPL rows of the row window.
// HUB: the tiles read their columns from col_hub, a copy of col_idx in which bit 31 marks the entries whose column is one
// of the most-referenced ("hub") columns of that half of the operator (graph.py: CSRGraph hub flags).  Hub rows are gathered
// with L1::evict_last, all other rows with L1::no_allocate: the SM's L1 then holds the hub rows only -- a demand-filled,
// per-SM staging of the hub rows in the L1 / shared-memory SRAM, shared by the CTAs resident on the SM.  Worth it when
// the working set is L2-resident (the launch is then bound by the L2 -> SM path, not by HBM).
template <bool HUB>
__device__ __forceinline__ float4 gather_row(const float4* __restrict__ Xl, int c, unsigned ldx4) {
  if constexpr (!HUB) {
    return __ldg(Xl + (unsigned long long)(unsigned)c * ldx4);
  } else {
    const float4* p = Xl + (unsigned long long)((unsigned)c & 0x7fffffffu) * ldx4;
    return c < 0 ? ldg_f4_l1_keep(p) : ldg_f4_l1_bypass(p);
  }
}

template <int LPR, int PPL, int UNR, bool GUARD, bool STASH, int MINB, int LONG_UNR, int LONG_PPL, bool HUB = false>
__global__ void __launch_bounds__(kWarpsPerBlock * 32, MINB)
spmm_flat_kernel(const int* __restrict__ col_idx, const float* __restrict__ vals, const float* __restrict__ X,
                 long long ldx, int dvec, const __grid_constant__ Epi ep, LongPlan lp, const int* __restrict__ row_ptr,
                 int long_blocks, int tile_blocks, TilePlan tp, const int* __restrict__ col_hub) {
  constexpr int RPW = 32 / LPR;
  constexpr int BS = LPR * PPL;            // entries per batch of one sub-warp
  constexpr int SLOTS = STASH ? UNR : 1;
  constexpr unsigned kFull = 0xffffffffu;
  static_assert(BS % UNR == 0 && UNR % 4 == 0, "batches are consumed UNR gathers at a time, four columns per LDS.128");
  __shared__ __align__(16) int col_slab[kWarpsPerBlock][RPW * BS];
  __shared__ __align__(16) float val_slab[kWarpsPerBlock][RPW * BS];
  __shared__ __align__(16) int mark_slab[kWarpsPerBlock][RPW * BS];
  __shared__ __align__(16) float4 stash[STASH ? kWarpsPerBlock : 1][SLOTS][32];
  __shared__ int stash_row[STASH ? kWarpsPerBlock : 1][SLOTS][RPW];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR;
  const int sl = lane % LPR;
  const unsigned mask = (LPR == 32) ? kFull : (((1u << LPR) - 1u) << (sub * LPR));

  if ((int)blockIdx.x < long_blocks) {
    float4 acc[1] = {f4_zero()};
    long_chunk_path<LPR, 1, LONG_UNR, GUARD, LONG_PPL>(row_ptr, col_idx, vals, X, ldx, dvec, ep, lp, warp, lane, sub, sl, mask, acc);
    return;
  }
  if ((int)blockIdx.x >= long_blocks + tile_blocks) {
    // rows without entries: the epilogue of a zero sum
    const int q = (((int)blockIdx.x - long_blocks - tile_blocks) * kWarpsPerBlock + warp) * RPW + sub;
    if (q >= tp.n_empty) return;
    float4 acc[1] = {f4_zero()};
    finish_row<LPR, 1, GUARD>(ep, __ldg(tp.empty_rows + q), sl, mask, dvec, acc);
    return;
  }
  const int tile = (((int)blockIdx.x - long_blocks) * kWarpsPerBlock + warp) * RPW + sub;
  const bool has_tile = tile < tp.n_tiles;
  int k = 0, k1 = 0, j = 0, jend = 0;      // running / last compact row, running / last entry
  if (has_tile) {
    const int2 kk = __ldg(tp.tiles + tile);
    k = kk.x;
    k1 = kk.y;
    j = __ldg(tp.rp + k);
    jend = __ldg(tp.rp + k1);
  }
  int nb = (jend - j + BS - 1) / BS;       // batches of this sub-warp; the warp runs the longest
  if (RPW > 1) nb = __reduce_max_sync(kFull, nb);
  int* cs = &col_slab[warp][sub * BS];
  float* vs = &val_slab[warp][sub * BS];
  int* ms = &mark_slab[warp][sub * BS];
  const bool col_ok = !GUARD || sl < dvec;
  const float4* Xl = reinterpret_cast<const float4*>(X) + sl;   // this lane's float4 column of X
  const unsigned ldx4 = (unsigned)(ldx >> 2);   // 32 x 32 -> 64-bit address products (one IMAD.WIDE.U32 each)
  const unsigned ldy4 = (unsigned)(ep.ldy >> 2);
  const bool renumbered = tp.ids != nullptr;
  const int* __restrict__ cols = HUB ? col_hub : col_idx;

  // (col, val) of batch 0 and the row window of batch 0.  Slots past the tile's last entry repeat its last column
  // with weight 0: their gathers hit L1 and cannot bring a non-finite value into a row that does not already hold it.
  int c[PPL], my_end[PPL], my_row[PPL];
  float v[PPL];
#pragma unroll
  for (int p = 0; p < PPL; ++p) {
    c[p] = 0; v[p] = 0.f; my_end[p] = 0; my_row[p] = 0;
    if (has_tile) {
      const int t = p * LPR + sl;
      const int jj = min(j + t, jend - 1);
      c[p] = ld_stream_i32(cols + jj);
      if (j + t < jend) v[p] = ld_stream_f32(vals + jj);
      my_end[p] = __ldg(tp.rp + min(k + 1 + t, k1));
      my_row[p] = renumbered ? __ldg(tp.ids + min(k + t, k1 - 1)) : k + t;
    }
  }
  float4 acc = f4_zero();
  for (int b = 0; b < nb; ++b, j += BS) {
    const int cnt = max(0, min(BS, jend - j));
    __syncwarp();                          // the previous batch has been read out of the slabs
#pragma unroll
    for (int p = 0; p < PPL; ++p) {
      cs[p * LPR + sl] = c[p];
      vs[p * LPR + sl] = v[p];
      ms[p * LPR + sl] = 0;
    }
    __syncwarp();
#pragma unroll
    for (int p = 0; p < PPL; ++p) {        // (col, val) of the next batch
      v[p] = 0.f;
      const int t = BS + p * LPR + sl;
      if (j + BS < jend) {
        const int jj = min(j + t, jend - 1);
        c[p] = ld_stream_i32(cols + jj);
        if (j + t < jend) v[p] = ld_stream_f32(vals + jj);
      }
    }
    float4 x[UNR];
#pragma unroll
    for (int t0 = 0; t0 < BS; t0 += UNR) {
      if (t0 < cnt) {
#pragma unroll
        for (int u = 0; u < UNR; u += 4) {
          const int4 cc = *reinterpret_cast<const int4*>(cs + t0 + u);   // four columns per broadcast LDS.128
          x[u + 0] = col_ok ? gather_row<HUB>(Xl, cc.x, ldx4) : f4_zero();
          x[u + 1] = col_ok ? gather_row<HUB>(Xl, cc.y, ldx4) : f4_zero();
          x[u + 2] = col_ok ? gather_row<HUB>(Xl, cc.z, ldx4) : f4_zero();
          x[u + 3] = col_ok ? gather_row<HUB>(Xl, cc.w, ldx4) : f4_zero();
        }
      }
      if (t0 == 0) {
        // with the first gathers in flight: rows that end inside this batch, and the window of the next one
        int n_done = 0;
#pragma unroll
        for (int p = 0; p < PPL; ++p) {
          const int t = p * LPR + sl;
          const int pos = my_end[p] - 1 - j;
          const bool ends_here = has_tile && k + t < k1 && pos < BS;
          if (ends_here) ms[pos] = my_row[p] + 1;
          n_done += __popc(__ballot_sync(kFull, ends_here) & mask);
        }
        k += n_done;
        if (has_tile) {
#pragma unroll
          for (int p = 0; p < PPL; ++p) {
            const int t = p * LPR + sl;
            my_end[p] = __ldg(tp.rp + min(k + 1 + t, k1));
            my_row[p] = renumbered ? __ldg(tp.ids + min(k + t, k1 - 1)) : k + t;
          }
        }
        __syncwarp();
      }
      int ns = 0;                          // rows parked in the stash by this group
      if (t0 < cnt) {
#pragma unroll
        for (int u = 0; u < UNR; u += 4) {
          const float4 ww = *reinterpret_cast<const float4*>(vs + t0 + u);
          const int4 mm = *reinterpret_cast<const int4*>(ms + t0 + u);
          const float w[4] = {ww.x, ww.y, ww.z, ww.w};
          const int m[4] = {mm.x, mm.y, mm.z, mm.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            f4_fma(acc, w[q], x[u + q]);
            if (m[q] != 0) {               // last entry of row m - 1
              if constexpr (STASH) {
                stash[warp][ns][lane] = acc;
                stash_row[warp][ns][sub] = m[q] - 1;
                if (col_ok) prefetch_epilogue(ep, m[q] - 1, sl);
                ++ns;
              } else {
                if (col_ok) (reinterpret_cast<float4*>(ep.Y) + sl)[(unsigned long long)(unsigned)(m[q] - 1) * ldy4] = acc;
              }
              acc = f4_zero();
            }
          }
        }
      }
      if constexpr (STASH) {
        const int ns_max = (RPW > 1) ? __reduce_max_sync(kFull, ns) : ns;
#pragma unroll 1
        for (int i = 0; i < ns_max; ++i) {
          const bool active = i < ns;
          float4 a[1] = {stash[warp][active ? i : 0][lane]};
          finish_row<LPR, 1, GUARD>(ep, stash_row[warp][active ? i : 0][sub], sl, kFull, dvec, a, active);
        }
      }
    }
  }
}

template <int LPR, int PPL, int UNR, bool GUARD, bool STASH, int MINB, int LONG_UNR, int LONG_PPL, bool HUB = false>
static int launch_flat(const gcf_csr_t* A, const float* X, long long ldx, int dvec, const Epi& ep, const LongPlan& lp,
                       cudaStream_t st) {
  constexpr int RPW = 32 / LPR;
  const int long_blocks = (int)cdiv(lp.n_chunks, kWarpsPerBlock);
  const long long tile_blocks = cdiv(A->n_tiles, (long long)kWarpsPerBlock * RPW);
  const long long empty_blocks = cdiv(A->n_empty, (long long)kWarpsPerBlock * RPW);
  const long long grid = long_blocks + tile_blocks + empty_blocks;
  if (grid <= 0) return GCF_OK;
  GCF_REQUIRE(grid < 2147483647LL, "gcf_spmm_csr_f32: grid too large");
  TilePlan tp{reinterpret_cast<const int2*>(A->tiles), A->n_tiles, A->n_empty > 0 ? A->nz_row_ptr : A->row_ptr,
              A->n_empty > 0 ? A->nz_rows : nullptr, A->empty_rows, A->n_empty};
  static const int pad_smem = getenv("GCF_SPMM_PAD_SMEM") ? atoi(getenv("GCF_SPMM_PAD_SMEM")) : 0;   // L1-capacity probe
  spmm_flat_kernel<LPR, PPL, UNR, GUARD, STASH, MINB, LONG_UNR, LONG_PPL, HUB><<<(unsigned)grid, kWarpsPerBlock * 32, pad_smem, st>>>(
      A->col_idx, A->vals, X, ldx, dvec, ep, lp, A->row_ptr, long_blocks, (int)tile_blocks, tp, A->hub_col_idx);
  GCF_LAUNCH_CHECK("spmm_flat_kernel");
  return GCF_OK;
}

// full-width rows: one (col, val) pair per lane, hub chunks on the r01 inner loop with 8 gathers in flight
template <int LPR, int UNR, bool GUARD, int MINB, bool HUB = false>
static int launch_flat_cls(bool plain, const gcf_csr_t* A, const float* X, long long ldx, int dvec, const Epi& ep,
                           const LongPlan& lp, cudaStream_t st) {
  if (plain) return launch_flat<LPR, 1, UNR, GUARD, false, MINB, 8, 1, HUB>(A, X, ldx, dvec, ep, lp, st);
  return launch_flat<LPR, 1, (UNR > 8 ? 8 : UNR), GUARD, true, MINB, 8, 1, HUB>(A, X, ldx, dvec, ep, lp, st);   // stash: UNR slots of a row each
}
// narrow rows: 16-entry batches (8-entry batches at 2 lanes per row: 16 sub-warps per warp share the 48 KB)
template <int LPR, int MINB, int LONG_UNR, int LONG_PPL>
static int launch_flat_narrow(bool plain, const gcf_csr_t* A, const float* X, long long ldx, int dvec, const Epi& ep,
                              const LongPlan& lp, cudaStream_t st) {
  constexpr int PPL = LPR == 2 ? 4 : 16 / LPR;
  if (plain) return launch_flat<LPR, PPL, 8, false, false, MINB, LONG_UNR, LONG_PPL>(A, X, ldx, dvec, ep, lp, st);
  return launch_flat<LPR, PPL, 8, false, true, MINB, LONG_UNR, LONG_PPL>(A, X, ldx, dvec, ep, lp, st);
}

__global__ void axpby_kernel(float4* __restrict__ out, const float4* __restrict__ a, float sa,
                             const float4* __restrict__ b, float sb, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    float4 x = a[i];
    float4 r = make_float4(sa * x.x, sa * x.y, sa * x.z, sa * x.w);
    if (b != nullptr) {
      float4 y = b[i];
      r.x = fmaf(sb, y.x, r.x); r.y = fmaf(sb, y.y, r.y); r.z = fmaf(sb, y.z, r.z); r.w = fmaf(sb, y.w, r.w);
    }
    out[i] = r;
  }
}

}  // namespace gcf

using namespace gcf;

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

extern "C" size_t gcf_spmm_counter_offset(const gcf_csr_t* A, int32_t d) {
  if (A == nullptr || A->n_long <= 0) return 0;
  return align_up((size_t)A->n_chunks * (size_t)d * sizeof(float));
}

extern "C" size_t gcf_spmm_workspace_bytes(const gcf_csr_t* A, int32_t d) {
  if (A == nullptr || A->n_long <= 0) return 0;
  return gcf_spmm_counter_offset(A, d) + align_up((size_t)A->n_long * sizeof(int32_t));
}

namespace gcf {
struct AdamFuse { float* param; float* m; float* v; AdamArgs args; };
}

static int spmm_impl(const gcf_csr_t* A, int32_t d, const float* X, int64_t ldx, float* Y, int64_t ldy, float* OUT,
                     int64_t ld_out, int32_t epilogue, float alpha, float post, int32_t n_addends,
                     const float* const* addends, const float* betas, void* workspace, size_t workspace_bytes,
                     int32_t variant, gcf_stream_t stream, const AdamFuse* adam) {
  GCF_REQUIRE(A != nullptr && X != nullptr, "gcf_spmm_csr_f32: null operator or X");
  GCF_REQUIRE(A->n_rows >= 0 && A->n_cols >= 0, "gcf_spmm_csr_f32: negative shape");
  GCF_REQUIRE(Y != nullptr || OUT != nullptr || adam != nullptr, "gcf_spmm_csr_f32: no output requested");
  if (d <= 0 || (d & 3) != 0 || d > 1024) {
    set_error("gcf_spmm_csr_f32: d=%d unsupported (need d %% 4 == 0 and d <= 1024)", d);
    return GCF_EUNSUPPORTED;
  }
  GCF_REQUIRE(ldx >= d && (ldx & 3) == 0 && aligned16(X), "gcf_spmm_csr_f32: X must be 16B aligned with ld %% 4 == 0");
  GCF_REQUIRE(Y == nullptr || (ldy >= d && (ldy & 3) == 0 && aligned16(Y)), "gcf_spmm_csr_f32: bad Y alignment/ld");
  GCF_REQUIRE(OUT == nullptr || (ld_out >= d && (ld_out & 3) == 0 && aligned16(OUT)), "gcf_spmm_csr_f32: bad OUT alignment/ld");
  GCF_REQUIRE(n_addends >= 0 && n_addends <= GCF_MAX_ADDENDS, "gcf_spmm_csr_f32: n_addends out of range");
  GCF_REQUIRE(n_addends == 0 || (addends != nullptr && betas != nullptr && (OUT != nullptr || adam != nullptr)),
              "gcf_spmm_csr_f32: addends need OUT, pointer and beta arrays");
  GCF_REQUIRE(adam == nullptr || (ld_out >= d && (ld_out & 3) == 0 && aligned16(adam->param) && aligned16(adam->m) && aligned16(adam->v)),
              "gcf_spmm_csr_f32: fused Adam needs 16B aligned tables and their leading dimension in ld_out");
  GCF_REQUIRE(epilogue == GCF_EPILOGUE_NONE || epilogue == GCF_EPILOGUE_L2NORM, "gcf_spmm_csr_f32: bad epilogue");
  if (A->n_rows == 0) return GCF_OK;
  GCF_REQUIRE(A->row_ptr != nullptr, "gcf_spmm_csr_f32: null row_ptr");

  Epi ep;
  ep.Y = Y; ep.ldy = ldy; ep.O = OUT; ep.ldo = ld_out;
  ep.epilogue = epilogue; ep.alpha = alpha; ep.post = post; ep.n_add = n_addends;
  ep.ad_p = ep.ad_m = ep.ad_v = nullptr;
  ep.adam = AdamArgs{};
  if (adam != nullptr) { ep.ad_p = adam->param; ep.ad_m = adam->m; ep.ad_v = adam->v; ep.adam = adam->args; }
  for (int q = 0; q < GCF_MAX_ADDENDS; ++q) { ep.add[q] = nullptr; ep.beta[q] = 0.f; }
  for (int q = 0; q < n_addends; ++q) {
    GCF_REQUIRE(addends[q] != nullptr && aligned16(addends[q]), "gcf_spmm_csr_f32: addend %d null or misaligned", q);
    ep.add[q] = addends[q];
    ep.beta[q] = betas[q];
  }

  LongPlan lp;
  lp.chunk = 0; lp.n_long = 0; lp.n_chunks = 0;
  lp.long_rows = nullptr; lp.long_chunk_ptr = nullptr; lp.chunk_long = nullptr; lp.partial = nullptr; lp.counters = nullptr;
  if (A->n_long > 0) {
    GCF_REQUIRE(A->chunk > 0 && A->n_chunks > 0 && A->long_rows && A->long_chunk_ptr && A->chunk_long,
                "gcf_spmm_csr_f32: incomplete long-row schedule");
    const size_t need = gcf_spmm_workspace_bytes(A, d);
    if (workspace == nullptr || workspace_bytes < need) {
      set_error("gcf_spmm_csr_f32: workspace too small (%zu < %zu)", workspace_bytes, need);
      return GCF_EWORKSPACE;
    }
    lp.chunk = A->chunk; lp.n_long = A->n_long; lp.n_chunks = A->n_chunks;
    lp.long_rows = A->long_rows; lp.long_chunk_ptr = A->long_chunk_ptr; lp.chunk_long = A->chunk_long;
    lp.partial = reinterpret_cast<float*>(workspace);
    lp.counters = reinterpret_cast<int*>(static_cast<char*>(workspace) + gcf_spmm_counter_offset(A, d));
  }

  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int dvec = d / 4;
  // flat-stream kernel: default whenever the operator carries a tile schedule and a row fits one float4 per lane.
  // variant 0 = default, 4 = the r01 row-walking kernel (kept for A/B runs), 10.. = flat-kernel tuning points
  // Rows of 8 / 16 floats (8- and 4-way feature-sharded slices) stay on the row-walking kernel unless a flat variant is
  // asked for: measured on cfg5 (gpurun_out/r02/exp_narrow_b_d*.log) flat vs row-walking 3.45 vs 2.54 ms (d = 8),
  // 4.10 vs 3.70 ms (d = 16), but 4.12 vs 5.59 ms at d = 32 and 7.02 vs 9.09 ms at d = 64.
  const bool flat_default = d == 32 || (d >= 36 && d <= 128);
  const bool flat_capable = flat_default || d == 8 || d == 16;
  // Launches with an epilogue (addends / normalise / Adam) on operators whose X table sits in the L2 go back to the row-walking
  // kernel: the flat kernel parks finished rows in a shared-memory stash and pays ~2x the instructions for it, which only
  // hides behind DRAM time.  Measured (gpurun_out/r03/exp_spmm_small_epilogue.log, one addend): cfg1 79.9 vs 63.6 us, cfg2
  // 184 vs 149 us, cfg3 (d = 128) 199 vs 125 us, cfg4 330 vs 301 us -- but cfg5 (3.84 GB table) 8.31 vs 10.1 ms.  Both kernels
  // sum in the same order: the choice does not change a bit of the result.
  const bool plain_launch = adam == nullptr && epilogue == GCF_EPILOGUE_NONE && OUT == nullptr;
  const bool table_in_l2 = (long long)A->n_cols * d * 4 <= (256LL << 20);
  const bool flat_ok = plain_launch || !table_in_l2;
  if (A->n_tiles > 0 && A->tiles != nullptr && ((variant == 0 && flat_default && flat_ok) || (variant >= 10 && flat_capable))) {
    GCF_REQUIRE(A->n_empty == 0 || (A->empty_rows != nullptr && A->nz_row_ptr != nullptr && A->nz_rows != nullptr),
                "gcf_spmm_csr_f32: operator has empty rows but no compact row numbering");
    const bool plain = adam == nullptr && epilogue == GCF_EPILOGUE_NONE && OUT == nullptr;   // Y = A X and nothing else
    // feature-sharded slices; hub chunks keep the r01 narrow-row inner loops (gathers in flight x pairs per lane)
    if (d == 8) {
      if (variant == 10) return launch_flat_narrow<2, 4, 2, 4>(plain, A, X, ldx, dvec, ep, lp, st);
      return launch_flat_narrow<2, 3, 2, 4>(plain, A, X, ldx, dvec, ep, lp, st);     // variant 11..
    }
    if (d == 16) {
      if (variant == 10) return launch_flat_narrow<4, 4, 4, 2>(plain, A, X, ldx, dvec, ep, lp, st);
      return launch_flat_narrow<4, 3, 4, 2>(plain, A, X, ldx, dvec, ep, lp, st);     // variant 11..
    }
    if (d == 32) {
      if (variant == 10) return launch_flat_narrow<8, 4, 8, 1>(plain, A, X, ldx, dvec, ep, lp, st);
      return launch_flat_narrow<8, 3, 8, 1>(plain, A, X, ldx, dvec, ep, lp, st);
    }
    // measured on cfg1 / cfg5 (profiles/r02_exp_spmm_*.log): 8 gathers in flight per sub-warp at 3 CTAs / SM
    // hub flags present (L2-resident operators, graph.py): hub rows kept in L1, cold rows bypass it.  variant 15 = the
    // same kernel without the flags (A/B)
    if (A->hub_col_idx != nullptr && variant != 15) {
      if (d == 64) {
        if (variant == 16) return launch_flat_cls<16, 8, false, 3, true>(plain, A, X, ldx, dvec, ep, lp, st);
        return launch_flat_cls<16, 8, false, 4, true>(plain, A, X, ldx, dvec, ep, lp, st);
      }
      if (d == 128) return launch_flat_cls<32, 8, false, 3, true>(plain, A, X, ldx, dvec, ep, lp, st);
    }
    if (d == 64) {
      if (variant == 10) return launch_flat_cls<16, 8, false, 4>(plain, A, X, ldx, dvec, ep, lp, st);
      if (variant == 11) return launch_flat_cls<16, 16, false, 3>(plain, A, X, ldx, dvec, ep, lp, st);
      if (variant == 13) return launch_flat_cls<16, 4, false, 5>(plain, A, X, ldx, dvec, ep, lp, st);
      if (variant == 14) return launch_flat_cls<16, 16, false, 2>(plain, A, X, ldx, dvec, ep, lp, st);
      return launch_flat_cls<16, 8, false, 3>(plain, A, X, ldx, dvec, ep, lp, st);
    }
    if (d == 128) {
      if (variant == 10) return launch_flat_cls<32, 8, false, 4>(plain, A, X, ldx, dvec, ep, lp, st);
      if (variant == 11) return launch_flat_cls<32, 16, false, 3>(plain, A, X, ldx, dvec, ep, lp, st);
      return launch_flat_cls<32, 8, false, 3>(plain, A, X, ldx, dvec, ep, lp, st);
    }
    if (d <= 64) return launch_flat_cls<16, 8, true, 3>(plain, A, X, ldx, dvec, ep, lp, st);
    return launch_flat_cls<32, 8, true, 3>(plain, A, X, ldx, dvec, ep, lp, st);
  }
  if (variant == 4) variant = 0;
  switch (d) {
    // feature-sharded slices (d / G columns per rank).  Measured on the cfg5 graph (profiles/r01_exp_narrow_cfg5.log):
    // d=8 4.23 -> 2.55 ms, d=16 4.65 -> 3.70 ms with 8 gathers in flight per sub-warp; d=32 6.85 -> 5.59 ms at 4 CTAs/SM
    case 8:
      if (variant == 1) return launch<2, 1, 2, false, 4>(A, X, ldx, dvec, ep, lp, st);      // 2 gathers in flight
      if (variant == 2) return launch<2, 1, 2, false, 2, 8>(A, X, ldx, dvec, ep, lp, st);   // 16
      if (variant == 3) return launch<2, 1, 2, false, 6, 2>(A, X, ldx, dvec, ep, lp, st);   // 4, 48 warps / SM
      if (variant == 5) return launch<2, 1, 2, false, 4, 4, true>(A, X, ldx, dvec, ep, lp, st);   // 8, gathers bypass the L1
      if (variant == 6) return launch<2, 1, 2, false, 3, 8, true>(A, X, ldx, dvec, ep, lp, st);   // 16, gathers bypass the L1
      return launch<2, 1, 2, false, 4, 4>(A, X, ldx, dvec, ep, lp, st);                     // 8
    case 16:
      if (variant == 1) return launch<4, 1, 4, false, 4>(A, X, ldx, dvec, ep, lp, st);      // 4 gathers in flight
      if (variant == 2) return launch<4, 1, 4, false, 2, 4>(A, X, ldx, dvec, ep, lp, st);   // 16
      if (variant == 3) return launch<4, 1, 4, false, 6>(A, X, ldx, dvec, ep, lp, st);      // 4, 48 warps / SM
      if (variant == 5) return launch<4, 1, 4, false, 4, 2, true>(A, X, ldx, dvec, ep, lp, st);   // 8, gathers bypass the L1
      if (variant == 6) return launch<4, 1, 4, false, 3, 4, true>(A, X, ldx, dvec, ep, lp, st);   // 16, gathers bypass the L1
      return launch<4, 1, 4, false, 4, 2>(A, X, ldx, dvec, ep, lp, st);                     // 8
    case 32:
      if (variant == 1) return launch<8, 1, 8, false, 2, 2>(A, X, ldx, dvec, ep, lp, st);   // 16 gathers in flight
      if (variant == 2) return launch<8, 1, 8, false, 3>(A, X, ldx, dvec, ep, lp, st);
      return launch<8, 1, 8, false, 4>(A, X, ldx, dvec, ep, lp, st);
    case 64:  // variants are tuning knobs (UNR loads in flight x resident blocks), same arithmetic
      if (variant == 1) return launch<16, 1, 4, false, 4>(A, X, ldx, dvec, ep, lp, st);
      if (variant == 2) return launch<16, 1, 16, false, 2>(A, X, ldx, dvec, ep, lp, st);
      if (variant == 3) return launch<16, 1, 8, false, 3>(A, X, ldx, dvec, ep, lp, st);
      return launch<16, 1, 8, false, 4>(A, X, ldx, dvec, ep, lp, st);  // measured best on cfg1 and cfg5
    case 128:
      if (variant == 1) return launch<32, 1, 4, false, 4>(A, X, ldx, dvec, ep, lp, st);
      if (variant == 2) return launch<32, 1, 16, false, 2>(A, X, ldx, dvec, ep, lp, st);
      if (variant == 3) return launch<32, 1, 8, false, 3>(A, X, ldx, dvec, ep, lp, st);
      return launch<32, 1, 8, false, 4>(A, X, ldx, dvec, ep, lp, st);
    case 256: return launch<32, 2, 4, false, 2>(A, X, ldx, dvec, ep, lp, st);
    default: break;
  }
  if (d <= 128) return launch<32, 1, 8, true, 3>(A, X, ldx, dvec, ep, lp, st);
  if (d <= 256) return launch<32, 2, 4, true, 2>(A, X, ldx, dvec, ep, lp, st);
  if (d <= 512) return launch<32, 4, 2, true, 1>(A, X, ldx, dvec, ep, lp, st);
  return launch<32, 8, 2, true, 1>(A, X, ldx, dvec, ep, lp, st);
}

extern "C" int gcf_spmm_csr_f32(const gcf_csr_t* A, int32_t d, const float* X, int64_t ldx, float* Y, int64_t ldy,
                                float* OUT, int64_t ld_out, int32_t epilogue, float alpha, float post,
                                int32_t n_addends, const float* const* addends, const float* betas, void* workspace,
                                size_t workspace_bytes, int32_t variant, gcf_stream_t stream) {
  return spmm_impl(A, d, X, ldx, Y, ldy, OUT, ld_out, epilogue, alpha, post, n_addends, addends, betas, workspace,
                   workspace_bytes, variant, stream, nullptr);
}

extern "C" int gcf_propagate_fwd(const gcf_csr_t* A, int32_t d, int32_t n_layers, const float* X0,
                                 float* const* layers, float* final_out, float scale, void* workspace,
                                 size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(A != nullptr && X0 != nullptr, "gcf_propagate_fwd: null operator or X0");
  GCF_REQUIRE(n_layers >= 1 && n_layers <= GCF_MAX_ADDENDS, "gcf_propagate_fwd: n_layers must be in [1, %d]", GCF_MAX_ADDENDS);
  GCF_REQUIRE(A->n_rows == A->n_cols, "gcf_propagate_fwd: operator must be square");
  GCF_REQUIRE(layers != nullptr, "gcf_propagate_fwd: null layers array");
  const float* cur = X0;
  for (int k = 1; k <= n_layers; ++k) {
    float* y = layers[k - 1];
    if (k < n_layers) {
      GCF_REQUIRE(y != nullptr, "gcf_propagate_fwd: layers[%d] is NULL (only the last may be)", k - 1);
      int rc = gcf_spmm_csr_f32(A, d, cur, d, y, d, nullptr, 0, GCF_EPILOGUE_NONE, 1.f, 1.f, 0, nullptr, nullptr,
                                workspace, workspace_bytes, 0, stream);
      if (rc != GCF_OK) return rc;
      cur = y;
    } else {
      GCF_REQUIRE(y != nullptr || final_out != nullptr, "gcf_propagate_fwd: nothing to compute for the last layer");
      const float* adds[GCF_MAX_ADDENDS];
      float betas[GCF_MAX_ADDENDS];
      int na = 0;
      if (final_out != nullptr) {
        adds[na] = X0; betas[na] = 1.f; ++na;
        for (int q = 0; q < n_layers - 1; ++q) { adds[na] = layers[q]; betas[na] = 1.f; ++na; }
      }
      int rc = gcf_spmm_csr_f32(A, d, cur, d, y, d, final_out, d, GCF_EPILOGUE_NONE, 1.f, scale, na, adds, betas,
                                workspace, workspace_bytes, 0, stream);
      if (rc != GCF_OK) return rc;
    }
  }
  return GCF_OK;
}

static int propagate_bwd_impl(const gcf_csr_t* At, int32_t d, int32_t n_layers, const float* g_final,
                              const float* const* extra, float scale, float* ping, float* pong, float* g_x0,
                              void* workspace, size_t workspace_bytes, gcf_stream_t stream, const AdamFuse* adam) {
  GCF_REQUIRE(At != nullptr && (g_x0 != nullptr || adam != nullptr), "gcf_propagate_bwd: null operator or output");
  GCF_REQUIRE(n_layers >= 1 && n_layers <= GCF_MAX_ADDENDS, "gcf_propagate_bwd: n_layers must be in [1, %d]", GCF_MAX_ADDENDS);
  GCF_REQUIRE(At->n_rows == At->n_cols, "gcf_propagate_bwd: operator must be square");
  GCF_REQUIRE(n_layers == 1 || (ping != nullptr && pong != nullptr), "gcf_propagate_bwd: ping/pong scratch required");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int K = n_layers;
  const long long n4 = (long long)At->n_rows * d / 4;
  auto ex = [&](int k) -> const float* { return extra != nullptr ? extra[k] : nullptr; };
  GCF_REQUIRE(g_final != nullptr || ex(K) != nullptr, "gcf_propagate_bwd: no gradient reaches the last layer");

  // G(K): only materialised when an extra gradient reaches E(K); otherwise fold scale*g into the first SpMM.
  const float* cur = nullptr;  // G(k+1)
  float cur_alpha = 1.f;
  float* bufs[2] = {ping, pong};
  int which = 0;
  if (ex(K) != nullptr) {
    GCF_REQUIRE(ping != nullptr, "gcf_propagate_bwd: ping scratch required when extra[K] is given");
    const int blocks = (int)std::min<long long>(cdiv(n4, 256), (long long)sm_count() * 8);
    if (g_final != nullptr)
      axpby_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<float4*>(ping), reinterpret_cast<const float4*>(g_final),
                                           scale, reinterpret_cast<const float4*>(ex(K)), 1.f, n4);
    else
      axpby_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<float4*>(ping), reinterpret_cast<const float4*>(ex(K)),
                                           1.f, nullptr, 0.f, n4);
    GCF_LAUNCH_CHECK("axpby_kernel");
    cur = ping;
    which = 1;
  } else {
    cur = g_final;
    cur_alpha = scale;
  }
  for (int k = K - 1; k >= 0; --k) {
    float* out = (k == 0) ? g_x0 : bufs[which];
    const float* adds[2];
    float betas[2];
    int na = 0;
    if (g_final != nullptr) { adds[na] = g_final; betas[na] = scale; ++na; }
    if (ex(k) != nullptr) { adds[na] = ex(k); betas[na] = 1.f; ++na; }
    int rc = spmm_impl(At, d, cur, d, nullptr, 0, out, d, GCF_EPILOGUE_NONE, cur_alpha, 1.f, na, adds, betas, workspace,
                       workspace_bytes, 0, stream, k == 0 ? adam : nullptr);
    if (rc != GCF_OK) return rc;
    cur = out;
    cur_alpha = 1.f;
    which ^= 1;
  }
  return GCF_OK;
}

extern "C" int gcf_propagate_bwd(const gcf_csr_t* At, int32_t d, int32_t n_layers, const float* g_final,
                                 const float* const* extra, float scale, float* ping, float* pong, float* g_x0,
                                 void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  return propagate_bwd_impl(At, d, n_layers, g_final, extra, scale, ping, pong, g_x0, workspace, workspace_bytes, stream, nullptr);
}

extern "C" int gcf_propagate_bwd_adam(const gcf_csr_t* At, int32_t d, int32_t n_layers, const float* g_final,
                                      const float* const* extra, float scale, float* ping, float* pong, float* g_x0,
                                      float* param, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2,
                                      float eps, float weight_decay, int32_t decoupled, int64_t step, void* workspace,
                                      size_t workspace_bytes, gcf_stream_t stream) {
  GCF_REQUIRE(param && exp_avg && exp_avg_sq, "gcf_propagate_bwd_adam: null optimiser state");
  GCF_REQUIRE(step >= 1, "gcf_propagate_bwd_adam: step must be >= 1");
  AdamFuse af{param, exp_avg, exp_avg_sq, make_adam_args(lr, beta1, beta2, eps, weight_decay, decoupled, step)};
  return propagate_bwd_impl(At, d, n_layers, g_final, extra, scale, ping, pong, g_x0, workspace, workspace_bytes, stream, &af);
}
