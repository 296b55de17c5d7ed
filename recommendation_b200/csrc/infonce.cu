// InfoNCE-family losses on the 5th-generation tensor cores (tcgen05 + TMEM), operands fed by TMA.
//
// Replaces the dense logits pipelines of the reference --
//   InfoNCE            ncl.py:125-130, ssl4rec.py:19-23      normalize -> matmul -> /tau -> log_softmax -> diag
//   ssl_layer_loss     ncl.py:358-367                        B x U and B x I logits, exp / sum
//   batch_softmax_loss ssl4rec.py:25-30
//   info_nce_loss      gcl.py:28-35                          full U x U (I x I) logits, row and column cross-entropy
//   DirectAU           directau.py:245-251                   pdist -> exp -> mean -> log
// -- none of which ever materialises here: a CTA owns a 128-row tile of one operand, streams 256-row tiles of
// the other through a TMA ring, one thread issues tcgen05.mma (bf16 x bf16 -> fp32 in TMEM, double-buffered),
// and eight epilogue warps (two per TMEM lane quadrant, one column half each) read the accumulator back with tcgen05.ld
// and fold it into an online log-sum-exp.
//
// Numerics: operands are L2-normalised (when cos != 0) in fp32, scaled by log2(e)/tau on the "query" side and
// rounded to bf16 once; accumulation, running max / sum and every reduction are fp32.  Logit error is bounded
// by the bf16 rounding of the operands (<= 2e-2 absolute for tau >= 0.05, the north-star tolerance).
#include "common.cuh"
#include "tc05.cuh"
#include <cuda_bf16.h>
#include <cublas_v2.h>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <mutex>

namespace gcf {

using namespace tc;

constexpr int kTileM = 128;      // rows of the stationary operand per CTA (= TMEM lanes)
constexpr int kTileN = 256;      // rows of the streamed operand per MMA tile (= TMEM columns per stage)
constexpr int kChunkK = 64;      // bf16 elements per 128-byte swizzled row
constexpr int kLseStages = 4;    // TMA ring depth (one stage = one [kTileN x 64] chunk = 32 KB)
constexpr int kLseThreads = 320; // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-9: epilogue (two per TMEM lane quadrant)
constexpr int kMaxChunks = 16;   // d <= 1024 (the tuner grids of ncl.py / ssl4rec.py / gcl.py go up to 1024)
#ifndef GCF_POLY_OF_8
#define GCF_POLY_OF_8 2           // measured best of 0 / 2 / 3 / 4 (r01) and of 2 / 3 / 4 / 5 (r03: tools/sweep_infonce_poly.sh, profiles/r03_infonce_poly_sweep.log)
#endif
constexpr int kPolyOf8 = GCF_POLY_OF_8;   // exponentials evaluated on the FMA pipe per 8 logits
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// ------------------------------------------------------------------------------------------------
// operand preparation: (optional) row L2-normalise, scale, round to bf16, zero-pad to [n_pad, d_pad]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
prep_rows_kernel(const float* __restrict__ x, long long ld, long long n, int d, int d_pad, long long n_pad, int cos,
                 float scale, __nv_bfloat16* __restrict__ out, float* __restrict__ inv_norm) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_pad) return;
  __nv_bfloat16* o = out + row * d_pad;
  if (row >= n) {
    for (int c = lane; c < d_pad; c += 32) o[c] = __float2bfloat16(0.f);
    return;
  }
  const float* xr = x + row * ld;
  float inv = 1.f;
  if (cos) {
    float ss = 0.f;
    for (int c = lane; c < d; c += 32) { const float v = xr[c]; ss = fmaf(v, v, ss); }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize eps
  }
  if (lane == 0 && inv_norm != nullptr) inv_norm[row] = inv;
  const float s = inv * scale;
  for (int c = lane; c < d_pad; c += 32) o[c] = __float2bfloat16(c < d ? xr[c] * s : 0.f);
}

// pos[i] = <qb_i, kb_{p_i}> * ln2   (same bf16 operands as the tensor-core logits; natural-log units)
__global__ void __launch_bounds__(256)
pos_logit_kernel(const __nv_bfloat16* __restrict__ qb, const __nv_bfloat16* __restrict__ kb, int d_pad, long long m,
                 long long n, const int64_t* __restrict__ pos_idx, float* __restrict__ pos) {
  const int lane = threadIdx.x & 31;
  const long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= m) return;
  long long p = pos_idx != nullptr ? pos_idx[row] : row;
  float acc = 0.f;
  if (p >= 0 && p < n) {
    const __nv_bfloat16* a = qb + row * d_pad;
    const __nv_bfloat16* b = kb + p * d_pad;
    for (int c = lane; c < d_pad; c += 32) acc = fmaf(__bfloat162float(a[c]), __bfloat162float(b[c]), acc);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
  if (lane == 0) pos[row] = acc * kLn2;
}

// lse[i] = ln2 * (M + log2 sum_s l_s 2^(m_s - M))   over the n_splits partial (max, sum) pairs
__global__ void __launch_bounds__(256)
combine_lse_kernel(const float* __restrict__ part_m, const float* __restrict__ part_l, int n_splits, long long m_pad,
                   long long m, float* __restrict__ lse) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  float mx = -INFINITY;
  for (int s = 0; s < n_splits; ++s) mx = fmaxf(mx, part_m[s * m_pad + i]);
  float l = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float ms = part_m[s * m_pad + i];
    if (ms > -INFINITY) l += part_l[s * m_pad + i] * exp2f(ms - mx);
  }
  lse[i] = (mx + log2f(l)) * kLn2;
}

// ------------------------------------------------------------------------------------------------
// streaming log-sum-exp kernel:  for the CTA's 128 rows a of A and its slice of B's row tiles,
//   (m_a, l_a) = online max / sum over b of 2^(<A_a, B_b>)
// ------------------------------------------------------------------------------------------------
struct LseSmem {
  // operand tiles first (1024-byte aligned for the 128B swizzle)
  static constexpr int kABytesPerChunk = kTileM * 128;  // 16 KB
  static constexpr int kBBytesPerStage = kTileN * 128;  // 32 KB
};

// BOUNDED: the operands are L2-normalised, so |S2| <= log2e / tau is known up front; when that bound is small enough for fp32
// (host checks <= 64) the running max and its rescaling are dropped: l = sum_b 2^S2, m = 0.  This is the reference's own
// non-stabilised exp / sum form (ncl.py:362-365) and halves the epilogue's instruction count.
// PROBS (wide embeddings, d > 256): instead of reducing the logits, the epilogue writes
//   P = w_r 2^(S2 - l_r) + w_c 2^(S2 - l_c)  as bf16 into a row-major [a_pad, ldp] buffer; the two gradient GEMMs P B and
//   P^T A then run as plain library GEMMs (the accumulator of the fused backward kernel does not fit the TMEM at d > 256).
struct ProbsArgs {
  const float* w_r; const float* lse_r; const float* w_c; const float* lse_c;
  __nv_bfloat16* P; long long ldp; long long n_a;
};

// KC > 0: the stationary operand (KC chunks of [128 x 64]) stays in shared memory.  KC == 0: d_pad > 256, its chunks are
// streamed through the ring together with the other operand's (kc_rt chunks per tile, read from L2 once per tile).
template <int KC, bool BOUNDED, int POLY, bool PROBS>
__global__ void __launch_bounds__(kLseThreads, 1)
lse_stream_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, long long n_b,
                  int n_tiles, int tiles_per_split, float* __restrict__ part_m, float* __restrict__ part_l,
                  long long m_pad, int skip_diag, int kc_rt, ProbsArgs pa) {
  constexpr bool kStreamA = KC == 0;
  const int n_kc = kStreamA ? kc_rt : KC;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                               // KC resident chunks | kLseStages streamed chunks of [128 x 64]
  uint8_t* smem_b = smem + (kStreamA ? kLseStages : KC) * LseSmem::kABytesPerChunk;   // kLseStages chunks of [256 x 64] bf16
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + kLseStages * LseSmem::kBBytesPerStage);
  uint64_t* full_bar = bars;                    // [kLseStages]
  uint64_t* empty_bar = bars + kLseStages;      // [kLseStages]
  uint64_t* a_bar = bars + 2 * kLseStages;      // [1]
  uint64_t* acc_full = a_bar + 1;               // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
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
      for (int s = 0; s < kLseStages; ++s) { mbar_init(full_bar + s, 1); mbar_init(empty_bar + s, 1); }
      mbar_init(a_bar, 1);
      for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 8); }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 2 * kTileN);  // 512 columns: two accumulator stages
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && my_tiles > 0) {
      if (!kStreamA) {
        mbar_expect_tx(a_bar, KC * LseSmem::kABytesPerChunk);
        for (int kc = 0; kc < KC; ++kc) tma_load_2d(&tm_a, a_bar, smem_a + kc * LseSmem::kABytesPerChunk, kc * kChunkK, m_tile * kTileM);
      }
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_expect_tx(full_bar + stage, LseSmem::kBBytesPerStage + (kStreamA ? LseSmem::kABytesPerChunk : 0));
          if (kStreamA) tma_load_2d(&tm_a, full_bar + stage, smem_a + stage * LseSmem::kABytesPerChunk, kc * kChunkK, m_tile * kTileM);
          tma_load_2d(&tm_b, full_bar + stage, smem_b + stage * LseSmem::kBBytesPerStage, kc * kChunkK, t * kTileN);
          if (++stage == kLseStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(kTileM, kTileN, 0, 0);
      if (!kStreamA) {
        mbar_wait(a_bar, 0);
        fence_after_sync();
      }
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        mbar_wait(acc_empty + acc, ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        fence_after_sync();
        for (int kc = 0; kc < n_kc; ++kc) {
          mbar_wait(full_bar + stage, phase);
          fence_after_sync();
          const uint32_t a_addr = smem_u32(smem_a + (kStreamA ? stage : kc) * LseSmem::kABytesPerChunk);
          const uint32_t b_addr = smem_u32(smem_b + stage * LseSmem::kBBytesPerStage);
#pragma unroll
          for (int kk = 0; kk < kChunkK / 16; ++kk) {  // UMMA_K = 16 bf16 = 32 bytes inside the swizzled row
            const uint64_t da = smem_desc_sw128(a_addr + kk * 32, 0, 1024);
            const uint64_t db = smem_desc_sw128(b_addr + kk * 32, 0, 1024);
            umma_bf16(tmem_base + acc * kTileN, da, db, idesc, (kc | kk) != 0);
          }
          umma_commit(empty_bar + stage);  // frees the smem stage once these MMAs have read it
          if (++stage == kLseStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full + acc);       // accumulator stage complete
      }
    }
  } else {
    // ===== epilogue: thread = one row x one half of the tile's columns; online base-2 log-sum-exp =====
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;                // columns [half * 128, half * 128 + 128) of every 256-column tile
    const int row = quad * 32 + lane;
    float m_run = BOUNDED ? 0.f : -INFINITY, l_run = 0.f;
    const long long row_g = (long long)m_tile * kTileM + row;
    float wr = 0.f, lr = 0.f;
    if (PROBS && pa.w_r != nullptr && row_g < pa.n_a) { wr = pa.w_r[row_g]; lr = pa.lse_r[row_g] * kLog2e; }
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const int t = t_begin + it;
      mbar_wait(acc_full + acc, (it >> 1) & 1);
      fence_after_sync();
      const long long col0 = (long long)t * kTileN + half * (kTileN / 2);
      if (PROBS) {
        const long long diag = row_g - col0;
        const bool has_diag = skip_diag && diag >= 0 && diag < kTileN / 2;
#pragma unroll 1
        for (int c = 0; c < kTileN / 64; ++c) {
          float v[32];
          tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileN + half * (kTileN / 2) + c * 32), v);
          const long long cc = col0 + c * 32;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sv = v[j];
            float pv = 0.f;
            if (pa.w_r != nullptr) pv = wr * ex2_mixed<POLY>(sv - lr, j);
            if (pa.w_c != nullptr && cc + j < n_b) pv = fmaf(__ldg(pa.w_c + cc + j), ex2_mixed<POLY>(sv - __ldg(pa.lse_c + cc + j) * kLog2e, j), pv);
            v[j] = pv;
          }
          if (has_diag && (int)(diag >> 5) == c) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j == (int)(diag & 31)) v[j] = 0.f;
          }
          // rows beyond n_a and columns beyond n_b meet zero-padded operand rows in the GEMMs: their values are irrelevant
          uint4* dst = reinterpret_cast<uint4*>(pa.P + row_g * pa.ldp + cc);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 pk;
            __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * u + 0], v[8 * u + 1]);
            __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * u + 2], v[8 * u + 3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * u + 4], v[8 * u + 5]);
            __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * u + 6], v[8 * u + 7]);
            pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
            pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
            dst[u] = pk;
          }
        }
        fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + acc);
        continue;
      }
      const bool ragged = col0 + kTileN / 2 > n_b;   // contains zero-padded rows of B: mask them out
      const long long diag = (long long)m_tile * kTileM + row - col0;  // column of this row's diagonal element
      const bool has_diag = skip_diag && diag >= 0 && diag < kTileN / 2;   // DirectAU: pairs i != j only
#pragma unroll 1
      for (int c = 0; c < kTileN / 64; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileN + half * (kTileN / 2) + c * 32), v);
        if (ragged) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + c * 32 + j >= n_b) v[j] = -INFINITY;
        }
        if (has_diag && (int)(diag >> 5) == c) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j == (int)(diag & 31)) v[j] = -INFINITY;
        }
        if (BOUNDED) {
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            s0 += ex2_mixed<POLY>(v[j], j); s1 += ex2_mixed<POLY>(v[j + 1], j + 1);
            s2 += ex2_mixed<POLY>(v[j + 2], j + 2); s3 += ex2_mixed<POLY>(v[j + 3], j + 3);
          }
          l_run += (s0 + s1) + (s2 + s3);
        } else {
          float cm = v[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) cm = fmaxf(cm, v[j]);
          if (cm > m_run) {  // lazy rescale: only when the running max moves
            l_run *= exp2f(m_run - cm);  // m_run = -inf -> factor 0 (l_run is 0 anyway)
            m_run = cm;
          }
          if (m_run > -INFINITY) {
            float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              s0 += ex2_mixed<POLY>(v[j] - m_run, j); s1 += ex2_mixed<POLY>(v[j + 1] - m_run, j + 1);
              s2 += ex2_mixed<POLY>(v[j + 2] - m_run, j + 2); s3 += ex2_mixed<POLY>(v[j + 3] - m_run, j + 3);
            }
            l_run += (s0 + s1) + (s2 + s3);
          }
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
    }
    if (!PROBS) {
      if (BOUNDED && l_run == 0.f) m_run = -INFINITY;  // nothing unmasked in this slice: neutral element for the merge
      // the two column halves are merged like two more splits
      const long long out = ((long long)split * 2 + half) * m_pad + (long long)m_tile * kTileM + row;
      part_m[out] = m_run;
      part_l[out] = l_run;
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 2 * kTileN);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
static inline int pad_d(int d) { return (int)round_up(d, kChunkK); }

struct LsePlan { long long a_pad, b_pad; int m_tiles, n_tiles, n_splits, tiles_per_split; };

static LsePlan plan_lse(long long n_a, long long n_b) {
  LsePlan p;
  p.a_pad = round_up(std::max<long long>(n_a, 1), kTileM);
  p.b_pad = round_up(std::max<long long>(n_b, 1), kTileN);
  p.m_tiles = (int)(p.a_pad / kTileM);
  p.n_tiles = (int)(p.b_pad / kTileN);
  // column splits so that the grid fills the SMs in ONE wave (a 160-CTA grid on 148 SMs costs a second, nearly empty wave)
  const int sms = sm_count();
  int splits = std::max(1, std::min(p.n_tiles, sms / std::max(p.m_tiles, 1)));
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

static size_t lse_smem_bytes(int kc) {  // kc > 4: the stationary operand's chunks ride in the ring (one per stage)
  const int a_chunks = kc > 4 ? kLseStages : kc;
  return 1024 + (size_t)a_chunks * LseSmem::kABytesPerChunk + (size_t)kLseStages * LseSmem::kBBytesPerStage + 256;
}

// (m, l) partials -> lse[n_a]; ab/bb are the prepared bf16 operands ([a_pad, d_pad], [b_pad, d_pad])
static int run_lse(const __nv_bfloat16* ab, long long n_a, const __nv_bfloat16* bb, long long n_b, int d_pad,
                   float* part_m, float* part_l, float* lse_out, cudaStream_t st, int skip_diag = 0, bool bounded = false) {
  const LsePlan p = plan_lse(n_a, n_b);
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap(&tm_a, ab, p.a_pad, d_pad, kTileM);
  if (rc != GCF_OK) return rc;
  rc = make_tmap(&tm_b, bb, p.b_pad, d_pad, kTileN);
  if (rc != GCF_OK) return rc;
  const int kc = d_pad / kChunkK;
  const size_t smem = lse_smem_bytes(kc);
  dim3 grid(p.m_tiles, p.n_splits);
  const ProbsArgs no_probs{nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0};
#define GCF_LSE_LAUNCH2(KC, B)                                                                                                       \
  do {                                                                                                                               \
    GCF_CUDA(cudaFuncSetAttribute(lse_stream_kernel<KC, B, kPolyOf8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    lse_stream_kernel<KC, B, kPolyOf8, false><<<grid, kLseThreads, smem, st>>>(tm_a, tm_b, n_b, p.n_tiles, p.tiles_per_split, part_m,   \
                                                                               part_l, p.a_pad, skip_diag, kc, no_probs);            \
  } while (0)
#define GCF_LSE_LAUNCH(KC)                                      \
  do {                                                          \
    if (bounded) GCF_LSE_LAUNCH2(KC, true);                     \
    else GCF_LSE_LAUNCH2(KC, false);                            \
  } while (0)
  switch (kc) {
    case 1: GCF_LSE_LAUNCH(1); break;
    case 2: GCF_LSE_LAUNCH(2); break;
    case 3: GCF_LSE_LAUNCH(3); break;
    case 4: GCF_LSE_LAUNCH(4); break;
    default:
      if (kc > kMaxChunks) { set_error("infonce: d_pad=%d unsupported (d <= %d)", d_pad, kMaxChunks * kChunkK); return GCF_EUNSUPPORTED; }
      GCF_LSE_LAUNCH(0);  // wide embeddings: both operands streamed
  }
#undef GCF_LSE_LAUNCH
#undef GCF_LSE_LAUNCH2
  GCF_LAUNCH_CHECK("lse_stream_kernel");
  combine_lse_kernel<<<(unsigned)cdiv(n_a, 256), 256, 0, st>>>(part_m, part_l, 2 * p.n_splits, p.a_pad, n_a, lse_out);
  GCF_LAUNCH_CHECK("combine_lse_kernel");
  return GCF_OK;
}

// ------------------------------------------------------------------------------------------------
// streaming gradient kernel:  for the CTA's 128 rows a of A and its slice of B's 128-row tiles
//   P_ab = w_r[a] 2^(S2_ab - l_r[a]) + w_c[b] 2^(S2_ab - l_c[b]),   S2 = A B^T (log2 units, recomputed on the tensor cores)
//   G_a  = sum_b P_ab B_b                                          (second MMA, accumulator stays in TMEM)
// warp 0: TMA, warp 1: MMA issuer, warps 2-9: S (TMEM) -> P (bf16, swizzled smem) conversion and the final G read-out.
// ------------------------------------------------------------------------------------------------
constexpr int kGradTileN = 128;   // B rows per tile = K extent of the second MMA
constexpr int kGradThreads = 320;  // warp 0: TMA, warp 1: MMA, warps 2-9: conversion (two per TMEM lane quadrant)
constexpr int kChunkBytes = 128 * 128;  // one [128 x 64] bf16 swizzled chunk = 16 KB

template <int KC> struct GradCfg {
  static constexpr int kBStages = (KC <= 1) ? 4 : (KC == 2 ? 3 : 2);
  static constexpr int kPBufs = (KC <= 3) ? 2 : 1;
  static constexpr int kCvecBufs = (KC <= 3) ? 2 : 1;  // d_pad = 256 is 512 bytes short of the 227 KB limit otherwise
  static constexpr size_t kSmem = 1024 + (size_t)KC * kChunkBytes + (size_t)kBStages * KC * kChunkBytes +
                                  (size_t)kPBufs * 2 * kChunkBytes + kCvecBufs * 2 * kGradTileN * sizeof(float) + 256;
  static_assert(kSmem <= 232448, "grad_stream_kernel: shared memory budget exceeded");
};

template <int KC, bool HAS_ROW, bool HAS_COL>
__global__ void __launch_bounds__(kGradThreads, 1)
grad_stream_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, long long n_a,
                   long long n_b, int n_tiles, int tiles_per_split, const float* __restrict__ w_r,
                   const float* __restrict__ lse_r, const float* __restrict__ w_c, const float* __restrict__ lse_c,
                   int skip_diag, float* __restrict__ part, long long a_pad) {
  using Cfg = GradCfg<KC>;
  constexpr int NS = Cfg::kBStages, NP = Cfg::kPBufs, NCV = Cfg::kCvecBufs;
  constexpr int DP = KC * kChunkK;          // padded width = N of the second MMA
  constexpr uint32_t kTmemG = 2 * kGradTileN;  // S stages at columns [0,128) and [128,256); G at [256, 256 + DP)
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem_a + KC * kChunkBytes;                 // NS stages of KC chunks
  uint8_t* smem_p = smem_b + NS * KC * kChunkBytes;            // NP buffers of 2 chunks
  float* cvec = reinterpret_cast<float*>(smem_p + NP * 2 * kChunkBytes);  // [NCV][w | lse][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(cvec + NCV * 2 * kGradTileN);
  uint64_t* b_full = bars;               // [NS]
  uint64_t* b_empty = b_full + NS;       // [NS]
  uint64_t* a_bar = b_empty + NS;        // [1]
  uint64_t* s_full = a_bar + 1;          // [2]
  uint64_t* s_empty = s_full + 2;        // [2]
  uint64_t* p_full = s_empty + 2;        // [NP]
  uint64_t* p_empty = p_full + NP;       // [NP]
  uint64_t* g_full = p_empty + NP;       // [1]
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(g_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, split = blockIdx.y;
  const int t_begin = split * tiles_per_split;
  const int t_end = min(t_begin + tiles_per_split, n_tiles);
  const int my_tiles = max(t_end - t_begin, 0);

  if (warp == 0 && lane == 0) { tma_prefetch_desc(&tm_a); tma_prefetch_desc(&tm_b); }
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < NS; ++i) { mbar_init(b_full + i, 1); mbar_init(b_empty + i, 1); }
      mbar_init(a_bar, 1);
      for (int i = 0; i < 2; ++i) { mbar_init(s_full + i, 1); mbar_init(s_empty + i, 8); }
      for (int i = 0; i < NP; ++i) { mbar_init(p_full + i, 8); mbar_init(p_empty + i, 1); }
      mbar_init(g_full, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc(tmem_holder, 512);
  }
  fence_before_sync();
  __syncthreads();
  fence_after_sync();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0 && my_tiles > 0) {
      mbar_expect_tx(a_bar, KC * kChunkBytes);
      for (int kc = 0; kc < KC; ++kc) tma_load_2d(&tm_a, a_bar, smem_a + kc * kChunkBytes, kc * kChunkK, m_tile * kTileM);
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        mbar_wait(b_empty + stage, phase ^ 1);
        mbar_expect_tx(b_full + stage, KC * kChunkBytes);
        for (int kc = 0; kc < KC; ++kc)
          tma_load_2d(&tm_b, b_full + stage, smem_b + (stage * KC + kc) * kChunkBytes, kc * kChunkK, t * kGradTileN);
        if (++stage == NS) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc_s = idesc_bf16_f32(kTileM, kGradTileN, 0, 0);  // S = A B^T : both operands K-major
      constexpr uint32_t idesc_g = idesc_bf16_f32(kTileM, DP, 0, 1);          // G += P B  : B operand MN-major
      mbar_wait(a_bar, 0);
      fence_after_sync();
      auto issue_s = [&](int it) {  // S(it) into TMEM stage it & 1
        const int stage = it % NS;
        const int acc = it & 1;
        mbar_wait(b_full + stage, (it / NS) & 1);
        mbar_wait(s_empty + acc, ((it >> 1) & 1) ^ 1);
        fence_after_sync();
#pragma unroll
        for (int kc = 0; kc < KC; ++kc) {
          const uint32_t a_addr = smem_u32(smem_a + kc * kChunkBytes);
          const uint32_t b_addr = smem_u32(smem_b + (stage * KC + kc) * kChunkBytes);
#pragma unroll
          for (int kk = 0; kk < kChunkK / 16; ++kk)
            umma_bf16(tmem_base + acc * kGradTileN, smem_desc_sw128(a_addr + kk * 32, 0, 1024),
                      smem_desc_sw128(b_addr + kk * 32, 0, 1024), idesc_s, (kc | kk) != 0);
        }
        umma_commit(s_full + acc);
      };
      issue_s(0);
      for (int it = 0; it < my_tiles; ++it) {
        if (it + 1 < my_tiles) issue_s(it + 1);     // the tensor core computes S(it+1) while the warps convert S(it)
        const int stage = it % NS, pb = it % NP;
        mbar_wait(p_full + pb, (it / NP) & 1);
        fence_after_sync();
        const uint32_t p_addr = smem_u32(smem_p + pb * 2 * kChunkBytes);
        const uint32_t b_addr = smem_u32(smem_b + stage * KC * kChunkBytes);
#pragma unroll
        for (int ks = 0; ks < kGradTileN / 16; ++ks) {  // K = the 128 B rows of the tile, 16 per MMA
          // A = P: K-major, chunk ks / 4, 32 bytes per K step inside the swizzled row
          const uint64_t da = smem_desc_sw128(p_addr + (ks >> 2) * kChunkBytes + (ks & 3) * 32, 0, 1024);
          // B = the same smem tile read MN-major: K rows are 128 B apart (8-row groups 1024 B), 64-wide N blocks
          // (the KC chunks) are kChunkBytes apart
          const uint64_t db = smem_desc_sw128(b_addr + ks * 16 * 128, kChunkBytes, 1024);
          umma_bf16(tmem_base + kTmemG, da, db, idesc_g, (it | ks) != 0);
        }
        umma_commit(p_empty + pb);
        umma_commit(b_empty + stage);
      }
      umma_commit(g_full);
    }
  } else {
    // ===== conversion warps: thread = one row of the tile x one 64-column half (= one swizzled P chunk) =====
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quad * 32 + lane;
    const int et = threadIdx.x - 64;  // 0..255; the first 128 stage the per-column vectors
    const long long row_g = (long long)m_tile * kTileM + row;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float wr = 0.f, lr = 0.f;
    if (HAS_ROW && row_g < n_a) { wr = w_r[row_g]; lr = lse_r[row_g] * kLog2e; }
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1, pb = it % NP;
      const long long col0 = (long long)(t_begin + it) * kGradTileN;
      float* cw = cvec + (it % NCV) * 2 * kGradTileN;
      float* cl = cw + kGradTileN;
      if (HAS_COL) {
        if (et < kGradTileN) {
          const long long j = col0 + et;
          cw[et] = (j < n_b) ? w_c[j] : 0.f;
          cl[et] = (j < n_b) ? lse_c[j] * kLog2e : 0.f;
        }
        named_bar_sync(1, 256);
      }
      mbar_wait(s_full + acc, (it >> 1) & 1);
      fence_after_sync();
      mbar_wait(p_empty + pb, ((it / NP) & 1) ^ 1);
      const long long diag = row_g - col0;
      const bool has_diag = skip_diag && diag >= 0 && diag < kGradTileN;
      uint8_t* prow = smem_p + pb * 2 * kChunkBytes + row * 128;
#pragma unroll 1
      for (int c = 2 * half; c < 2 * half + 2; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + lane_off + (uint32_t)(acc * kGradTileN + c * 32), v);
#pragma unroll
        for (int j4 = 0; j4 < 32; j4 += 4) {
          float4 w4 = make_float4(0.f, 0.f, 0.f, 0.f), l4 = w4;
          if (HAS_COL) {
            w4 = *reinterpret_cast<const float4*>(cw + c * 32 + j4);
            l4 = *reinterpret_cast<const float4*>(cl + c * 32 + j4);
          }
          const float wv[4] = {w4.x, w4.y, w4.z, w4.w}, lv[4] = {l4.x, l4.y, l4.z, l4.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float sv = v[j4 + q];
            float pv = 0.f;
            if (HAS_ROW) pv = wr * ex2_mixed<kPolyOf8>(sv - lr, j4 + q);
            if (HAS_COL) pv = fmaf(wv[q], ex2_mixed<kPolyOf8>(sv - lv[q], j4 + q), pv);
            v[j4 + q] = pv;
          }
        }
        if (has_diag && (int)(diag >> 5) == c) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j == (int)(diag & 31)) v[j] = 0.f;
        }
        // bf16 pack + 128B-swizzled store: chunk c / 2, 16-byte units (c % 2) * 4 + u, XOR (row % 8)
        uint8_t* pchunk = prow + (c >> 1) * kChunkBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 pk;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * u + 0], v[8 * u + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * u + 2], v[8 * u + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * u + 4], v[8 * u + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * u + 6], v[8 * u + 7]);
          pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
          pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
          const int unit = ((c & 1) * 4 + u) ^ (row & 7);
          *reinterpret_cast<uint4*>(pchunk + unit * 16) = pk;
        }
      }
      if (HAS_COL && NCV == 1) named_bar_sync(1, 256);  // single staging buffer: everyone is done reading it
      fence_proxy_async_smem();   // generic-proxy P writes -> visible to the tensor core's async-proxy reads
      fence_before_sync();
      __syncwarp();
      if (lane == 0) { mbar_arrive(p_full + pb); mbar_arrive(s_empty + acc); }
    }
    // final read-out of G (raw, unscaled) into this split's partial buffer
    float* out = part + ((long long)split * a_pad + row_g) * DP;
    if (my_tiles > 0) {
      mbar_wait(g_full, 0);
      fence_after_sync();
#pragma unroll 1
      for (int c = half; c < DP / 32; c += 2) {
        float v[32];
        tmem_ld_32x32(tmem_base + lane_off + kTmemG + (uint32_t)(c * 32), v);
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(out + c * 32 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      }
    } else if (half == 0) {
      for (int c = 0; c < DP; c += 4) *reinterpret_cast<float4*>(out + c) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  fence_before_sync();
  __syncthreads();
  if (warp == 1) {
    fence_after_sync();
    tmem_dealloc(tmem_base, 512);
  }
}

// g_hat[i] = scale * sum_s part[s][i] (+ pos term) (+ extra[i]);  then the chain rule of x^ = x / max(|x|, eps).
// One warp per row.
template <int NQ>  // 32 * NQ >= d columns per row, NQ values per lane
__global__ void __launch_bounds__(256)
grad_finish_kernel_t(const float* __restrict__ part, int n_splits, long long a_pad, int d_pad, float scale,
                   const float* __restrict__ x, long long ldx, const float* __restrict__ x_inv, long long n, int d,
                   int cos, const float* __restrict__ w_pos, const int64_t* __restrict__ pos_idx, float pos_scale,
                   const float* __restrict__ other, long long ld_other, const float* __restrict__ other_inv,
                   long long n_other, const float* __restrict__ extra, float* __restrict__ g, long long ldg) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  float acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float* p = part + ((long long)s * a_pad + i) * d_pad;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const int c = lane + 32 * q;
      if (c < d) acc[q] += p[c];
    }
  }
  long long pp = -1;
  float wp = 0.f;
  if (w_pos != nullptr) {
    pp = pos_idx != nullptr ? pos_idx[i] : i;
    if (pp < 0 || pp >= n_other) pp = -1;
    else wp = w_pos[i] * pos_scale * (cos ? other_inv[pp] : 1.f);
  }
  const float inv = cos ? x_inv[i] : 1.f;
  float dot = 0.f;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = lane + 32 * q;
    if (c < d) {
      float v = acc[q] * scale;
      if (pp >= 0) v = fmaf(wp, other[pp * ld_other + c], v);
      if (extra != nullptr) v += extra[i * (long long)d + c];
      acc[q] = v;
      if (cos) dot = fmaf(v, x[i * ldx + c] * inv, dot);
    }
  }
  if (cos) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
    if (inv >= 1e12f) dot = 0.f;  // |x| below the F.normalize eps: x^ = x / eps, plain scaling
  }
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    const int c = lane + 32 * q;
    if (c < d) {
      float v = acc[q];
      if (cos) v = inv * (v - x[i * ldx + c] * inv * dot);
      g[i * ldg + c] = v;
    }
  }
}

static void grad_finish_kernel_launch(unsigned blocks, cudaStream_t st, const float* part, int n_splits, long long a_pad, int d_pad,
                                      float scale, const float* x, long long ldx, const float* x_inv, long long n, int d, int cos,
                                      const float* w_pos, const int64_t* pos_idx, float pos_scale, const float* other,
                                      long long ld_other, const float* other_inv, long long n_other, const float* extra, float* g,
                                      long long ldg) {
#define GCF_FINISH(NQ) grad_finish_kernel_t<NQ><<<blocks, 256, 0, st>>>(part, n_splits, a_pad, d_pad, scale, x, ldx, x_inv, n, d, cos, \
                                                                        w_pos, pos_idx, pos_scale, other, ld_other, other_inv, n_other, extra, g, ldg)
  if (d <= 256) GCF_FINISH(8);
  else if (d <= 512) GCF_FINISH(16);
  else GCF_FINISH(32);
#undef GCF_FINISH
}

// extra[pos_i] += w_pos[i] * pos_scale * q^_i   (gradient of the positive logits w.r.t. the key rows)
__global__ void __launch_bounds__(256)
pos_scatter_kernel(const float* __restrict__ q, long long ldq, const float* __restrict__ q_inv, long long m, int d, int cos,
                   const float* __restrict__ w_pos, const int64_t* __restrict__ pos_idx, float pos_scale, long long n,
                   float* __restrict__ extra) {
  const int lane = threadIdx.x & 31;
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= m) return;
  const long long p = pos_idx != nullptr ? pos_idx[i] : i;
  if (p < 0 || p >= n) return;
  const float w = w_pos[i] * pos_scale * (cos ? q_inv[i] : 1.f);
  for (int c = lane; c < d; c += 32) atomicAdd(extra + p * (long long)d + c, w * q[i * ldq + c]);
}

struct GradPlan { long long a_pad; int m_tiles, n_tiles, n_splits, tiles_per_split; };

static GradPlan plan_grad(long long n_a, long long n_b) {
  GradPlan p;
  p.a_pad = round_up(std::max<long long>(n_a, 1), kTileM);
  p.m_tiles = (int)(p.a_pad / kTileM);
  p.n_tiles = (int)cdiv(std::max<long long>(n_b, 1), kGradTileN);
  const int sms = sm_count();
  int splits = std::max(1, std::min(p.n_tiles, sms / std::max(p.m_tiles, 1)));
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

// part[n_splits][a_pad][d_pad] = raw G of each split
static int run_grad(const __nv_bfloat16* ab, long long n_a, const __nv_bfloat16* bb, long long n_b, int d_pad,
                    const float* w_r, const float* lse_r, const float* w_c, const float* lse_c, int skip_diag,
                    float* part, GradPlan* plan_out, cudaStream_t st) {
  const GradPlan p = plan_grad(n_a, n_b);
  *plan_out = p;
  const long long b_rows = round_up(std::max<long long>(n_b, 1), kTileN);  // operands are padded to 256 rows
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap(&tm_a, ab, round_up(std::max<long long>(n_a, 1), kTileN), d_pad, kTileM);
  if (rc != GCF_OK) return rc;
  rc = make_tmap(&tm_b, bb, b_rows, d_pad, kGradTileN);
  if (rc != GCF_OK) return rc;
  const bool has_row = w_r != nullptr, has_col = w_c != nullptr;
  if (!has_row && !has_col) { set_error("infonce backward: no weights given"); return GCF_EINVAL; }
  dim3 grid(p.m_tiles, p.n_splits);
#define GCF_GRAD_LAUNCH2(KC, R, C)                                                                                    \
  do {                                                                                                                \
    GCF_CUDA(cudaFuncSetAttribute(grad_stream_kernel<KC, R, C>, cudaFuncAttributeMaxDynamicSharedMemorySize,          \
                                  (int)GradCfg<KC>::kSmem));                                                          \
    grad_stream_kernel<KC, R, C><<<grid, kGradThreads, GradCfg<KC>::kSmem, st>>>(                                     \
        tm_a, tm_b, n_a, n_b, p.n_tiles, p.tiles_per_split, w_r, lse_r, w_c, lse_c, skip_diag, part, p.a_pad);        \
  } while (0)
#define GCF_GRAD_LAUNCH(KC)                                                \
  do {                                                                     \
    if (has_row && has_col) GCF_GRAD_LAUNCH2(KC, true, true);              \
    else if (has_row) GCF_GRAD_LAUNCH2(KC, true, false);                   \
    else GCF_GRAD_LAUNCH2(KC, false, true);                                \
  } while (0)
  switch (d_pad / kChunkK) {
    case 1: GCF_GRAD_LAUNCH(1); break;
    case 2: GCF_GRAD_LAUNCH(2); break;
    case 3: GCF_GRAD_LAUNCH(3); break;
    case 4: GCF_GRAD_LAUNCH(4); break;
    default: set_error("infonce backward: d_pad=%d unsupported (d <= 256)", d_pad); return GCF_EUNSUPPORTED;
  }
#undef GCF_GRAD_LAUNCH
#undef GCF_GRAD_LAUNCH2
  GCF_LAUNCH_CHECK("grad_stream_kernel");
  return GCF_OK;
}

// ------------------------------------------------------------------------------------------------
// wide embeddings (256 < d <= 1024): the d_pad-column gradient accumulator of grad_stream_kernel does not fit the TMEM
// next to the logits, so the backward materialises P block-wise (bf16, PROBS mode of the streaming kernel) and runs the two
// gradient products P B and P^T A as plain library GEMMs (cuBLAS, bf16 x bf16 -> fp32).  One P serves both gradients.
// ------------------------------------------------------------------------------------------------
constexpr size_t kProbsBlockBytes = size_t(1) << 30;

static cublasHandle_t cublas_handle() {  // one handle per host thread and device, created on first use
  static thread_local cublasHandle_t handles[64] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  if (handles[dev] == nullptr && cublasCreate(&handles[dev]) != CUBLAS_STATUS_SUCCESS) handles[dev] = nullptr;
  return handles[dev];
}

static long long probs_block_rows(long long a_rows, long long b_rows) {
  long long blk = (long long)(kProbsBlockBytes / ((size_t)b_rows * 2)) / kTileM * kTileM;
  return std::max<long long>(kTileM, std::min(a_rows, blk));
}

// Ga[n_a.., d_pad] = P B (nullable), Gb[b_rows, d_pad] = P^T A (nullable); ab / bb: prepared operands with a_rows / b_rows
// (multiples of 256) zero-padded rows; P: scratch of probs_block_rows(a_rows, b_rows) x b_rows bf16.
static int run_grad_wide(const __nv_bfloat16* ab, long long n_a, long long a_rows, const __nv_bfloat16* bb, long long n_b,
                         long long b_rows, int d_pad, const float* w_r, const float* lse_r, const float* w_c,
                         const float* lse_c, int skip_diag, float* Ga, float* Gb, __nv_bfloat16* P, cudaStream_t st) {
  cublasHandle_t h = cublas_handle();
  if (h == nullptr) { set_error("infonce backward: cublasCreate failed"); return GCF_ECUDA; }
  if (cublasSetStream(h, st) != CUBLAS_STATUS_SUCCESS) { set_error("infonce backward: cublasSetStream failed"); return GCF_ECUDA; }
  const int kc = d_pad / kChunkK;
  const size_t smem = lse_smem_bytes(kc);
  const long long blk = probs_block_rows(a_rows, b_rows);
  const int n_tiles = (int)(b_rows / kTileN);
  const float one = 1.f, zero = 0.f;
  CUtensorMap tm_b;
  int rc = make_tmap(&tm_b, bb, b_rows, d_pad, kTileN);
  if (rc != GCF_OK) return rc;
  const long long a_used = round_up(std::max<long long>(n_a, 1), kTileM);
  for (long long r0 = 0, bi = 0; r0 < a_used; r0 += blk, ++bi) {
    const long long rows = std::min(blk, a_used - r0);
    CUtensorMap tm_a;
    rc = make_tmap(&tm_a, ab + r0 * d_pad, rows, d_pad, kTileM);
    if (rc != GCF_OK) return rc;
    const int m_tiles = (int)(rows / kTileM);
    int splits = std::max(1, std::min(n_tiles, sm_count() / std::max(m_tiles, 1)));
    const int tiles_per_split = (n_tiles + splits - 1) / splits;
    splits = (n_tiles + tiles_per_split - 1) / tiles_per_split;
    dim3 grid(m_tiles, splits);
    ProbsArgs pa{w_r != nullptr ? w_r + r0 : nullptr, lse_r != nullptr ? lse_r + r0 : nullptr, w_c, lse_c, P, b_rows,
                 std::max<long long>(0, n_a - r0)};
    // the diagonal of a row block starts r0 columns in: shift it through the column origin seen by the kernel
    const int sd = skip_diag;
#define GCF_PROBS_LAUNCH(KC)                                                                                                        \
    do {                                                                                                                            \
      GCF_CUDA(cudaFuncSetAttribute(lse_stream_kernel<KC, false, kPolyOf8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      lse_stream_kernel<KC, false, kPolyOf8, true><<<grid, kLseThreads, smem, st>>>(tm_a, tm_b, n_b, n_tiles, tiles_per_split, nullptr, \
                                                                                   nullptr, rows, sd, kc, pa);                     \
    } while (0)
    GCF_REQUIRE(!(skip_diag && r0 != 0), "infonce backward: diagonal masking needs the operand in one row block");
    GCF_PROBS_LAUNCH(0);  // only reached for d_pad > 256: both operands streamed
#undef GCF_PROBS_LAUNCH
    GCF_LAUNCH_CHECK("lse_stream_kernel (probs)");
    if (Ga != nullptr) {  // row-major Ga[r0 : r0 + rows] = P [rows x b_rows] * B [b_rows x d_pad]
      if (cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_N, d_pad, (int)rows, (int)b_rows, &one, bb, CUDA_R_16BF, d_pad, P, CUDA_R_16BF,
                       (int)b_rows, &zero, Ga + r0 * d_pad, CUDA_R_32F, d_pad, CUBLAS_COMPUTE_32F, CUBLAS_GEMM_DEFAULT) !=
          CUBLAS_STATUS_SUCCESS) { set_error("infonce backward: cublasGemmEx (P B) failed"); return GCF_ECUDA; }
    }
    if (Gb != nullptr) {  // row-major Gb += P^T [b_rows x rows] * A[r0 : r0 + rows] [rows x d_pad]
      if (cublasGemmEx(h, CUBLAS_OP_N, CUBLAS_OP_T, d_pad, (int)b_rows, (int)rows, &one, ab + r0 * d_pad, CUDA_R_16BF, d_pad, P,
                       CUDA_R_16BF, (int)b_rows, bi == 0 ? &zero : &one, Gb, CUDA_R_32F, d_pad, CUBLAS_COMPUTE_32F,
                       CUBLAS_GEMM_DEFAULT) != CUBLAS_STATUS_SUCCESS) { set_error("infonce backward: cublasGemmEx (P^T A) failed"); return GCF_ECUDA; }
    }
  }
  return GCF_OK;
}

struct InfoWs {
  __nv_bfloat16 *qb, *kb, *probs;
  float *q_inv, *k_inv, *part_m, *part_l, *gpart, *gpart_b, *extra;
  long long q_rows, k_rows;
  size_t bytes;
};

static InfoWs carve_ws(void* ws, long long m, long long n, int d) {
  const int d_pad = pad_d(d);
  const LsePlan pq = plan_lse(m, n), pk = plan_lse(n, m);
  const GradPlan gq = plan_grad(m, n), gk = plan_grad(n, m);
  // operands are padded for every role (stationary: multiple of 128, streamed: multiple of 256 / 128)
  const long long q_rows = round_up(std::max<long long>(m, 1), kTileN), k_rows = round_up(std::max<long long>(n, 1), kTileN);
  const size_t part = 2 * std::max((size_t)pq.n_splits * pq.a_pad, (size_t)pk.n_splits * pk.a_pad);  // x2: column halves
  const bool wide = d_pad > 4 * kChunkK;
  // d_pad <= 256: per-split partial gradients of the TMEM kernel; wider: the two full fp32 products + one P block
  const size_t gpart = wide ? (size_t)q_rows * d_pad
                            : std::max((size_t)gq.n_splits * gq.a_pad, (size_t)gk.n_splits * gk.a_pad) * d_pad;
  const size_t gpart_b = wide ? (size_t)k_rows * d_pad : 0;
  const size_t probs = wide ? (size_t)probs_block_rows(q_rows, k_rows) * k_rows : 0;
  InfoWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t o_qb = take((size_t)q_rows * d_pad * 2), o_kb = take((size_t)k_rows * d_pad * 2);
  const size_t o_qi = take((size_t)q_rows * 4), o_ki = take((size_t)k_rows * 4);
  const size_t o_pm = take(part * 4), o_pl = take(part * 4);
  const size_t o_gp = take(gpart * 4), o_ex = take((size_t)std::max<long long>(std::max(m, n), 1) * d * 4);
  const size_t o_gb = take(gpart_b * 4), o_pr = take(probs * 2);
  char* b = static_cast<char*>(ws);
  w.qb = reinterpret_cast<__nv_bfloat16*>(b + o_qb); w.kb = reinterpret_cast<__nv_bfloat16*>(b + o_kb);
  w.q_inv = reinterpret_cast<float*>(b + o_qi); w.k_inv = reinterpret_cast<float*>(b + o_ki);
  w.part_m = reinterpret_cast<float*>(b + o_pm); w.part_l = reinterpret_cast<float*>(b + o_pl);
  w.gpart = reinterpret_cast<float*>(b + o_gp); w.extra = reinterpret_cast<float*>(b + o_ex);
  w.gpart_b = reinterpret_cast<float*>(b + o_gb); w.probs = reinterpret_cast<__nv_bfloat16*>(b + o_pr);
  w.q_rows = q_rows; w.k_rows = k_rows;
  w.bytes = off;
  return w;
}

static int prep(const float* x, long long ld, long long n, int d, int cos, float scale, __nv_bfloat16* out, float* inv,
                cudaStream_t st) {
  const long long n_pad = round_up(std::max<long long>(n, 1), kTileN);
  const long long blocks = cdiv(n_pad * 32, 256);
  prep_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, ld, n, d, pad_d(d), n_pad, cos, scale, out, inv);
  GCF_LAUNCH_CHECK("prep_rows_kernel");
  return GCF_OK;
}

static void* align_ws(void* workspace) {
  return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
}

// ------------------------------------------------------------------------------------------------
// DirectAU helpers
// ------------------------------------------------------------------------------------------------
// out3[0] = mean_b |x^_b - y^_b|^2 ; out3[1 + which] = log( sum_i exp(lse_i - 2t) / (B (B-1)) + 1e-8 )
__global__ void __launch_bounds__(256)
directau_finish_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ x_inv,
                       const float* __restrict__ y, long long ldy, const float* __restrict__ y_inv, long long B, int d,
                       float t, const float* __restrict__ lse_x, const float* __restrict__ lse_y, float* __restrict__ out3) {
  __shared__ double sh[3][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double al = 0.0, ux = 0.0, uy = 0.0;
  for (long long b = warp; b < B; b += 8) {
    const float ix = x_inv[b], iy = y_inv[b];
    float s = 0.f;
    for (int c = lane; c < d; c += 32) { const float df = x[b * ldx + c] * ix - y[b * ldy + c] * iy; s = fmaf(df, df, s); }
    al += (double)s;
  }
  for (long long b = threadIdx.x; b < B; b += 256) {
    ux += exp((double)lse_x[b] - 2.0 * (double)t);
    uy += exp((double)lse_y[b] - 2.0 * (double)t);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    al += __shfl_xor_sync(0xffffffffu, al, off); ux += __shfl_xor_sync(0xffffffffu, ux, off); uy += __shfl_xor_sync(0xffffffffu, uy, off);
  }
  if (lane == 0) { sh[0][warp] = al; sh[1][warp] = ux; sh[2][warp] = uy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0, u = 0, v = 0;
    for (int w = 0; w < 8; ++w) { a += sh[0][w]; u += sh[1][w]; v += sh[2][w]; }
    const double pairs = (double)B * (double)(B - 1);
    out3[0] = (float)(a / (double)B);
    out3[1] = (float)log(u / pairs + 1e-8);
    out3[2] = (float)log(v / pairs + 1e-8);
  }
}

// per-row weights of the uniformity gradient: w[i] = w3[k] * 4t / ((m + 1e-8) B (B-1)), l[i] = 2t  with m + 1e-8 = exp(out3[k])
__global__ void __launch_bounds__(256)
directau_weights_kernel(const float* __restrict__ out3, const float* __restrict__ w3, int which, long long B, float t,
                        float* __restrict__ w, float* __restrict__ l) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B) return;
  const float mean_eps = expf(out3[1 + which]);
  w[i] = w3[1 + which] * 4.f * t / (mean_eps * (float)B * (float)(B - 1));
  l[i] = 2.f * t;
}

// extra_x[b] = sign * (2 / B) * w3[0] * (x^_b - y^_b)
__global__ void __launch_bounds__(256)
directau_align_grad_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ x_inv,
                           const float* __restrict__ y, long long ldy, const float* __restrict__ y_inv, long long B, int d,
                           const float* __restrict__ w3, float sign, float* __restrict__ extra) {
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * d) return;
  const long long b = idx / d; const int c = (int)(idx % d);
  extra[idx] = sign * 2.f / (float)B * w3[0] * (x[b * ldx + c] * x_inv[b] - y[b * ldy + c] * y_inv[b]);
}

}  // namespace gcf

using namespace gcf;

extern "C" size_t gcf_infonce_workspace_bytes(int64_t M, int64_t N, int32_t d) {
  if (M < 0 || N < 0 || d <= 0 || d > kMaxChunks * kChunkK) return 0;
  return carve_ws(nullptr, M, N, d).bytes + 1024;
}

static int infonce_check(const char* who, const float* Q, int64_t ldq, int64_t M, const float* Kmat, int64_t ldk, int64_t N,
                         int32_t d, float tau, void* workspace, size_t workspace_bytes) {
  GCF_REQUIRE(M >= 0 && N >= 0, "%s: negative sizes", who);
  if (d <= 0 || d > kMaxChunks * kChunkK) { set_error("%s: d=%d unsupported (1..%d)", who, d, kMaxChunks * kChunkK); return GCF_EUNSUPPORTED; }
  GCF_REQUIRE(tau > 0.f, "%s: temperature must be positive", who);
  if (M == 0) return GCF_OK;
  GCF_REQUIRE(N > 0, "%s: empty key set", who);
  GCF_REQUIRE(Q && Kmat && ldq >= d && ldk >= d, "%s: null operands / bad leading dims", who);
  GCF_REQUIRE(M < (1LL << 31) && N < (1LL << 31), "%s: sizes must fit int32", who);
  const size_t need = gcf_infonce_workspace_bytes(M, N, d);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  return GCF_OK;
}

extern "C" int gcf_infonce_fwd(const float* Q, int64_t ldq, int64_t M, const float* Kmat, int64_t ldk, int64_t N,
                               int32_t d, int32_t cos, float tau, const int64_t* pos_idx, float* row_lse,
                               float* col_lse, float* pos, void* workspace, size_t workspace_bytes,
                               gcf_stream_t stream) {
  int rc = infonce_check("gcf_infonce_fwd", Q, ldq, M, Kmat, ldk, N, d, tau, workspace, workspace_bytes);
  if (rc != GCF_OK || M == 0) return rc;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  InfoWs w = carve_ws(align_ws(workspace), M, N, d);
  const int d_pad = pad_d(d);
  rc = prep(Q, ldq, M, d, cos, kLog2e / tau, w.qb, w.q_inv, st);
  if (rc != GCF_OK) return rc;
  rc = prep(Kmat, ldk, N, d, cos, 1.f, w.kb, w.k_inv, st);
  if (rc != GCF_OK) return rc;
  const bool bounded = cos != 0 && kLog2e / tau <= 64.f;  // |S2| <= log2e / tau: 2^S2 and its sums stay inside fp32
  if (row_lse != nullptr) {
    rc = run_lse(w.qb, M, w.kb, N, d_pad, w.part_m, w.part_l, row_lse, st, 0, bounded);
    if (rc != GCF_OK) return rc;
  }
  if (col_lse != nullptr) {  // column log-sum-exp = row log-sum-exp of the transposed product
    rc = run_lse(w.kb, N, w.qb, M, d_pad, w.part_m, w.part_l, col_lse, st, 0, bounded);
    if (rc != GCF_OK) return rc;
  }
  if (pos != nullptr) {
    pos_logit_kernel<<<(unsigned)cdiv(M * 32, 256), 256, 0, st>>>(w.qb, w.kb, d_pad, M, N, pos_idx, pos);
    GCF_LAUNCH_CHECK("pos_logit_kernel");
  }
  return GCF_OK;
}

extern "C" int gcf_infonce_bwd(const float* Q, int64_t ldq, int64_t M, const float* Kmat, int64_t ldk, int64_t N,
                               int32_t d, int32_t cos, float tau, const int64_t* pos_idx, const float* row_lse,
                               const float* col_lse, const float* w_row, const float* w_col, const float* w_pos,
                               float* gQ, int64_t ldgq, float* gK, int64_t ldgk, void* workspace,
                               size_t workspace_bytes, gcf_stream_t stream) {
  int rc = infonce_check("gcf_infonce_bwd", Q, ldq, M, Kmat, ldk, N, d, tau, workspace, workspace_bytes);
  if (rc != GCF_OK || M == 0) return rc;
  GCF_REQUIRE(w_row == nullptr || row_lse != nullptr, "gcf_infonce_bwd: w_row needs row_lse");
  GCF_REQUIRE(w_col == nullptr || col_lse != nullptr, "gcf_infonce_bwd: w_col needs col_lse");
  GCF_REQUIRE(w_row || w_col || w_pos, "gcf_infonce_bwd: no upstream gradient given");
  GCF_REQUIRE(gQ == nullptr || ldgq >= d, "gcf_infonce_bwd: bad gQ leading dim");
  GCF_REQUIRE(gK == nullptr || ldgk >= d, "gcf_infonce_bwd: bad gK leading dim");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  InfoWs w = carve_ws(align_ws(workspace), M, N, d);
  const int d_pad = pad_d(d);
  rc = prep(Q, ldq, M, d, cos, kLog2e / tau, w.qb, w.q_inv, st);
  if (rc != GCF_OK) return rc;
  rc = prep(Kmat, ldk, N, d, cos, 1.f, w.kb, w.k_inv, st);
  if (rc != GCF_OK) return rc;
  const bool dense = w_row != nullptr || w_col != nullptr;
  const unsigned fin_q = (unsigned)cdiv(M * 32, 256), fin_k = (unsigned)cdiv(N * 32, 256);
  if (d_pad > 4 * kChunkK) {
    // wide embeddings: one block-wise P, both gradient products as library GEMMs, then the same finishing kernel
    if (dense) {
      rc = run_grad_wide(w.qb, M, w.q_rows, w.kb, N, w.k_rows, d_pad, w_row, row_lse, w_col, col_lse, 0,
                         gQ != nullptr ? w.gpart : nullptr, gK != nullptr ? w.gpart_b : nullptr, w.probs, st);
      if (rc != GCF_OK) return rc;
    }
    const int ns = dense ? 1 : 0;
    if (gQ != nullptr) {
      grad_finish_kernel_launch(fin_q, st, w.gpart, ns, w.q_rows, d_pad, 1.f / tau, Q, ldq, w.q_inv, M, d, cos, w_pos, pos_idx,
                                                1.f / tau, Kmat, ldk, w.k_inv, N, nullptr, gQ, ldgq);
      GCF_LAUNCH_CHECK("grad_finish_kernel");
    }
    if (gK != nullptr) {
      const float* extra = nullptr;
      if (w_pos != nullptr) {
        GCF_CUDA(cudaMemsetAsync(w.extra, 0, (size_t)N * d * sizeof(float), st));
        pos_scatter_kernel<<<fin_q, 256, 0, st>>>(Q, ldq, w.q_inv, M, d, cos, w_pos, pos_idx, 1.f / tau, N, w.extra);
        GCF_LAUNCH_CHECK("pos_scatter_kernel");
        extra = w.extra;
      }
      grad_finish_kernel_launch(fin_k, st, w.gpart_b, ns, w.k_rows, d_pad, kLn2, Kmat, ldk, w.k_inv, N, d, cos, nullptr, nullptr,
                                                0.f, nullptr, 0, nullptr, 0, extra, gK, ldgk);
      GCF_LAUNCH_CHECK("grad_finish_kernel");
    }
    return GCF_OK;
  }
  if (gQ != nullptr) {
    // g q^_i = (1/tau) sum_j P_ij k^_j  (+ w_pos[i]/tau k^_pos(i))
    GradPlan gp{0, 0, 0, 0, 0};
    if (dense) {
      rc = run_grad(w.qb, M, w.kb, N, d_pad, w_row, row_lse, w_col, col_lse, 0, w.gpart, &gp, st);
      if (rc != GCF_OK) return rc;
    }
    grad_finish_kernel_launch(fin_q, st, w.gpart, gp.n_splits, gp.a_pad, d_pad, 1.f / tau, Q, ldq, w.q_inv, M, d, cos,
                                              w_pos, pos_idx, 1.f / tau, Kmat, ldk, w.k_inv, N, nullptr, gQ, ldgq);
    GCF_LAUNCH_CHECK("grad_finish_kernel");
  }
  if (gK != nullptr) {
    // g k^_j = (1/tau) sum_i P_ij q^_i = ln2 * sum_i P_ij qb_i   (qb = q^ log2e / tau)   (+ scatter of the positive term)
    GradPlan gp{0, 0, 0, 0, 0};
    if (dense) {
      rc = run_grad(w.kb, N, w.qb, M, d_pad, w_col, col_lse, w_row, row_lse, 0, w.gpart, &gp, st);
      if (rc != GCF_OK) return rc;
    }
    const float* extra = nullptr;
    if (w_pos != nullptr) {
      GCF_CUDA(cudaMemsetAsync(w.extra, 0, (size_t)N * d * sizeof(float), st));
      pos_scatter_kernel<<<fin_q, 256, 0, st>>>(Q, ldq, w.q_inv, M, d, cos, w_pos, pos_idx, 1.f / tau, N, w.extra);
      GCF_LAUNCH_CHECK("pos_scatter_kernel");
      extra = w.extra;
    }
    grad_finish_kernel_launch(fin_k, st, w.gpart, gp.n_splits, gp.a_pad, d_pad, kLn2, Kmat, ldk, w.k_inv, N, d, cos,
                                              nullptr, nullptr, 0.f, nullptr, 0, nullptr, 0, extra, gK, ldgk);
    GCF_LAUNCH_CHECK("grad_finish_kernel");
  }
  return GCF_OK;
}

// ---- DirectAU (directau.py:240-251) on the same two kernels: S = 2t x^ x^T with the diagonal masked -------------
extern "C" size_t gcf_directau_workspace_bytes(int64_t B, int32_t d) {
  if (B < 0 || d <= 0 || d > kMaxChunks * kChunkK) return 0;
  return carve_ws(nullptr, B, B, d).bytes + 1024 + 4 * align_up((size_t)std::max<int64_t>(B, 1) * 4, 1024);
}

struct DauWs { float *lse_x, *lse_y, *w, *l; };
static DauWs carve_dau(void* ws_aligned, int64_t B, int32_t d) {
  char* b = static_cast<char*>(ws_aligned) + carve_ws(nullptr, B, B, d).bytes;
  const size_t step = align_up((size_t)std::max<int64_t>(B, 1) * 4, 1024);
  DauWs r;
  r.lse_x = reinterpret_cast<float*>(b); r.lse_y = reinterpret_cast<float*>(b + step);
  r.w = reinterpret_cast<float*>(b + 2 * step); r.l = reinterpret_cast<float*>(b + 3 * step);
  return r;
}

static int directau_check(const char* who, const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t B, int32_t d,
                          float t, void* workspace, size_t workspace_bytes) {
  if (d <= 0 || d > kMaxChunks * kChunkK) { set_error("%s: d=%d unsupported (1..%d)", who, d, kMaxChunks * kChunkK); return GCF_EUNSUPPORTED; }
  GCF_REQUIRE(B >= 2, "%s: needs at least two rows (pdist of fewer is empty)", who);
  GCF_REQUIRE(B < (1LL << 31), "%s: B must fit int32", who);
  GCF_REQUIRE(x && y && ldx >= d && ldy >= d, "%s: null operands / bad leading dims", who);
  GCF_REQUIRE(t > 0.f, "%s: t must be positive", who);
  const size_t need = gcf_directau_workspace_bytes(B, d);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("%s: workspace too small (%zu < %zu)", who, workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  return GCF_OK;
}

extern "C" int gcf_directau_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t B, int32_t d, float t,
                                float* out3, void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  int rc = directau_check("gcf_directau_fwd", x, ldx, y, ldy, B, d, t, workspace, workspace_bytes);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(out3 != nullptr, "gcf_directau_fwd: null output");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* wsa = align_ws(workspace);
  InfoWs w = carve_ws(wsa, B, B, d);
  DauWs dw = carve_dau(wsa, B, d);
  const int d_pad = pad_d(d);
  const float* src[2] = {x, y};
  const long long lds[2] = {ldx, ldy};
  float* lses[2] = {dw.lse_x, dw.lse_y};
  float* invs[2] = {w.q_inv, w.k_inv};  // kept for the finish kernel: q_inv = 1/|x|, k_inv = 1/|y|
  for (int which = 0; which < 2; ++which) {
    // qb = x^ * 2t log2e, kb = x^   ->  S2 = 2t <x^_i, x^_j> in log2 units
    rc = prep(src[which], lds[which], B, d, 1, 2.f * t * kLog2e, w.qb, invs[which], st);
    if (rc != GCF_OK) return rc;
    rc = prep(src[which], lds[which], B, d, 1, 1.f, w.kb, invs[which], st);
    if (rc != GCF_OK) return rc;
    rc = run_lse(w.qb, B, w.kb, B, d_pad, w.part_m, w.part_l, lses[which], st, 1, 2.f * t * kLog2e <= 64.f);
    if (rc != GCF_OK) return rc;
  }
  directau_finish_kernel<<<1, 256, 0, st>>>(x, ldx, w.q_inv, y, ldy, w.k_inv, B, d, t, dw.lse_x, dw.lse_y, out3);
  GCF_LAUNCH_CHECK("directau_finish_kernel");
  return GCF_OK;
}

extern "C" int gcf_directau_bwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t B, int32_t d, float t,
                                const float* out3, const float* w3, float* gx, int64_t ldgx, float* gy, int64_t ldgy,
                                void* workspace, size_t workspace_bytes, gcf_stream_t stream) {
  int rc = directau_check("gcf_directau_bwd", x, ldx, y, ldy, B, d, t, workspace, workspace_bytes);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(out3 && w3 && gx && gy && ldgx >= d && ldgy >= d, "gcf_directau_bwd: null pointers / bad leading dims");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* wsa = align_ws(workspace);
  InfoWs w = carve_ws(wsa, B, B, d);
  DauWs dw = carve_dau(wsa, B, d);
  const int d_pad = pad_d(d);
  const float* src[2] = {x, y};
  const long long lds[2] = {ldx, ldy};
  float* outs[2] = {gx, gy};
  const long long ldo[2] = {ldgx, ldgy};
  // both inverse norms first (the alignment term needs x^ and y^ together)
  rc = prep(x, ldx, B, d, 1, 1.f, w.qb, w.q_inv, st);
  if (rc != GCF_OK) return rc;
  rc = prep(y, ldy, B, d, 1, 1.f, w.kb, w.k_inv, st);
  if (rc != GCF_OK) return rc;
  float* inv_x = dw.lse_x;  // reuse: the forward's lse vectors are not needed here
  float* inv_y = dw.lse_y;
  GCF_CUDA(cudaMemcpyAsync(inv_x, w.q_inv, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  GCF_CUDA(cudaMemcpyAsync(inv_y, w.k_inv, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
  const float* invs[2] = {inv_x, inv_y};
  const unsigned fin = (unsigned)cdiv(B * 32, 256);
  for (int which = 0; which < 2; ++which) {
    rc = prep(src[which], lds[which], B, d, 1, 2.f * t * kLog2e, w.qb, w.q_inv, st);
    if (rc != GCF_OK) return rc;
    rc = prep(src[which], lds[which], B, d, 1, 1.f, w.kb, w.k_inv, st);
    if (rc != GCF_OK) return rc;
    directau_weights_kernel<<<(unsigned)cdiv(B, 256), 256, 0, st>>>(out3, w3, which, B, t, dw.w, dw.l);
    GCF_LAUNCH_CHECK("directau_weights_kernel");
    GradPlan gp{0, 0, 0, 0, 0};
    if (d_pad > 4 * kChunkK) {
      rc = run_grad_wide(w.qb, B, w.q_rows, w.kb, B, w.k_rows, d_pad, dw.w, dw.l, nullptr, nullptr, 1, w.gpart, nullptr, w.probs, st);
      gp.n_splits = 1; gp.a_pad = w.q_rows;
    } else {
      rc = run_grad(w.qb, B, w.kb, B, d_pad, dw.w, dw.l, nullptr, nullptr, 1, w.gpart, &gp, st);
    }
    if (rc != GCF_OK) return rc;
    directau_align_grad_kernel<<<(unsigned)cdiv(B * d, 256), 256, 0, st>>>(x, ldx, inv_x, y, ldy, inv_y, B, d, w3,
                                                                          which == 0 ? 1.f : -1.f, w.extra);
    GCF_LAUNCH_CHECK("directau_align_grad_kernel");
    grad_finish_kernel_launch(fin, st, w.gpart, gp.n_splits, gp.a_pad, d_pad, 1.f, src[which], lds[which], invs[which],
                                            B, d, 1, nullptr, nullptr, 0.f, nullptr, 0, nullptr, 0, w.extra, outs[which],
                                            ldo[which]);
    GCF_LAUNCH_CHECK("grad_finish_kernel");
  }
  return GCF_OK;
}
