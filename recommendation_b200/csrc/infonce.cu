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
// and four epilogue warps read the accumulator back with tcgen05.ld and fold it into an online log-sum-exp.
//
// Numerics: operands are L2-normalised (when cos != 0) in fp32, scaled by log2(e)/tau on the "query" side and
// rounded to bf16 once; accumulation, running max / sum and every reduction are fp32.  Logit error is bounded
// by the bf16 rounding of the operands (<= 2e-2 absolute for tau >= 0.05, the north-star tolerance).
#include "common.cuh"
#include "tc05.cuh"
#include <cuda_bf16.h>
#include <algorithm>
#include <cmath>
#include <mutex>

namespace gcf {

using namespace tc;

constexpr int kTileM = 128;      // rows of the stationary operand per CTA (= TMEM lanes)
constexpr int kTileN = 256;      // rows of the streamed operand per MMA tile (= TMEM columns per stage)
constexpr int kChunkK = 64;      // bf16 elements per 128-byte swizzled row
constexpr int kLseStages = 4;    // TMA ring depth (one stage = one [kTileN x 64] chunk = 32 KB)
constexpr int kLseThreads = 192; // warp 0: TMA, warp 1: MMA + TMEM alloc, warps 2-5: epilogue
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

template <int KC>  // number of 64-wide K chunks (d_pad = 64 * KC)
__global__ void __launch_bounds__(kLseThreads, 1)
lse_stream_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, long long n_b,
                  int n_tiles, int tiles_per_split, float* __restrict__ part_m, float* __restrict__ part_l,
                  long long m_pad) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;                                               // KC chunks of [128 x 64] bf16
  uint8_t* smem_b = smem + KC * LseSmem::kABytesPerChunk;               // kLseStages chunks of [256 x 64] bf16
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
      for (int s = 0; s < 2; ++s) { mbar_init(acc_full + s, 1); mbar_init(acc_empty + s, 4); }
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
      mbar_expect_tx(a_bar, KC * LseSmem::kABytesPerChunk);
      for (int kc = 0; kc < KC; ++kc) tma_load_2d(&tm_a, a_bar, smem_a + kc * LseSmem::kABytesPerChunk, kc * kChunkK, m_tile * kTileM);
      int stage = 0; uint32_t phase = 0;
      for (int t = t_begin; t < t_end; ++t) {
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(empty_bar + stage, phase ^ 1);
          mbar_expect_tx(full_bar + stage, LseSmem::kBBytesPerStage);
          tma_load_2d(&tm_b, full_bar + stage, smem_b + stage * LseSmem::kBBytesPerStage, kc * kChunkK, t * kTileN);
          if (++stage == kLseStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0 && my_tiles > 0) {
      constexpr uint32_t idesc = idesc_bf16_f32(kTileM, kTileN, 0, 0);
      mbar_wait(a_bar, 0);
      fence_after_sync();
      int stage = 0; uint32_t phase = 0;
      for (int it = 0; it < my_tiles; ++it) {
        const int acc = it & 1;
        mbar_wait(acc_empty + acc, ((it >> 1) & 1) ^ 1);  // epilogue has drained this accumulator stage
        fence_after_sync();
        for (int kc = 0; kc < KC; ++kc) {
          mbar_wait(full_bar + stage, phase);
          fence_after_sync();
          const uint32_t a_addr = smem_u32(smem_a + kc * LseSmem::kABytesPerChunk);
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
    // ===== epilogue: thread = one row of the tile; online base-2 log-sum-exp =====
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    float m_run = -INFINITY, l_run = 0.f;
    for (int it = 0; it < my_tiles; ++it) {
      const int acc = it & 1;
      const int t = t_begin + it;
      mbar_wait(acc_full + acc, (it >> 1) & 1);
      fence_after_sync();
      const long long col0 = (long long)t * kTileN;
      const bool ragged = col0 + kTileN > n_b;       // tile contains zero-padded rows of B: mask them out
#pragma unroll 1
      for (int c = 0; c < kTileN / 32; ++c) {
        float v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(acc * kTileN + c * 32), v);
        if (ragged) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col0 + c * 32 + j >= n_b) v[j] = -INFINITY;
        }
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
            s0 += exp2f(v[j] - m_run); s1 += exp2f(v[j + 1] - m_run);
            s2 += exp2f(v[j + 2] - m_run); s3 += exp2f(v[j + 3] - m_run);
          }
          l_run += (s0 + s1) + (s2 + s3);
        }
      }
      fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(acc_empty + acc);
    }
    const long long out = (long long)split * m_pad + (long long)m_tile * kTileM + row;
    part_m[out] = m_run;
    part_l[out] = l_run;
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
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// bf16 row-major [rows, d_pad] matrix, box = [box_rows x 64] with the 128-byte swizzle
static int make_tmap(CUtensorMap* map, const void* base, long long rows, int d_pad, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (fn == nullptr) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return GCF_ECUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)d_pad, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)d_pad * sizeof(__nv_bfloat16)};
  cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return GCF_ECUDA; }
  return GCF_OK;
}

static inline long long round_up(long long x, long long m) { return (x + m - 1) / m * m; }
static inline int pad_d(int d) { return (int)round_up(d, kChunkK); }

struct LsePlan { long long a_pad, b_pad; int m_tiles, n_tiles, n_splits, tiles_per_split; };

static LsePlan plan_lse(long long n_a, long long n_b) {
  LsePlan p;
  p.a_pad = round_up(std::max<long long>(n_a, 1), kTileM);
  p.b_pad = round_up(std::max<long long>(n_b, 1), kTileN);
  p.m_tiles = (int)(p.a_pad / kTileM);
  p.n_tiles = (int)(p.b_pad / kTileN);
  const int sms = sm_count();
  int splits = std::max(1, std::min(p.n_tiles, (sms + p.m_tiles - 1) / p.m_tiles));
  p.tiles_per_split = (p.n_tiles + splits - 1) / splits;
  p.n_splits = (p.n_tiles + p.tiles_per_split - 1) / p.tiles_per_split;
  return p;
}

static size_t lse_smem_bytes(int kc) {
  return 1024 + (size_t)kc * LseSmem::kABytesPerChunk + (size_t)kLseStages * LseSmem::kBBytesPerStage + 256;
}

// (m, l) partials -> lse[n_a]; ab/bb are the prepared bf16 operands ([a_pad, d_pad], [b_pad, d_pad])
static int run_lse(const __nv_bfloat16* ab, long long n_a, const __nv_bfloat16* bb, long long n_b, int d_pad,
                   float* part_m, float* part_l, float* lse_out, cudaStream_t st) {
  const LsePlan p = plan_lse(n_a, n_b);
  CUtensorMap tm_a, tm_b;
  int rc = make_tmap(&tm_a, ab, p.a_pad, d_pad, kTileM);
  if (rc != GCF_OK) return rc;
  rc = make_tmap(&tm_b, bb, p.b_pad, d_pad, kTileN);
  if (rc != GCF_OK) return rc;
  const int kc = d_pad / kChunkK;
  const size_t smem = lse_smem_bytes(kc);
  dim3 grid(p.m_tiles, p.n_splits);
#define GCF_LSE_LAUNCH(KC)                                                                                          \
  do {                                                                                                              \
    GCF_CUDA(cudaFuncSetAttribute(lse_stream_kernel<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    lse_stream_kernel<KC><<<grid, kLseThreads, smem, st>>>(tm_a, tm_b, n_b, p.n_tiles, p.tiles_per_split, part_m,   \
                                                           part_l, p.a_pad);                                       \
  } while (0)
  switch (kc) {
    case 1: GCF_LSE_LAUNCH(1); break;
    case 2: GCF_LSE_LAUNCH(2); break;
    case 3: GCF_LSE_LAUNCH(3); break;
    case 4: GCF_LSE_LAUNCH(4); break;
    default: set_error("infonce: d_pad=%d unsupported (d <= 256)", d_pad); return GCF_EUNSUPPORTED;
  }
#undef GCF_LSE_LAUNCH
  GCF_LAUNCH_CHECK("lse_stream_kernel");
  combine_lse_kernel<<<(unsigned)cdiv(n_a, 256), 256, 0, st>>>(part_m, part_l, p.n_splits, p.a_pad, n_a, lse_out);
  GCF_LAUNCH_CHECK("combine_lse_kernel");
  return GCF_OK;
}

struct InfoWs {
  __nv_bfloat16 *qb, *kb;
  float *q_inv, *k_inv, *part_m, *part_l;
  size_t bytes;
};

static InfoWs carve_ws(void* ws, long long m, long long n, int d) {
  const int d_pad = pad_d(d);
  const LsePlan pq = plan_lse(m, n), pk = plan_lse(n, m);
  // operands are padded for BOTH roles (stationary: multiple of 128, streamed: multiple of 256)
  const long long q_rows = round_up(std::max<long long>(m, 1), kTileN), k_rows = round_up(std::max<long long>(n, 1), kTileN);
  const size_t part = std::max((size_t)pq.n_splits * pq.a_pad, (size_t)pk.n_splits * pk.a_pad);
  InfoWs w;
  size_t off = 0;
  auto take = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 1024); return o; };
  const size_t o_qb = take((size_t)q_rows * d_pad * 2), o_kb = take((size_t)k_rows * d_pad * 2);
  const size_t o_qi = take((size_t)q_rows * 4), o_ki = take((size_t)k_rows * 4);
  const size_t o_pm = take(part * 4), o_pl = take(part * 4);
  char* b = static_cast<char*>(ws);
  w.qb = reinterpret_cast<__nv_bfloat16*>(b + o_qb); w.kb = reinterpret_cast<__nv_bfloat16*>(b + o_kb);
  w.q_inv = reinterpret_cast<float*>(b + o_qi); w.k_inv = reinterpret_cast<float*>(b + o_ki);
  w.part_m = reinterpret_cast<float*>(b + o_pm); w.part_l = reinterpret_cast<float*>(b + o_pl);
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

}  // namespace gcf

using namespace gcf;

extern "C" size_t gcf_infonce_workspace_bytes(int64_t M, int64_t N, int32_t d) {
  if (M < 0 || N < 0 || d <= 0 || d > 256) return 0;
  return carve_ws(nullptr, M, N, d).bytes + 1024;
}

extern "C" int gcf_infonce_fwd(const float* Q, int64_t ldq, int64_t M, const float* Kmat, int64_t ldk, int64_t N,
                               int32_t d, int32_t cos, float tau, const int64_t* pos_idx, float* row_lse,
                               float* col_lse, float* pos, void* workspace, size_t workspace_bytes,
                               gcf_stream_t stream) {
  GCF_REQUIRE(M >= 0 && N >= 0, "gcf_infonce_fwd: negative sizes");
  if (d <= 0 || d > 256) { set_error("gcf_infonce_fwd: d=%d unsupported (1..256)", d); return GCF_EUNSUPPORTED; }
  GCF_REQUIRE(tau > 0.f, "gcf_infonce_fwd: temperature must be positive");
  if (M == 0) return GCF_OK;
  GCF_REQUIRE(N > 0, "gcf_infonce_fwd: empty key set");
  GCF_REQUIRE(Q && Kmat && ldq >= d && ldk >= d, "gcf_infonce_fwd: null operands / bad leading dims");
  GCF_REQUIRE(M < (1LL << 31) && N < (1LL << 31), "gcf_infonce_fwd: sizes must fit int32");
  const size_t need = gcf_infonce_workspace_bytes(M, N, d);
  if (workspace == nullptr || workspace_bytes < need) {
    set_error("gcf_infonce_fwd: workspace too small (%zu < %zu)", workspace_bytes, need);
    return GCF_EWORKSPACE;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  void* ws_aligned = reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  InfoWs w = carve_ws(ws_aligned, M, N, d);
  const int d_pad = pad_d(d);
  int rc = prep(Q, ldq, M, d, cos, kLog2e / tau, w.qb, w.q_inv, st);
  if (rc != GCF_OK) return rc;
  rc = prep(Kmat, ldk, N, d, cos, 1.f, w.kb, w.k_inv, st);
  if (rc != GCF_OK) return rc;
  if (row_lse != nullptr) {
    rc = run_lse(w.qb, M, w.kb, N, d_pad, w.part_m, w.part_l, row_lse, st);
    if (rc != GCF_OK) return rc;
  }
  if (col_lse != nullptr) {  // column log-sum-exp = row log-sum-exp of the transposed product
    rc = run_lse(w.kb, N, w.qb, M, d_pad, w.part_m, w.part_l, col_lse, st);
    if (rc != GCF_OK) return rc;
  }
  if (pos != nullptr) {
    pos_logit_kernel<<<(unsigned)cdiv(M * 32, 256), 256, 0, st>>>(w.qb, w.kb, d_pad, M, N, pos_idx, pos);
    GCF_LAUNCH_CHECK("pos_logit_kernel");
  }
  return GCF_OK;
}

// ---- backward / DirectAU: implemented below in later commits (fail loudly until then) ----------
extern "C" int gcf_infonce_bwd(const float*, int64_t, int64_t, const float*, int64_t, int64_t, int32_t, int32_t, float,
                               const int64_t*, const float*, const float*, const float*, const float*, const float*,
                               float*, int64_t, float*, int64_t, void*, size_t, gcf_stream_t) {
  set_error("gcf_infonce_bwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
extern "C" size_t gcf_directau_workspace_bytes(int64_t, int32_t) { return 0; }
extern "C" int gcf_directau_fwd(const float*, int64_t, const float*, int64_t, int64_t, int32_t, float, float*, void*,
                                size_t, gcf_stream_t) {
  set_error("gcf_directau_fwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
extern "C" int gcf_directau_bwd(const float*, int64_t, const float*, int64_t, int64_t, int32_t, float, const float*,
                                const float*, float*, int64_t, float*, int64_t, void*, size_t, gcf_stream_t) {
  set_error("gcf_directau_bwd: not implemented in this build");
  return GCF_EUNSUPPORTED;
}
