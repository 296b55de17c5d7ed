// Peer-memory exchange of the feature-sharded multi-GPU step (the reference has no distributed code).
//
// In the feature-sharded layout (dist.py) every rank owns d/G columns of all [N, d] tables and only the BPR loss needs
// full-width rows.  r01/r02 moved the column slices with NCCL (all-gather / all-to-all / reduce-scatter) and converted the
// layout with a separate pass on either side.  Here ONE kernel per direction reads the peers' buffers directly over
// NVLink / NVSwitch (CUDA IPC mappings of cudaMalloc'd buffers, every GPU of an NVSwitch box is a peer of every other)
// and writes the layout the consumer wants:
//
//   gcf_peer_gather_cols : dst[r, g*w : (g+1)*w] = src_g[r, 0:w]          -- "all-gather + slices -> rows" in one pass
//   gcf_peer_sum_cols    : dst[r, 0:w] = sum_g src_g[r, 0:w] (g ascending) -- "rows -> slices + reduce-scatter" in one pass
//   gcf_peer_copy_blocks : dst[off_g + r * ld_dst + 0:w] = src_g[r, 0:w]   -- "rows -> slices + all-to-all" in one pass
//
// src_g may be local or a peer mapping; all sources share one leading dimension.  Remote reads are 16-byte loads with
// several independent requests in flight per thread (NVLink round trips are ~2 us: the link is filled by the number of
// outstanding requests, not by the issue rate); consecutive threads read consecutive addresses of ONE source, so requests
// leave as full 128-byte lines whenever the source rows are contiguous (ld_src == w).
// Ordering between ranks (data ready / buffer free) is the caller's: a stream-ordered barrier before each call.
#include "common.cuh"
#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace gcf {

constexpr int kMaxPeers = 16;
struct PeerSrc { const float4* p[kMaxPeers]; };
struct PeerBlocks { const float4* src[kMaxPeers]; float4* dst[kMaxPeers]; long long rows[kMaxPeers]; };

// idx -> (row, float4 column) of a [*, w4] block; w4 is a power of two for every d/G the trainers use
__device__ __forceinline__ void split_rc(long long idx, int w4, int shift, long long& r, int& k) {
  if (shift >= 0) { r = idx >> shift; k = (int)(idx & (w4 - 1)); }
  else { r = idx / w4; k = (int)(idx - r * w4); }
}
static int pow2_shift(int w4) { int s = 0; while ((1 << s) < w4) ++s; return (1 << s) == w4 ? s : -1; }

__device__ __forceinline__ float4 ld_peer(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// n blocks, block g = rows[g] x w4 float4: dst[g][r * ld_dst4 + k] = src[g][r * ld_src4 + k].  blockIdx.y = g, so every
// block (every peer) is being moved at the same time; either side of a block may be peer memory (pull: remote loads,
// push: remote stores -- stores are posted, loads pay the NVLink round trip).
template <int UNR>
__global__ void __launch_bounds__(256)
peer_copy2d_kernel(PeerBlocks b, long long ld_src4, long long ld_dst4, int w4, int shift) {
  const int g = blockIdx.y;
  const long long total = b.rows[g] * w4;
  const float4* __restrict__ src = b.src[g];
  float4* __restrict__ dst = b.dst[g];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * UNR) {
    float4 v[UNR];
    long long o[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long idx = base + u * stride;
      o[u] = -1;
      if (idx < total) {
        long long r;
        int k;
        split_rc(idx, w4, shift, r, k);
        v[u] = ld_peer(src + r * ld_src4 + k);
        o[u] = r * ld_dst4 + k;
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (o[u] >= 0) dst[o[u]] = v[u];
  }
}

// Stream-ordered barrier between the GPUs of the group, on the device: every rank stores the epoch into slot `rank` of every
// peer's flag array (system-scope release: all earlier writes of this stream, local or remote, are visible first) and then
// waits until its own array shows the epoch in every slot.  One warp; the kernel of a rank can only finish once every other
// rank has LAUNCHED the same barrier, so all ranks must issue their barriers in the same order.
struct PeerFlags { unsigned* p[kMaxPeers]; };
__global__ void peer_barrier_kernel(PeerFlags flags, int n, int rank, unsigned epoch) {
  const int g = threadIdx.x;
  if (g < n) {
    __threadfence_system();
    // max, not a plain store: barriers issued from two streams of one rank may execute out of issue order, and a flag
    // that stepped back from epoch e + 1 to e would strand a peer that waits for e + 1
    asm volatile("red.release.sys.global.max.u32 [%0], %1;" ::"l"(flags.p[g] + rank), "r"(epoch) : "memory");
    unsigned v;
    unsigned long long t0 = 0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    unsigned spins = 0;
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags.p[rank] + g) : "memory");
      if ((++spins & 0xfffu) == 0) {            // a peer that never arrives (crashed process) must not hang the GPU for ever
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 30ull * 1000000000ull) __trap();
      }
    } while ((int)(v - epoch) < 0);
    __threadfence_system();
  }
}

template <int UNR>
__global__ void __launch_bounds__(256)
peer_sum_kernel(PeerSrc src, int n_src, long long ld_src4, float4* __restrict__ dst, long long ld_dst4, int w4, int shift,
                long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long base = (long long)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * UNR) {
    long long so[UNR], doff[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const long long idx = base + u * stride;
      doff[u] = -1;
      so[u] = 0;
      if (idx < total) {
        long long r;
        int k;
        split_rc(idx, w4, shift, r, k);
        so[u] = r * ld_src4 + k;
        doff[u] = r * ld_dst4 + k;
      }
    }
    float4 acc[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) acc[u] = f4_zero();
    // two sources at a time: 2 * UNR independent remote loads in flight, sums taken in ascending source order
    for (int g = 0; g < n_src; g += 2) {
      float4 a[UNR], b[UNR];
#pragma unroll
      for (int u = 0; u < UNR; ++u) {
        a[u] = f4_zero(); b[u] = f4_zero();
        if (doff[u] >= 0) {
          a[u] = ld_peer(src.p[g] + so[u]);
          if (g + 1 < n_src) b[u] = ld_peer(src.p[g + 1] + so[u]);
        }
      }
#pragma unroll
      for (int u = 0; u < UNR; ++u) { f4_add(acc[u], a[u]); if (g + 1 < n_src) f4_add(acc[u], b[u]); }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u)
      if (doff[u] >= 0) dst[doff[u]] = acc[u];
  }
}

static int peer_args_ok(const char* who, const float* const* src, int32_t n_src, int64_t ld_src, const float* dst, int64_t ld_dst,
                        int32_t w) {
  GCF_REQUIRE(src != nullptr && dst != nullptr, "%s: null source array or destination", who);
  GCF_REQUIRE(n_src >= 1 && n_src <= kMaxPeers, "%s: n_src must be in [1, %d]", who, kMaxPeers);
  GCF_REQUIRE(w > 0 && (w & 3) == 0 && ld_src >= w && (ld_src & 3) == 0 && (ld_dst & 3) == 0, "%s: w, ld_src, ld_dst must be multiples of 4 floats", who);
  GCF_REQUIRE((reinterpret_cast<uintptr_t>(dst) & 15u) == 0, "%s: destination must be 16-byte aligned", who);
  for (int g = 0; g < n_src; ++g)
    GCF_REQUIRE(src[g] != nullptr && (reinterpret_cast<uintptr_t>(src[g]) & 15u) == 0, "%s: source %d null or misaligned", who, g);
  return GCF_OK;
}

// Grid sizing.  Measured on 4 B200s (tools/peer_bw.py, profiles/r03_peer_bw.md): with 8 CTAs per SM of remote loads in flight
// the links deliver 365-420 GB/s per GPU, with ~100 CTAs in total 620 GB/s -- past a few MB of outstanding requests the
// NVLink request queues thrash.  So the movers run a FIXED, small number of CTAs (GCF_PEER_CTAS, default 128, shared by the
// blocks of a call) instead of filling the machine.
static int peer_cta_budget() {
  static const int v = [] { const char* e = getenv("GCF_PEER_CTAS"); const int x = e ? atoi(e) : 0; return x > 0 ? x : 128; }();
  return v;
}
static int peer_grid(long long total, int unr) {
  const long long want = cdiv(total, 256LL * unr);
  return (int)std::max<long long>(1, std::min<long long>(want, (long long)peer_cta_budget()));
}
static int peer_ctas_per_block(long long elems_of_largest_block, int unr, int n_blocks) {
  const long long want = cdiv(elems_of_largest_block, 256LL * unr);
  const long long cap = std::max(8, peer_cta_budget() / std::max(1, n_blocks - 1));   // one of the blocks is usually local
  return (int)std::max<long long>(1, std::min(want, cap));
}

static int launch_blocks(const PeerBlocks& b, int n_blocks, long long most_rows, int64_t ld_src, int64_t ld_dst, int32_t w,
                         int ctas_per_block, gcf_stream_t stream) {
  if (most_rows == 0) return GCF_OK;
  constexpr int UNR = 8;
  const long long elems = most_rows * (w / 4);
  const long long want = cdiv(elems, 256LL * UNR);
  const long long cap = ctas_per_block > 0 ? ctas_per_block : peer_ctas_per_block(elems, UNR, n_blocks);
  const dim3 grid((unsigned)std::max<long long>(1, std::min(want, cap)), (unsigned)n_blocks);
  peer_copy2d_kernel<UNR><<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(b, ld_src / 4, ld_dst / 4, w / 4, pow2_shift(w / 4));
  GCF_LAUNCH_CHECK("peer_copy2d_kernel");
  return GCF_OK;
}

}  // namespace gcf

using namespace gcf;

extern "C" int gcf_peer_alloc(size_t bytes, void** dev_ptr) {
  GCF_REQUIRE(dev_ptr != nullptr && bytes > 0, "gcf_peer_alloc: null output or zero size");
  *dev_ptr = nullptr;
  GCF_CUDA(cudaMalloc(dev_ptr, bytes));     // a plain cudaMalloc block: exportable with cudaIpcGetMemHandle at offset 0
  GCF_CUDA(cudaMemset(*dev_ptr, 0, bytes));
  return GCF_OK;
}

extern "C" int gcf_peer_free(void* dev_ptr) {
  if (dev_ptr != nullptr) GCF_CUDA(cudaFree(dev_ptr));
  return GCF_OK;
}

extern "C" int gcf_peer_export(const void* dev_ptr, void* handle64) {
  GCF_REQUIRE(dev_ptr != nullptr && handle64 != nullptr, "gcf_peer_export: null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == GCF_PEER_HANDLE_BYTES, "cudaIpcMemHandle_t is 64 bytes");
  cudaIpcMemHandle_t h;
  GCF_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(dev_ptr)));
  memcpy(handle64, &h, sizeof(h));
  return GCF_OK;
}

extern "C" int gcf_peer_open(const void* handle64, void** peer_ptr) {
  GCF_REQUIRE(handle64 != nullptr && peer_ptr != nullptr, "gcf_peer_open: null argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  *peer_ptr = nullptr;
  GCF_CUDA(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GCF_OK;
}

extern "C" int gcf_peer_close(void* peer_ptr) {
  if (peer_ptr != nullptr) GCF_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return GCF_OK;
}

extern "C" int gcf_peer_barrier(void* const* flag_arrays, int32_t n_ranks, int32_t rank, uint64_t epoch, gcf_stream_t stream) {
  GCF_REQUIRE(flag_arrays != nullptr && n_ranks >= 1 && n_ranks <= kMaxPeers && rank >= 0 && rank < n_ranks,
              "gcf_peer_barrier: bad rank / n_ranks (<= %d)", kMaxPeers);
  PeerFlags f{};
  for (int g = 0; g < n_ranks; ++g) {
    GCF_REQUIRE(flag_arrays[g] != nullptr, "gcf_peer_barrier: null flag array %d", g);
    f.p[g] = static_cast<unsigned*>(flag_arrays[g]);
  }
  peer_barrier_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(f, n_ranks, rank, (unsigned)epoch);
  GCF_LAUNCH_CHECK("peer_barrier_kernel");
  return GCF_OK;
}

extern "C" int gcf_peer_gather_cols(const float* const* src, int32_t n_src, int64_t ld_src, float* dst, int64_t ld_dst,
                                    int64_t n_rows, int32_t w, gcf_stream_t stream) {
  int rc = peer_args_ok("gcf_peer_gather_cols", src, n_src, ld_src, dst, ld_dst, w);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(n_rows >= 0 && ld_dst >= (int64_t)n_src * w, "gcf_peer_gather_cols: ld_dst < n_src * w");
  PeerBlocks b{};
  for (int g = 0; g < n_src; ++g) {
    b.src[g] = reinterpret_cast<const float4*>(src[g]);
    b.dst[g] = reinterpret_cast<float4*>(dst + (long long)g * w);
    b.rows[g] = n_rows;
  }
  return launch_blocks(b, n_src, n_rows, ld_src, ld_dst, w, 0, stream);
}

extern "C" int gcf_peer_copy_blocks(const float* const* src, const int64_t* rows_per_src, const int64_t* dst_offsets,
                                    int32_t n_src, int64_t ld_src, float* dst, int64_t ld_dst, int32_t w, gcf_stream_t stream) {
  int rc = peer_args_ok("gcf_peer_copy_blocks", src, n_src, ld_src, dst, ld_dst, w);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(rows_per_src != nullptr && ld_dst >= w, "gcf_peer_copy_blocks: null row counts or ld_dst < w");
  PeerBlocks b{};
  long long rows = 0, most = 0;
  for (int g = 0; g < n_src; ++g) {
    GCF_REQUIRE(rows_per_src[g] >= 0, "gcf_peer_copy_blocks: negative row count");
    GCF_REQUIRE(dst_offsets == nullptr || (dst_offsets[g] >= 0 && (dst_offsets[g] & 3) == 0),
                "gcf_peer_copy_blocks: destination offsets must be non-negative multiples of 4 floats");
    b.src[g] = reinterpret_cast<const float4*>(src[g]);
    b.dst[g] = reinterpret_cast<float4*>(dst + (dst_offsets != nullptr ? dst_offsets[g] : rows * ld_dst));
    b.rows[g] = rows_per_src[g];
    rows += rows_per_src[g];
    most = std::max<long long>(most, rows_per_src[g]);
  }
  return launch_blocks(b, n_src, most, ld_src, ld_dst, w, 0, stream);
}

extern "C" int gcf_peer_copy2d(const float* const* src, float* const* dst, const int64_t* rows, int32_t n_blocks, int64_t ld_src,
                               int64_t ld_dst, int32_t w, int32_t ctas_per_block, gcf_stream_t stream) {
  GCF_REQUIRE(src != nullptr && dst != nullptr && rows != nullptr, "gcf_peer_copy2d: null arrays");
  GCF_REQUIRE(n_blocks >= 1 && n_blocks <= kMaxPeers, "gcf_peer_copy2d: n_blocks must be in [1, %d]", kMaxPeers);
  GCF_REQUIRE(w > 0 && (w & 3) == 0 && ld_src >= w && ld_dst >= w && (ld_src & 3) == 0 && (ld_dst & 3) == 0,
              "gcf_peer_copy2d: w, ld_src, ld_dst must be multiples of 4 floats with ld >= w");
  PeerBlocks b{};
  long long most = 0;
  for (int g = 0; g < n_blocks; ++g) {
    GCF_REQUIRE(rows[g] >= 0, "gcf_peer_copy2d: negative row count");
    GCF_REQUIRE(rows[g] == 0 || (src[g] != nullptr && dst[g] != nullptr && ((reinterpret_cast<uintptr_t>(src[g]) | reinterpret_cast<uintptr_t>(dst[g])) & 15u) == 0),
                "gcf_peer_copy2d: block %d null or misaligned", g);
    b.src[g] = reinterpret_cast<const float4*>(src[g]);
    b.dst[g] = reinterpret_cast<float4*>(dst[g]);
    b.rows[g] = rows[g];
    most = std::max<long long>(most, rows[g]);
  }
  return launch_blocks(b, n_blocks, most, ld_src, ld_dst, w, ctas_per_block, stream);
}

extern "C" int gcf_peer_sum_cols(const float* const* src, int32_t n_src, int64_t ld_src, float* dst, int64_t ld_dst,
                                 int64_t n_rows, int32_t w, gcf_stream_t stream) {
  int rc = peer_args_ok("gcf_peer_sum_cols", src, n_src, ld_src, dst, ld_dst, w);
  if (rc != GCF_OK) return rc;
  GCF_REQUIRE(n_rows >= 0 && ld_dst >= w, "gcf_peer_sum_cols: ld_dst < w");
  if (n_rows == 0) return GCF_OK;
  PeerSrc ps{};
  for (int g = 0; g < n_src; ++g) ps.p[g] = reinterpret_cast<const float4*>(src[g]);
  const long long total = n_rows * (w / 4);
  constexpr int UNR = 4;
  peer_sum_kernel<UNR><<<peer_grid(total, UNR), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      ps, n_src, ld_src / 4, reinterpret_cast<float4*>(dst), ld_dst / 4, w / 4, pow2_shift(w / 4), total);
  GCF_LAUNCH_CHECK("peer_sum_kernel");
  return GCF_OK;
}
