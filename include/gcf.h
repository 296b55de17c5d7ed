/*
 * gcf.h -- C-ABI of libgcf.so: the B200 (sm_100a) graph-collaborative-filtering hot path.
 *
 * The reference (Cmint22/Recommendation) is pure Python/PyTorch and has no FFI of its own
 * (SURVEY.md section 8b); its boundary is the Python class/function surface of each script.
 * Every entry point below therefore cites the reference *Python* code whose arithmetic it
 * replaces (file:line relative to the reference tree).  The Python shims in
 * recommendation_b200/ keep the reference signatures and call these symbols through
 * ctypes (recommendation_b200/_lib.py); INTEGRATION.md shows the binding.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes only; no torch types.
 *   - every pointer is a CUDA *device* pointer owned by the caller unless the parameter
 *     is documented "host".  The library never allocates or frees tensors; scratch is a
 *     caller-provided workspace sized by the matching *_workspace_bytes() query.
 *   - every call is asynchronous on the given stream (a cudaStream_t passed as void*),
 *     has no implicit device synchronisation and no global mutable state.
 *   - return value: 0 = ok, <0 = error; gcf_last_error() returns a thread-local message.
 *   - dense matrices are fp32 row-major with an explicit leading dimension (in elements).
 *   - node / batch indices coming from the reference API are int64 (torch.LongTensor);
 *     CSR structure built by this library is int32.
 */
#ifndef GCF_H_
#define GCF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GCF_OK            0
#define GCF_EINVAL      (-1)
#define GCF_ECUDA       (-2)
#define GCF_EWORKSPACE  (-3)
#define GCF_EUNSUPPORTED (-4)

typedef void* gcf_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------------- */
const char* gcf_version(void);
const char* gcf_last_error(void);

/* ---- (1) adjacency build: integer kernels --------------------------------------------- */

/* deg[idx[e]] += 1 for e in [0,n).  deg is int32[n_nodes], zeroed by the callee.
 * Replaces PyG gcn_norm's scatter_add degree (lightgcn.py:17,25) and scipy's
 * adj.sum(1) for 0/1 adjacencies (selfcf.py:243, ssl4rec.py:85). */
int gcf_degree_count(const int64_t* idx, int64_t n, int32_t* deg, int64_t n_nodes, gcf_stream_t stream);

/* rows = [u | i+U], cols = [i+U | u]  (each int64[2E]).
 * Replaces load_data's edge_index build (lightgcn.py:36-39, gcl.py:72-77). */
int gcf_bipartite_edge_index(const int64_t* users, const int64_t* items, int64_t n_edges, int64_t n_users,
                             int64_t* rows, int64_t* cols, gcf_stream_t stream);

/* COO (int64 rows/cols, optional fp32 vals; NULL vals = all ones) -> canonical CSR:
 * sorted by (row, col) with a hand-written stable LSD radix sort, duplicate (row,col)
 * entries summed in their original order.  Outputs: row_ptr int32[n_rows+1],
 * col_idx int32[capacity nnz], out_vals fp32[capacity nnz], nnz_out int64 device scalar
 * (number of distinct entries; -1 when a row or column index lies outside the matrix --
 * scipy and torch raise on such input, so must the caller).  Replaces scipy's csr_matrix((v,(r,c))) + tmp+tmp.T
 * canonicalisation (selfcf.py:297-306, ssl4rec.py:79-84) and torch's coalescing of the
 * uncoalesced COO tensors (ncl.py:76-85,203-209). */
size_t gcf_coo_to_csr_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols);
int gcf_coo_to_csr_stable(const int64_t* rows, const int64_t* cols, const float* vals, int64_t nnz,
                          int64_t n_rows, int64_t n_cols,
                          int32_t* row_ptr, int32_t* col_idx, float* out_vals, int64_t* nnz_out,
                          void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* Normalise CSR values.  mode 0 = none (copy), 1 = sym: (dinv[r]*a)*dinv[c] with
 * dinv = rowsum^-1/2 and inf -> 0 (square only), 2 = row: rowsum^-1 * a, inf -> 0.
 * rowsum_out fp32[n_rows] receives the row sums (integer valued for 0/1 graphs: the
 * degrees), dinv_out fp32[n_rows] the scaling vector.  Replaces
 * Graph.normalize_graph_mat (selfcf.py:240-255, ncl.py:30-44, mhcn.py:70-84),
 * ssl4rec.py:85-88 and PyG gcn_norm (lightgcn.py:25). */
int gcf_norm_values(int32_t mode, const int32_t* row_ptr, const int32_t* col_idx, const float* vals_in,
                    int64_t n_rows, int64_t n_cols, float* vals_out, float* rowsum_out, float* dinv_out,
                    gcf_stream_t stream);

/* vals_out[j] = (row_scale[r] * vals_in[j]) * col_scale[col_idx[j]]  -- the same two-multiply order as mode 1 above,
 * with the scaling vectors supplied by the caller.  Used for the row blocks of a row-sharded adjacency, where
 * dinv of remote column nodes comes from the global degree vector.  Either vector may be NULL (= 1). */
int gcf_scale_csr_values(const int32_t* row_ptr, const int32_t* col_idx, const float* vals_in, int64_t n_rows,
                         const float* row_scale, const float* col_scale, float* vals_out, gcf_stream_t stream);

/* CSR -> CSR of the transpose (stable: rows ascending inside each output row).
 * Needed for the backward of non-symmetric operators (mhcn.py:440-456 R / R^T,
 * diffnet.py:1127,1131 S and A).  nnz is the exact entry count. */
size_t gcf_csr_transpose_workspace_bytes(int64_t nnz, int64_t n_rows, int64_t n_cols);
int gcf_csr_transpose(const int32_t* row_ptr, const int32_t* col_idx, const float* vals,
                      int64_t n_rows, int64_t n_cols, int64_t nnz,
                      int32_t* t_row_ptr, int32_t* t_col_idx, float* t_vals,
                      void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* ---- (2) SpMM / propagation ----------------------------------------------------------- */

/* CSR operator plus its (optional) long-row schedule.  Rows with more than `chunk`
 * entries are split into chunks of `chunk` entries, each processed by one warp; the
 * last-arriving chunk of a row reduces the partial sums in chunk order (deterministic). */
typedef struct gcf_csr {
  int64_t n_rows, n_cols, nnz;
  const int32_t* row_ptr;        /* [n_rows+1] */
  const int32_t* col_idx;        /* [nnz] */
  const float*   vals;           /* [nnz] */
  int32_t chunk;                 /* long-row threshold / chunk length (0 = no schedule) */
  int32_t n_long;                /* number of rows with degree > chunk */
  int32_t n_chunks;              /* total chunks over all long rows */
  const int32_t* long_rows;      /* [n_long] row ids */
  const int32_t* long_chunk_ptr; /* [n_long+1] first chunk of each long row */
  const int32_t* chunk_long;     /* [n_chunks] index into long_rows */
  /* flat-stream schedule (optional, n_tiles = 0: none).  The rows that are neither long nor empty are cut into tiles
   * of consecutive rows holding about tile_nnz entries each; one sub-warp streams a tile's entries in fixed batches,
   * so the number of row gathers in flight does not depend on the row lengths.  Tiles are expressed in the COMPACT
   * numbering of the non-empty rows: nz_rows[k] is the id of the k-th non-empty row and nz_row_ptr[k] its first
   * entry (nz_row_ptr[n_rows - n_empty] = nnz); both may be NULL when n_empty = 0 (compact = plain numbering).
   * Empty rows receive the epilogue of a zero sum.  Requires the long-row schedule above whenever a row exceeds
   * `chunk` entries. */
  const int32_t* tiles;          /* [2 * n_tiles]: first row, one-past-last row of each tile (compact numbering) */
  int32_t n_tiles;
  int32_t n_empty;
  const int32_t* empty_rows;     /* [n_empty] */
  const int32_t* nz_row_ptr;     /* [n_rows - n_empty + 1] */
  const int32_t* nz_rows;        /* [n_rows - n_empty] */
  /* hub flags (optional, NULL: none): a copy of col_idx whose bit 31 marks the entries that reference a "hub" column
   * (one of the most-referenced columns of that half of the operator).  The flat-stream kernel keeps the rows of hub
   * columns resident in the SM's L1 / shared-memory SRAM (L1::evict_last) and lets all other gathers bypass it
   * (L1::no_allocate): per-SM staging of the hub rows for operators whose working set is L2-resident. */
  const int32_t* hub_col_idx;    /* [nnz] */
} gcf_csr_t;

#define GCF_MAX_ADDENDS 8
#define GCF_EPILOGUE_NONE   0
#define GCF_EPILOGUE_L2NORM 1

/* workspace: n_chunks*d floats of partial sums + n_long int32 counters.  The counter
 * region (the LAST n_long*4 bytes, 256-byte aligned start, see gcf_spmm_counter_offset)
 * must be zero before the first call; the kernel leaves it zero again. */
size_t gcf_spmm_workspace_bytes(const gcf_csr_t* A, int32_t d);
size_t gcf_spmm_counter_offset(const gcf_csr_t* A, int32_t d);

/* T = A * X  (fp32, [n_rows, d]);  then
 *   Y   (nullable) = T
 *   OUT (nullable) = post * ( alpha * f(T) + sum_j betas[j] * addends[j] ),  f = identity | row-L2-normalise
 * addends / betas are HOST arrays (n_addends <= GCF_MAX_ADDENDS) of device pointers / scalars,
 * every addend has leading dimension ld_out.
 * Replaces torch.sparse.mm (ncl.py:419, selfcf.py:479, directau.py:290, mhcn.py:440-456,
 * diffnet.py:1127,1131), PyG LGConv.propagate (lightgcn.py:25) and the per-layer
 * F.normalize / stack().mean() / x += out epilogues (ncl.py:421, selfcf.py:481-482,
 * lightgcn.py:26, sept.py:224, mhcn.py:441-457). */
int gcf_spmm_csr_f32(const gcf_csr_t* A, int32_t d, const float* X, int64_t ldx,
                     float* Y, int64_t ldy, float* OUT, int64_t ld_out,
                     int32_t epilogue, float alpha, float post,
                     int32_t n_addends, const float* const* addends, const float* betas,
                     void* workspace, size_t workspace_bytes, int32_t variant, gcf_stream_t stream);

/* K-layer LightGCN propagation: E0 = X0, E(k) = A * E(k-1);
 *   final = scale * sum_{k=0..K} E(k)        (scale = 1/(K+1) for mean, 1 for sum)
 * layers: HOST array of K device pointers [n_rows, d] (ld = d) receiving E(1)..E(K);
 * layers[K-1] may be NULL when E(K) itself is not needed (it is folded into `final`).
 * Replaces LGCNEncoder.forward (ncl.py:415-422, directau.py:286-293),
 * LGCN_Encoder.forward (selfcf.py:475-485), LightGCN.forward (lightgcn.py:21-27). */
int gcf_propagate_fwd(const gcf_csr_t* A, int32_t d, int32_t n_layers, const float* X0,
                      float* const* layers, float* final_out, float scale,
                      void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* Backward of gcf_propagate_fwd for a symmetric operator A (or pass A^T):
 *   G(K) = scale*g_final + extra[K];  G(k) = A^T * G(k+1) + scale*g_final + extra[k];  g_X0 = G(0)
 * extra: HOST array of K+1 nullable device pointers (gradients that reached individual
 * layer outputs E(k), e.g. NCL's ssl_layer_loss, ncl.py:318-323); may be NULL.
 * ping/pong: two [n_rows, d] scratch buffers.  Replaces autograd's sparse addmm backward
 * (SURVEY.md row a10). */
int gcf_propagate_bwd(const gcf_csr_t* At, int32_t d, int32_t n_layers, const float* g_final,
                      const float* const* extra, float scale, float* ping, float* pong, float* g_x0,
                      void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* gcf_propagate_bwd with the optimiser fused into the epilogue of its LAST SpMM: G(0), the gradient of the embedding
 * table, is consumed row by row by an in-place Adam / AdamW update of param / exp_avg / exp_avg_sq ([n_rows, d], ld = d;
 * torch.optim.Adam numerics, see gcf_adam_step) instead of being written out and read back.  g_x0 may be NULL (the
 * gradient is then never materialised) or a buffer that additionally receives it. */
int gcf_propagate_bwd_adam(const gcf_csr_t* At, int32_t d, int32_t n_layers, const float* g_final,
                           const float* const* extra, float scale, float* ping, float* pong, float* g_x0,
                           float* param, float* exp_avg, float* exp_avg_sq, float lr, float beta1, float beta2, float eps,
                           float weight_decay, int32_t decoupled, int64_t step,
                           void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* ---- (4) gather / scatter-add / sampler ----------------------------------------------- */

/* out[t, :] = table[idx[t], :]   (x[idx], ncl.py:314-316, lightgcn.py:95-102, ...) */
int gcf_gather_rows(const float* table, int64_t ld, int64_t n_table_rows, int32_t d,
                    const int64_t* idx, int64_t n, float* out, int64_t ld_out, gcf_stream_t stream);

/* table_grad[idx[t], :] += src[t, :]   (backward of x[idx]: index_put_(accumulate=True)).
 * mode 0: warp-aggregated atomics (match.any groups equal indices inside a warp, one
 *         red.global.add.v4.f32 per distinct row per warp) -- fp32 order not deterministic;
 * mode 1: deterministic -- indices radix-sorted (stable) and each destination row summed
 *         in source order by one sub-warp.  Needs the workspace. */
size_t gcf_scatter_add_workspace_bytes(int64_t n, int64_t n_table_rows, int32_t mode);
int gcf_scatter_add_rows(const float* src, int64_t ld_src, int32_t d, const int64_t* idx, int64_t n,
                         float* table_grad, int64_t ld, int64_t n_table_rows, int32_t mode,
                         void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* Layout conversion around the collectives of the feature-sharded multi-GPU trainers (the reference has
 * no distributed code).  `blocked` = [n_slices][n_rows][w] fp32 -- what an all-gather / all-to-all of per-rank
 * [n_rows, w] column slices delivers -- and `rows` = the row-major [n_rows, n_slices * w] table (leading dimension
 * ld_rows) that the fused loss kernels gather from.  w % 4 == 0, 16-byte aligned buffers. */
int gcf_slices_to_rows(const float* blocked, float* rows, int64_t ld_rows, int64_t n_rows, int32_t n_slices, int32_t w,
                       gcf_stream_t stream);
int gcf_rows_to_slices(const float* rows, int64_t ld_rows, float* blocked, int64_t n_rows, int32_t n_slices, int32_t w,
                       gcf_stream_t stream);

/* Peer-memory exchange of the feature-sharded multi-GPU step over NVLink / NVSwitch (the reference has no
 * distributed code -- these replace the NCCL all-gather / all-to-all / reduce-scatter + layout-pass pairs around the loss).
 * Buffers that peers read are cudaMalloc blocks of their own (gcf_peer_alloc: zero-filled, exportable at offset 0);
 * gcf_peer_export writes the 64-byte CUDA IPC handle that another PROCESS on the same box turns into a device pointer with
 * gcf_peer_open (peer access enabled lazily).  The three movers take an array of n_src (<= 16) source pointers, local or
 * peer mappings, all with leading dimension ld_src; w % 4 == 0, 16-byte aligned pointers:
 *   gather_cols: dst[r, g*w : (g+1)*w] = src[g][r, 0:w]             r < n_rows   (column slices -> full-width rows)
 *   sum_cols   : dst[r, 0:w] = src[0][r, 0:w] + ... + src[n_src-1][r, 0:w]       (fixed order: deterministic)
 *   copy_blocks: dst[off_g + r * ld_dst + 0:w] = src[g][r, 0:w], r < rows_per_src[g]; off_g = dst_offsets[g] floats, or, with
 *                dst_offsets == NULL, the blocks stacked: off_g = (rows_per_src[0] + ... + rows_per_src[g-1]) * ld_dst
 * The caller orders the ranks (data ready before the call, buffers not rewritten while peers read): the kernels themselves
 * do not synchronise across devices. */
#define GCF_PEER_HANDLE_BYTES 64
int gcf_peer_alloc(size_t bytes, void** dev_ptr);
int gcf_peer_free(void* dev_ptr);
int gcf_peer_export(const void* dev_ptr, void* handle64);
int gcf_peer_open(const void* handle64, void** peer_ptr);
int gcf_peer_close(void* peer_ptr);
/* Device-side barrier of the n_ranks processes on `stream`: flag_arrays[g] = rank g's zero-initialised array of >= n_ranks
 * uint32 (peer mappings except flag_arrays[rank]); epoch = 1, 2, 3, ... (compared modulo 2^32), the same on every rank for the same barrier.  Work
 * enqueued on the stream before the call (on any rank) is complete and visible before work enqueued after it starts. */
int gcf_peer_barrier(void* const* flag_arrays, int32_t n_ranks, int32_t rank, uint64_t epoch, gcf_stream_t stream);
int gcf_peer_gather_cols(const float* const* src, int32_t n_src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t n_rows,
                         int32_t w, gcf_stream_t stream);
int gcf_peer_sum_cols(const float* const* src, int32_t n_src, int64_t ld_src, float* dst, int64_t ld_dst, int64_t n_rows,
                      int32_t w, gcf_stream_t stream);
/* the general mover: n_blocks (<= 16) independent 2-D blocks, dst[g][r * ld_dst + 0:w] = src[g][r * ld_src + 0:w], r < rows[g];
 * either side of a block may be peer memory (pull = remote loads, push = remote stores).  ctas_per_block = 0: the library's default
 * (a small fixed budget shared by the blocks: NVLink throughput DROPS when too many requests are outstanding). */
int gcf_peer_copy2d(const float* const* src, float* const* dst, const int64_t* rows, int32_t n_blocks, int64_t ld_src,
                    int64_t ld_dst, int32_t w, int32_t ctas_per_block, gcf_stream_t stream);
int gcf_peer_copy_blocks(const float* const* src, const int64_t* rows_per_src, const int64_t* dst_offsets, int32_t n_src,
                         int64_t ld_src, float* dst, int64_t ld_dst, int32_t w, gcf_stream_t stream);

/* Philox4x32-10 counter-based negative sampler.  For triple t and negative slot j the
 * candidate stream is Philox(key=seed, counter=(t*n_negs+j, trial/4, offset_lo, offset_hi)),
 * candidate = mulhi32(word, n_items).  Without a positives CSR (pos_row_ptr == NULL) the
 * first candidate is taken (lightgcn.py:91-94: torch.randint, no rejection); with it,
 * candidates found in the user's sorted positive list are rejected, up to max_trials
 * (ncl.py:91-114 caps at 100; selfcf.py:188-211 / directau.py:14-32 are unbounded).
 * out int64[n * n_negs]. */
int gcf_sample_negatives(uint64_t seed, uint64_t offset, const int64_t* users, int64_t n, int32_t n_negs,
                         int64_t n_items, const int32_t* pos_row_ptr, const int32_t* pos_col_idx,
                         int32_t max_trials, int64_t* out, gcf_stream_t stream);

/* The same stream for a WINDOW of the slot numbering: local slot s of this call draws what slot slot_base + s of a
 * call over the whole list draws (users / out are indexed locally).  A rank that owns triples [t0, t1) of the global
 * list passes slot_base = t0 * n_negs, so the negatives do not depend on how the triples are sharded. */
int gcf_sample_negatives_at(uint64_t seed, uint64_t offset, int64_t slot_base, const int64_t* users, int64_t n,
                            int32_t n_negs, int64_t n_items, const int32_t* pos_row_ptr, const int32_t* pos_col_idx,
                            int32_t max_trials, int64_t* out, gcf_stream_t stream);

/* The same stream for an ARBITRARY subset of the triples: slot_pos[t] = position of local triple t in the global list, so
 * ranks that own interleaved subsets (users dealt out cyclically) draw what the single-GPU run draws.  No rejection. */
int gcf_sample_negatives_pos(uint64_t seed, uint64_t offset, const int64_t* slot_pos, int64_t n, int32_t n_negs,
                             int64_t n_items, int64_t* out, gcf_stream_t stream);

/* Stochastic edge dropout of a sparse operator's values (buir.py:300-309 sparse_dropout):
 *   out[j] = keep(e) ? vals[e] / (1 - rate) : 0,   e = index ? index[j] : j,
 *   keep(e) <=> Philox4x32-10(counter = (e, offset), key = seed)[0] < (1 - rate) * 2^32
 * Every stored entry is kept independently with probability 1 - rate (rate = 0 keeps all: thresh = 2^32 - 1 drops one
 * value in 2^32).  With index = the transposition permutation of the CSR the same call yields the values of the dropped
 * operator's TRANSPOSE (same mask per entry), which the backward of the propagation needs. */
int gcf_csr_dropout_values(const float* vals, int64_t n, const int32_t* index, float rate, uint64_t seed, uint64_t offset,
                           float* out, gcf_stream_t stream);

/* out[j] = Philox4x32-10(counter = (j, offset), key = seed)[0] in [0, 2^32): one uniform key per entry.  The `keep`
 * smallest keys select a uniformly random subset of exact size (GraphAugmentor.edge_dropout, sept.py:53-62:
 * np.random.choice(idx, int(n * (1 - drop_rate)), replace=False)). */
int gcf_philox_keys(int64_t n, uint64_t seed, uint64_t offset, int64_t* out, gcf_stream_t stream);

/* ---- (3) losses ----------------------------------------------------------------------- */

#define GCF_BPR_LOG_EPS_SIGMOID 0 /* -log(eps + sigmoid(x))      ncl.py:116-120, mhcn.py:35-39 */
#define GCF_BPR_SOFTPLUS        1 /* -log(sigmoid(x))            lightgcn.py:108, gcl.py:221   */
#define GCF_BPR_RAW_SCORE       2 /* gcf_bpr_fwd only: coef_out[t] = x_t (the score itself), loss_out = the reg terms.
                                     For tables sharded along the feature dimension: every rank holds d/G columns, the
                                     partial scores are summed across ranks and gcf_bpr_coef_from_scores applies the loss. */
#define GCF_REDUCE_MEAN 0
#define GCF_REDUCE_SUM  1         /* diffnet.py:1113 */

/* Fused gather + BPR (+ squared-L2 reg of the gathered rows):
 *   x_t   = <u_t, p_t> - mean_j <u_t, n_tj>
 *   loss  = reduce_t l(x_t) + reg_u*sum|u_t|^2 + reg_p*sum|p_t|^2 + reg_n*sum|n_tj|^2
 * coef_out[t] = dl/dx_t (already divided by n for the mean) is kept for the backward.
 * lightgcn.py:95-118: (reg, reg, 0); gcl.py:216-223: reg/B each; ncl.py:116-120: zeros. */
size_t gcf_bpr_workspace_bytes(int64_t n_triples);
int gcf_bpr_fwd(const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item, int32_t d,
                const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n_triples, int32_t n_negs,
                int32_t variant, float eps, int32_t reduction, float reg_u, float reg_p, float reg_n,
                float* loss_out, float* coef_out, void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* Backward: accumulates (warp-aggregated red.add) into g_user / g_item, which the caller
 * zero-initialises (or which already hold other gradient terms).  grad_out: device scalar
 * dL/dloss (NULL = 1). */
int gcf_bpr_bwd(const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item, int32_t d,
                const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n_triples, int32_t n_negs,
                const float* coef, const float* grad_out, float reg_u, float reg_p, float reg_n,
                float* g_user, int64_t ldg_user, float* g_item, int64_t ldg_item, gcf_stream_t stream);

/* loss = reduce_t l(x_t), coef[t] = dl/dx_t (already divided by n for the mean) from complete scores x[n]:
 * the pointwise half of gcf_bpr_fwd, used after the partial scores of feature-sharded tables were all-reduced. */
int gcf_bpr_coef_from_scores(const float* x, int64_t n_triples, int32_t variant, float eps, int32_t reduction,
                             float* loss_out, float* coef_out, void* workspace, size_t workspace_bytes,
                             gcf_stream_t stream);

/* Forward and backward in ONE pass (the three row gathers are done once): the loss is a scalar, so its upstream
 * gradient is known before the backward starts (lightgcn.py:119 `loss.backward()`: 1) and is passed as the host
 * scalar grad_scale.  Accumulates grad_scale * dloss/d(rows) into g_user / g_item exactly like gcf_bpr_bwd and
 * writes the loss like gcf_bpr_fwd; coef_out is optional (NULL = not kept). */
int gcf_bpr_fwd_bwd(const float* user_emb, int64_t ld_user, const float* item_emb, int64_t ld_item, int32_t d,
                    const int64_t* u_idx, const int64_t* p_idx, const int64_t* n_idx, int64_t n_triples, int32_t n_negs,
                    int32_t variant, float eps, int32_t reduction, float reg_u, float reg_p, float reg_n,
                    float grad_scale, float* loss_out, float* coef_out,
                    float* g_user, int64_t ldg_user, float* g_item, int64_t ldg_item,
                    void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* Fused dense Adam / AdamW step (torch.optim.Adam semantics, ncl.py:305, lightgcn.py:80).
 * step = 1-based step count. */
int gcf_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                  float lr, float beta1, float beta2, float eps, float weight_decay, int32_t decoupled,
                  int64_t step, gcf_stream_t stream);

/* Fused torch.optim.SGD step (selfcf.py:544, directau.py:214, univariate/selfcf_univariate.py:553: lr, momentum = 0.9):
 *   g' = g + weight_decay*p;  buf = g' when first_step != 0, else momentum*buf + (1 - dampening)*g';
 *   p -= lr * (nesterov ? g' + momentum*buf : buf)           (momentum = 0: p -= lr*g', momentum_buf may be NULL)
 * one pass over param / grad / momentum_buf, same operation order as torch's single-tensor implementation. */
int gcf_sgd_momentum_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                          float dampening, float weight_decay, int32_t nesterov, int32_t first_step, gcf_stream_t stream);

/* Row-sparse Adam for the mini-batch models: only rows[0..n_rows) of param / exp_avg / exp_avg_sq
 * ([*, d], leading dimension ld) are updated, with the gradient rows given densely in list order ([n_rows, d], ld_grad).
 * rows must be distinct.  Same arithmetic as gcf_adam_step on the touched rows (untouched rows keep their moments:
 * torch.optim.SparseAdam semantics, NOT the dense torch.optim.Adam the reference uses -- see DESIGN.md). */
int gcf_adam_rows_step(float* param, int64_t ld, const float* grad_rows, int64_t ld_grad, float* exp_avg, float* exp_avg_sq,
                       const int64_t* rows, int64_t n_rows, int32_t d, float lr, float beta1, float beta2, float eps,
                       float weight_decay, int32_t decoupled, int64_t step, gcf_stream_t stream);

/* Lloyd's k-means on the device: NCL's E-step (ncl.py:339-356, faiss.Kmeans(d, k).train(x) + index.search(x, 1), which the
 * reference runs on the CPU inside every batch, ncl.py:313,324).  x [n, d] fp32 (leading dimension ldx); centroids [k, d]
 * fp32 contiguous, IN: the initial centroids, OUT: the trained ones.  n_iter Lloyd iterations (faiss default 25), then one
 * more assignment pass: assign[n] (int64, nullable) = nearest centroid, dist[n] (nullable) = squared distance to it.
 * Distances on the tcgen05 tensor cores with bf16 hi/lo split operands (~fp32 accuracy), ties to the lowest index;
 * centroid sums in a fixed order (deterministic); empty clusters re-seeded from the largest ones, all on the device:
 * no host synchronisation.  d % 4 == 0, d <= 1024, 1 <= k <= n. */
size_t gcf_kmeans_workspace_bytes(int64_t n, int32_t k, int32_t d);
int gcf_kmeans_lloyd(const float* x, int64_t ldx, int64_t n, int32_t d, int32_t k, int32_t n_iter, float* centroids,
                     int64_t* assign, float* dist, void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* x[0..n) *= *g (device scalar); returns without touching memory when *g == 1.  Used to apply the upstream
 * gradient to tables whose gradient was produced in the forward pass (gcf_bpr_fwd_bwd). */
int gcf_scale_by_device_scalar(float* x, int64_t n, const float* g, gcf_stream_t stream);

/* ---- InfoNCE family on tcgen05 tensor cores (bf16 operands, fp32 accumulate + LSE) ----
 *
 * Q [M, d] fp32, Kmat [N, d] fp32 (rows are L2-normalised inside when cos != 0).
 *   s_ij = <q_i, k_j> / tau
 *   row_lse[i] = log sum_j exp(s_ij)                      (always)
 *   col_lse[j] = log sum_i exp(s_ij)                      (when col_lse != NULL; gcl.py:28-35)
 *   pos[i]     = s_{i, pos_idx[i]}  (pos_idx NULL -> diagonal j = i)
 * The B x N logits are never materialised.  The scalar losses are assembled from these
 * vectors by the shim:
 *   InfoNCE          ncl.py:125-130, ssl4rec.py:19-23 : mean_i(row_lse - pos)
 *   ssl_layer_loss   ncl.py:358-367                   : sum_i(row_lse - pos)
 *   batch_softmax    ssl4rec.py:25-30                 : mean_i -log(exp(pos-row_lse)+1e-6)
 *   info_nce_loss    gcl.py:28-35                     : (mean(row_lse-pos)+mean(col_lse-pos))/2
 */
/* d <= 1024 (the reference's tuner grids reach embedding.size = 1024).  d <= 256 runs the fully fused TMEM kernels; wider
 * embeddings stream both operands through the TMA ring in the forward, and the backward materialises P block-wise in bf16
 * (<= 1 GiB of workspace at a time) and runs the two gradient products as cuBLAS bf16 GEMMs with fp32 accumulation. */
size_t gcf_infonce_workspace_bytes(int64_t M, int64_t N, int32_t d);
int gcf_infonce_fwd(const float* Q, int64_t ldq, int64_t M, const float* Kmat, int64_t ldk, int64_t N, int32_t d,
                    int32_t cos, float tau, const int64_t* pos_idx,
                    float* row_lse, float* col_lse, float* pos,
                    void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* Backward: given w_row[i] = dL/d row_lse[i], w_col[j] = dL/d col_lse[j] (nullable),
 * w_pos[i] = dL/d pos[i]:   dS_ij = w_row[i] softmax_row_ij + w_col[j] softmax_col_ij + w_pos[i] [j == pos_i]
 * and gQ += dS K / tau, gK += dS^T Q / tau, chained through the row normalisation when
 * cos != 0.  gQ [M,d] / gK [N,d] are overwritten (not accumulated). */
int gcf_infonce_bwd(const float* Q, int64_t ldq, int64_t M, const float* Kmat, int64_t ldk, int64_t N, int32_t d,
                    int32_t cos, float tau, const int64_t* pos_idx,
                    const float* row_lse, const float* col_lse,
                    const float* w_row, const float* w_col, const float* w_pos,
                    float* gQ, int64_t ldgq, float* gK, int64_t ldgk,
                    void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* DirectAU (directau.py:240-251).  x, y: [B, d] fp32 rows (normalised inside).
 *   align   = mean_b |x^_b - y^_b|^2
 *   unif(x) = log( mean_{i<j} exp(-t * |x^_i - x^_j|^2) + 1e-8 ),  Gram matrix on tensor cores
 * out[0] = align, out[1] = unif(x), out[2] = unif(y). */
size_t gcf_directau_workspace_bytes(int64_t B, int32_t d);
int gcf_directau_fwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t B, int32_t d, float t,
                     float* out3, void* workspace, size_t workspace_bytes, gcf_stream_t stream);
/* w3 = device [3] upstream gradients of (align, unif(x), unif(y)); gx, gy overwritten. */
int gcf_directau_bwd(const float* x, int64_t ldx, const float* y, int64_t ldy, int64_t B, int32_t d, float t,
                     const float* out3, const float* w3, float* gx, int64_t ldgx, float* gy, int64_t ldgy,
                     void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* ---- sparse x sparse products for the motif-induced adjacency matrices ----------------------
 * univariate/mhcn.py:340-368 (build_hyper_adj_mats): sixteen (P.dot(Q)).multiply(M) terms plus Y.dot(Y.T), evaluated
 * by scipy on the host in the reference.  All operands are canonical CSR (rows sorted, columns ascending, int32 indices).
 * The long-row schedule fields of gcf_csr_t are ignored here. */

/* out[e] = X[i_e, j_e] for every stored entry e = (i_e, j_e) of the pattern P, 0 where X stores nothing.
 * With it: S.multiply(S.T) = S.vals * sample(S^T, S); S - B on S's pattern; alignment of two matrices on one pattern. */
int gcf_csr_sample(const gcf_csr_t* X, const gcf_csr_t* P, float* out, gcf_stream_t stream);

/* (A . B).multiply(M) on M's pattern without forming A . B:
 *   out[e] = M.vals[e] * sum_k A[i_e, k] * B[k, j_e],   B given by the CSR of its transpose (Bt: row j = column j of B).
 * A: [m, K], Bt: [n, K], mask: [m, n]; out: float[mask.nnz]. */
int gcf_spgemm_masked(const gcf_csr_t* A, const gcf_csr_t* Bt, const gcf_csr_t* mask, float* out, gcf_stream_t stream);

/* Full product C = A . B by expand-sort-compress, over the stored entries [entry_begin, entry_end) of A (cut at row
 * boundaries by the caller to bound memory):
 *   gcf_spgemm_count   *n_products (device int64) = number of scalar products; their offsets stay in the workspace;
 *   gcf_spgemm_expand  writes them as COO (row of A, column of B, a * b) in emission order -- requires
 *                      *n_products < 2^32 (split the range otherwise); feed the COO to gcf_coo_to_csr_stable, which
 *                      sums duplicates in that order (scipy's Y.dot(Y.T) up to fp32 summation order; exact for integer data). */
size_t gcf_spgemm_workspace_bytes(int64_t n_entries);
int gcf_spgemm_count(const gcf_csr_t* A, const gcf_csr_t* B, int64_t entry_begin, int64_t entry_end, int64_t* n_products,
                     void* workspace, size_t workspace_bytes, gcf_stream_t stream);
int gcf_spgemm_expand(const gcf_csr_t* A, const gcf_csr_t* B, int64_t entry_begin, int64_t entry_end, int64_t* rows_out,
                      int64_t* cols_out, float* vals_out, void* workspace, size_t workspace_bytes, gcf_stream_t stream);

/* ---- text ingest -----------------------------------------------------------------------------
 * `user item rating` text files -> id arrays -> dense indices, for the loaders of the reference: load_data
 * (ncl.py:542-543: `line.strip().split()[:2]`, blank lines skipped), Interaction._build (ncl.py:55-70: ids numbered by
 * sorted() of the id STRINGS; selfcf.py:281-288: by first appearance) and lightgcn.py:29-33 (integer ids, pandas).
 * `text` is the file content in device memory.  A record is a line with at least one non-blank byte. */

/* *n_records (device int64) = number of records; per-segment record numbers stay in the workspace for the parse. */
size_t gcf_text_workspace_bytes(int64_t n_bytes);
int gcf_text_count_records(const uint8_t* text, int64_t n_bytes, int64_t* n_records, void* workspace, size_t workspace_bytes,
                           gcf_stream_t stream);
/* first[r], second[r] = the first two whitespace-separated tokens of record r as keys.
 *   mode 0: string keys -- up to 8 bytes, left-aligned big-endian, zero padded: unsigned key order == byte-wise
 *           lexicographic order of the ids (what Python's sorted() gives for ASCII strings: "10" < "2");
 *   mode 1: unsigned decimal integers (lightgcn.py's integer ids).
 * *status (device int32) collects bit 1 = a record with fewer than two tokens, bit 2 = a token outside the key format. */
int gcf_text_parse_pairs(const uint8_t* text, int64_t n_bytes, int32_t mode, uint64_t* first, uint64_t* second,
                         int32_t* status, void* workspace, size_t workspace_bytes, gcf_stream_t stream);
/* uniq[0 .. *n_uniq) = the distinct keys in ascending order (stable LSD radix sort + run heads);
 * first_pos[j] (nullable) = position in `keys` of the first occurrence of uniq[j]. */
size_t gcf_sort_unique_workspace_bytes(int64_t n);
int gcf_sort_unique_u64(const uint64_t* keys, int64_t n, uint64_t* uniq, int64_t* first_pos, int64_t* n_uniq,
                        void* workspace, size_t workspace_bytes, gcf_stream_t stream);
/* idx[i] = position of keys[i] in the ascending table, -1 when absent (ids that only occur in the test file). */
int gcf_lookup_sorted_u64(const uint64_t* table, int64_t n_table, const uint64_t* keys, int64_t n, int64_t* idx,
                          gcf_stream_t stream);

/* The same three steps for ids of up to 8 * n_words bytes (ncl.py:60-61 sorts arbitrary id strings): a key is n_words
 * 64-bit words (left-aligned, big-endian, zero padded), stored word-major: word w of record r at keys[w * n + r]; the
 * lexicographic order of the word tuples is the byte-wise order of the strings.
 *   gcf_text_parse_pairs_words  first / second [n_words][n_records]; status[0]: bit 1 = a record with fewer than two fields,
 *                               bit 2 = a token longer than 8 * n_words bytes; status[1] = length of the longest token
 *                               (call with n_words = 1 first, widen if bit 2 is set).  Workspace: as gcf_text_parse_pairs.
 *   gcf_sort_unique_words       stable LSD radix sort word by word + run heads: uniq [n_words][n] (word-major with stride n)
 *                               distinct keys ascending, first_pos / n_uniq as above, rank[n] (nullable) = position of every
 *                               input key in the distinct table (its dense "sorted" index).
 *   gcf_lookup_sorted_words     binary search with word-tuple comparison; table is word-major with stride table_stride. */
int gcf_text_parse_pairs_words(const uint8_t* text, int64_t n_bytes, int32_t n_words, int64_t n_records, uint64_t* first,
                               uint64_t* second, int32_t* status, void* workspace, size_t workspace_bytes, gcf_stream_t stream);
size_t gcf_sort_unique_words_workspace_bytes(int64_t n);
int gcf_sort_unique_words(const uint64_t* keys, int32_t n_words, int64_t n, uint64_t* uniq, int64_t* first_pos, int64_t* n_uniq,
                          int64_t* rank, void* workspace, size_t workspace_bytes, gcf_stream_t stream);
int gcf_lookup_sorted_words(const uint64_t* table, int32_t n_words, int64_t n_table, int64_t table_stride, const uint64_t* keys,
                            int64_t n, int64_t* idx, gcf_stream_t stream);

/* ---- batched evaluation ---------------------------------------------------
 *
 * scores [n_queries, n_items] fp32 (ld elements per row) is the dense score block user_emb[q] . item_emb^T -- a plain
 * GEMM, produced by the caller (cuBLAS).  For every row: the training items of user users[q] (NULL -> q), given as a
 * per-user sorted CSR, are overwritten IN PLACE with mask_value (-1e8 in ncl.py:257-259), then the n_top (<= 128) best
 * items are selected exactly (score descending, ties by ascending item id) into out_idx / out_val [n_queries, n_top].
 * Replaces the predict / mask / torch.topk loop of test() (ncl.py:253-266, lightgcn.py:48-74). */
int gcf_masked_topn(float* scores, int64_t ld, int64_t n_queries, int64_t n_items, const int64_t* users,
                    const int32_t* pos_row_ptr, const int32_t* pos_col_idx, float mask_value, int32_t n_top,
                    int64_t* out_idx, float* out_val, gcf_stream_t stream);

/* hits[q, c] = |top-cutoffs[c] of row q  ∩  test items of users[q]|, dcg[q, c] = sum_{hit at rank i < cutoffs[c]} 1/log2(i+2)
 * (Metric.hits / Metric.NDCG, ncl.py:136-162); cutoffs ascending int32 device array.  test CSR: per-user sorted items. */
int gcf_ranking_hits(const int64_t* topn, int64_t n_queries, int32_t n_top, const int64_t* users,
                     const int32_t* test_row_ptr, const int32_t* test_col_idx, const int32_t* cutoffs, int32_t n_cutoffs,
                     int32_t* hits, float* dcg, gcf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GCF_H_ */
