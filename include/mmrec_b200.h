/* mmrec_b200 -- C ABI of the B200-native hot path of MMRec.
 *
 * The reference (enoche/MMRec, /root/reference) is pure Python on PyTorch; it has no FFI of
 * its own.  Its "plugin boundary" is the model class contract of
 * src/common/abstract_recommender.py:10-52,71-103, and the hot path below that boundary is the
 * set of PyTorch library calls listed next to each entry point.  This header is what a
 * maintainer binds (ctypes stub in INTEGRATION.md) to replace exactly those calls.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (normally a torch tensor); the
 *    library never allocates or frees device memory: scratch is passed in as `ws`, sized by the
 *    matching *_workspace_bytes() call;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*; 0 = legacy default);
 *  - return value 0 = ok, negative = MMREC_E* below; mmrec_last_error() gives the message
 *    (thread-local);
 *  - all matrices are row-major; floating point is fp32 (the reference computes in fp32),
 *    CSR indices are int32, COO indices / gather indices / result indices are int64 as in the
 *    reference's tensors.
 */
#ifndef MMREC_B200_H
#define MMREC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMREC_OK 0
#define MMREC_EINVAL (-1)      /* bad argument (null pointer, negative size, unsupported d or k) */
#define MMREC_EWORKSPACE (-2)  /* ws_bytes smaller than *_workspace_bytes() */
#define MMREC_ECUDA (-3)       /* a CUDA runtime call failed; see mmrec_last_error() */
#define MMREC_EUNSUPPORTED (-4)/* device is not sm_100 (the library carries sm_100a code only) */

#define MMREC_ABI_VERSION 3

int mmrec_abi_version(void);
const char* mmrec_last_error(void);
/* 0 if the current device can run this library (compute capability 10.x), else MMREC_EUNSUPPORTED */
int mmrec_device_check(void);
/* Kernels of this library launched by the calling process so far (every launch site counts itself; library
 * primitives -- CUB sort/scan, memsets -- are not counted).  bench.py reports the difference over its timed region. */
int64_t mmrec_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * K1c  COO -> CSR.   Replaces the per-call `coalesce()` + COO->CSR conversion hidden inside every
 * `torch.sparse.mm(adj, x)` on the path (src/models/freedom.py:167,172; bm3.py:90;
 * mgcn.py:162,172,176,180,184; layergcn.py:131; lightgcn.py:120; common/encoders.py:99,122).
 * The reference's matrices arrive un-coalesced and, for FREEDOM's mm_adj, with duplicate
 * coordinates that must add (freedom.py:74): duplicates are summed in input order (stable sort),
 * which is what coalesce() does.  Output rows are sorted by column.
 *   row/col  int64[nnz]; val fp32[nnz] or NULL (= all ones)
 *   rowptr   int32[n_rows+1]; colidx int32[nnz]; vals fp32[nnz]  (first nnz_out[0] entries valid)
 *   nnz_out  int64[1] on the device: number of entries after merging duplicates
 * ------------------------------------------------------------------------------------------- */
size_t mmrec_csr_from_coo_workspace_bytes(int64_t nnz, int64_t n_rows);
int mmrec_csr_from_coo(int64_t nnz, const int64_t* row, const int64_t* col, const float* val,
                       int64_t n_rows, int64_t n_cols, int sum_duplicates,
                       int32_t* rowptr, int32_t* colidx, float* vals, int64_t* nnz_out,
                       void* ws, size_t ws_bytes, void* stream);

/* Work plan for mmrec_spmm_f32.  A task is a whole row or, for rows longer than `seg` non-zeros, one segment of it.
 * Tasks longer than `light_max` are run by a whole CTA (its lane groups split the task and reduce through shared
 * memory), the others by one lane group each -- the power-law item rows neither serialise on one warp nor sit on
 * the critical path.
 *   tasks      int32[4 * max_tasks]  {row, begin, end, split_id(-1 = whole row)}, sorted longest first, so the
 *                                    CTA-run tasks are the first counts[4] entries
 *   split_rows int32[4 * max_split]  {first_slot, n_seg, row_begin, seg}
 *   counts     int64[8] on the device: {n_tasks, n_split_rows, n_slots, longest_row, n_cta_tasks, 0, 0, 0}
 * max_tasks = n_rows + nnz / seg + 1 and max_split = nnz / seg + 1 are always enough. */
size_t mmrec_spmm_plan_workspace_bytes(int64_t n_rows, int64_t max_tasks);
int mmrec_spmm_plan(int64_t n_rows, const int32_t* rowptr, int seg, int light_max, int64_t max_tasks,
                    int32_t* tasks, int32_t* split_rows, int64_t* counts,
                    void* ws, size_t ws_bytes, void* stream);

/* Per-edge symmetric normalisation of a bipartite edge list, fp32, computed like
 * src/models/freedom.py:145-154 (`_normalize_adj_m`): deg counted exactly, then
 * val[e] = rsqrt(deg_u[u] + eps) * rsqrt(deg_i[i] + eps) with IEEE sqrt and divide.
 *   ws: int32[n_users + n_items] */
size_t mmrec_bipartite_norm_workspace_bytes(int64_t n_users, int64_t n_items);
int mmrec_bipartite_norm_f32(int64_t n_edges, const int64_t* users, const int64_t* items,
                             int64_t n_users, int64_t n_items, float eps, float* vals,
                             void* ws, size_t ws_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K1  CSR SpMM with fused layer-combination epilogue.   Replaces `torch.sparse.mm(adj, x)` and the
 * `stack(...).mean(dim=1)` / `+ h` that follow it (src/models/freedom.py:164-178, bm3.py:84-95,
 * lightgcn.py:115-128, mgcn.py:157-185, layergcn.py:125-138); its backward is the same call on the
 * transposed CSR (autograd of torch.sparse.mm, triggered at src/common/trainer.py:185).
 *
 *   y[r,:]   = sum_j vals[j] * X[colidx[j], :]           j in rowptr[r] .. rowptr[r+1]
 *   if gate_ref: y[r,:] *= cos(y[r,:], gate_ref[r,:])    (LayerGCN, layergcn.py:132-133)
 *   if Y:        Y[r,:] = y[r,:]
 *   if acc_out:  acc_out[r,:] = ((acc_in ? acc_in[r,:] : 0) + y[r,:]) / acc_div
 *
 * acc_in may alias acc_out.  d in {32, 64, 128, 256} runs the vectorised kernel, any other d >= 1
 * the generic one.  tasks/n_tasks/n_cta_tasks/split_rows/counters/partial come from mmrec_spmm_plan
 * (counters: int32[n_split_rows], zero on entry and zero again on exit; partial:
 * fp32[n_slots * d]); tasks == NULL selects one warp per row.  Summation order is fixed, so the
 * result is bit-reproducible run to run.
 * ------------------------------------------------------------------------------------------- */
/* tuning knob: lanes that cooperate on one row (0 = default min(32, d/4); a power of two, d/(4*lanes) float4 per lane) */
int mmrec_spmm_set_lanes(int lanes_per_row);
int mmrec_spmm_f32(int64_t n_rows, int64_t n_cols, int d,
                   const int32_t* rowptr, const int32_t* colidx, const float* vals,
                   const int32_t* tasks, int64_t n_tasks, int64_t n_cta_tasks, const int32_t* split_rows,
                   int32_t* counters, float* partial,
                   const float* X, int64_t ldx,
                   float* Y, int64_t ldy,
                   const float* acc_in, float* acc_out, int64_t ldacc, float acc_div,
                   const float* gate_ref, int64_t ldgate,
                   void* stream);

/* mmrec_spmm_f32 without the gate and with an ACCUMULATING Y (y_accumulate != 0: Y[r,:] += y[r,:]; the running-sum epilogue
 * sees this call's y only, so the division belongs to the last call).  For graphs whose dense operand X does not fit the L2:
 * the matrix is cut into column panels whose share of X does, the panels are multiplied one after the other and add up in Y
 * -- every row of X is fetched from HBM once per layer instead of once per non-zero (ops.PanelCSR). */
int mmrec_spmm_acc_f32(int64_t n_rows, int64_t n_cols, int d,
                       const int32_t* rowptr, const int32_t* colidx, const float* vals,
                       const int32_t* tasks, int64_t n_tasks, int64_t n_cta_tasks, const int32_t* split_rows,
                       int32_t* counters, float* partial,
                       const float* X, int64_t ldx, float* Y, int64_t ldy,
                       const float* acc_in, float* acc_out, int64_t ldacc, float acc_div, int y_accumulate, void* stream);

/* The SpMMs of one propagation as ONE persistent cooperative launch (src/models/freedom.py:164-178: n_ui_layers products with
 * A_hat, the item-item product, the layer mean and `+ h`): the steps run in order on one resident grid, with a grid-wide
 * barrier before every step whose `sync_before` is set (= it reads what an earlier step wrote).  Every step needs its work
 * plan; d in {32, 64, 128, 256}, 16-byte aligned operands -- otherwise MMREC_EUNSUPPORTED and the caller launches the steps
 * one by one with mmrec_spmm_f32.  Extra epilogue term: acc_out[r,:] += post[r - post_row0,:] for r >= post_row0, applied
 * after the division (FREEDOM's `i_g_embeddings + h`, freedom.py:178). */
typedef struct {
    int64_t n_rows, n_cols;
    const int32_t* rowptr; const int32_t* colidx; const float* vals;
    const int32_t* tasks; int64_t n_tasks, n_cta_tasks; const int32_t* split_rows; int32_t* counters; float* partial;
    const float* X; int64_t ldx;
    float* Y; int64_t ldy;                                   /* nullable */
    const float* acc_in; float* acc_out; int64_t ldacc; float acc_div;   /* as mmrec_spmm_f32 */
    const float* post; int64_t ldpost; int64_t post_row0;    /* nullable */
    int sync_before;
} mmrec_spmm_step;
int mmrec_spmm_chain_f32(int d, int n_steps, const mmrec_spmm_step* steps, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K2  fused gather -> linear(+bias) -> optional row L2-normalise.   Replaces
 * `self.image_trs(self.image_embedding.weight)[items]` (src/models/freedom.py:205-209,
 * bm3.py:102-104, mgcn.py:148-150) and `F.normalize(MLP(features))` (mmgcn.py:165-168).
 *   Y[n,:] = table[idx ? idx[n] : n, :] @ W^T + bias      table [n_table, F], W [d, F], bias [d]|NULL
 *   l2_normalize: Y[n,:] /= max(||Y[n,:]||_2, 1e-12)
 * Two arithmetic paths: tcgen05 3xTF32 with the table split in-kernel (default when ws is given and d <= 256;
 * ws from mmrec_project_workspace_bytes holds the re-tiled weights and the K-split partials) and exact fp32 on CUDA
 * cores (ws == NULL, or mmrec_project_set_path(0) / env MMREC_PROJECT_PATH=simt).
 * ------------------------------------------------------------------------------------------- */
int mmrec_project_set_path(int tensor_core);
size_t mmrec_project_workspace_bytes(int64_t n_out, int64_t F, int d);
int mmrec_project_f32(int64_t n_out, const int64_t* idx, const float* table, int64_t n_table, int64_t F,
                      const float* W, const float* bias, int d, int l2_normalize,
                      float* Y, int64_t ldy, void* ws, size_t ws_bytes, void* stream);

/* measurement aid (tools/probe_mma.py): cycles of `iters` back-to-back tcgen05.mma (cta_group::1, operands in shared memory,
 * K-major no swizzle, M = 128, K = 32 bytes) per CTA, one CTA per SM; kind 0 = tf32, 1 = bf16 */
int mmrec_debug_mma_rate(int kind, int N, int iters, int distinct, long long* cycles_per_cta, void* stream);

/* measurement aid (tools/probe_stream.py): stream a [n_rows, F] fp32 table the way K2 does -- every CTA visits R rows
 * round-robin for burst_bytes contiguous bytes each, R * burst_bytes = 64 KB in flight per CTA -- to see what the DRAM
 * delivers for a given burst length */
int mmrec_debug_stream_probe(const float* table, int64_t n_rows, int64_t F, int R, int burst_bytes, int ctas_per_sm, float* sink,
                             void* stream);

/* ---------------------------------------------------------------------------------------------
 * K3  full-catalog scoring, train-positive mask and per-user top-k.   Replaces
 * `torch.matmul(u_embeddings, restore_item_e.transpose(0, 1))` (src/models/freedom.py:219,
 * bm3.py:153, mgcn.py:262, layergcn.py:185, lightgcn.py:162, mmgcn.py:104) and
 * `scores[mask[0], mask[1]] = -1e10; torch.topk(scores, k)` (src/common/trainer.py:304-309).
 *
 * mmrec_score_f32:    S[b, i] = <Ue[users ? users[b] : b, :], Ie[i, :]>            S [B, ldS]
 *                     ws (mmrec_score_workspace_bytes) holds the re-tiled hi/lo operands of the tensor-core
 *                     path; ws == NULL selects the CUDA-core fp32 path.
 * mmrec_mask_f32:     S[mask_rows[j], mask_cols[j] - item_offset] = -1e10 for columns inside
 *                     [item_offset, item_offset + n_items)
 * mmrec_topk_rows_f32: out_idx/out_val [B, k], descending value, ties -> lower index, index
 *                     reported as column + item_offset.  1 <= k <= 1024, k <= n_items.
 * mmrec_score_topk_f32: all three fused, scores never materialised in HBM (score_cf.cu): the tensor cores compute
 *                     approximate scores (operands rounded to fp16 after a power-of-two scaling: 11 significand bits, as
 *                     tf32) with a proven error bound and only FILTER (per-row certified threshold = the
 *                     (k + masked)-th largest group maximum minus the bound); every candidate that survives is scored
 *                     again in fp32 (fmaf) from the original tables and ranked on that value, ties -> lower index.
 *                     Nothing depends on timing.  ws from mmrec_score_topk_workspace_bytes.
 * mmrec_catalog_pack_f32 / mmrec_score_topk_cat_f32: the same with the item operand prepared once per embedding
 *                     table instead of once per batch (the reference calls full_sort_predict per eval batch of 4096
 *                     users against the same item table, src/common/trainer.py:302-310).  `cat`: 1024-byte aligned
 *                     buffer of mmrec_catalog_bytes(n_items, d) bytes; it must be re-packed whenever Ie changes and
 *                     passed together with the same (n_items, Ie, ldi, d).  cat == NULL packs into ws.
 * mmrec_topk_merge:   merge `parts` sorted lists per user ([parts, B, k] values + indices) into
 *                     one (the per-user top-k reduction across item shards).
 * ------------------------------------------------------------------------------------------- */
/* path of mmrec_score_f32 / mmrec_score_topk_f32 (env MMREC_SCORE_PATH = simt | tc | auto | fused sets the start value):
 *   0 simt   exact fp32 on CUDA cores
 *   1 tc     tcgen05 3xTF32 GEMM into an L2-resident score block, then mask + radix-select top-k kernels
 *   2 auto   (default) fused wherever its shape rules allow (k <= 256, d <= 128, at least 2k item groups of 16..128
 *            items), else tc
 *   3 fused  certified-filter path, no score matrix */
int mmrec_score_set_path(int path);
size_t mmrec_score_workspace_bytes(int64_t B, int64_t n_items, int d);
int mmrec_score_f32(int64_t B, const int64_t* users, const float* Ue, int64_t ldu,
                    int64_t n_items, const float* Ie, int64_t ldi, int d,
                    float* S, int64_t ldS, void* ws, size_t ws_bytes, void* stream);
int mmrec_mask_f32(int64_t mask_nnz, const int64_t* mask_rows, const int64_t* mask_cols,
                   int64_t B, int64_t n_items, int64_t item_offset, float* S, int64_t ldS, void* stream);
int mmrec_topk_rows_f32(int64_t B, int64_t n_items, const float* S, int64_t ldS, int k,
                        int64_t item_offset, int64_t* out_idx, float* out_val, void* stream);
size_t mmrec_score_topk_workspace_bytes(int64_t B, int64_t n_items, int d, int k);
int mmrec_score_topk_f32(int64_t B, const int64_t* users, const float* Ue, int64_t ldu,
                         int64_t n_items, const float* Ie, int64_t ldi, int d,
                         int64_t mask_nnz, const int64_t* mask_rows, const int64_t* mask_cols,
                         int k, int64_t item_offset, int64_t* out_idx, float* out_val,
                         void* ws, size_t ws_bytes, void* stream);
size_t mmrec_catalog_bytes(int64_t n_items, int d);
int mmrec_catalog_pack_f32(int64_t n_items, const float* Ie, int64_t ldi, int d,
                           void* cat, size_t cat_bytes, void* stream);
int mmrec_score_topk_cat_f32(int64_t B, const int64_t* users, const float* Ue, int64_t ldu,
                             int64_t n_items, const float* Ie, int64_t ldi, int d, const void* cat,
                             int64_t mask_nnz, const int64_t* mask_rows, const int64_t* mask_cols,
                             int k, int64_t item_offset, int64_t* out_idx, float* out_val,
                             void* ws, size_t ws_bytes, void* stream);
/* diagnostic, synchronising: rows of the last row block of the last fused call on `ws` that needed the exact kernel
 * (with_cat: the call packed its catalogue into ws, i.e. cat was NULL) */
int64_t mmrec_debug_fused_fallback_rows(const void* ws, int64_t B, int64_t n_items, int d, int k, int64_t mask_nnz, int with_cat);
/* tuning aid: with env MMREC_CF_TIMING set, device time in microseconds of the stages of the last fused call (host array
 * us[cap]; stages: catalogue pack | prep + mask | pass 1 | threshold | pass 2 | finalists | exact rows); returns the count */
int mmrec_debug_cf_timing(float* us, int cap);
int mmrec_topk_merge(int parts, int64_t B, int k, const float* vals, const int64_t* idx,
                     int64_t* out_idx, float* out_val, void* stream);
/* the same merge over lists left where each rank wrote them (peer-mapped memory): vals[p] / idx[p] are host
 * arrays of `parts` (<= 16, parts * k <= 1024) device pointers to [B, k] lists, each sorted (value desc, index asc); an
 * index becomes idx * idx_mul + p * idx_add (round-robin item shards: idx_mul = world, idx_add = 1; must stay below
 * 2^32).  Only rows [row0, row0 + n_rows) are merged, into out_idx / out_val [n_rows, k]: in the sharded evaluation
 * every rank merges its own slice of the batch.  The ranks must be synchronised before the lists are read: either by
 * the caller, or inside the kernel when `flags` / `state` / `rank` are given (see mmrec_peer_exchange_f32). */
int mmrec_topk_merge_peers(int parts, int64_t B, int k, const void* const* vals, const void* const* idx,
                           int64_t idx_mul, int64_t idx_add, int64_t row0, int64_t n_rows,
                           int64_t* out_idx, float* out_val,
                           void* const* flags /* nullable */, int32_t* state /* nullable */, int rank, void* stream);

/* ---------------------------------------------------------------------------------------------
 * K4  user-embedding exchange of the item-sharded propagation, over peer memory (no reference counterpart:
 * the reference is single-GPU, src/utils/configurator.py:114-118; the sum it distributes is the user half of
 * `torch.sparse.mm(adj, ego)`, src/models/freedom.py:172, plus the layer mean of freedom.py:175-176).
 *
 * mmrec_peer_sum_f32 (all ranks read all partials; right for world = 2):
 *   parts   host array of `world` device pointers, rank order: partial user sums R_g E_Ig, [n] floats each,
 *           peer-mapped (CUDA IPC / symmetric memory) -- the caller synchronises the ranks before the call
 *   sum_out = sum over ranks in rank order (identical bits on every rank), may be NULL
 *   acc_out = (acc_in + sum) / acc_div, may be NULL (acc_in NULL = 0; in place allowed)
 *   n % 4 == 0, all pointers 16-byte aligned, world <= 16.
 *
 * mmrec_peer_reduce_push_f32 (reduce-scatter + all-gather in one kernel; what scales to 8 GPUs): rank `rank` owns the
 *   slice [rank * per, (rank + 1) * per) of the n / 4 float4 elements, per = ceil(n / 4 / world).  It sums that slice of
 *   all partials in rank order and stores the result into the same slice of every dst[p] (peer-mapped, its own
 *   included): a non-final layer stores the sum and, if acc_in, acc_out = acc_in + sum; the final layer stores
 *   (acc_in + sum) / acc_div.  acc_in / acc_out hold THIS RANK'S SLICE only (per float4, may alias).  The caller
 *   synchronises the ranks before (partials complete) and after (stores visible) the call.
 *
 * mmrec_peer_gather_f32: dst[p * n_each + i] = src[p][i] -- all-gather of a sharded table by peer loads (the item-id
 *   embeddings the item-item layer of the sharded FREEDOM needs, src/models/freedom.py:166-167).
 */
int mmrec_peer_sum_f32(int64_t n, int world, const void* const* parts, const float* acc_in,
                       float* acc_out, float acc_div, float* sum_out, void* stream);
int mmrec_peer_reduce_push_f32(int64_t n, int world, int rank, const void* const* parts, void* const* dst,
                               const float* acc_in, float* acc_out, float acc_div, int final_layer, void* stream);
int mmrec_peer_gather_f32(int64_t n_each, int world, const void* const* src, float* dst, void* stream);
/* The same with both synchronisations INSIDE the kernel -- one launch per layer: wait until every rank's partial is
 * complete, reduce my slice, store it to every rank, wait until every rank's stores have landed.
 *   flags  host array of `world` device pointers to each rank's flag array (int32[2 * world], peer-mapped, zero at start)
 *   state  this rank's int32[4] in ordinary device memory, zero at start: {calls completed, release word, block counter, -}
 * Calls are numbered on the device (state[0]), so the launch can be replayed from a CUDA graph; every rank must issue
 * the same sequence of calls that use the same flags (mmrec_peer_exchange_f32, mmrec_peer_barrier, mmrec_topk_merge_peers
 * with flags).  A rank that never arrives makes the others trap after 4 s instead of hanging. */
int mmrec_peer_exchange_f32(int64_t n, int world, int rank, const void* const* parts, void* const* dst, void* const* flags,
                            int32_t* state, const float* acc_in, float* acc_out, float acc_div, int final_layer, void* stream);
int mmrec_peer_barrier(int world, int rank, void* const* flags, int32_t* state, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f2  evaluator on the device.   Replaces the host loop that builds the hit matrix from `.cpu().numpy()` of the index
 * matrix and the metric functions (src/utils/topk_evaluator.py:70-102, src/utils/metrics.py:12-105).
 *   topk_idx  int64 [n_users, K]  the trainer's index matrix (K = max(topk) <= 128)
 *   pos_ptr   int64 [n_users + 1], pos_items int64 sorted ascending per user: the ground-truth items
 *   disc      float64 [K] = 1 / log2(j + 2);  idcg_all float64 [K] = cumsum(disc)      (computed by the host, as numpy does)
 *   sums      float64 [4, K], ADDED to: per position j the sum over users of recall / ndcg / precision / map at j + 1
 *             (zero it before the first batch; divide by the number of users afterwards)
 * ------------------------------------------------------------------------------------------- */
int mmrec_topk_metrics_f64(int64_t n_users, int K, const int64_t* topk_idx, const int64_t* pos_ptr, const int64_t* pos_items,
                           const double* disc, const double* idcg_all, double* sums, void* stream);

/* ---------------------------------------------------------------------------------------------
 * f1  feature-table gradient path and optimiser step.   Replaces, per training batch and modality, the backward of
 * `self.image_trs(self.image_embedding.weight)` (src/models/freedom.py:58-62,205-209; bm3.py:102-104; mgcn.py:148-150:
 * the tables are `nn.Embedding.from_pretrained(.., freeze=False)`, i.e. trainable [n_items, F] parameters) -- cuBLAS
 * GEMMs dW = g^T X, dX = g W plus a dense [n_items, F] gradient -- and `torch.optim.Adam.step` over it
 * (src/common/trainer.py:117-118,185-189).  IEEE fp32 on CUDA cores, operation order of torch's `_multi_tensor_adam`.
 *
 * mmrec_index_sum_rows_f32     G[i,:] = sum over j ascending with idx[j] == i of g[j,:]    G [n_rows, ldG], g [n_idx, ldg];
 *                              rows nobody points at become zero; indices outside [0, n_rows) are ignored.
 *                              The gradient of a gathered projection `Linear(table)[idx]` w.r.t. the table is G W.
 * mmrec_linear_wgrad_f32       dW[k,f] = sum_j g[j,k] table[idx ? idx[j] : j, f]   dW [d, F];   db[k] = sum_j g[j,k]  (db nullable)
 *                              ws from mmrec_linear_wgrad_workspace_bytes (per-CTA partials, reduced in a fixed order:
 *                              bit-reproducible).  16-byte accesses when F % 4 == 0 and table / dW / ws are aligned, else 4-byte.
 * mmrec_linear_dgrad_f32       dX = G W   dX [n_rows, F] (leading dimension F), G [n_rows, ldG], W [d, F]; any d (128 k at a
 *                              time), any F.
 * mmrec_linear_dgrad_adam_f32  one Adam step of `param` [n_rows, F] whose gradient is G W, WITHOUT materialising it:
 *                                grad = G W (+ weight_decay * param);  exp_avg += (1 - beta1)(grad - exp_avg);
 *                                exp_avg_sq = beta2 exp_avg_sq + (1 - beta2) grad^2;
 *                                param += step_size * exp_avg / (sqrt(exp_avg_sq) / bc2_sqrt + eps)
 *                              with step_size = -lr / (1 - beta1^t) and bc2_sqrt = sqrt(1 - beta2^t) computed by the caller in
 *                              double, as torch/optim/adam.py does.  W must still hold the values the forward used.
 *                              d <= 128, F % 4 == 0, 16-byte aligned pointers (else MMREC_EUNSUPPORTED / MMREC_EINVAL).
 * mmrec_adam_f32               the same update for `n_tensors` ordinary (param, grad) pairs, one launch per 24 tensors;
 *                              `tensors` is a HOST array.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    float* param; const float* grad; float* exp_avg; float* exp_avg_sq;
    int64_t n;
    double step_size, bc2_sqrt;
} mmrec_adam_tensor;
int mmrec_index_sum_rows_f32(int64_t n_idx, const int64_t* idx, const float* g, int64_t ldg, int d, int64_t n_rows, float* G,
                             int64_t ldG, void* stream);
size_t mmrec_linear_wgrad_workspace_bytes(int64_t n, int64_t F, int d);
int mmrec_linear_wgrad_f32(int64_t n, const int64_t* idx, const float* g, int64_t ldg, int d, const float* table, int64_t n_table,
                           int64_t F, float* dW, float* db, void* ws, size_t ws_bytes, void* stream);
int mmrec_linear_dgrad_f32(int64_t n_rows, const float* G, int64_t ldG, int d, const float* W, int64_t F, float* dX, void* stream);
int mmrec_linear_dgrad_adam_f32(int64_t n_rows, const float* G, int64_t ldG, int d, const float* W, int64_t F, float* param,
                                float* exp_avg, float* exp_avg_sq, double beta1, double beta2, double eps, double weight_decay,
                                double step_size, double bc2_sqrt, void* stream);
int mmrec_adam_f32(int n_tensors, const mmrec_adam_tensor* tensors, double beta1, double beta2, double eps, double weight_decay,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * a5b  MGCN's row-wise fusion (src/models/mgcn.py:153-154 purifier gates, :187-201 two-view attention, preference gates,
 * side + content), inference form.  d in {32, 64, 128}; all matrices row-major with leading dimension d; W are
 * `nn.Linear` weights [d, d] (y = x W^T + b), biases nullable.
 * mmrec_gate_rows_f32   out[n,:] = mul[n,:] * sigmoid(X[n,:] W^T + b)                  (mul nullable: the gate alone)
 * mmrec_mgcn_fuse_f32   a_v = wq2 . tanh(Wq img + bq), a_t likewise on txt; (w0, w1) = softmax(a_v, a_t);
 *                       common = w0 img + w1 txt; sep_v = sigmoid(Wgi content + bgi) (img - common), sep_t likewise;
 *                       side = (sep_v + sep_t + common) / 3 (stored if side != NULL); out = content + side
 * ------------------------------------------------------------------------------------------- */
int mmrec_gate_rows_f32(int64_t n, int d, const float* X, const float* W, const float* b, const float* mul, float* out, void* stream);
int mmrec_mgcn_fuse_f32(int64_t n, int d, const float* img, const float* txt, const float* content, const float* Wq, const float* bq,
                        const float* wq2, const float* Wgi, const float* bgi, const float* Wgt, const float* bgt, float* out,
                        float* side, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMREC_B200_H */
