/*
 * b200rec.h -- C ABI of libb200rec.so: the B200 (sm_100a) hot path of the inductive graph-convolution
 * recommender (LightGCN / IGCN / IMF / MF of umaDosey/Inductive-Recommendation).
 *
 * The reference has no FFI; its replaceable boundary is the DGL / ATen leaf calls made from model.py and
 * trainer.py.  Every entry point below cites the reference lines whose arithmetic it replaces.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch types.  Unless a parameter is marked HOST, every pointer is a
 *     DEVICE pointer into caller-owned memory (the Python host passes tensor.data_ptr()).
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it; nothing allocates,
 *     synchronises or keeps global state, so calls can be captured into a CUDA graph.
 *   - return 0 on success, negative b200rec_status otherwise; b200rec_last_error() gives the text
 *     (thread-local).  There is no CPU fallback: without a CUDA device every compute call fails.
 *   - indices are int32 (all BASELINE shapes have nnz < 2^31), embeddings fp32 row-major [rows, D],
 *     D in {8, 16, 32, 64, 128, 256} (8 and 16 exist for embedding-dimension sharding: D/P columns per GPU;
 *     the scoring entry points need D >= 16).
 */
#ifndef B200REC_H
#define B200REC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  B200REC_OK = 0,
  B200REC_ERR_ARG = -1,      /* bad argument (null pointer, unsupported D, K too large ...) */
  B200REC_ERR_CUDA = -2,     /* a CUDA runtime call or launch failed */
  B200REC_ERR_UNSUPPORTED = -3
} b200rec_status;

const char* b200rec_last_error(void);
int b200rec_version(void);
/* number of kernels this library has launched from the calling process (bench.py "gpu_launches") */
uint64_t b200rec_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * Sparse operand: CSR + a load-balanced work decomposition ("work items").
 *
 * A work item is a slice [item_start, item_end) of the nnz arrays that belongs to one row.  Rows with
 * at most `chunk` entries are one item whose item_dst is the row id; longer rows (hubs of the
 * power-law graph, IGCN's two shared template columns) are split into chunk-sized items whose item_dst
 * is ~slot (negative) -- they write a partial row into `partial[slot]`, and whichever lane group parks the last
 * chunk of a row adds that row's slots in slot order (deterministic) and writes the row.  Items are sorted by length, longest first.
 * The decomposition is built once per graph by b200rec_plan_build().
 *
 * Column-blocked plans (tables larger than L2, e.g. 3M x 128 fp32): the source rows are cut into blocks that fit L2 and
 * the items into one PASS per block -- a pass holds, for every row (or hub piece) with entries in that block, the slice
 * of its entries that falls into it.  CSR columns ascend inside a row, so the slices of a row are consecutive; each
 * continues the row's running fmaf chain (bit 29 of the item code = "reload the running sum", bit 30 = "park it again
 * instead of finishing", the low 29 bits = row id or slot) -- same chain, same bits as the single-pass kernel.
 * b200rec_stream_build re-lays such an operand out pass-major as one 8-byte record stream
 *   [carry-in row, 1.0] [col, val] ... [end of segment: park / finish row]      record.x = type<<30 | hub<<29 | id
 * cut into windows of equal record count that hold whole segments; b200rec_spmm_f32_blocked gives one lane group per
 * window and launches the passes back to back.  The other entry points take single-pass plans only.
 * ------------------------------------------------------------------------------------------------ */
typedef struct {
  int32_t n_rows;            /* output rows */
  int32_t n_cols;            /* rows of the gathered table */
  int32_t nnz;
  const int32_t* rowptr;     /* [n_rows+1] */
  const int32_t* colidx;     /* [nnz] ascending inside a row; with col_hint != 0 bit 31 marks a "hot" source row */
  int32_t col_hint;          /* 0: plain ids.  1 / 2: hot rows are gathered L2::evict_last (kept in the persisting set-aside,
                                b200rec_l2_persist), the others with the default policy (1) or L2::evict_first (2) */
  const float* vals;         /* [nnz] per-edge value, or NULL (all ones) */
  const float* nbr_scale;    /* [n_cols] multiplies each gathered row, or NULL */
  const float* row_scale;    /* [n_rows] multiplies the finished row sum, or NULL */
  const int32_t* eid;        /* [nnz] edge id used to look up keep_bits (transposed operand), or NULL = position */
  /* work decomposition */
  int32_t n_items;
  const int32_t* item_start; /* [n_items] */
  const int32_t* item_end;   /* [n_items] */
  const int32_t* item_dst;   /* [n_items] row id, or ~slot for a piece of a long row (column-blocked plans: | bits 29/30) */
  const int32_t* item_row;   /* [n_items] the row the item belongs to (needed only with dst_flags) */
  int32_t n_long;            /* rows that were split */
  const int32_t* long_row;   /* [n_long] */
  const int32_t* long_slot0; /* [n_long] first slot */
  const int32_t* long_nslot; /* [n_long] */
  int32_t* long_cnt;         /* [n_long] arrival counters, zero-initialised once by the caller (self-resetting) */
  int32_t n_slots;
  const int32_t* slot_long;  /* [n_slots] index (into long_*) of the split row a slot belongs to */
  float* partial;            /* [n_slots, D] scratch for split rows (caller-owned, D = largest D used) */
  int32_t n_passes;          /* 1 (or 0) = single-pass plan; > 1 = column-blocked plan */
  const int32_t* pass_ptr;   /* HOST [n_passes+1]: items [pass_ptr[b], pass_ptr[b+1]) form pass b (NULL for single-pass plans) */
  /* record stream of a column-blocked operand (b200rec_stream_build), what b200rec_spmm_f32_blocked walks */
  const void* records;       /* [n_records] 8-byte records, pass-major */
  const int32_t* win_start;  /* [n_windows+1] record offset of every window */
  const int32_t* pass_win_ptr; /* HOST [n_passes+1]: windows [pass_win_ptr[b], pass_win_ptr[b+1]) form pass b */
  int32_t* win_counter;      /* [n_passes] device scratch: the passes' "next window" counters (zeroed by every call) */
} b200rec_csr;

/* COO -> coalesced int32 CSR on the device (utils.py:42-50 generate_daj_mat: scipy COO->CSR sums duplicates;
 * utils.py:33-39 get_sparse_tensor: coalesced, row-major, columns ascending).  Entry i is (rows[i] + row_offset,
 * cols[i] + col_offset); int64 ids like the reference's train_array.  Outputs (caller-allocated): rowptr [n_rows+1],
 * colidx [n] and mult [n] (capacity n = the upper bound; *nnz_out (HOST) entries are valid; mult = multiplicity of each
 * entry as fp32, may be NULL), first_pos [n] (optional: list position of the first entry merged into each output
 * entry -- a transposed operand uses it as its edge-id map).  *has_dup_out (HOST) = 1 if any pair repeated.
 * Setup-time call: allocates scratch and synchronises `stream`. */
int b200rec_csr_build(const int64_t* rows, const int64_t* cols, int64_t n, int64_t row_offset, int64_t col_offset,
                      int32_t n_rows, int32_t n_cols, int32_t* rowptr, int32_t* colidx, float* mult,
                      int32_t* first_pos, int32_t* nnz_out /*HOST*/, int32_t* has_dup_out /*HOST*/, void* stream);
/* The symmetric bipartite adjacency of the E train pairs (utils.py:42-50): rows/cols [users | items + n_users],
 * both directions, duplicates summed.  colidx / mult capacity 2E. */
int b200rec_adj_build(const int64_t* users, const int64_t* items, int64_t n_pairs, int32_t n_users, int32_t n_items,
                      int32_t* rowptr, int32_t* colidx, float* mult, int32_t* nnz_out /*HOST*/,
                      int32_t* has_dup_out /*HOST*/, void* stream);
/* Build the work decomposition on the device (one-time, per graph and -- for column-blocked plans -- per table width).
 * order_split: rows >= it are scheduled after the others (n_users: each side of the bipartite graph gathers from the
 * other side only, so each phase has the caches to itself); 0 = one phase.
 * col_bounds (HOST [n_blocks+1], ascending from 0 to n_cols) cuts the source rows into n_blocks L2-sized blocks;
 * n_blocks = 1 (col_bounds may be NULL) builds a single-pass plan.
 * row_order: 0 = items of a pass sorted longest first (work items of the single-pass kernel); 1 = rows ascending
 * (input of b200rec_stream_build: carried sums and outputs then stream through memory in order).
 * Call once with item_start == NULL to get sizes[3] (HOST) = {n_items, n_long, n_slots}, allocate, call again.
 * pass_ptr: HOST out [max(n_blocks,1)+1].  Allocates scratch and synchronises `stream`. */
int b200rec_plan_build(const int32_t* rowptr, const int32_t* colidx, int32_t n_rows, int32_t chunk, int32_t order_split,
                       const int32_t* col_bounds /*HOST*/, int32_t n_blocks, int32_t row_order,
                       int32_t* sizes /*HOST out [3]*/,
                       int32_t* item_start, int32_t* item_end, int32_t* item_dst, int32_t* item_row,
                       int32_t* long_row, int32_t* long_slot0, int32_t* long_nslot, int32_t* slot_long,
                       int32_t* pass_ptr /*HOST out*/, void* stream);

/* Carve `bytes` of L2 out for persisting (evict_last) lines on the current device (cudaLimitPersistingL2CacheSize, capped at
 * the device maximum; 0 gives it back).  Without a set-aside the evict_last hints of a col_hint operand do nothing.
 * *granted_out (HOST, optional) = the size in effect.  Device-wide setting; call once at setup. */
int b200rec_l2_persist(int64_t bytes, int64_t* granted_out /*HOST*/);

/* Symmetric-normalised adjacency values (model.py:89-98 LightGCN.generate_graph; utils.py:42-50).
 * deg[r] = max(1, sum of multiplicities in row r); dinv = deg^-1/2 (fp32, correctly rounded);
 * vals[e] = fl(fl(dinv[r] * mult[e]) * dinv[col[e]]).  mult may be NULL (= 1). */
int b200rec_adj_normalize(const int32_t* rowptr, const int32_t* colidx, const float* mult, int32_t n_rows,
                          float* dinv /*out [n_rows]*/, float* vals /*out [nnz]*/, void* stream);

/* ------------------------------------------------------------------------------------------------
 * SpMM with fused epilogue -- replaces dgl.ops.gspmm(g,'mul','sum',X,vals) (model.py:106, :4184, :4196).
 *   s[r]  = row_scale[r] * post_scale * sum_e keep(e) * vals[e] * nbr_scale[col[e]] * X[col[e], :]
 *   y[r]   = s[r]                                   (if y   != NULL)
 *   out[r] = (addend[r] + s[r]) * out_scale         (if out != NULL; addend may be NULL = 0)
 * keep_bits: optional edge-dropout bitmask (bit e of word e>>5; NGCF.dropout_sp_mat model.py:4016-4028),
 * post_scale carries its 1/(1-p).  `out` may alias `addend`.  Row-sum order is CSR order (deterministic).
 * ------------------------------------------------------------------------------------------------ */
int b200rec_spmm_f32(const b200rec_csr* a /*HOST struct of device pointers*/, const float* x, int32_t d,
                     const uint32_t* keep_bits, float post_scale,
                     float* y, const float* addend, float* out, float out_scale, void* stream);
/* Same, with two per-row byte masks that let a caller skip work whose result it does not need / knows to be zero:
 *   dst_flags [n_rows]: only rows with a non-zero flag are computed and written (others are left untouched);
 *   src_flags [n_cols]: gathered rows with a zero flag are all-zero in x and are not read.
 * Results on the computed rows are bit-identical to the unmasked call. */
int b200rec_spmm_f32_ex(const b200rec_csr* a, const float* x, int32_t d, const uint32_t* keep_bits, float post_scale,
                        float* y, const float* addend, float* out, float out_scale,
                        const uint8_t* dst_flags, const uint8_t* src_flags, void* stream);
/* Same (no dropout mask), plus a byte mask of the rows where the fused `out` / `addend` matter to the caller -- a training
 * step reads the layer sum only at the <= 3B sampled rows, and its gradient G (the addend of every backward hop) is zero
 * outside them:
 *   out_mode 1: `out` is written (and `addend` read) only at rows with a non-zero out_rows flag; `y` is written everywhere
 *               (forward layers 1 .. L-1: saves 2 x N x D x 4 bytes per layer);
 *   out_mode 2: `addend` is known to be zero outside the flagged rows and is not read there; `out` is written everywhere
 *               (backward hops: saves N x D x 4 bytes per hop).
 * Bit-identical to b200rec_spmm_f32_ex wherever something is written. */
int b200rec_spmm_f32_sel(const b200rec_csr* a, const float* x, int32_t d, float post_scale, float* y, const float* addend,
                         float* out, float out_scale, const uint8_t* dst_flags, const uint8_t* src_flags,
                         const uint8_t* out_rows, int32_t out_mode, void* stream);

/* Record stream of a column-blocked plan (items sorted by pass, row_order = 1).  window: records per window (>= 32; 128).
 * Call with records == NULL for *n_records_out / *n_windows_out (HOST), allocate records [n_records] x 8 B and win_start
 * [n_windows + 1], call again; pass_win_ptr: HOST out [n_passes+1].  Allocates scratch, synchronises `stream`. */
int b200rec_stream_build(const int32_t* colidx, const float* vals, int32_t n_items, const int32_t* item_start,
                         const int32_t* item_end, const int32_t* item_dst, const int32_t* pass_ptr /*HOST*/,
                         int32_t n_passes, int32_t window, int64_t* n_records_out /*HOST*/, int32_t* n_windows_out /*HOST*/,
                         void* records, int32_t* win_start, int32_t* pass_win_ptr /*HOST out*/, void* stream);
/* The same sum through a column-blocked operand (a->records): one launch per block of source rows, running sums carried
 * in `carry` ([n_rows, ld] fp32 scratch).  Plain valued operand only (no masks / scales).  `ld` is the row stride in
 * floats of x, y, addend, out and carry (>= d): with ld > d the call works on d columns of wider tables -- a D = 128
 * table can be swept as two independent 64-column halves, each with half as many passes.  Bit-identical to
 * b200rec_spmm_f32 on a single-pass plan with the same chunk. */
int b200rec_spmm_f32_blocked(const b200rec_csr* a, const float* x, int32_t d, int32_t ld, float post_scale,
                             float* y, const float* addend, float* out, float out_scale, float* carry, void* stream);

/* L-layer propagation + layer mean (model.py:100-110 LightGCN.get_rep; :4193-4199 IGCN.get_rep):
 *   X_{k+1} = A X_k,  mean_out = (X_0 + ... + X_L) / (L+1).  buf0/buf1: [n_rows, D] scratch (L >= 2 needs both).
 * The running sum lives in mean_out; the last layer writes only the mean (no (L+1) x N x D stack).
 * needed_rows (optional byte mask [n_rows]): the caller will only read these rows of mean_out (a BPR step reads the
 * <= 3B sampled rows, model.py:118-119), so the LAST layer is computed for them only; other rows of mean_out then
 * hold an unfinished sum.  NULL = all rows. */
int b200rec_propagate_fwd(const b200rec_csr* a, const float* x0, int32_t d, int32_t n_layers,
                          float* buf0, float* buf1, float* mean_out, const uint8_t* needed_rows, void* stream);
/* Backward of the above w.r.t. X_0 given G = dLoss/d(mean_out) (the autograd of model.py:100-110):
 *   dX0 = (1/(L+1)) sum_k A^k G, evaluated as H <- G + A H (A is bit-wise symmetric, so the forward CSR serves).
 * nonzero_rows (optional byte mask [n_rows]): rows of G outside it are all-zero (G has <= 3B non-zero rows), so the
 * first hop does not gather them.  NULL = dense G. */
int b200rec_propagate_bwd(const b200rec_csr* a, const float* g, int32_t d, int32_t n_layers,
                          float* buf0, float* buf1, float* dx0_out, const uint8_t* nonzero_rows, void* stream);

/* ------------------------------------------------------------------------------------------------
 * BPR sampling -- the distribution of BasicDataset.__getitem__ (dataset.py:119-131): user uniform over
 * users with a non-empty train row; positive uniform over that row; negative uniform over items,
 * rejected while it is in the (sorted) row.  The reference's MT19937 / NumPy streams cannot be matched
 * on device; the stated stream is Philox4x32-10, key = (seed_lo, seed_hi), counter = (slot, step, draw, purpose),
 * restated bit-exactly in oracle/oracle_c.c.  `step` is read from device memory so a captured graph advances.
 * out_batch: int64 [B,3] = (user, pos, neg), the layout trainer.py:414-415 slices.
 * ------------------------------------------------------------------------------------------------ */
int b200rec_bpr_sample(const int32_t* user_ptr, const int32_t* user_items, int32_t n_users, int32_t n_items,
                       uint64_t seed, const int64_t* step /*device scalar*/, int32_t batch,
                       int64_t* out_batch, void* stream);

/* Row-restricted SpMM through a compacted work list (the last forward layer of a BPR step, which is needed only at the
 * <= 3B sampled rows): b200rec_live_items writes the plan items whose row has a non-zero flag to live_items (int32
 * [n_items], any order) and their number to *live_count (device scalar); b200rec_spmm_f32_live runs exactly those items
 * with a grid sized for max_live (an upper bound of the count known to the caller, e.g. 3B + the number of hub pieces).
 * Rows outside the list are left untouched; the rows computed are bit-identical to b200rec_spmm_f32. */
int b200rec_live_items(const b200rec_csr* a, const uint8_t* row_flags, int32_t* live_items, int32_t* live_count, void* stream);
int b200rec_spmm_f32_live(const b200rec_csr* a, const float* x, int32_t d, float post_scale, float* y, const float* addend,
                          float* out, float out_scale, const int32_t* live_items, const int32_t* live_count,
                          int32_t max_live, void* stream);

/* Byte mask of the table rows a batch touches: flags[user] = flags[item_offset+pos] = flags[item_offset+neg] = 1.
 * flags [n_rows] must be zeroed by the caller first.  Feeds needed_rows / nonzero_rows of the propagation calls. */
int b200rec_mark_rows(const int64_t* batch /*[B,3]*/, int32_t n_batch, int64_t item_offset, uint8_t* flags, void* stream);

/* Byte mask of the rows WITHIN ONE HOP of a batch (new; no reference counterpart -- the reference propagates every row in
 * every layer, model.py:100-110): the sampled rows and every train item of a sampled user (user_ptr / user_items = the
 * sampler's CSR by user).  Only flags[user], flags[item_offset + item] entries are written (set to 1): the caller zeroes
 * the ITEM half before the call and keeps the USER half all-ones (a superset is always valid; the users within a hop of
 * the sampled items are ~98 % of the users of the C4 graph).  A BPR step reads the propagated table only at the sampled
 * rows, so layer L-1 of the forward pass is needed only inside this mask (dst_flags) and after the first backward hop the
 * gradient is non-zero only inside it (dst_flags of hop 1, src_flags of hop 2); rows computed are bit-identical. */
int b200rec_mark_reach(const int64_t* batch /*[B,3]*/, int32_t n_batch, int64_t item_offset, const int32_t* user_ptr,
                       const int32_t* user_items, uint8_t* flags, void* stream);

/* Row gather / scatter-add used by the autograd-compatible bpr_forward (model.py:118-119 rep[idx,:] and its
 * index_put_(accumulate) backward).  idx int64 [n]; offset is added to every index (n_users for items). */
int b200rec_gather_rows(const float* table, int32_t d, const int64_t* idx, int64_t offset, int32_t n,
                        float* out /*[n,D]*/, float* sqnorm /*[n] or NULL*/, void* stream);
int b200rec_scatter_add_rows(float* table, int32_t d, const int64_t* idx, int64_t offset, int32_t n,
                             const float* src /*[n,D]*/, void* stream);

/* Fused BPR forward + backward (trainer.py:414-424 with model.py:112-120 / :4046-4052 / :66-71):
 *   pos = <u,p>, neg = <u,n>;  *loss_out += loss_scale * (mean softplus(neg-pos) + [reg_mode 1] l2_reg * mean l2);
 *   the gradient w.r.t. the gathered rows is scatter-added into g_rep (pre-zeroed [rows, D]) with
 *   red.global.add.v4.f32 (one 128-bit reduction per lane).
 * reg_mode 0: no L2 here (LightGCN regularises layer-0 rows: b200rec_bpr_l2_emb0);
 *          1: L2 on the gathered `rep` rows (NGCF.bpr_forward used by IGCN/IMF, and MF), gradient into g_rep.
 * w (optional, [D]): elementwise score weight of the IGCN auxiliary loss (trainer.py:542-549); g_w [D] is
 *   OVERWRITTEN with dLoss/dw (deterministic block-ordered sum).
 * loss_scale multiplies the loss term and all gradients (aux_reg for the auxiliary term, else 1).
 * item_offset: added to pos / neg ids (n_users for the joint table, template-user count for the aux loss).
 * block_scratch: b200rec_bpr_scratch_floats(B, D) floats, zero-initialised once by the caller (the ticket
 *   counter in it resets itself); the loss sum is taken in block order, so it is run-to-run deterministic. */
int64_t b200rec_bpr_scratch_floats(int32_t n_batch, int32_t d);
int b200rec_bpr_fwd_bwd(const float* rep, int32_t d, const int64_t* batch /*[B,3]*/, int32_t n_batch,
                        int64_t item_offset, float l2_reg, int32_t reg_mode, const float* w,
                        float loss_scale, float* g_rep, float* g_w, float* loss_out,
                        float* block_scratch, void* stream);
/* The same step when the embedding DIMENSION is sharded over P GPUs (each rank holds D/P columns of every table:
 * propagation, Adam and all gradients are column-separable, only the two dot products couple the ranks):
 *   phase 1: dots[B,3] <- this rank's partial (pos, neg, sum of squares) per sample;
 *   caller : all-reduce(dots, SUM) over the ranks (48 KB at B = 2048);
 *   phase 2: gradients for this rank's columns from the reduced dots; *loss_out += loss_weight * (loss terms), so a
 *            caller that gives rank 0 weight 1 and the others 0 obtains the global loss by summing the ranks. */
int b200rec_bpr_fwd_bwd_sharded(const float* rep, int32_t d, const int64_t* batch, int32_t n_batch,
                                int64_t item_offset, float l2_reg, int32_t reg_mode, const float* w,
                                float loss_scale, float* g_rep, float* g_w, float* loss_out, float* block_scratch,
                                float* dots /*[B,3]*/, int32_t phase, float loss_weight, void* stream);
/* Ordered (deterministic, aggregated) form of the scatter.  Batches repeat ids (popular items), and one atomic per sample
 * sums a row's contributions in arrival order.  b200rec_bpr_group_rows ranks the 3B (row, slot) pairs of a batch once
 * (slot = 3 * sample + {0 user, 1 pos, 2 neg}; 3B <= 12288, rows < 2^31): order int32 [3B] = the slots in (row, slot)
 * order, bit 31 set on every position that continues the row of the position before it (so a clear bit 31 marks the
 * first slot of a distinct row).
 * b200rec_bpr_fwd_bwd_ordered is b200rec_bpr_fwd_bwd (phase 0) / _sharded (phase 1, 2) whose gradient rows are formed by
 * ONE lane group per distinct row, adding that row's contributions in slot order and storing the row once
 * (accumulate = 0: overwrite, g_rep must be zero elsewhere; 1: add to the row) -- no atomics, run-to-run identical bits.
 * coef: float [B] scratch.  emb0 (optional) folds LightGCN's layer-0 regulariser into the same launch's loss:
 * *loss_out += l2_emb0 * mean(||e_u||^2 + ||e_p||^2 + ||e_n||^2) over the layer-0 rows (model.py:114-117).
 * b200rec_bpr_l2_emb0_ordered adds that term's gradient to g_emb0, one lane group per distinct row (loss_out NULL when
 * the loss was folded in as above; clear_table, optional: the same rows of that [rows, D] table are zeroed in the same
 * launch -- G after the backward chain has consumed it).  b200rec_clear_rows zeroes the rows a batch touches on its own
 * (so g_rep can stay all-zero between steps without a full memset). */
int b200rec_bpr_group_rows(const int64_t* batch, int32_t n_batch, int64_t item_offset, int32_t* order, void* stream);
int b200rec_bpr_fwd_bwd_ordered(const float* rep, int32_t d, const int64_t* batch, int32_t n_batch,
                                int64_t item_offset, float l2_reg, int32_t reg_mode, const float* w,
                                float loss_scale, float* g_rep, float* g_w, float* loss_out, float* block_scratch,
                                float* dots, int32_t phase, float loss_weight, float* coef, const int32_t* order,
                                int32_t accumulate, const float* emb0, float l2_emb0, void* stream);
int b200rec_bpr_l2_emb0_ordered(const float* emb0, int32_t d, const int64_t* batch, int32_t n_batch,
                                int64_t item_offset, float l2_reg, float* g_emb0, float* loss_out,
                                float* block_scratch, const int32_t* order, float* clear_table, void* stream);
int b200rec_clear_rows(const int64_t* batch, int32_t n_batch, int64_t item_offset, float* table, int32_t d, void* stream);
/* LightGCN's layer-0 regulariser (model.py:114-117): g_emb0 += l2_reg * d(mean ||e_u||^2+||e_p||^2+||e_n||^2)/de,
 * *loss_out += l2_reg * mean(l2). */
int b200rec_bpr_l2_emb0(const float* emb0, int32_t d, const int64_t* batch, int32_t n_batch, int64_t item_offset,
                        float l2_reg, float* g_emb0, float* loss_out, float* block_scratch, void* stream);

/* Dense Adam, the arithmetic of torch.optim.Adam defaults (trainer.py:44-46): no weight decay / amsgrad.
 * step (device int64 scalar) is the count BEFORE this update; the kernel uses step+1 for the bias corrections.
 * Hyper-parameters are doubles: torch forms 1-beta, lr/bias_correction1 and sqrt(bias_correction2) in double before
 * the fp32 kernels see them, and so does this one. */
int b200rec_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                      double lr, double beta1, double beta2, double eps, const int64_t* step, void* stream);
/* End of an optimiser step: *step += 1, *step_b += 1 (each if non-NULL; e.g. the Adam count and the sampler count),
 * and AverageMeter.update(loss, B) (utils.py:286-289) on device: loss_accum[0] += *loss * n_batch,
 * loss_accum[1] += n_batch (if both given). */
int b200rec_step_advance(int64_t* step, int64_t* step_b, const float* loss, double* loss_accum /*[2]*/,
                         int32_t n_batch, void* stream);

/* ---- contrastive views (SGL model.py:130-243, HALF :246-365) ----
 * InfoNCE of the `info_nce` package as the reference calls it (cal_loss, model.py:206-214 / :322-332): 'unpaired'
 * negatives with negative_keys == positive_key, reduction 'mean':
 *   qn = q/max(|q|,1e-12), kn likewise; logits_i = [qn_i.kn_i, qn_i.kn_0 .. qn_i.kn_{n-1}] / temperature; CE against index 0.
 * Forward and backward in one call; the n x n logits are never stored.
 *   rows == NULL: q, k are [n,d]; gq, gk [n,d] are OVERWRITTEN with loss_scale * dLoss/dq, dLoss/dk.
 *   rows != NULL: q, k are tables [*,d]; sample i is row rows[i*row_stride] of both (the users of a [B,3] BPR batch:
 *                 rows = batch, row_stride = 3); gq, gk are tables and the gradients are ADDED at those rows.
 * loss_out[0] += loss_scale * mean_i CE_i.  workspace: b200rec_infonce_workspace_floats(n, d) floats. */
int64_t b200rec_infonce_workspace_floats(int32_t n, int32_t d);
int b200rec_infonce_fwd_bwd(const float* q, const float* k, const int64_t* rows, int32_t row_stride, int32_t n, int32_t d,
                            float temperature, float loss_scale, float* loss_out, float* gq, float* gk, float* workspace,
                            void* stream);

/* Edge-dropout keep mask of NGCF.dropout_sp_mat (model.py:4016-4021) drawn on device: bit e = floor((1-p) + r[e]),
 * r uniform [0,1) at 24-bit resolution from Philox4x32-10 (key = seed, counter = (e>>2, step_lo, step_hi, 0xD0),
 * lane e&3); restated in oracle/oracle_c.c.  keep_bits: uint32 [ceil(nnz/32)], edge order = CSR position. */
int b200rec_dropout_mask(int32_t nnz, float p, uint64_t seed, const int64_t* step /*device scalar*/,
                         uint32_t* keep_bits, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Full-rank evaluation (model.py:122-127 predict; trainer.py:146-170 eval).
 * ------------------------------------------------------------------------------------------------ */
/* Dense scores[b, n_items] = U[b,:] . I[j,:]  in fp32, accumulated with fmaf in d = 0..D-1 order (the order the
 * oracle restates), for API parity with predict().  users: int64 [b] row ids into rep_users. */
int b200rec_score_dense_f32(const float* rep_users, const int64_t* users, int32_t n_batch_users,
                            const float* rep_items, int32_t n_items, int32_t d, float* scores, void* stream);
/* Fused score + mask + top-K (never materialises the score matrix):
 *   for each listed user: score all items, drop items in its exclusion rows (train, and val when testing:
 *   excl_ptr_a/excl_idx_a and optional excl_ptr_b/excl_idx_b, sorted CSR by user, trainer.py:155-165) and items in
 *   [banned_lo, banned_hi) (trainer.py:166-167), return the K best as (score desc, item id asc).
 * precision 0: exact fp32 CUDA-core path;  1 (D = 64 / 128): TMA-fed tcgen05 bf16 candidate pass + exact fp32 re-score
 *   (same ids and scores as precision 0 by construction: the candidate margin bounds the bf16 error).
 * out_overflow [b] (required for precision 1, optional otherwise): 1 for a user whose candidate list overflowed --
 *   its output row is not valid and must be recomputed with precision 0; always 0 for precision 0.
 * workspace: device scratch, size from b200rec_score_topk_workspace(). */
int64_t b200rec_score_topk_workspace(int32_t n_batch_users, int32_t n_items, int32_t d, int32_t k, int32_t precision);
int b200rec_score_topk(const float* rep_users, const int64_t* users, int32_t n_batch_users,
                       const float* rep_items, int32_t n_items, int32_t d,
                       const int32_t* excl_ptr_a, const int32_t* excl_idx_a,
                       const int32_t* excl_ptr_b, const int32_t* excl_idx_b,
                       int32_t banned_lo, int32_t banned_hi, int32_t k, int32_t precision,
                       int32_t* out_ids /*[b,K]*/, float* out_scores /*[b,K]*/, int32_t* out_overflow /*[b]*/,
                       void* workspace, void* stream);
/* Precision / Recall / NDCG at every cut-off of `topks` (device int32 [n_topks], ascending, each <= k, n_topks <= 32) in one
 * pass over the recommended ids (calculate_metrics, trainer.py:115-144): hit = rec[u,j] in eval row u, gains 1/log2(j+2),
 * ideal gains over min(len, k), users without eval items excluded.  block_sums: double [ceil(n_rows/128)][3*n_topks+1],
 * per block the sums of (precision, recall, ndcg) per cut-off and the number of counted users; the caller adds the blocks
 * and divides.  fp64, fixed reduction order. */
int b200rec_rank_metrics(const int32_t* rec_ids, int32_t n_rows, int32_t k, int64_t user0, const int32_t* eval_ptr,
                         const int32_t* eval_idx, const int32_t* topks, int32_t n_topks, double* block_sums, void* stream);
/* hit[u,j] = rec[u,j] in eval row u (trainer.py:117-121), eval rows sorted CSR. */
int b200rec_hit_matrix(const int32_t* rec_ids, int32_t n_rows, int32_t k, int64_t user0,
                       const int32_t* eval_ptr, const int32_t* eval_idx, float* hit, void* stream);

/* Global top-m of n candidate values (the flattened-matrix torch.topk of DOSE's similarity mining, model.py:503-560):
 * out_idx int32 [m] = positions of the m largest values, largest first (ties: lower position first), out_vals [m]
 * optional.  Setup-time call (once per epoch): allocates scratch, synchronises `stream`. */
int b200rec_topk_global(const float* vals, int64_t n, int32_t m, int32_t* out_idx, float* out_vals, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Peer memory for the row-partitioned multi-GPU propagation (new: the reference has no multi-device code).
 * b200rec_peer_alloc: cudaMalloc + zero + CUDA IPC handle (HOST, 64 bytes) that the other ranks of the node open with
 * b200rec_peer_open.  b200rec_spmm_f32_peer is b200rec_spmm_f32_ex whose epilogue also stores every finished row of
 * y / out into the same offset of up to 8 peer copies (an all-gather by push, fused into the SpMM).
 * b200rec_peer_signal / _wait: per-layer hand-shake -- flags[my_rank] on every peer <- *epoch + epoch_add, then spin
 * (bounded) until all flags of this rank reach that value, then *epoch += epoch_add.  `peer_flags` is a DEVICE array of
 * n_peers pointers (one flag array per rank, this rank's own included).
 * ------------------------------------------------------------------------------------------------ */
int b200rec_spmm_f32_peer(const b200rec_csr* a, const float* x, int32_t d, const uint32_t* keep_bits, float post_scale,
                          float* y, const float* addend, float* out, float out_scale,
                          const uint8_t* dst_flags, const uint8_t* src_flags, int32_t n_peers,
                          float* const* peer_y /*HOST [n_peers] device pointers, or NULL*/,
                          float* const* peer_out /*HOST [n_peers], or NULL*/,
                          const uint8_t* out_rows /*or NULL*/, int32_t out_mode /*as b200rec_spmm_f32_sel*/, void* stream);
/* b200rec_adam_step on a row block whose updated parameters are also stored into the peers' tables. */
int b200rec_adam_step_peer(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           double lr, double beta1, double beta2, double eps, const int64_t* step,
                           int32_t n_peers, float* const* peer_param /*HOST [n_peers]*/, void* stream);
int b200rec_peer_alloc(int64_t bytes, void** ptr_out /*HOST out*/, uint8_t* handle_out /*HOST [64]*/);
int b200rec_peer_open(const uint8_t* handle /*HOST [64]*/, void** ptr_out /*HOST out*/);
int b200rec_peer_close(void* ptr);
int b200rec_peer_free(void* ptr);
int b200rec_peer_signal(void* const* peer_flags, int32_t n_peers, int32_t my_rank, const int64_t* epoch, int64_t epoch_add,
                        void* stream);
int b200rec_peer_wait(const void* my_flags, int32_t n_peers, int64_t* epoch /*advanced by epoch_add on return*/,
                      int64_t epoch_add, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B200REC_H */
