/*
 * libqst -- C ABI of the B200 (sm_100a) retrieval-scoring + quadruplet-loss hot path.
 *
 * The reference (lucastrefezza/quadruplet-sentence-transformer) is pure Python and has no
 * FFI of its own; the boundary it exposes for this path is two Python protocols
 * (SURVEY.md section 8b).  This header is the native boundary underneath the drop-in Python
 * classes: every entry point cites the reference call it replaces (paths relative to
 * /root/reference).  INTEGRATION.md shows the reference-side ctypes stub.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch/C++ types.
 *   - every function returns 0 on success, <0 on error; qst_last_error() gives the message
 *     (thread-local).  No exception crosses the boundary.
 *   - ALL data pointers are DEVICE pointers owned by the caller (inputs, outputs, workspace).
 *     Workspace sizes come from the matching *_workspace_bytes() call.
 *   - every launch is asynchronous on the given stream (a cudaStream_t passed as void*).
 *   - no hidden global state, no hidden allocation, no CPU fallback.
 */
#ifndef QST_H_
#define QST_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QST_VERSION 100

typedef void* qst_stream_t; /* cudaStream_t */

enum qst_dtype { QST_F32 = 0, QST_F16 = 1, QST_BF16 = 2 };
enum qst_reduction { QST_RED_NONE = 0, QST_RED_SUM = 1, QST_RED_MEAN = 2 };
/* score functions of ir_evauation_script.py:70 */
enum qst_score { QST_SCORE_COS = 0, QST_SCORE_DOT = 1, QST_SCORE_EUCLID = 2 };

enum qst_status {
  QST_OK = 0,
  QST_ERR_INVALID = -1,  /* bad argument (the Python layer raises ValueError before this) */
  QST_ERR_CUDA = -2,     /* a CUDA runtime/driver call failed */
  QST_ERR_WORKSPACE = -3,/* workspace too small */
  QST_ERR_UNSUPPORTED = -4
};

int qst_version(void);
const char* qst_last_error(void);
/* Number of kernels this library has launched so far in this process (all entry points, all
 * threads): the difference around a region is the number of OUR kernels that ran in it. */
long long qst_launch_count(void);
/* SM count / compute capability of the current device. */
int qst_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------
 * K5  gamma-quadruplet loss.   Replaces models/losses/losses.py:9-69 (gamma_quadruplet_loss:
 * three F.triplet_margin_loss calls :35-61 + reductions :64-69) and its autograd backward.
 *
 *   d(u,v) = || u - v + eps ||_p          (eps added to the difference, torch pairwise_distance)
 *   A = max(0, m_pos_neg  + d(a,pos)  - dn(a,pos ,neg ))
 *   B = max(0, m_part_neg + d(a,part) - dn(a,part,neg ))
 *   C = max(0, m_pos_part + d(a,pos)  - dn(a,pos ,part))
 *   dn(a,x,y) = d(a,y), or min(d(a,y), d(x,y)) when swap
 *   row loss = A + gamma*B + (1-gamma)*C
 *
 * Inputs are row-major [B, D] of `dtype`; all arithmetic is fp32; loss outputs are fp32;
 * gradients are written in `dtype`.
 * ------------------------------------------------------------------------------------------ */
typedef struct qst_quad_params {
  float gamma;
  float one_minus_gamma; /* float(1.0 - gamma) evaluated in double, as Python does at losses.py:65 */
  float margin_pos_neg;
  float margin_pos_part;
  float margin_part_neg;
  float p;    /* > 0; INFINITY allowed */
  float eps;  /* 1e-6 in the reference (torch default) */
  int32_t swap;
} qst_quad_params;

#define QST_QUAD_SAVED_PER_ROW 8 /* 6 distances + 2 pad floats saved by fwd for bwd */

/* Scratch for the cross-row reduction ('mean' / 'sum'): must be zero-filled ONCE when allocated; the
 * kernels leave it ready for the next launch.  One workspace serves one stream at a time.  Size does
 * not depend on B.  The result is bitwise reproducible: per-CTA partial sums are combined either in a
 * fixed order or (fwd_bwd on rows of <= 1024 fp32 / 2048 half elements) as fixed-point integers carried
 * in atomics, whose sum does not depend on arrival order.  A NaN or inf in the inputs makes the loss
 * NaN / inf as torch's clamp_min and float addition do. */
size_t qst_quadruplet_workspace_bytes(void);

/* Forward.  loss_out: [B] (QST_RED_NONE) or [1].  saved: [B, QST_QUAD_SAVED_PER_ROW] fp32 or
 * NULL (no backward wanted, e.g. under torch.no_grad(): models/evaluators.py:82). */
int qst_quadruplet_fwd(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                       int dtype, int64_t B, int64_t D, const qst_quad_params* prm, int reduction,
                       float* loss_out, float* saved, void* workspace, qst_stream_t stream);

/* Backward from the distances saved by fwd.  grad_out: fp32 [B] when reduction==NONE else [1]
 * (device pointer: the upstream gradient of the returned loss).  Any grad pointer may be NULL. */
int qst_quadruplet_bwd(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                       int dtype, int64_t B, int64_t D, const qst_quad_params* prm, int reduction,
                       const float* saved, const float* grad_out,
                       void* g_anchor, void* g_pos, void* g_part, void* g_neg, qst_stream_t stream);

/* One-launch forward+backward for the training step (loss.backward() with upstream 1, scaled by
 * `upstream`): reads each input once from HBM, writes each gradient once.  8*B*D*sizeof(dtype)
 * algorithmic bytes. */
int qst_quadruplet_fwd_bwd(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                           int dtype, int64_t B, int64_t D, const qst_quad_params* prm, int reduction,
                           float upstream, float* loss_out,
                           void* g_anchor, void* g_pos, void* g_part, void* g_neg,
                           void* workspace, qst_stream_t stream);

/* Paired distances of the QuadrupletEvaluator (models/evaluators.py:130-389: three ST 2.2.2
 * TripletEvaluators pos/part, pos/neg, part/neg with sklearn paired cosine / manhattan / euclidean
 * distances).  One pass over the four [B, D] matrices.
 *   out_dist   [B, 9] fp32 or NULL: (cosine, manhattan, euclidean) x d(anchor, {pos, part, neg})
 *   out_counts [9] uint64: (cosine, manhattan, euclidean) x #{pos<part, pos<neg, part<neg}
 *              (zeroed by the call) */
int qst_quadruplet_eval(const void* x_anchor, const void* x_pos, const void* x_part, const void* x_neg,
                        int dtype, int64_t B, int64_t D, float* out_dist, unsigned long long* out_counts,
                        qst_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K1  row preparation.  Replaces the F.normalize(p=2, dim=1, eps=1e-12) half of
 * sentence_transformers.util.cos_sim (called at ir_evauation_script.py:70,
 * models/evaluators.py:545) and produces the bf16 operand of the tensor-core pass.
 *
 *   x: [n, d] of `dtype`, row stride d.
 *   out_bf16: [n, d_pad] bf16, d_pad = qst_padded_dim_for(d, mode) (zero padded), see qst_prep_mode.
 *             euclidean_score (models/evaluators.py:392-405) ranks like 2q.c - ||c||^2, which the
 *             two euclid modes turn into a plain dot product of augmented rows.
 *   out_inv_norm[n]: 1/max(||x||_2, 1e-12)            (NULL to skip)
 *   out_sq_norm[n] : ||x||_2^2                        (NULL to skip; euclid score)
 *   out_err[n]     : || bf16(row) - row_used ||_2     (NULL to skip; rounding residual used by
 *                    the exactness certificate), row_used = normalised or raw row
 *   stats (2 floats, device, caller zero-fills before the first call over a corpus):
 *                    stats[0] = max over rows of out_err, stats[1] = max over rows of ||row_used||
 *                    (NULL to skip)
 * ------------------------------------------------------------------------------------------ */
/* prep modes: how a row becomes a tensor-core operand */
enum qst_prep_mode {
  QST_PREP_RAW = 0,           /* dot_score: bf16(x) */
  QST_PREP_COS = 1,           /* cos_sim: bf16(x / max(||x||, 1e-12)) */
  QST_PREP_EUCLID_CORPUS = 2, /* euclid_score corpus: bf16(x) and ||x||^2 split exactly over 3 extra columns */
  QST_PREP_EUCLID_QUERY = 3   /* euclid_score query: bf16(2x) and -1 in those columns: dot = 2q.c - ||c||^2 */
};
int64_t qst_padded_dim(int64_t d);
int64_t qst_padded_dim_for(int64_t d, int mode);
int qst_prep_rows(const void* x, int dtype, int64_t n, int64_t d, int mode, void* out_bf16,
                  float* out_inv_norm, float* out_sq_norm, float* out_err, float* stats,
                  qst_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K2+K3  score + top-k.  Replaces, for one corpus shard, the chunk loop of
 * InformationRetrievalEvaluator.compute_metrices [sentence-transformers 2.2.2]:
 *   pair_scores = score_function(query_embeddings, sub_corpus_embeddings)      -> K2 (tcgen05)
 *   torch.topk(pair_scores, min(max_k, chunk), dim=1, largest=True, sorted=False)
 *   ... sorted(hits, key=score, reverse=True)                                  -> K3
 * constructed at ir_evauation_script.py:107-123 / models/evaluators.py:572-588.
 *
 * K2: bf16 tensor-core scores  S = Qb * Cb^T  tile by tile (TMA -> smem -> tcgen05.mma -> TMEM);
 *     the epilogue keeps, per query row, only scores above a running threshold (k'-th best seen)
 *     in a small candidate buffer -- the [Q, N] score matrix never reaches HBM.
 * K3: per query, selects the k' best candidates, rescoring them exactly in fp32 from the fp32
 *     masters, sorts descending (ties: lower corpus index first) and emits the top k with a
 *     certificate:  margin = (k-th exact score) - (k'-th bf16 score) - eps_q, where eps_q is a
 *     rigorous bound on |bf16 score - exact score|.  margin > 0 proves no document outside the
 *     candidate set can belong to the exact top k.
 * ------------------------------------------------------------------------------------------ */
typedef struct qst_topk_plan {
  int64_t Q, N, D, D_pad;
  int32_t k, kprime;             /* kprime = candidates rescored exactly per query */
  int32_t kunit, cap;            /* entries one work unit keeps per row; its buffer capacity */
  int32_t m_tiles, n_tiles, stripes, tiles_per_stripe, units;
  int32_t grid, score;           /* grid = CTA groups launched (x ctas CTAs each) */
  int32_t ctas, rows_per_unit;   /* CTAs per tile (2 = cta_group::2 pairs), query rows per work unit */
  size_t ws_bytes;
  size_t off_thr, off_cnt, off_uthr, off_cand; /* layout inside the workspace */
  int32_t qs;                    /* 1: query-stationary kernel (query block resident in TMEM; CTA pairs, D_pad <= 768) */
  int32_t reserved_;
} qst_topk_plan;

/* Fills `plan` for (Q queries, N corpus rows, D dims, top k).  kprime <= 0 picks the default
 * head-room.  sm_count <= 0 queries the current device.  Tile shape: CTA pairs (256 query rows per
 * work unit) by default, single-CTA tiles (128 rows) for batches of at most 128 queries and for
 * batches of at most 2048 whose last 256-row block would be at most half full; the environment
 * variable QST_SCORE_CTAS=1|2 forces one. */
int qst_topk_plan_make(int64_t Q, int64_t N, int64_t D, int k, int kprime, int score, int sm_count,
                       qst_topk_plan* plan);

/* Overrides the number of entries a work unit keeps per row (and the derived buffer capacity and
 * workspace layout).  Used by the corpus-sharded path, where the relevant count is k' of the WHOLE
 * corpus spread over the units of ALL shards. */
int qst_topk_plan_set_kunit(qst_topk_plan* plan, int kunit);

/* K2.  q_bf16 [Q, D_pad], c_bf16 [N, D_pad] from qst_prep_rows.  Fills plan->ws. */
int qst_score_select(const qst_topk_plan* plan, const void* q_bf16, const void* c_bf16,
                     void* workspace, qst_stream_t stream);

/* Corpus-sharded variant of K2: every rank scores the SAME queries against its
 * own shard; the per-row thresholds are shared between ranks THROUGH PEER MEMORY while the kernels
 * run -- a rank that establishes a better threshold for a query pushes it with a remote atomicMax
 * (NVLink/NVSwitch) into the hint array of every peer, so all shards filter against the best
 * threshold found anywhere.  hint_local / peer_hints[i] are [plan->m_tiles * plan->rows_per_unit]
 * uint32 arrays in buffers from qst_peer_buffer_create / _open; the caller zeroes hint_local before
 * the launch (and keeps two generations so a fast rank never pushes into a buffer being cleared).
 * Purely opportunistic: no rank ever waits on another inside the kernel. */
#define QST_MAX_PEERS 15
#define QST_MAX_WORLD (QST_MAX_PEERS + 1)
#define QST_IPC_HANDLE_BYTES 64
int qst_score_select_peers(const qst_topk_plan* plan, const void* q_bf16, const void* c_bf16, void* workspace,
                           void* hint_local, void* const* peer_hints, int n_peers, qst_stream_t stream);
/* Device buffers other processes on the node can map (CUDA IPC).  create: cudaMalloc + zero fill +
 * export handle (64 bytes, to be exchanged by the caller, e.g. with an all-gather); open: map a
 * peer's buffer into this process; close / destroy undo them. */
int qst_peer_buffer_create(size_t bytes, void** dev_ptr, unsigned char* handle64);
int qst_peer_buffer_open(const unsigned char* handle64, void** dev_ptr);
int qst_peer_buffer_clear(void* dev_ptr, size_t offset, size_t bytes, qst_stream_t stream);
/* Copy-engine transfer into / out of a peer-mapped buffer (cudaMemcpyAsync, device to device): uses no SM,
 * so it proceeds underneath the persistent K2.  The sharded path pushes the fp32 queries of a rank into
 * every peer's gather buffer this way while K2 runs. */
int qst_peer_copy(void* dst, const void* src, size_t bytes, qst_stream_t stream);
int qst_peer_buffer_close(void* peer_ptr);
int qst_peer_buffer_destroy(void* dev_ptr);

/* Dense exact score matrix out[Q, N] fp32 with K3's arithmetic (fp32 dot products on CUDA cores): what
 * sentence_transformers.util.cos_sim / dot_score and the reference's euclidean_score
 * (models/evaluators.py:392-405) return when called directly, e.g. at
 * dataset/positive_examples_selection.py:55 and dataset/quadruplet_dataset.py:229-234.  For small inputs;
 * retrieval-sized inputs go through qst_score_select / qst_finalize_topk and never build the matrix.
 * q_inv / c_inv: inverse norms from qst_prep_rows (cos_sim) or NULL.  Q <= 65535 per call. */
int qst_dense_scores(int64_t Q, int64_t N, int64_t D, int score, const float* q_f32, const float* q_inv,
                     const float* c_f32, const float* c_inv, float* out, qst_stream_t stream);

/* Debug/validation aid: raw tensor-core scores of one call written densely, out[Q, N] fp32.
 * Same kernel, same tiles, epilogue stores instead of selecting.  Small shapes only. */
int qst_score_dense(const void* q_bf16, int64_t Q, const void* c_bf16, int64_t N, int64_t D_pad,
                    float* out, qst_stream_t stream);

/* Debug aid: per-tile trace of the last qst_score_select run with QST_SCORE_DEBUG bit 32 set
 * (time stamps in ns, candidate count and threshold of each CTA's first row; 64 slots per CTA,
 * slot 0 = start of the CTA's first unit, slot 63 = its end).  Host pointers, n <= 160*64. */
int qst_debug_read_trace(long long* ns, int* cnt, float* thr, int n);

/* K3.  Exact fp32 rescoring of the selected candidates and final ordering.
 *   q_f32 [Q, D], c_f32 [N, D]: fp32 masters; q_inv/c_inv: inverse norms (cos) or NULL (dot).
 *   q_err [Q] and c_stats (2 floats) from qst_prep_rows feed the certificate; either may be NULL
 *   (then margin is reported without the eps term).
 *   idx_offset is added to every emitted corpus index (global id of the shard's row 0).
 *   out_val [Q, k] fp32 descending, out_idx [Q, k] int64 (-1 / -inf padded when N < k),
 *   out_margin [Q] fp32 (NULL to skip). */
int qst_finalize_topk(const qst_topk_plan* plan, const void* workspace,
                      const float* q_f32, const float* q_inv, const float* q_err,
                      const float* c_f32, const float* c_inv, const float* c_stats,
                      int64_t idx_offset, float* out_val, int64_t* out_idx, float* out_margin,
                      qst_stream_t stream);

/* K3 in two passes: kprime_first (< plan->kprime) candidates are rescored for every query; the queries whose
 * certificate then fails are finalised again with the plan's full k' (the others are left alone).  Same
 * output as qst_finalize_topk, less data gathered (config 3: k' 176 then 224, 0.3 % of the queries take
 * the second pass).  Needs q_err / c_stats / out_margin (the certificate decides who takes the second pass). */
int qst_finalize_topk_adaptive(const qst_topk_plan* plan, int kprime_first, const void* workspace,
                               const float* q_f32, const float* q_inv, const float* q_err,
                               const float* c_f32, const float* c_inv, const float* c_stats,
                               int64_t idx_offset, float* out_val, int64_t* out_idx, float* out_margin,
                               qst_stream_t stream);

/* ---- corpus-sharded retrieval, candidate exchange ----------------------
 * Step 1 on every shard, after qst_score_select: the m best candidates per query BY bf16 KEY, no
 * rescoring.  out_lists [Q, m+1] of 8-byte entries (ordered-uint key, uint32 global row id =
 * local row + idx_offset; id 0xffffffff = empty); entry m of each row is the trailer
 * (bound key of everything the shard did NOT list, number of valid entries).
 * Step 2, on the rank that owns a query, after the lists of all G shards have been exchanged
 * (lists [G, Q, m+1]): select the k' best overall by key, rescore them exactly from the fp32 master
 * (global row ids -> c_f32 must be the full corpus), order, emit top k + certificate, as
 * qst_finalize_topk does.  scratch: qst_finalize_lists_scratch_bytes(Q, G). */
int qst_select_candidates(const qst_topk_plan* plan, const void* workspace, int m, int64_t idx_offset,
                          void* out_lists, qst_stream_t stream);
size_t qst_finalize_lists_scratch_bytes(int64_t Q, int G);
int qst_finalize_lists(int64_t Q, int G, int m, int k, int kprime, int score, int64_t D, const void* lists,
                       const float* q_f32, const float* q_inv, const float* q_err, const float* c_f32,
                       const float* c_inv, const float* c_stats, float* out_val, int64_t* out_idx,
                       float* out_margin, void* scratch, qst_stream_t stream);

/* ---- corpus-sharded retrieval, fp32 master SHARDED too (each GPU keeps its shard
 * only).  After qst_select_candidates on every shard and ONE exchange of the lists (as above):
 *   owner :  qst_select_requests   the k' best candidates overall by bf16 key, grouped by the shard that
 *                                   holds them: out_req [G, Q, m] int32 row ids LOCAL to shard g (-1 = none),
 *                                   out_bound [Q] = bf16 key bounding every document that is not requested;
 *                                   scratch: qst_finalize_lists_scratch_bytes(Q, G).  Shards are the balanced
 *                                   contiguous ranges of n_total rows (first n_total % G shards one row longer).
 *            -- exchange: block g of out_req goes to shard g --
 *   shard :  qst_rescore_requests  exact fp32 scores of the requested rows from the shard's OWN fp32 rows:
 *                                   req [rows, m] (rows = G owners x Q queries, as received), q_f32 [rows, D]
 *                                   the all-gathered fp32 queries in the same order, out [rows, m] fp32
 *                                   (cos_sim / dot_score: the score; euclid_score: ||q-c||^2; -inf where req < 0)
 *            -- exchange back: block o of out goes to owner o --
 *   owner :  qst_finalize_exact    orders the exact scores (ties: lower global id), emits top k + certificate
 *                                   exactly as qst_finalize_topk does; exact [G, Q, m] as received, req the
 *                                   array qst_select_requests wrote; c_stats = max over shards of the two
 *                                   statistics of qst_prep_rows. */
int qst_select_requests(int64_t Q, int G, int m, int kprime, int64_t n_total, const void* lists,
                        int32_t* out_req, uint32_t* out_bound, void* scratch, qst_stream_t stream);
int qst_rescore_requests(int64_t rows, int m, int64_t D, int score, const int32_t* req, const float* q_f32,
                         const float* q_inv, const float* c_f32, const float* c_inv, float* out,
                         qst_stream_t stream);

/* ---- the three exchanges of the sharded path FUSED into their producer kernels over peer memory.
 * Each exchange is "rows grouped in G blocks of rows_per_block rows, block b is meant for rank b"
 * (candidate lists -> owners, requests -> shards, exact scores -> owners).  With a qst_scatter the
 * producing kernel stores every finished row straight into rank b's receive buffer (a peer-mapped
 * buffer from qst_peer_buffer_create / _open, base[b]; base[rank] is the caller's own) at row
 * (rank * rows_per_block + i) -- whole rows, coalesced, over NVLink while the kernel is still
 * producing the next ones -- and the all-to-all that followed it (qst_comm_alltoall, i.e. what replaces
 * the sequential chunk loop of ir_evauation_script.py:161) shrinks to qst_peer_barrier: every rank signals
 * every rank's flag word (release, system scope) and waits for all of theirs.  `flags` reuses the
 * descriptor: base[r] = rank r's flag array (QST_MAX_WORLD uint32, zeroed once), rows_per_block ignored;
 * `epoch` must grow by one per barrier, identically on all ranks.  The receive layout is exactly what the
 * all-to-all would have delivered, so the consumers (qst_select_requests, qst_rescore_requests,
 * qst_finalize_exact) are unchanged.  A rank that never arrives makes the barrier trap after 120 s
 * (environment QST_BARRIER_TIMEOUT_S) instead of hanging the GPU.
 *   qst_select_candidates_scatter : as qst_select_candidates, lists go to dst (G * rows_per_block = plan->Q)
 *   qst_select_requests_scatter   : as qst_select_requests; out_req is still written locally
 *                                   (qst_finalize_exact reads it) and block g also goes to shard g
 *   qst_rescore_requests_scatter  : as qst_rescore_requests; `staging` [rows, m] fp32 is scratch */
typedef struct qst_scatter {
  void* base[QST_MAX_WORLD];
  int32_t world, rank;
  int64_t rows_per_block;
} qst_scatter;
int qst_select_candidates_scatter(const qst_topk_plan* plan, const void* workspace, int m, int64_t idx_offset,
                                  const qst_scatter* dst, qst_stream_t stream);
int qst_select_requests_scatter(int64_t Q, int G, int m, int kprime, int64_t n_total, const void* lists,
                                int32_t* out_req, uint32_t* out_bound, void* scratch, const qst_scatter* dst,
                                qst_stream_t stream);
int qst_rescore_requests_scatter(int64_t rows, int m, int64_t D, int score, const int32_t* req, const float* q_f32,
                                 const float* q_inv, const float* c_f32, const float* c_inv, float* staging,
                                 const qst_scatter* dst, qst_stream_t stream);
int qst_peer_barrier(const qst_scatter* flags, uint32_t epoch, qst_stream_t stream);

int qst_finalize_exact(int64_t Q, int G, int m, int k, int score, int64_t D, int64_t n_total,
                       const int32_t* req, const float* exact, const uint32_t* bound, const float* q_f32,
                       const float* q_err, const float* c_stats, float* out_val, int64_t* out_idx,
                       float* out_margin, qst_stream_t stream);

/* Exact fp32 brute-force re-scan for the queries whose certificate failed (margin <= 0):
 * every corpus row is scored in fp32 against each flagged query and rows scoring at least the
 * current k-th best are collected, which yields the exact top k regardless of bf16 error.
 * Runs entirely on device (no host read of the flags).  scratch from
 * qst_exact_rescan_workspace_bytes().  A pass serves at most 8192 flagged queries (16 per corpus pass);
 * the call queues ceil(Q / 8192) passes so that every flagged query is served (a pass with nothing left
 * to do is three empty launches).  At most 2048 rows are collected per query: a query with more than
 * 2048 rows tied at or above its k-th score cannot be repaired and is left with margin == 0 exactly --
 * the only way a query can come out of this call uncertified. */
size_t qst_exact_rescan_workspace_bytes(int64_t Q, int k);
int qst_exact_rescan(int64_t Q, int64_t N, int64_t D, int k, int score,
                     const float* q_f32, const float* q_inv, const float* c_f32, const float* c_inv,
                     int64_t idx_offset, float* out_val, int64_t* out_idx, float* margin_inout,
                     void* scratch, qst_stream_t stream);
/* The same scan on ONE SHARD of a sharded corpus: for every flagged query (margin[q] < 0; margins are not
 * modified) the shard's rows scoring at least kth_val[q] (the owner's current k-th exact score) are
 * written as a descending list padded with (-inf, -1) to out_val / out_idx [Q, k]; rows of unflagged
 * queries are left untouched.  Merging the G shards' lists (qst_merge_topk) gives the exact top k.
 * overflow [Q] (or NULL): 1 where more than 2048 rows of this shard qualified. */
int qst_exact_rescan_lists(int64_t Q, int64_t N, int64_t D, int k, int score,
                           const float* q_f32, const float* q_inv, const float* c_f32, const float* c_inv,
                           int64_t idx_offset, const float* kth_val, const float* margin,
                           float* out_val, int64_t* out_idx, int* overflow, void* scratch, qst_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K6  merge of per-shard top-k lists (after the NCCL all-gather of SURVEY.md section 8e).
 *   vals [G, Q, k] fp32 descending per shard, idx [G, Q, k] int64 global ids (-1 = empty).
 *   out_val/out_idx [Q, k]: global top-k, descending, ties -> lower global id.
 * ------------------------------------------------------------------------------------------ */
int qst_merge_topk(const float* vals, const int64_t* idx, int G, int64_t Q, int k,
                   float* out_val, int64_t* out_idx, qst_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Collectives of the corpus-sharded path: NCCL over NVLink / NVSwitch, called
 * directly (libnccl.so.2 is resolved with dlopen at first use; the copy already loaded in the process
 * is reused).  One communicator per process and device; qst_comm_init is collective.  The unique id
 * (128 bytes) is created on one rank and distributed by the caller (file, socket, MPI, torch.distributed
 * ...).  Every call is stream-ordered.  What the exchanges replace is the sequential corpus chunk loop
 * of ir_evauation_script.py:161, cut in space instead of time.
 *
 * Sharded top-k from C, per rank (G ranks, q_own queries owned by each, all buffers on the device):
 *   qst_prep_rows (own queries) -> qst_comm_allgather (fp32 queries) -> qst_prep_rows (all G*q_own) ->
 *   qst_topk_plan_make / qst_score_select[_peers] -> qst_select_candidates -> qst_exchange_candidates ->
 *   qst_select_requests -> qst_comm_alltoall (requests) -> qst_rescore_requests -> qst_comm_alltoall
 *   (exact scores) -> qst_finalize_exact  [-> qst_exact_rescan_lists + qst_comm_alltoall + qst_merge_topk
 *   for queries left uncertified];  or, list exchange: qst_finalize_topk locally -> qst_allgather_topk ->
 *   qst_merge_topk.
 * ------------------------------------------------------------------------------------------ */
#define QST_COMM_ID_BYTES 128
typedef struct qst_comm qst_comm;
int qst_comm_available(void);      /* 1 when libnccl.so.2 could be loaded */
int qst_comm_nccl_version(void);   /* e.g. 22809, 0 when unavailable */
int qst_comm_unique_id(unsigned char* id128);
int qst_comm_init(const unsigned char* id128, int world, int rank, qst_comm** out);
int qst_comm_destroy(qst_comm* comm);
int qst_comm_world(const qst_comm* comm);
int qst_comm_rank(const qst_comm* comm);
/* recv [world * bytes_per_rank]: rank-major concatenation of every rank's `send`. */
int qst_comm_allgather(qst_comm* comm, const void* send, void* recv, size_t bytes_per_rank, qst_stream_t stream);
/* block r of send (bytes_per_peer bytes) goes to rank r; block r of recv came from rank r. */
int qst_comm_alltoall(qst_comm* comm, const void* send, void* recv, size_t bytes_per_peer, qst_stream_t stream);
int qst_comm_allreduce_max_f32(qst_comm* comm, const float* send, float* recv, size_t n, qst_stream_t stream);
/* K6's exchange: vals/idx [Q, k] per rank -> out [G, Q, k] on every rank (then qst_merge_topk). */
int qst_allgather_topk(qst_comm* comm, const float* vals, const int64_t* idx, int64_t Q, int k,
                       float* out_vals, int64_t* out_idx, qst_stream_t stream);
/* candidate lists of qst_select_candidates over all G * q_own queries (grouped by owner rank) ->
 * recv [G, q_own, m + 1]: the G shards' lists of the queries this rank owns. */
int qst_exchange_candidates(qst_comm* comm, const void* lists, void* recv, int64_t q_own, int m,
                            qst_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * K4  IR metrics.  Replaces InformationRetrievalEvaluator.compute_metrics / compute_dcg_at_k
 * [sentence-transformers 2.2.2] (Python float64 loops; 5.9 s of 6.7 s at the script's default
 * k-lists) for rankings that are already on the device.
 *
 *   ranked_idx [Q, K] int64 (descending score; -1 = no hit), relevant docs as CSR
 *   (rel_rowptr [Q+1] int64, rel_cols sorted ascending per row, int64).
 *   ks [n_ks] int32 cut-offs (a cut-off may exceed K: the list is then simply shorter, as
 *   top_hits[0:k] is in the reference).  With T = max(K, max ks): log2_tab [T] fp64 =
 *   np.log2(i+2) computed on the host, so that 1/log2_tab[i] is the same IEEE division the
 *   reference performs; idcg_tab [T+1] fp64, idcg_tab[j] = sequential sum of the first j terms
 *   1/log2_tab[i].
 *   out [6, n_ks, Q] fp64 per-query values in the order
 *     0 accuracy (0/1)  1 precision  2 recall  3 reciprocal rank  4 ndcg  5 average precision
 *   The cross-query means are taken on the host with the reference's own reductions
 *   (numpy.mean / sequential +=) so the final numbers are bit-identical given identical rankings.
 * ------------------------------------------------------------------------------------------ */
int qst_ir_metrics(const int64_t* ranked_idx, int64_t Q, int K, const int64_t* rel_rowptr,
                   const int64_t* rel_cols, const int32_t* ks, int n_ks, const double* log2_tab,
                   const double* idcg_tab, double* out, qst_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* QST_H_ */
