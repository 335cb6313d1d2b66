/*
 * tristage.h -- C ABI of libtristage.so: the B200 (sm_100a) candidate-scoring
 * hot path of TriStage-RAG.
 *
 * The reference (pure Python) has no FFI of its own; the boundary it does have
 * is two library calls inside two classes:
 *
 *   Stage 1  faiss.IndexFlatIP(d) / .add(x) / .search(q, k)
 *            /root/reference/src/stage1_retriever.py:263,270,276-277,313,380
 *            + numpy row normalisation                         :285-288
 *   Stage 2  F.normalize / matmul / max / mean (softmax-sum)
 *            /root/reference/src/stage2_rescorer.py:173-183, :188-199
 *            + stable descending sort and truncate             :294-297
 *
 * Every entry point below names the reference line(s) it stands in for.  The
 * ctypes binding a maintainer of the reference would add is in
 * INTEGRATION.md; the shipped binding is tristage_rag_b200/_lib.py.
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch/C++ types.  `stream` is a
 *     cudaStream_t passed as void* (NULL = legacy default stream).
 *   - every function returns TS_OK (0) or a negative ts_status; the message
 *     for the calling thread's last failure is ts_last_error().
 *   - "dev" pointers are CUDA device pointers on the handle's device; "host"
 *     pointers are ordinary host memory (pinned or pageable).
 *   - the caller owns every in/out buffer; the library owns the corpus /
 *     token shards and its scratch space behind the opaque handles.
 *   - a handle is not re-entrant: one call at a time per handle.
 *   - there is NO CPU fallback: without a usable sm_100 device every compute
 *     entry point fails with TS_ERR_CUDA / TS_ERR_UNSUPPORTED.
 */
#ifndef TRISTAGE_H_
#define TRISTAGE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TS_ABI_VERSION 1

typedef enum ts_status {
  TS_OK = 0,
  TS_ERR_INVALID = -1,     /* bad argument                                   */
  TS_ERR_CUDA = -2,        /* a CUDA runtime/driver call failed              */
  TS_ERR_NOMEM = -3,       /* device or host allocation failed               */
  TS_ERR_UNSUPPORTED = -4, /* valid request this build cannot serve          */
  TS_ERR_IO = -5,          /* save/load failure                              */
  TS_ERR_EMPTY = -6        /* search on an empty index (reference raises
                              ValueError, stage1_retriever.py:370-371)       */
} ts_status;

typedef enum ts_dtype { TS_F32 = 0, TS_BF16 = 1, TS_F16 = 2 } ts_dtype;

/* TS_METRIC_IP: rows are scored as stored (the reference normalises BEFORE
 * add, so IP == cosine).  TS_METRIC_COSINE: rows are stored un-normalised and
 * the scan epilogue multiplies each score by 1/(|x|+1e-8) of the stored row
 * (fp32, computed at add time) -- L2-norm scaling fused into the top-k.     */
typedef enum ts_metric { TS_METRIC_IP = 0, TS_METRIC_COSINE = 1 } ts_metric;

/* which Stage-1 scan kernel serves a search call */
typedef enum ts_path {
  TS_PATH_AUTO = 0,   /* umma for bf16/fp16 storage (faster at every B on B200), stream for fp32 */
  TS_PATH_STREAM = 1, /* CUDA-core 128-bit streaming scan (bandwidth path)   */
  TS_PATH_UMMA = 2    /* TMA -> smem -> tcgen05.mma -> TMEM (tensor path); on fp32
                         storage the operands are read as tf32 (10-bit mantissa,
                         fp32 accumulate)                                     */
} ts_path;

/* search / maxsim flags */
#define TS_FLAG_NORMALIZE_Q 1u /* normalise queries with the reference formula
                                  x/(|x|+1e-8) (stage1_retriever.py:377) or, in
                                  Stage 2, F.normalize (stage2_rescorer.py:173) */

typedef enum ts_s2_mode {
  TS_S2_MAXSIM = 0, /* mean_i max_j  (stage2_rescorer.py:180-183)            */
  TS_S2_COLBERT = 1 /* sum_i softmax(m)_i m_i  (stage2_rescorer.py:195-199)  */
} ts_s2_mode;

#define TS_MAX_K 512        /* largest fused top-k (reference default k = 500) */
#define TS_S2_MAX_LQ 128    /* query tokens per query (reference: <= 192 after
                               truncation; BASELINE config: 32)               */
#define TS_S2_MAX_LD 256    /* doc tokens per doc (reference max_seq_length 192) */

typedef struct ts_index ts_index;       /* one Stage-1 corpus shard            */
typedef struct ts_tokstore ts_tokstore; /* one Stage-2 token-embedding shard   */

/* ------------------------------------------------------------------ misc -- */
int ts_abi_version(void);
const char* ts_last_error(void);
/* number of visible CUDA devices with compute capability 10.x; <0 on error  */
int ts_device_count(void);

/* ---------------------------------------------------------------- Stage 1 -- */

/* faiss.IndexFlatIP(d)   (stage1_retriever.py:263,276).
 * storage: TS_BF16 / TS_F16 / TS_F32 (both scan kernels; fp32 rows are read
 * as tf32 on the tensor path).
 * reserve_rows: rows to pre-allocate (grows by doubling beyond that).        */
int ts_index_create(ts_index** out, int device, int dim, int storage_dtype, int metric,
                    int64_t reserve_rows);
int ts_index_destroy(ts_index* h);

/* index.add(x) preceded by _normalize_embeddings
 * (stage1_retriever.py:307 + :270,277,313).
 * rows: [n, dim] row-major, src_dtype TS_F32 or the storage dtype, on the host
 * (src_on_device = 0) or the device.  normalize != 0 applies x/(|x|+1e-8) in
 * fp32 before the cast to the storage dtype (METRIC_IP) or records the
 * inverse norms (METRIC_COSINE).  Row ids are positional: ntotal .. ntotal+n. */
int ts_index_add(ts_index* h, const void* rows, int64_t n, int src_dtype, int src_on_device,
                 int normalize, void* stream);

int64_t ts_index_ntotal(const ts_index* h);
int ts_index_dim(const ts_index* h);
/* mcp clear_index (src/mcp_retrieval_server.py:243-247): drop all rows       */
int ts_index_reset(ts_index* h);
/* global id of local row 0 (row-sharded corpus: rank r owns [base, base+n))  */
int ts_index_set_id_base(ts_index* h, int64_t id_base);

/* index.search(q, k)   (stage1_retriever.py:380), batched.
 * q_dev:  [B, dim] row-major, q_dtype TS_F32 or the storage dtype (device).
 * out:    scores [B, k] fp32 descending, ids [B, k] int64 (global ids);
 *         unused slots hold id -1 and the lowest float, as FAISS does.
 * Ties: score descending, then id ascending (deterministic).
 * Everything is enqueued on `stream`; nothing synchronises.                  */
int ts_index_search(ts_index* h, const void* q_dev, int q_dtype, int B, int k, unsigned flags,
                    int path, float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* Same call with HOST buffers (the reference's numpy in / numpy out): copies
 * q host->device, searches, copies results back and synchronises `stream`.   */
int ts_index_search_host(ts_index* h, const void* q_host, int q_dtype, int B, int k,
                         unsigned flags, int path, float* out_scores_host,
                         int64_t* out_ids_host, void* stream);

/* Merge n_lists per-shard results (what an all-gather of ts_index_search
 * outputs delivers) into one top-k per query: scores/ids [n_lists, B, k] ->
 * [B, k].  Lists must be ordered by ascending id range for the id-ascending
 * tie rule to hold across shards.  Device pointers.                          */
int ts_topk_merge(int device, const float* scores_dev, const int64_t* ids_dev, int n_lists, int B,
                  int k, float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* Same merge over ONE packed buffer per list, the layout a single all-gather
 * delivers: list l lives at blob + l*list_stride_bytes and holds its [B, k]
 * fp32 scores at offset 0 and its [B, k] int64 ids at ids_offset_bytes.      */
int ts_topk_merge_packed(int device, const void* blob_dev, int64_t list_stride_bytes,
                         int64_t ids_offset_bytes, int n_lists, int B, int k,
                         float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* Multi-GPU exchange over PEER MEMORY (NVLink / NVSwitch) instead of an all-gather followed by
 * ts_topk_merge_packed.  Every rank owns one symmetric receive buffer, mapped into all ranks:
 *   [parity 0: n_ranks slots of slot_bytes][parity 1: n_ranks slots] ... [flags: 2 * n_ranks u32 at flags_offset]
 * Step s uses parity s & 1 and sequence number seq = s + 1 (never 0).  ts_exchange_push writes this
 * rank's packed [B,k] result (the ts_index_search packed layout: fp32 scores at 0, int64 ids at
 * ids_offset) into slot `rank` of EVERY rank's buffer with 16-byte stores and then publishes seq in
 * that rank's flag (system-scope release).  ts_exchange_wait_merge spins (system-scope acquire,
 * bounded: a missing peer traps after ~4 s) until all n_ranks flags of its own buffer show seq and
 * merges the n_ranks lists into out [B,k].  Two parities suffice: a rank can only be one step ahead
 * of the slowest peer, because its next wait needs that peer's next push.
 * peer_bases_dev: device array [n_ranks] of the buffers' base addresses as seen from this rank.    */
int ts_exchange_push(int device, const void* blob_dev, int64_t nbytes, const int64_t* peer_bases_dev, int n_ranks,
                     int rank, int64_t slot_bytes, int64_t flags_offset, int parity, uint32_t seq, void* stream);
int ts_exchange_wait_merge(int device, const void* local_base_dev, int n_ranks, int B, int k, int64_t slot_bytes,
                           int64_t ids_offset, int64_t flags_offset, int parity, uint32_t seq,
                           float* out_scores_dev, int64_t* out_ids_dev, void* stream);

/* FUSED exchange (the default multi-GPU data plane when peer memory is available): no separate push kernel and
 * no collective-library call in the step.  The final pass of the local top-k selection stores query b's k
 * (score, id) pairs straight into slot `rank` of EVERY rank's receive buffer (NVLink stores; the own buffer is one
 * of them) and publishes flag[parity][rank][b] = seq with a system-scope release; the merge kernel (one CTA per
 * query) waits with system-scope acquires -- bounded: a missing peer traps after ~4 s -- until all n_ranks rows of
 * ITS query have arrived and merges them.  A step is therefore: query prep, scan, select+push, wait+merge.
 * Receive buffer of one rank (ts_exchange_buffer_bytes; allocate it zeroed as symmetric / peer-mapped memory):
 *   [parity 2][rank n] slots of { scores f32[B_max*k_max] | ids i64[B_max*k_max] }, then flags u32 [parity 2][rank n][B_max].
 * All ranks must make the same sequence of ts_index_search_push / ts_exchange_merge calls (SPMD); two parities
 * suffice because a rank can be at most one step ahead of its slowest peer.
 * peer_bases_host[n_ranks]: base address of every rank's buffer as seen from THIS device (own buffer at [rank]). */
typedef struct ts_exchange ts_exchange;
int64_t ts_exchange_buffer_bytes(int n_ranks, int B_max, int k_max);
int ts_exchange_create(ts_exchange** out, int device, int rank, int n_ranks, const int64_t* peer_bases_host,
                       int B_max, int k_max);
int ts_exchange_destroy(ts_exchange* x);
/* local search (as ts_index_search) whose result goes to every rank's buffer instead of a local output          */
int ts_index_search_push(ts_index* h, ts_exchange* x, const void* q_dev, int q_dtype, int B, int k, unsigned flags,
                         int path, void* stream);
/* wait for all ranks' rows of this step and merge them: identical [B, k] result on every rank                    */
int ts_exchange_merge(ts_exchange* x, int B, int k, float* out_scores_dev, int64_t* out_ids_dev, void* stream);
/* the whole step in one call.  When every CTA of the select kernel is co-resident (B <= SM count) and the n_ranks * k
 * keys of a query fit its buffer (k <= 128, n_ranks * k <= 2048) the exchange is ONE kernel: CTA b pushes query b's row,
 * then waits for the peers' rows of query b and merges them (every rank pushes before it waits: no cycle); otherwise
 * ts_index_search_push + ts_exchange_merge back to back.  Same results either way (TS_XFUSE=0 forces two kernels).
 * The _host variant takes pinned or pageable HOST buffers (H2D, step, D2H, synchronise) -- the sharded counterpart
 * of ts_index_search_host                                                                                        */
int ts_index_search_sharded(ts_index* h, ts_exchange* x, const void* q_dev, int q_dtype, int B, int k, unsigned flags,
                            int path, float* out_scores_dev, int64_t* out_ids_dev, void* stream);
int ts_index_search_sharded_host(ts_index* h, ts_exchange* x, const void* q_host, int q_dtype, int B, int k,
                                 unsigned flags, int path, float* out_scores_host, int64_t* out_ids_host, void* stream);

/* The same exchange for the Stage-2 score matrices (replaces the all-reduce(SUM) of the per-rank
 * ts_maxsim outputs): every rank pushes its [B, C] fp32 matrix with ts_exchange_push (nbytes = B*C*4),
 * ts_exchange_wait_sum waits for all n_ranks slots of the step and writes their element-wise sum.   */
int ts_exchange_wait_sum(int device, const void* local_base_dev, int n_ranks, int64_t n_floats, int64_t slot_bytes,
                         int64_t flags_offset, int parity, uint32_t seq, float* out_dev, void* stream);

/* Stage 2 with the exchange FUSED into the scoring kernel (replaces ts_maxsim + all-reduce(SUM); the loop it stands
 * for is stage2_rescorer.py:268-291 run on every shard): every candidate has exactly one owner, so the owner's kernel
 * stores each finalised score at its flat position b*C + j of the [B, C] fp32 matrix at matrix_offset inside EVERY
 * rank's receive buffer (peer_bases_dev[n_ranks]: the buffers as seen from this GPU, NVLink stores) -- nothing is left
 * to reduce.  The last CTA of the grid then publishes seq (never 0) in u32 flag `rank` at flags_offset of every
 * buffer (system-scope release).  Needs the tile-layout tensor kernel (bf16/fp16 store, dim % 16 == 0, dim <= 256,
 * lq_stride <= 128), else TS_ERR_UNSUPPORTED -- decide for the whole group before the first call.
 * ts_exchange_wait_take is the consumer: it waits (system-scope acquire, bounded) until the n_ranks flags at
 * flags_dev show seq, copies the n_floats of matrix_dev to out_dev and leaves ZEROS behind: positions nobody owns
 * read 0.0 like ts_maxsim's output, and the matrix is clean when it is used again.  Use two (matrix, flags) pairs
 * alternately (step parity): a rank can be at most one step ahead of its slowest peer.                            */
int ts_maxsim_scatter(ts_tokstore* h, const void* q_tok_dev, int q_dtype, const int32_t* q_len_dev, int B,
                      int lq_stride, const int64_t* cand_dev, const int32_t* n_cand_dev, int C, int mode,
                      unsigned flags, const int64_t* peer_bases_dev, int n_ranks, int rank,
                      int64_t matrix_offset, int64_t flags_offset, uint32_t seq, void* stream);
int ts_exchange_wait_take(int device, void* matrix_dev, const void* flags_dev, int n_ranks, uint32_t seq,
                          int64_t n_floats, float* out_dev, void* stream);

/* Measurement aid: with TS_DBG_TIMELINE=1 in the environment the kernels of a tensor-path search stamp the GPU's
 * globaltimer (ns) into 16 slots of the handle; this call synchronises the device and copies the slots of the LAST step:
 * [0] query prep starts, [1] ends, [2] first scan CTA enters its epilogue, [3] last scan CTA leaves, [4] select kernel
 * (CTA 0) starts, [5] local rows sorted, [6] pushed to every rank + published, [7] the peers' rows of query 0 are in,
 * [8] merged result written (0 = not reached in this configuration).                                                 */
int ts_index_debug_timeline(ts_index* h, uint64_t* out16);

/* faiss.write_index / read_index  (stage1_retriever.py:436,463): one shard
 * file per handle (layout: see "shard files" below).  save synchronises the
 * device; load verifies the file's checksums.                                */
int ts_index_save(const ts_index* h, const char* path);
int ts_index_load(ts_index** out, int device, const char* path);
/* Append rows [row_lo, row_lo + n_rows) of an index shard file to h (same dim,
 * storage dtype and metric).  This is the primitive behind re-sharding: a
 * corpus saved by G ranks is loaded by G' ranks, each appending the pieces of
 * the old files that intersect its new row range.  id_base is NOT taken from
 * the file (use ts_index_set_id_base).  Synchronises `stream`.               */
int ts_index_append_file(ts_index* h, const char* path, int64_t row_lo, int64_t n_rows, void* stream);
int ts_index_dtype(const ts_index* h);
int ts_index_metric(const ts_index* h);

/* copy stored rows [start, start+n) back as fp32 [n, dim] to host (tests,
 * migration); not on the hot path.                                           */
int ts_index_get_rows(const ts_index* h, int64_t start, int64_t n, float* out_host);

/* kernels launched by this handle since creation (bench.py gpu_launches)     */
int64_t ts_index_launch_count(const ts_index* h);

/* Measurement aid (bench.py roofline): while enabled, every search brackets its
 * SCAN kernel (stream or umma, not the query prep / merge) with CUDA events on
 * the launching stream.  ts_index_scan_time reports the mean duration (ms) of
 * the scans recorded since the last call (at most 256) and clears the record;
 * it synchronises on the last event.                                         */
int ts_index_set_profiling(ts_index* h, int enable);
int ts_index_scan_time(ts_index* h, float* mean_ms_out, int* n_out);

/* ---------------------------------------------------------------- Stage 2 -- */

/* Token-embedding shard: the reference re-encodes every candidate per query
 * (stage2_rescorer.py:255-259); the store keeps encode_documents_batch's
 * outputs ([Ld_i, dim] per doc, :226-231) resident, L2-normalised at add.    */
int ts_tokstore_create(ts_tokstore** out, int device, int dim, int storage_dtype,
                       int64_t reserve_docs, int64_t reserve_tokens);
int ts_tokstore_destroy(ts_tokstore* h);

/* tok: concatenated [sum(lens), dim] rows, src_dtype TS_F32 or storage dtype,
 * host or device; lens_host[n_docs] (1 <= len <= TS_S2_MAX_LD).
 * normalize != 0 applies F.normalize (x / max(|x|, 1e-12), :174) per token
 * in fp32 before the cast.  Doc ids are positional.                          */
int ts_tokstore_add(ts_tokstore* h, const void* tok, int src_dtype, int src_on_device,
                    const int32_t* lens_host, int n_docs, int normalize, void* stream);
int64_t ts_tokstore_ndocs(const ts_tokstore* h);
int64_t ts_tokstore_ntokens(const ts_tokstore* h);
int ts_tokstore_reset(ts_tokstore* h);
int ts_tokstore_set_id_base(ts_tokstore* h, int64_t id_base);
int64_t ts_tokstore_launch_count(const ts_tokstore* h);
/* persistence of one token shard (no reference equivalent: the reference
 * re-encodes at query time, stage2_rescorer.py:255-259).                     */
int ts_tokstore_save(const ts_tokstore* h, const char* path);
int ts_tokstore_load(ts_tokstore** out, int device, const char* path);
/* append docs [doc_lo, doc_lo + n_docs) of a token shard file (re-sharding)  */
int ts_tokstore_append_file(ts_tokstore* h, const char* path, int64_t doc_lo, int64_t n_docs, void* stream);
int ts_tokstore_dim(const ts_tokstore* h);
int ts_tokstore_dtype(const ts_tokstore* h);
/* HBM layout of the shard, fixed at creation: 0 = row-major rows with zero pad rows; 1 = tile layout
 * (tok[row/8][dim/8][row%8][8], pad rows repeat the doc's last token) -- the tcgen05 operand image the
 * Stage-2 tensor kernel copies with one cp.async.bulk per doc; chosen for 2-byte dtypes with dim % 16 == 0,
 * dim <= 256.  Shard FILES always hold the row-major image (save / load / append_file convert).          */
int ts_tokstore_layout(const ts_tokstore* h);
/* same measurement aid for the MaxSim kernel                                 */
int ts_tokstore_set_profiling(ts_tokstore* h, int enable);
int ts_tokstore_scan_time(ts_tokstore* h, float* mean_ms_out, int* n_out);

/* _maxsim_score / _colbert_score for every (query, candidate) pair in one
 * launch (replaces the loop at stage2_rescorer.py:268-273).
 * q_tok_dev:  [B, lq_stride, dim], q_dtype TS_F32 or storage dtype
 * q_len_dev:  [B] int32 real query-token counts (NULL -> lq_stride each)
 * cand_dev:   [B, C] int64 global doc ids; entries outside this shard's
 *             [id_base, id_base+ndocs) -- including -1 -- score 0.0, so
 *             per-shard outputs can be summed (all-reduce) across ranks
 * n_cand_dev: [B] int32 valid candidates per query (NULL -> C each)
 * out:        [B, C] fp32 scores, 0.0 beyond n_cand                          */
int ts_maxsim(ts_tokstore* h, const void* q_tok_dev, int q_dtype, const int32_t* q_len_dev, int B,
              int lq_stride, const int64_t* cand_dev, const int32_t* n_cand_dev, int C, int mode,
              unsigned flags, float* out_scores_dev, void* stream);

/* host-buffer variant (copies in, scores, copies out, synchronises)          */
int ts_maxsim_host(ts_tokstore* h, const void* q_tok_host, int q_dtype, const int32_t* q_len_host,
                   int B, int lq_stride, const int64_t* cand_host, const int32_t* n_cand_host,
                   int C, int mode, unsigned flags, float* out_scores_host, void* stream);

/* scored_candidates.sort(reverse=True)[:top_k]  (stage2_rescorer.py:294-297):
 * per query, the positions of the top_k scores in STABLE descending order
 * (equal scores keep incoming order).  scores [B, C], n_cand [B] or NULL;
 * out_pos [B, top_k] int32 (-1 padded), out_scores [B, top_k].               */
int ts_rank_desc(int device, const float* scores_dev, const int32_t* n_cand_dev, int B, int C,
                 int top_k, float* out_scores_dev, int32_t* out_pos_dev, void* stream);

/* --------------------------------------------------- hybrid (BM25 + dense) -- */
/* Device-side BM25 search and rank fusion.  The
 * reference scores every document with a Python loop per query and sorts all
 * of them (BM25Index.search, stage1_retriever.py:84-112); here the fitted
 * index lives on the device as CSR postings with one fp64 weight per posting,
 *   w = idf * tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl))          (:93-99)
 * precomputed by the caller (it depends on the fit, not on the query), and a
 * query only touches the documents that share a token with it.  Arithmetic
 * and order are the reference's: fp64 sums in query-token order, stable
 * descending rank (ties and the zero-score tail by ascending document index). */
#define TS_BM25_MAX_K 1024
typedef struct ts_bm25 ts_bm25;
/* term_off_host [n_terms + 1], post_doc_host / post_w_host [term_off[n_terms]];
 * every weight must be > 0 (TS_ERR_UNSUPPORTED otherwise: the reference's
 * stale-refit quirk can make idf <= 0 -- keep the host search for that index) */
int ts_bm25_create(ts_bm25** out, int device, int64_t n_docs, int64_t n_terms, const int64_t* term_off_host,
                   const int32_t* post_doc_host, const double* post_w_host);
int ts_bm25_destroy(ts_bm25* h);
int64_t ts_bm25_ndocs(const ts_bm25* h);
int64_t ts_bm25_launch_count(const ts_bm25* h);
/* BM25Index.search(query, top_k) for B queries.  q_terms_host: term ids of each
 * query in query-token order (repeated tokens repeated, tokens unknown to the
 * index left out), query b = [q_off_host[b], q_off_host[b+1]).
 * out: [B, top_k] fp64 scores and int64 document indices, -1 beyond n_docs.  */
int ts_bm25_search_host(ts_bm25* h, const int32_t* q_terms_host, const int64_t* q_off_host, int B, int top_k,
                        double* out_scores_host, int64_t* out_ids_host, void* stream);
/* _reciprocal_rank_fusion (method 0, :326-343) / _weighted_fusion (method 1,
 * :345-366) of a dense list ([B, k1] ids + fp32 scores, n_dense[B] valid or
 * NULL) and a BM25 list ([B, k2] ids + fp64 scores) in fp64; entries keep dict
 * insertion order (dense first, then BM25-only), stable descending sort, the
 * best top_k per query go to out_ids/out_scores [B, top_k] (-1 padded), their
 * number to out_n[B].  k1 + k2 <= 2048.                                      */
int ts_hybrid_fuse_host(int device, int method, int rrf_k, double w_dense, double w_bm25,
                        const int64_t* dense_ids_host, const float* dense_scores_host,
                        const int32_t* n_dense_host, int k1, const int64_t* bm25_ids_host,
                        const double* bm25_scores_host, const int32_t* n_bm25_host, int k2, int B,
                        int top_k, int64_t* out_ids_host, double* out_scores_host, int32_t* out_n_host,
                        void* stream);

/* ------------------------------------------- approximate mode (inverted lists) -- */
/* faiss.IndexIVFFlat(IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT), the index
 * the reference builds when its first batch has more than 1000 rows
 * (stage1_retriever.py:262-273), as a view over the rows of an existing
 * ts_index: the corpus is NOT copied, the lists hold row numbers (SURVEY.md
 * section 8f-4).  `base` must outlive the ts_ivf.  Scores are those of the
 * exact scan restricted to the probed lists; nprobe == nlist gives the exact
 * result.  Ties: score descending, then id ascending.                        */
#define TS_IVF_MAX_NLIST 4096
typedef struct ts_ivf ts_ivf;
int ts_ivf_create(ts_ivf** out, ts_index* base, int nlist);
int ts_ivf_destroy(ts_ivf* h);
int ts_ivf_nlist(const ts_ivf* h);
int ts_ivf_is_trained(const ts_ivf* h);
/* rows of `base` that are in a list (ts_ivf_sync brings it up to ntotal)     */
int64_t ts_ivf_nassigned(const ts_ivf* h);
int64_t ts_ivf_launch_count(const ts_ivf* h);
/* index.train(x) (:267): installs the coarse centroids [nlist, dim] fp32 (host).
 * The k-means itself runs in the binding (tristage_rag_b200/ivf.py), as FAISS
 * runs it on the host for the reference.  Drops all lists.                   */
int ts_ivf_set_centroids(ts_ivf* h, const float* centroids_host, void* stream);
int ts_ivf_get_centroids(const ts_ivf* h, float* out_host);
/* index.add(x) (:270,313), list part: every row of `base` not yet in a list
 * goes to the list of the centroid with the largest inner product (fp32, ties
 * to the lowest list); the lists are rebuilt.  Called by ts_ivf_search when
 * rows were added since the last call.  Synchronises `stream`.               */
int ts_ivf_sync(ts_ivf* h, void* stream);
/* install the list of every row directly (n == ntotal of base): lists saved
 * by ts_ivf_get_assignments, or taken from an index file written by the
 * reference (tristage_rag_b200/faiss_io.py)                                  */
int ts_ivf_set_assignments(ts_ivf* h, const int32_t* assign_host, int64_t n, void* stream);
int ts_ivf_get_assignments(const ts_ivf* h, int32_t* out_host, int64_t n);
int ts_ivf_list_sizes(const ts_ivf* h, int64_t* out_host /* [nlist] */);
/* quantizer.search(q, nprobe): the nprobe lists with the largest <q, centroid>
 * (fp32 queries on the host), best first: out_lists [B, nprobe] int32,
 * out_scores [B, nprobe] fp32.                                               */
int ts_ivf_coarse_host(ts_ivf* h, const void* q_host, int B, int nprobe, unsigned flags,
                       int32_t* out_lists_host, float* out_scores_host, void* stream);
/* index.nprobe = nprobe; index.search(q, k)  (:273,380).  Arguments and
 * results as ts_index_search; slots beyond the rows of the probed lists hold
 * id -1 and the lowest float.  nprobe is clamped to 1..nlist.                */
int ts_ivf_search(ts_ivf* h, const void* q_dev, int q_dtype, int B, int k, int nprobe, unsigned flags,
                  float* out_scores_dev, int64_t* out_ids_dev, void* stream);
int ts_ivf_search_host(ts_ivf* h, const void* q_host, int q_dtype, int B, int k, int nprobe,
                       unsigned flags, float* out_scores_host, int64_t* out_ids_host, void* stream);

/* ------------------------------------------------------------ shard files -- */
/* On-disk form of one shard, little endian, sections
 * 4096-byte aligned so the payload can be mmap'ed:
 *   header page | table (index: inv_norm f32[n] for TS_METRIC_COSINE;
 *   tokstore: doc_off i64[n] + doc_len i32[n]) | payload (index: rows[n][ld];
 *   tokstore: tok[nrows][dim], docs padded to 8 rows), storage dtype.
 * Each section carries an XXH64 digest; files are written to "<path>.tmp" and
 * renamed into place.  The functions in this block are HOST ONLY (no CUDA
 * call): tools and CPU tests can inspect, check and produce shard files.     */
typedef enum ts_file_kind { TS_FILE_INDEX = 1, TS_FILE_TOKSTORE = 2 } ts_file_kind;
typedef struct ts_file_info {
  int32_t kind, version, dim, ld, dtype, metric;
  int64_t n;       /* rows (index) or docs (tokstore)                          */
  int64_t nrows;   /* payload rows (index: n; tokstore: 8-row padded tokens)   */
  int64_t ntokens; /* tokstore: real tokens                                    */
  int64_t id_base; /* global id of row / doc 0 when the shard was saved        */
  uint64_t table_offset, table_bytes, payload_offset, payload_bytes;
  uint64_t table_hash, payload_hash;
} ts_file_info;
/* parse + validate the header and the section bounds                         */
int ts_file_probe(const char* path, ts_file_info* out);
/* recompute both digests (and, for a token shard, check that the doc table
 * adds up to the payload); TS_ERR_IO on any mismatch                         */
int ts_file_verify(const char* path);
/* Write a shard file from HOST memory already in the storage dtype -- no
 * arithmetic, only layout (migration of an existing matrix, tests).
 * rows_storage: [n][ld] with ld = dim rounded up to 16 bytes, pad columns 0;
 * inv_norm: [n] for TS_METRIC_COSINE, else NULL.                             */
int ts_file_write_index_host(const char* path, int dim, int storage_dtype, int metric, int64_t n,
                             int64_t id_base, const void* rows_storage, const float* inv_norm);
/* tok_storage: [sum(lens)][dim] un-padded, already normalised; the writer
 * inserts the zero rows that pad every doc to a multiple of 8.               */
int ts_file_write_tokstore_host(const char* path, int dim, int storage_dtype, int64_t n_docs,
                                int64_t id_base, const int32_t* lens, const void* tok_storage);

#ifdef __cplusplus
}
#endif
#endif /* TRISTAGE_H_ */
