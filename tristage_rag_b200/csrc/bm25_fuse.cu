// Hybrid (lexical + dense) candidate generation on the device.
//
// Replaces, for a batch of queries:
//   * BM25Index.search (/root/reference/src/stage1_retriever.py:84-112): the reference scores
//     EVERY document with a Python loop per query and sorts all N of them;
//   * Stage1Retriever._reciprocal_rank_fusion (:326-343) and _weighted_fusion (:345-366).
//
// BM25 arithmetic is fp64 like the reference's Python floats.  The per-posting contribution
//     w = idf * tf*(k1+1) / (tf + k1*(1 - b + b*dl/avgdl))                      (:93-99)
// depends only on the fitted index, never on the query, so the caller precomputes it once per
// fit with the reference's expression (tristage_rag_b200/stage1_retriever.py::BM25Index) and
// the device keeps CSR postings (term_off, post_doc, post_w).  A query is a list of term ids in
// query-token order (a token that occurs twice is listed twice, :93-99); the kernel walks the
// tokens in that order with a block barrier in between, so every document's sum is accumulated
// in the reference's order and is bit-identical to it.
//
// bm25_score_kernel   grid (query, document range): scores[b][doc] += w over the postings of each
//                     query token that fall into the CTA's document range; a document's first hit (score still 0.0 -- all weights are > 0, the
//                     caller falls back to the host otherwise) appends it to the query's
//                     "touched" list.  Bound: HBM/L2 latency of the scattered fp64 updates;
//                     bytes = postings touched * (4 + 8 + 16).
// (each (query, range) CTA also selects its own best top_k and hands only those on)
// bm25_topk_kernel    one CTA per query: stable descending top-k of the ranges' candidates
//                     (score desc, doc index asc == CPython's stable sort with reverse=True over
//                     enumerate(scores)), then zero-score documents in ascending index order
//                     until k entries are filled, exactly what ranking all N scores gives.
// fuse_kernel         one CTA per query: RRF / weighted fusion of the dense top-k and the BM25
//                     list in fp64, entries kept in dict insertion order (dense first, then
//                     BM25-only), stable descending sort, truncate.
#include <stdlib.h>
#include <string.h>

#include "ts_common.cuh"
#include "ts_handles.h"

namespace ts {
namespace {

constexpr int kBmThreads = 512;
constexpr int kBmBuf = 2048;   // (score, doc) pairs of shared memory per CTA in the top-k kernel

// IEEE double arithmetic without contraction, so device sums equal CPython's
__device__ __forceinline__ double dadd(double a, double b) {
#ifdef TS_CUDASIM
  volatile double r = a + b; return r;
#else
  return __dadd_rn(a, b);
#endif
}
__device__ __forceinline__ double dmul(double a, double b) {
#ifdef TS_CUDASIM
  volatile double r = a * b; return r;
#else
  return __dmul_rn(a, b);
#endif
}
__device__ __forceinline__ double ddiv(double a, double b) {
#ifdef TS_CUDASIM
  volatile double r = a / b; return r;
#else
  return __ddiv_rn(a, b);
#endif
}

// "a ranks before b": score descending, then tie-break key ascending (doc index / insertion order)
__device__ __forceinline__ bool before(double sa, int ka, double sb, int kb) {
  return sa > sb || (sa == sb && ka < kb);
}

// block-wide bitonic sort of n = 2^m (score, key) pairs in shared memory, best first
__device__ __forceinline__ void block_sort_pairs(double* s, int* k, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int e = ((i / stride) * (stride << 1)) + (i % stride);
        const int p = e + stride;
        const bool asc_block = ((e & size) == 0);          // this run wants "best first"
        const bool p_first = before(s[p], k[p], s[e], k[e]);
        if (asc_block ? p_first : !p_first) {
          const double ts_ = s[e]; s[e] = s[p]; s[p] = ts_;
          const int tk = k[e]; k[e] = k[p]; k[p] = tk;
        }
      }
      __syncthreads();
    }
  }
}

constexpr double kWorst = -1.7976931348623157e308;   // pads the sort buffer: ranks after everything
constexpr int kWorstKey = 0x7fffffff;

// exclusive rank of this thread's flag among the block's flags (block-wide scan over one chunk
// of blockDim.x flags); *total = flags set in the chunk.  Ends with a barrier (s_wsum reusable).
__device__ __forceinline__ int block_rank_of_flag(int z, int* s_wsum, int* total) {
  int incl = z;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((int)(threadIdx.x & 31) >= o) incl += v;
  }
  if ((threadIdx.x & 31) == 31) s_wsum[threadIdx.x >> 5] = incl;
  __syncthreads();
  int before_me = 0, tot = 0;
  for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
    const int v = s_wsum[w];
    if (w < (int)(threadIdx.x >> 5)) before_me += v;
    tot += v;
  }
  __syncthreads();
  *total = tot;
  return before_me + incl - z;
}

// Chunked block-wide selection: best `keep` of the T documents listed in tl[] (scores looked up in sc[]) end up in
// s_sc / s_doc [0, keep), best first.  Buffer = [best so far | next chunk], sort, repeat.  All threads call it.
__device__ __forceinline__ void block_select_docs(const double* __restrict__ sc, const int32_t* __restrict__ tl, int T, int keep,
                                                  double* s_sc, int* s_doc) {
  const int room = kBmBuf - keep;
  for (int i = threadIdx.x; i < keep; i += blockDim.x) { s_sc[i] = kWorst; s_doc[i] = kWorstKey; }
  __syncthreads();
  for (int off = 0; off < T; off += room) {
    const int take = min(room, T - off);
    for (int i = threadIdx.x; i < room; i += blockDim.x) {
      if (i < take) { const int d = tl[off + i]; s_doc[keep + i] = d; s_sc[keep + i] = sc[d]; }
      else { s_doc[keep + i] = kWorstKey; s_sc[keep + i] = kWorst; }
    }
    __syncthreads();
    block_sort_pairs(s_sc, s_doc, kBmBuf);
  }
}

// first index in post_doc[lo, hi) whose document is >= d (postings of a term are ascending in document)
__device__ __forceinline__ int64_t lower_bound_doc(const int32_t* __restrict__ post_doc, int64_t lo, int64_t hi, int64_t d) {
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (post_doc[mid] < d) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// grid (B, R): CTA (b, r) owns the documents [r*span, (r+1)*span) of query b.  A document belongs to
// exactly one CTA, so walking the query tokens in order with a block barrier in between keeps every
// document's fp64 sum in the reference's accumulation order without any cross-CTA synchronisation.
// The CTA lists the documents it touched in its own segment of `touched`, selects its best top_k of them
// (anything outside a range's best top_k cannot be in the query's best top_k) and appends those to the
// query's candidate list; touched_cnt[b] accumulates the TRUE number of touched documents.
__global__ void __launch_bounds__(kBmThreads)
    bm25_score_kernel(const int64_t* __restrict__ term_off, const int32_t* __restrict__ post_doc,
                      const double* __restrict__ post_w, const int32_t* __restrict__ q_terms,
                      const int64_t* __restrict__ q_off, int64_t n_docs, int64_t span, int64_t seg_cap, int top_k,
                      double* __restrict__ scores, int32_t* __restrict__ touched, int32_t* __restrict__ touched_cnt,
                      int32_t* __restrict__ cand, int64_t cand_cap, int32_t* __restrict__ cand_cnt) {
  __shared__ double s_sc[kBmBuf];
  __shared__ int s_doc[kBmBuf];
  __shared__ int s_n, s_at;
  const int b = blockIdx.x;
  const int64_t d_lo = (int64_t)blockIdx.y * span;
  const int64_t d_hi = (d_lo + span < n_docs) ? d_lo + span : n_docs;
  double* sc = scores + (size_t)b * n_docs;
  int32_t* tl = touched + ((size_t)b * gridDim.y + blockIdx.y) * seg_cap;     // this CTA's segment
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  for (int64_t ti = q_off[b]; ti < q_off[b + 1]; ++ti) {
    const int t = q_terms[ti];
    int64_t p0 = term_off[t], p1 = term_off[t + 1];
    if (gridDim.y > 1) {                              // this CTA's slice of the term's postings
      p0 = lower_bound_doc(post_doc, p0, p1, d_lo);
      p1 = lower_bound_doc(post_doc, p0, p1, d_hi);
    }
    for (int64_t p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      const int d = post_doc[p];                      // a document occurs once per term: no atomics needed
      const double old = sc[d];
      sc[d] = dadd(old, post_w[p]);
      if (old == 0.0) {                               // first query token that hits this document
        const int at = atomicAdd(&s_n, 1);
        if (at < seg_cap) tl[at] = d;
      }
    }
    __syncthreads();                                  // the next token may hit the same documents
    __threadfence_block();
  }
  const int T = min(s_n, (int)min((int64_t)0x7fffffff, seg_cap));
  const int keep = min(top_k, kBmBuf / 2);
  block_select_docs(sc, tl, T, keep, s_sc, s_doc);
  const int n_out = min(T, keep);
  if (threadIdx.x == 0) {
    s_at = atomicAdd(cand_cnt + b, n_out);
    atomicAdd(touched_cnt + b, T);
  }
  __syncthreads();
  int32_t* cl = cand + (size_t)b * cand_cap + s_at;
  for (int i = threadIdx.x; i < n_out; i += blockDim.x) cl[i] = s_doc[i];
}

__global__ void __launch_bounds__(kBmThreads)
    bm25_topk_kernel(const double* __restrict__ scores, int64_t n_docs, const int32_t* __restrict__ cand, int64_t cand_cap,
                     const int32_t* __restrict__ cand_cnt, const int32_t* __restrict__ touched_cnt, int top_k,
                     double* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  __shared__ double s_sc[kBmBuf];
  __shared__ int s_doc[kBmBuf];
  __shared__ int s_wsum[kBmThreads / 32];
  const int b = blockIdx.x;
  const double* sc = scores + (size_t)b * n_docs;
  const int32_t* tl = cand + (size_t)b * cand_cap;
  const int n_cand = min(cand_cnt[b], (int)min((int64_t)0x7fffffff, cand_cap));
  const int T = touched_cnt[b];                        // documents with a non-zero score (all ranges)
  const int keep = min(top_k, kBmBuf / 2);
  block_select_docs(sc, tl, n_cand, keep, s_sc, s_doc);
  const int n_hit = min(T, top_k);
  double* os = out_scores + (size_t)b * top_k;
  int64_t* oi = out_ids + (size_t)b * top_k;
  for (int r = threadIdx.x; r < n_hit; r += blockDim.x) { os[r] = s_sc[r]; oi[r] = s_doc[r]; }
  // fewer hits than top_k: the ranking continues with zero-score documents in ascending index
  // order.  Among the first top_k + T indices at least top_k are untouched.
  const int need = (int)min((int64_t)(top_k - n_hit), n_docs - T);
  int filled = 0;
  for (int64_t base = 0; need > 0 && filled < need && base < n_docs; base += blockDim.x) {
    const int64_t d = base + threadIdx.x;
    const int z = (d < n_docs && sc[d] == 0.0) ? 1 : 0;
    int total = 0;
    const int rank = filled + block_rank_of_flag(z, s_wsum, &total);
    if (z && rank < need) { os[n_hit + rank] = 0.0; oi[n_hit + rank] = d; }
    filled += total;
  }
  const int got = n_hit + (need > 0 ? need : 0);
  for (int r = got + threadIdx.x; r < top_k; r += blockDim.x) { os[r] = 0.0; oi[r] = -1; }   // top_k > N
}

struct FuseParams {
  int method;            // 0 = rrf, 1 = weighted
  int rrf_k;
  double w_dense, w_bm25;
  const int64_t* dense_ids; const float* dense_scores; const int32_t* n_dense; int k1;
  const int64_t* bm_ids; const double* bm_scores; const int32_t* n_bm; int k2;
  int top_k;
  int64_t* out_ids; double* out_scores; int32_t* out_n;
};

constexpr int kFuseCap = 2048;   // k1 + k2 entries at most

__global__ void __launch_bounds__(kBmThreads) fuse_kernel(const FuseParams p) {
  __shared__ double f_sc[kFuseCap];
  __shared__ int f_ord[kFuseCap];          // insertion order == index into f_id
  __shared__ long long f_id[kFuseCap];
  __shared__ double s_mx[2];
  const int b = blockIdx.x;
  const int nd = min(max(p.n_dense ? p.n_dense[b] : p.k1, 0), p.k1);
  const int nb = min(max(p.n_bm ? p.n_bm[b] : p.k2, 0), p.k2);
  const int64_t* did = p.dense_ids + (size_t)b * p.k1;
  const float* dsc = p.dense_scores + (size_t)b * p.k1;
  const int64_t* bid = p.bm_ids + (size_t)b * p.k2;
  const double* bsc = p.bm_scores + (size_t)b * p.k2;
  if (threadIdx.x == 0) {
    // max() of each list (weighted fusion normalises by it, :349-350,:356-357)
    double md = nd ? (double)dsc[0] : 0.0, mb = nb ? bsc[0] : 0.0;
    for (int i = 1; i < nd; ++i) md = fmax(md, (double)dsc[i]);
    for (int i = 1; i < nb; ++i) mb = fmax(mb, bsc[i]);
    s_mx[0] = md; s_mx[1] = mb;
  }
  __syncthreads();
  // dense entries take the first slots in rank order (dict insertion order)
  for (int i = threadIdx.x; i < nd; i += blockDim.x) {
    const double c = p.method == 0 ? ddiv(1.0, (double)(p.rrf_k + i + 1)) : dmul(p.w_dense, ddiv((double)dsc[i], s_mx[0]));
    f_id[i] = did[i]; f_sc[i] = dadd(0.0, c); f_ord[i] = i;
  }
  __syncthreads();
  // a BM25 entry adds to its dense twin if there is one, else it becomes a new key; new keys keep
  // BM25 rank order (dict insertion order).  One thread per BM25 entry looks its twin up; a block
  // scan over the "new key" flags gives every new key its slot.
  __shared__ int s_wsum[kBmThreads / 32];
  int n = nd;
  for (int j0 = 0; j0 < nb; j0 += blockDim.x) {
    const int j = j0 + threadIdx.x;
    int twin = -1;
    long long id = 0;
    if (j < nb) {
      id = bid[j];
      for (int i = 0; i < nd; ++i)
        if (f_id[i] == id) { twin = i; break; }         // ids inside one list are unique
    }
    int total = 0;
    const int is_new = (j < nb && twin < 0) ? 1 : 0;
    const int slot = n + block_rank_of_flag(is_new, s_wsum, &total);
    if (j < nb) {
      const double c = p.method == 0 ? ddiv(1.0, (double)(p.rrf_k + j + 1)) : dmul(p.w_bm25, ddiv(bsc[j], s_mx[1]));
      if (twin >= 0) f_sc[twin] = dadd(f_sc[twin], c);  // at most one BM25 entry per dense twin
      else { f_id[slot] = id; f_sc[slot] = dadd(0.0, c); f_ord[slot] = slot; }
    }
    n += total;
    __syncthreads();
  }
  // stable descending sort over insertion order, then truncate
  int np2 = 2;
  while (np2 < n) np2 <<= 1;
  for (int i = n + threadIdx.x; i < np2; i += blockDim.x) { f_sc[i] = kWorst; f_ord[i] = kWorstKey; }
  __syncthreads();
  block_sort_pairs(f_sc, f_ord, np2);
  const int m = min(n, p.top_k);
  for (int r = threadIdx.x; r < p.top_k; r += blockDim.x) {
    const size_t o = (size_t)b * p.top_k + r;
    if (r < m) { p.out_ids[o] = f_id[f_ord[r]]; p.out_scores[o] = f_sc[r]; }
    else { p.out_ids[o] = -1; p.out_scores[o] = 0.0; }
  }
  if (threadIdx.x == 0) p.out_n[b] = m;
}

}  // namespace
}  // namespace ts

using namespace ts;

struct ts_bm25 {
  int device, sm_count;
  int64_t n_docs, n_terms, nnz;
  int64_t* term_off;
  int32_t* post_doc;
  double* post_w;
  int64_t* term_off_host;   // host copy: per-query bound of the documents a query can touch
  // per-call scratch (grow-only)
  void* scratch; size_t scratch_b;
  int64_t launches;
};

extern "C" {

int ts_bm25_create(ts_bm25** out, int device, int64_t n_docs, int64_t n_terms, const int64_t* term_off_host,
                   const int32_t* post_doc_host, const double* post_w_host) {
  if (!out || n_docs < 0 || n_terms < 0 || !term_off_host || n_docs > 0x7fffffffll) { set_error("ts_bm25_create: invalid argument"); return TS_ERR_INVALID; }
  const int64_t nnz = term_off_host[n_terms];
  if (term_off_host[0] != 0 || nnz < 0 || (nnz > 0 && (!post_doc_host || !post_w_host))) { set_error("ts_bm25_create: bad postings"); return TS_ERR_INVALID; }
  for (int64_t t = 0; t < n_terms; ++t)
    if (term_off_host[t + 1] < term_off_host[t]) { set_error("ts_bm25_create: term offsets must not decrease"); return TS_ERR_INVALID; }
  for (int64_t i = 0; i < nnz; ++i) {
    if (post_doc_host[i] < 0 || post_doc_host[i] >= n_docs) { set_error("ts_bm25_create: posting %lld names document %d of %lld", (long long)i, post_doc_host[i], (long long)n_docs); return TS_ERR_INVALID; }
    if (!(post_w_host[i] > 0.0)) { set_error("ts_bm25_create: posting weights must be > 0 (stale-fit idf <= 0: use the host search)"); return TS_ERR_UNSUPPORTED; }
  }
  DeviceInfo info;
  int rc = check_device(device, &info);
  if (rc) return rc;
  ts_bm25* h = new ts_bm25();
  memset(h, 0, sizeof(*h));
  h->device = device; h->sm_count = info.sm_count; h->n_docs = n_docs; h->n_terms = n_terms; h->nnz = nnz;
  size_t sizes[3] = {(size_t)(n_terms + 1) * 8, (size_t)nnz * 4, (size_t)nnz * 8};
  void** ptrs[3] = {(void**)&h->term_off, (void**)&h->post_doc, (void**)&h->post_w};
  const void* srcs[3] = {term_off_host, post_doc_host, post_w_host};
  for (int i = 0; i < 3; ++i) {
    if (sizes[i] == 0) continue;
    if (cudaMalloc(ptrs[i], sizes[i]) != cudaSuccess) { cudaGetLastError(); set_error("ts_bm25_create: cudaMalloc(%zu) failed", sizes[i]); ts_bm25_destroy(h); return TS_ERR_NOMEM; }
    if (cudaMemcpy(*ptrs[i], srcs[i], sizes[i], cudaMemcpyHostToDevice) != cudaSuccess) { set_error("ts_bm25_create: upload failed"); ts_bm25_destroy(h); return TS_ERR_CUDA; }
  }
  h->term_off_host = new int64_t[(size_t)n_terms + 1];
  memcpy(h->term_off_host, term_off_host, (size_t)(n_terms + 1) * 8);
  *out = h;
  return TS_OK;
}

int ts_bm25_destroy(ts_bm25* h) {
  if (!h) return TS_OK;
  cudaSetDevice(h->device);
  delete[] h->term_off_host;
  void* ptrs[] = {h->term_off, h->post_doc, h->post_w, h->scratch};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h;
  return TS_OK;
}

int64_t ts_bm25_ndocs(const ts_bm25* h) { return h ? h->n_docs : -1; }
int64_t ts_bm25_launch_count(const ts_bm25* h) { return h ? h->launches : -1; }

int ts_bm25_search_host(ts_bm25* h, const int32_t* q_terms_host, const int64_t* q_off_host, int B, int top_k,
                        double* out_scores_host, int64_t* out_ids_host, void* stream) {
  if (!h || !q_off_host || !out_scores_host || !out_ids_host || B <= 0 || top_k <= 0 || top_k > TS_BM25_MAX_K) {
    set_error("ts_bm25_search_host: invalid argument (1 <= top_k <= %d)", TS_BM25_MAX_K);
    return TS_ERR_INVALID;
  }
  const int64_t nq = q_off_host[B];
  if (q_off_host[0] != 0 || nq < 0 || (nq > 0 && !q_terms_host)) { set_error("ts_bm25_search_host: bad query offsets"); return TS_ERR_INVALID; }
  // upper bound of the documents one query can touch: sum of its terms' document frequencies
  const int64_t* toff = h->term_off_host;
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  int64_t cap = 1;
  for (int b = 0; b < B; ++b) {
    if (q_off_host[b + 1] < q_off_host[b]) { set_error("ts_bm25_search_host: query offsets must not decrease"); return TS_ERR_INVALID; }
    int64_t sum = 0;
    for (int64_t i = q_off_host[b]; i < q_off_host[b + 1]; ++i) {
      const int t = q_terms_host[i];
      if (t < 0 || t >= h->n_terms) { set_error("ts_bm25_search_host: term id %d outside 0..%lld", t, (long long)h->n_terms - 1); return TS_ERR_INVALID; }
      sum += toff[(size_t)t + 1] - toff[(size_t)t];
    }
    if (sum > h->n_docs) sum = h->n_docs;
    if (sum > cap) cap = sum;
  }
  // enough CTAs to fill the GPU: split every query's documents into R contiguous ranges
  int R = (4 * h->sm_count + B - 1) / B;
  if (R > 64) R = 64;
  if ((int64_t)R * 4096 > h->n_docs) R = (int)(h->n_docs / 4096);
  if (const char* e = getenv("TS_BM25_SPLIT")) { if (atoi(e) > 0) R = atoi(e); }      // tests force a split on tiny corpora
  if (R < 1) R = 1;
  const int64_t span = (h->n_docs + R - 1) / R;
  const int64_t seg_cap = cap < span ? cap : span;     // a range cannot touch more documents than it holds
  const int64_t cand_cap = (int64_t)R * (top_k < kBmBuf / 2 ? top_k : kBmBuf / 2);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t b_scores = up((size_t)B * h->n_docs * 8), b_touched = up((size_t)B * R * seg_cap * 4), b_cnt = up((size_t)B * 8);
  const size_t b_cand = up((size_t)B * cand_cap * 4);
  const size_t b_qt = up((size_t)(nq > 0 ? nq : 1) * 4), b_qo = up((size_t)(B + 1) * 8);
  const size_t b_os = up((size_t)B * top_k * 8), b_oi = up((size_t)B * top_k * 8);
  int rc = ensure_bytes(&h->scratch, &h->scratch_b, b_scores + b_touched + b_cnt + b_cand + b_qt + b_qo + b_os + b_oi);
  if (rc) return rc;
  char* base = (char*)h->scratch;
  double* d_scores = (double*)base; base += b_scores;
  int32_t* d_touched = (int32_t*)base; base += b_touched;
  int32_t* d_cnt = (int32_t*)base; base += b_cnt;        // [B] touched totals, then [B] candidate counts
  int32_t* d_cand = (int32_t*)base; base += b_cand;
  int32_t* d_qt = (int32_t*)base; base += b_qt;
  int64_t* d_qo = (int64_t*)base; base += b_qo;
  double* d_os = (double*)base; base += b_os;
  int64_t* d_oi = (int64_t*)base;
  TS_CUDA_OK(cudaMemsetAsync(d_scores, 0, (size_t)B * h->n_docs * 8, st));
  TS_CUDA_OK(cudaMemsetAsync(d_cnt, 0, (size_t)B * 8, st));
  if (nq > 0) TS_CUDA_OK(cudaMemcpyAsync(d_qt, q_terms_host, (size_t)nq * 4, cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaMemcpyAsync(d_qo, q_off_host, (size_t)(B + 1) * 8, cudaMemcpyHostToDevice, st));
  TS_LAUNCH(bm25_score_kernel, dim3(B, R), kBmThreads, 0, st, h->term_off, h->post_doc, h->post_w, d_qt, d_qo, h->n_docs, span,
            seg_cap, top_k, d_scores, d_touched, d_cnt, d_cand, cand_cap, d_cnt + B);
  TS_CUDA_OK(cudaGetLastError());
  TS_LAUNCH(bm25_topk_kernel, B, kBmThreads, 0, st, d_scores, h->n_docs, d_cand, cand_cap, d_cnt + B, d_cnt, top_k, d_os, d_oi);
  TS_CUDA_OK(cudaGetLastError());
  h->launches += 2;
  TS_CUDA_OK(cudaMemcpyAsync(out_scores_host, d_os, (size_t)B * top_k * 8, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaMemcpyAsync(out_ids_host, d_oi, (size_t)B * top_k * 8, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  return TS_OK;
}

int ts_hybrid_fuse_host(int device, int method, int rrf_k, double w_dense, double w_bm25, const int64_t* dense_ids_host,
                        const float* dense_scores_host, const int32_t* n_dense_host, int k1, const int64_t* bm25_ids_host,
                        const double* bm25_scores_host, const int32_t* n_bm25_host, int k2, int B, int top_k,
                        int64_t* out_ids_host, double* out_scores_host, int32_t* out_n_host, void* stream) {
  if (B <= 0 || k1 < 0 || k2 < 0 || k1 + k2 <= 0 || k1 + k2 > kFuseCap || top_k <= 0 || top_k > kFuseCap || (method != 0 && method != 1) ||
      !out_ids_host || !out_scores_host || !out_n_host || (k1 > 0 && (!dense_ids_host || !dense_scores_host)) ||
      (k2 > 0 && (!bm25_ids_host || !bm25_scores_host))) {
    set_error("ts_hybrid_fuse_host: invalid argument (k1 + k2 <= %d)", kFuseCap);
    return TS_ERR_INVALID;
  }
  DeviceInfo info;
  int rc = check_device(device, &info);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t sz[] = {up((size_t)B * k1 * 8 + 8), up((size_t)B * k1 * 4 + 8), up((size_t)B * 4), up((size_t)B * k2 * 8 + 8),
                       up((size_t)B * k2 * 8 + 8), up((size_t)B * 4), up((size_t)B * top_k * 8), up((size_t)B * top_k * 8), up((size_t)B * 4)};
  size_t total = 0;
  for (size_t s : sz) total += s;
  char* buf = nullptr;
  if (cudaMalloc((void**)&buf, total) != cudaSuccess) { cudaGetLastError(); set_error("ts_hybrid_fuse_host: cudaMalloc(%zu) failed", total); return TS_ERR_NOMEM; }
  char* at = buf;
  void* d[9];
  for (int i = 0; i < 9; ++i) { d[i] = at; at += sz[i]; }
  auto fail = [&](int code) { cudaFree(buf); return code; };
  const void* src[6] = {dense_ids_host, dense_scores_host, n_dense_host, bm25_ids_host, bm25_scores_host, n_bm25_host};
  const size_t nb[6] = {(size_t)B * k1 * 8, (size_t)B * k1 * 4, (size_t)B * 4, (size_t)B * k2 * 8, (size_t)B * k2 * 8, (size_t)B * 4};
  for (int i = 0; i < 6; ++i)
    if (src[i] && nb[i] && cudaMemcpyAsync(d[i], src[i], nb[i], cudaMemcpyHostToDevice, st) != cudaSuccess) { set_error("ts_hybrid_fuse_host: upload failed"); return fail(TS_ERR_CUDA); }
  FuseParams p{};
  p.method = method; p.rrf_k = rrf_k; p.w_dense = w_dense; p.w_bm25 = w_bm25;
  p.dense_ids = (const int64_t*)d[0]; p.dense_scores = (const float*)d[1]; p.n_dense = n_dense_host ? (const int32_t*)d[2] : nullptr; p.k1 = k1;
  p.bm_ids = (const int64_t*)d[3]; p.bm_scores = (const double*)d[4]; p.n_bm = n_bm25_host ? (const int32_t*)d[5] : nullptr; p.k2 = k2;
  p.top_k = top_k; p.out_ids = (int64_t*)d[6]; p.out_scores = (double*)d[7]; p.out_n = (int32_t*)d[8];
  TS_LAUNCH(fuse_kernel, B, kBmThreads, 0, st, p);
  if (cudaGetLastError() != cudaSuccess) { set_error("fuse_kernel launch failed"); return fail(TS_ERR_CUDA); }
  if (cudaMemcpyAsync(out_ids_host, d[6], (size_t)B * top_k * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaMemcpyAsync(out_scores_host, d[7], (size_t)B * top_k * 8, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaMemcpyAsync(out_n_host, d[8], (size_t)B * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
      cudaStreamSynchronize(st) != cudaSuccess) { set_error("ts_hybrid_fuse_host: download failed"); return fail(TS_ERR_CUDA); }
  cudaFree(buf);
  return TS_OK;
}

}  // extern "C"
