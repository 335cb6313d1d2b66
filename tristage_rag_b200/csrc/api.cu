// C ABI of libtristage.so (see include/tristage.h for the contract and the
// reference call sites each entry point replaces).
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <vector>

#include "ts_handles.h"

namespace ts {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

int get_device_info(int device, DeviceInfo* out) {
  cudaDeviceProp prop;
  TS_CUDA_OK(cudaGetDeviceProperties(&prop, device));
  out->sm_count = prop.multiProcessorCount;
  out->cc_major = prop.major;
  out->cc_minor = prop.minor;
  out->smem_optin = prop.sharedMemPerBlockOptin;
  return TS_OK;
}

// grow-only device scratch
int ensure_bytes(void** p, size_t* cur, size_t need) {
  if (need <= *cur) return TS_OK;
  if (*p) { cudaFree(*p); *p = nullptr; *cur = 0; }
  cudaError_t e = cudaMalloc(p, need);
  if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc(%zu) failed: %s", need, cudaGetErrorString(e)); return TS_ERR_NOMEM; }
  *cur = need;
  return TS_OK;
}

}  // namespace ts

using namespace ts;

namespace ts {

int check_device(int device, DeviceInfo* info) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    set_error("no CUDA device available (%s); libtristage has no CPU fallback", e == cudaSuccess ? "count=0" : cudaGetErrorString(e));
    return TS_ERR_CUDA;
  }
  if (device < 0 || device >= n) { set_error("device %d out of range (%d visible)", device, n); return TS_ERR_INVALID; }
  int rc = get_device_info(device, info);
  if (rc) return rc;
  if (info->cc_major != 10) {
    set_error("device %d is sm_%d%d; libtristage is built for sm_100a only", device, info->cc_major, info->cc_minor);
    return TS_ERR_UNSUPPORTED;
  }
  TS_CUDA_OK(cudaSetDevice(device));
  return TS_OK;
}

int index_reserve(ts_index* h, int64_t rows, cudaStream_t st) {
  if (rows <= h->cap) return TS_OK;
  int64_t ncap = h->cap > 0 ? h->cap : 1024;
  while (ncap < rows) ncap *= 2;
  const size_t row_b = (size_t)h->ld * dtype_size(h->dtype);
  void* nrows = nullptr;
  cudaError_t e = cudaMalloc(&nrows, (size_t)ncap * row_b);
  if (e != cudaSuccess) {
    // retry with the exact size before giving up (doubling can overshoot HBM)
    cudaGetLastError();   // a failed cudaMalloc must not surface at the next launch check
    ncap = rows;
    e = cudaMalloc(&nrows, (size_t)ncap * row_b);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of %lld corpus rows failed: %s", (long long)ncap, cudaGetErrorString(e)); return TS_ERR_NOMEM; }
  }
  float* ninv = nullptr;
  if (h->metric == TS_METRIC_COSINE) {
    e = cudaMalloc((void**)&ninv, (size_t)ncap * sizeof(float));
    if (e != cudaSuccess) { cudaGetLastError(); cudaFree(nrows); set_error("cudaMalloc inv_norm failed"); return TS_ERR_NOMEM; }
  }
  if (h->n > 0) {
    TS_CUDA_OK(cudaMemcpyAsync(nrows, h->rows, (size_t)h->n * row_b, cudaMemcpyDeviceToDevice, st));
    if (ninv) TS_CUDA_OK(cudaMemcpyAsync(ninv, h->inv_norm, (size_t)h->n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TS_CUDA_OK(cudaStreamSynchronize(st));
  }
  if (h->rows) cudaFree(h->rows);
  if (h->inv_norm) cudaFree(h->inv_norm);
  h->rows = nrows; h->inv_norm = ninv; h->cap = ncap;
  return TS_OK;
}

}  // namespace ts

namespace {
bool valid_storage(int dt) { return dt == TS_F32 || dt == TS_BF16 || dt == TS_F16; }
}  // namespace

extern "C" {

int ts_abi_version(void) { return TS_ABI_VERSION; }
const char* ts_last_error(void) { return ts::get_error(); }

int ts_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, d) == cudaSuccess && p.major == 10) ++ok;
  }
  return ok;
}

// ------------------------------------------------------------------ Stage 1
int ts_index_create(ts_index** out, int device, int dim, int storage_dtype, int metric, int64_t reserve_rows) {
  if (!out || dim <= 0 || dim > 8192 || !valid_storage(storage_dtype) ||
      (metric != TS_METRIC_IP && metric != TS_METRIC_COSINE) || reserve_rows < 0) {
    set_error("ts_index_create: invalid argument");
    return TS_ERR_INVALID;
  }
  DeviceInfo info;
  int rc = check_device(device, &info);
  if (rc) return rc;
  ts_index* h = new ts_index();
  memset(h, 0, sizeof(*h));
  h->device = device; h->dim = dim; h->dtype = storage_dtype; h->metric = metric;
  h->ld = row_pitch(dim, storage_dtype);
  h->info = info;
  h->timer = new ScanTimer();
  {
    int coop = 0;
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device);
    h->coop = coop;
    // [0] arrival counter of the fused scan's grid barrier, [1 + mt] next-tile counter of query tile mt: all 16 words
    // are zeroed by the query-prep kernel of every search
    // (+ 192 bytes behind them: the TS_DBG_TIMELINE slots, ts_index_debug_timeline)
    if (cudaMalloc((void**)&h->grid_bar, 256) == cudaSuccess) cudaMemset(h->grid_bar, 0, 256);
    else { cudaGetLastError(); h->grid_bar = nullptr; }
  }
  if (reserve_rows > 0) {
    h->cap = 0;
    // exact reservation (no doubling) for the first allocation
    const size_t row_b = (size_t)h->ld * dtype_size(h->dtype);
    cudaError_t e = cudaMalloc(&h->rows, (size_t)reserve_rows * row_b);
    if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of %lld rows failed: %s", (long long)reserve_rows, cudaGetErrorString(e)); h->rows = nullptr; ts_index_destroy(h); return TS_ERR_NOMEM; }
    if (metric == TS_METRIC_COSINE) {
      e = cudaMalloc((void**)&h->inv_norm, (size_t)reserve_rows * sizeof(float));
      if (e != cudaSuccess) { cudaGetLastError(); h->inv_norm = nullptr; set_error("cudaMalloc inv_norm failed"); ts_index_destroy(h); return TS_ERR_NOMEM; }
    }
    h->cap = reserve_rows;
  }
  *out = h;
  return TS_OK;
}

int ts_index_destroy(ts_index* h) {
  if (!h) return TS_OK;
  cudaSetDevice(h->device);
  void* ptrs[] = {h->rows, h->inv_norm, h->qbuf, h->lists, h->partial, h->tmp0, h->tmp1, h->counts, h->pub, h->stage, h->hout, h->grid_bar};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->timer) { h->timer->destroy(); delete h->timer; }
  delete h;
  return TS_OK;
}

int ts_index_set_profiling(ts_index* h, int enable) { if (!h) return TS_ERR_INVALID; h->timer->on = enable != 0; h->timer->n = 0; return TS_OK; }
int ts_index_scan_time(ts_index* h, float* mean_ms, int* n) { if (!h) return TS_ERR_INVALID; cudaSetDevice(h->device); return h->timer->report(mean_ms, n); }

int ts_index_add(ts_index* h, const void* rows, int64_t n, int src_dtype, int src_on_device, int normalize, void* stream) {
  if (!h || n < 0 || (n > 0 && !rows)) { set_error("ts_index_add: invalid argument"); return TS_ERR_INVALID; }
  if (src_dtype != TS_F32 && src_dtype != h->dtype) { set_error("ts_index_add: src dtype must be f32 or the storage dtype"); return TS_ERR_INVALID; }
  if (n == 0) return TS_OK;
  if (h->n + n > 0xFFFFFFF0ll) { set_error("ts_index_add: shard limited to 2^32 rows"); return TS_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  int rc = index_reserve(h, h->n + n, st);
  if (rc) return rc;
  const int ssz = dtype_size(src_dtype);
  const size_t row_b = (size_t)h->ld * dtype_size(h->dtype);
  const int norm_mode = (h->metric == TS_METRIC_COSINE || normalize) ? kNormStage1 : kNormNone;
  const int64_t chunk = src_on_device ? n : (int64_t)((256ull << 20) / ((size_t)h->dim * ssz) + 1);
  for (int64_t s = 0; s < n; s += chunk) {
    const int64_t m = (n - s) < chunk ? (n - s) : chunk;
    const char* src = (const char*)rows + (size_t)s * h->dim * ssz;
    const void* dsrc = src;
    if (!src_on_device) {
      rc = ensure_bytes(&h->stage, &h->stage_b, (size_t)m * h->dim * ssz);
      if (rc) return rc;
      TS_CUDA_OK(cudaMemcpyAsync(h->stage, src, (size_t)m * h->dim * ssz, cudaMemcpyHostToDevice, st));
      dsrc = h->stage;
    }
    float* inv = (h->metric == TS_METRIC_COSINE) ? h->inv_norm + h->n + s : nullptr;
    rc = launch_convert_rows(dsrc, src_dtype, h->dim, (char*)h->rows + (size_t)(h->n + s) * row_b, h->dtype, h->ld, m,
                             h->dim, norm_mode, inv, st);
    if (rc) return rc;
    ++h->launches;
    if (!src_on_device) TS_CUDA_OK(cudaStreamSynchronize(st));  // staging buffer is reused
  }
  h->n += n;
  return TS_OK;
}

int64_t ts_index_ntotal(const ts_index* h) { return h ? h->n : -1; }
int ts_index_dim(const ts_index* h) { return h ? h->dim : -1; }
int ts_index_dtype(const ts_index* h) { return h ? h->dtype : -1; }
int ts_index_metric(const ts_index* h) { return h ? h->metric : -1; }
int ts_index_reset(ts_index* h) { if (!h) return TS_ERR_INVALID; h->n = 0; ++h->reset_gen; return TS_OK; }
int ts_index_set_id_base(ts_index* h, int64_t b) { if (!h) return TS_ERR_INVALID; h->id_base = b; return TS_OK; }
int64_t ts_index_launch_count(const ts_index* h) { return h ? h->launches : -1; }

}  // extern "C"

// fused multi-GPU exchange state (ts_exchange_*): the receive buffers are allocated and mapped by the caller
struct ts_exchange {
  int device, rank, n_ranks, B_max, k_max;
  long long* peer_bases_dev;     // [n_ranks]
  char* local_base;              // this rank's own buffer
  int64_t slot_bytes, ids_in_slot, flags_off;
  uint64_t step;                 // calls made so far: parity = step & 1, seq = step + 1
};

namespace ts {
static int index_search_impl(ts_index* h, const void* q_dev, int q_dtype, int B, int k, unsigned flags, int path,
                             float* out_scores, int64_t* out_ids, void* stream, const PushTarget* push);
}

extern "C" {

int ts_index_search(ts_index* h, const void* q_dev, int q_dtype, int B, int k, unsigned flags, int path,
                    float* out_scores, int64_t* out_ids, void* stream) {
  if (!out_scores || !out_ids) { set_error("ts_index_search: invalid argument"); return TS_ERR_INVALID; }
  return ts::index_search_impl(h, q_dev, q_dtype, B, k, flags, path, out_scores, out_ids, stream, nullptr);
}

}  // extern "C"

namespace ts {
static int index_search_impl(ts_index* h, const void* q_dev, int q_dtype, int B, int k, unsigned flags, int path,
                             float* out_scores, int64_t* out_ids, void* stream, const PushTarget* push) {
  if (!h || !q_dev || B <= 0) { set_error("ts_index_search: invalid argument"); return TS_ERR_INVALID; }
  if (push && B > 1024) { set_error("ts_index_search: the fused exchange takes at most 1024 queries per call"); return TS_ERR_UNSUPPORTED; }
  if (k <= 0 || k > TS_MAX_K) { set_error("ts_index_search: k=%d outside 1..%d", k, TS_MAX_K); return TS_ERR_INVALID; }
  if (q_dtype != TS_F32 && q_dtype != h->dtype) { set_error("ts_index_search: query dtype must be f32 or the storage dtype"); return TS_ERR_INVALID; }
  if (h->n == 0) { set_error("No documents indexed. Call add_documents() first."); return TS_ERR_EMPTY; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  int use = path;
  // measured on B200: the TMA/tcgen05 scan streams faster than the CUDA-core scan at every batch size
  // fp32 storage: the CUDA-core scan (exact fp32 products) unless the tf32 tensor path is switched on (TS_TF32) and
  // the batch needs more than one CUDA-core pass (4 queries per pass)
  if (use == TS_PATH_AUTO) use = (h->dtype != TS_F32 || (B > 4 && env_flag("TS_TF32", kDefaultTf32))) ? TS_PATH_UMMA : TS_PATH_STREAM;
  if (use != TS_PATH_STREAM && use != TS_PATH_UMMA) { set_error("ts_index_search: bad path %d", path); return TS_ERR_INVALID; }

  const int esz = dtype_size(h->dtype);
  const int chunkB = 1024;
  unsigned long long* tl = (h->grid_bar && env_on("TS_DBG_TIMELINE")) ? reinterpret_cast<unsigned long long*>(h->grid_bar + 16) : nullptr;
  for (int b0 = 0; b0 < B; b0 += chunkB) {
    const int Bc = (B - b0) < chunkB ? (B - b0) : chunkB;
    int rc = ensure_bytes(&h->qbuf, &h->qbuf_b, (size_t)chunkB * h->ld * esz);
    if (rc) return rc;
    rc = launch_convert_rows((const char*)q_dev + (size_t)b0 * h->dim * dtype_size(q_dtype), q_dtype, h->dim, h->qbuf,
                             h->dtype, h->ld, Bc, h->dim, (flags & TS_FLAG_NORMALIZE_Q) ? kNormStage1 : kNormNone,
                             nullptr, st, h->grid_bar, tl);
    if (rc) return rc;
    ++h->launches;
    ScanArgs a{};
    a.rows = h->rows; a.n = h->n; a.dim = h->dim; a.ld = h->ld; a.dtype = h->dtype;
    a.inv_norm = (h->metric == TS_METRIC_COSINE) ? h->inv_norm : nullptr;
    a.q = h->qbuf; a.B = Bc; a.k = k; a.sm_count = h->info.sm_count;
    int launches = 0;
    if (use == TS_PATH_STREAM) {
      int L = 0; size_t lists_keys = 0;
      if ((rc = s1_stream_plan(a, &L, &lists_keys))) return rc;
      if ((rc = ensure_bytes(&h->lists, &h->lists_b, lists_keys * 8))) return rc;
      const size_t partial_keys = (size_t)L * Bc * k;
      if ((rc = ensure_bytes(&h->partial, &h->partial_b, partial_keys * 8))) return rc;
      const size_t tmpk = merge_tmp_keys(L, Bc, k);
      if (tmpk) {
        if ((rc = ensure_bytes(&h->tmp0, &h->tmp0_b, tmpk * 8))) return rc;
        if ((rc = ensure_bytes(&h->tmp1, &h->tmp1_b, tmpk * 8))) return rc;
      }
      a.lists = (uint64_t*)h->lists; a.lists_keys = lists_keys;
      a.partial = (uint64_t*)h->partial; a.partial_keys = partial_keys;
      h->timer->begin(st);
      rc = launch_s1_stream(a, st, &launches);
      h->timer->end(st);
      if (rc) return rc;
      rc = launch_merge_keys((const uint64_t*)h->partial, L, Bc, k, h->id_base, (uint64_t*)h->tmp0, (uint64_t*)h->tmp1,
                             out_scores ? out_scores + (size_t)b0 * k : nullptr, out_ids ? out_ids + (size_t)b0 * k : nullptr, st,
                             &launches, push);
      if (rc) return rc;
    } else {
      UmmaLayout lay{};
      if ((rc = s1_umma_plan(a, &lay))) return rc;
      if ((rc = ensure_bytes(&h->lists, &h->lists_b, lay.lists_keys * 8))) return rc;
      if ((rc = ensure_bytes(&h->counts, &h->counts_b, lay.counts_n * sizeof(int)))) return rc;
      if ((rc = ensure_bytes(&h->pub, &h->pub_b, lay.pub_n * sizeof(float)))) return rc;
      a.lists = (uint64_t*)h->lists; a.lists_keys = lay.lists_keys;
      a.counts = (int*)h->counts; a.pub = (float*)h->pub; a.grid_bar = h->grid_bar; a.coop = h->coop; a.tl = tl;
      h->timer->begin(st);
      rc = launch_s1_umma(a, lay, st, &launches);
      h->timer->end(st);
      if (rc) return rc;
      lay.tl = tl;
      rc = launch_merge_lists((const uint64_t*)h->lists, (const int*)h->counts, (const float*)h->pub, lay, Bc, k, h->id_base,
                              out_scores ? out_scores + (size_t)b0 * k : nullptr, out_ids ? out_ids + (size_t)b0 * k : nullptr, st,
                              &launches, push);
      if (rc) return rc;
    }
    h->launches += launches;
  }
  return TS_OK;
}
}  // namespace ts

extern "C" {

// ---- fused multi-GPU exchange -------------------------------------------------
static int64_t xchg_slot_bytes(int B_max, int k_max, int64_t* ids_in_slot) {
  const int64_t sc = ((int64_t)B_max * k_max * 4 + 15) / 16 * 16;
  if (ids_in_slot) *ids_in_slot = sc;
  return sc + (int64_t)B_max * k_max * 8;
}

int64_t ts_exchange_buffer_bytes(int n_ranks, int B_max, int k_max) {
  if (n_ranks < 1 || B_max < 1 || k_max < 1 || k_max > TS_MAX_K) return -1;
  const int64_t slot = xchg_slot_bytes(B_max, k_max, nullptr);
  return 2 * n_ranks * slot + 2 * (int64_t)n_ranks * B_max * 4;
}

int ts_exchange_create(ts_exchange** out, int device, int rank, int n_ranks, const int64_t* peer_bases_host, int B_max, int k_max) {
  if (!out || !peer_bases_host || n_ranks < 1 || n_ranks > 256 || rank < 0 || rank >= n_ranks || B_max < 1 || B_max > 1024 || k_max < 1 || k_max > TS_MAX_K) {
    set_error("ts_exchange_create: invalid argument");
    return TS_ERR_INVALID;
  }
  for (int r = 0; r < n_ranks; ++r)
    if (!peer_bases_host[r] || (peer_bases_host[r] & 15)) { set_error("ts_exchange_create: buffer %d is null or not 16-byte aligned", r); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(device));
  ts_exchange* x = new ts_exchange();
  memset(x, 0, sizeof(*x));
  x->device = device; x->rank = rank; x->n_ranks = n_ranks; x->B_max = B_max; x->k_max = k_max;
  x->slot_bytes = xchg_slot_bytes(B_max, k_max, &x->ids_in_slot);
  x->flags_off = 2 * n_ranks * x->slot_bytes;
  x->local_base = reinterpret_cast<char*>(peer_bases_host[rank]);
  if (cudaMalloc((void**)&x->peer_bases_dev, (size_t)n_ranks * 8) != cudaSuccess) { cudaGetLastError(); delete x; set_error("ts_exchange_create: cudaMalloc failed"); return TS_ERR_NOMEM; }
  if (cudaMemcpy(x->peer_bases_dev, peer_bases_host, (size_t)n_ranks * 8, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaGetLastError(); cudaFree(x->peer_bases_dev); delete x; set_error("ts_exchange_create: copy failed"); return TS_ERR_CUDA;
  }
  *out = x;
  return TS_OK;
}

int ts_exchange_destroy(ts_exchange* x) {
  if (!x) return TS_OK;
  cudaSetDevice(x->device);
  if (x->peer_bases_dev) cudaFree(x->peer_bases_dev);
  delete x;
  return TS_OK;
}

int ts_index_search_push(ts_index* h, ts_exchange* x, const void* q_dev, int q_dtype, int B, int k, unsigned flags, int path,
                         void* stream) {
  if (!h || !x || B < 1 || B > x->B_max || k < 1 || k > x->k_max) { set_error("ts_index_search_push: B / k outside the exchange's capacity"); return TS_ERR_INVALID; }
  if (x->device != h->device) { set_error("ts_index_search_push: index and exchange live on different devices"); return TS_ERR_INVALID; }
  const int parity = (int)(x->step & 1);
  PushTarget pt{};
  pt.peer_bases = x->peer_bases_dev; pt.n_ranks = x->n_ranks;
  pt.scores_off = ((long long)parity * x->n_ranks + x->rank) * x->slot_bytes;
  pt.ids_off = pt.scores_off + x->ids_in_slot;
  pt.flags_off = x->flags_off + ((long long)parity * x->n_ranks + x->rank) * x->B_max * 4;
  pt.seq = (unsigned int)((x->step % 0x7FFFFFFFull) + 1);
  return ts::index_search_impl(h, q_dev, q_dtype, B, k, flags, path, nullptr, nullptr, stream, &pt);
}

int ts_exchange_merge(ts_exchange* x, int B, int k, float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
  if (!x || !out_scores_dev || !out_ids_dev || B < 1 || B > x->B_max || k < 1 || k > x->k_max) { set_error("ts_exchange_merge: invalid argument"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(x->device));
  const int parity = (int)(x->step & 1);
  const unsigned int seq = (unsigned int)((x->step % 0x7FFFFFFFull) + 1);
  ++x->step;                                       // the pair (push, merge) is one step
  const char* slots = x->local_base + (size_t)parity * x->n_ranks * x->slot_bytes;
  const unsigned int* fl = reinterpret_cast<const unsigned int*>(x->local_base + x->flags_off) + (size_t)parity * x->n_ranks * x->B_max;
  return launch_merge_pairs_wait((const float*)slots, (const int64_t*)(slots + x->ids_in_slot), x->slot_bytes / 4, x->slot_bytes / 8,
                                 x->n_ranks, B, k, fl, seq, out_scores_dev, out_ids_dev, (cudaStream_t)stream, x->B_max);
}

int ts_index_search_sharded(ts_index* h, ts_exchange* x, const void* q_dev, int q_dtype, int B, int k, unsigned flags, int path,
                            float* out_scores_dev, int64_t* out_ids_dev, void* stream) {
  if (!h || !x || !out_scores_dev || !out_ids_dev || B < 1 || B > x->B_max || k < 1 || k > x->k_max) { set_error("ts_index_search_sharded: B / k outside the exchange's capacity"); return TS_ERR_INVALID; }
  // One kernel for the whole exchange (select + push + wait + merge) when every CTA of it is co-resident (one per
  // query, B <= SM count) and the G * k keys of a query fit the select kernel's small buffer; TS_XFUSE=0: two kernels.
  // (The emulator runs in-process "ranks" one after the other -- a kernel that waits for its peers cannot run there --
  // so its build fuses only when the ranks are separate emulator PROCESSES sharing the buffers: TS_SIM_XFUSE=1.)
#ifdef TS_CUDASIM
  const bool may_fuse = env_on("TS_SIM_XFUSE");
#else
  const bool may_fuse = true;
#endif
  if (may_fuse && B <= h->info.sm_count && k <= 128 && (long long)x->n_ranks * k <= 2048 && env_flag("TS_XFUSE", true)) {
    if (x->device != h->device) { set_error("ts_index_search_sharded: index and exchange live on different devices"); return TS_ERR_INVALID; }
    const int parity = (int)(x->step & 1);
    PushTarget pt{};
    pt.peer_bases = x->peer_bases_dev; pt.n_ranks = x->n_ranks;
    pt.scores_off = ((long long)parity * x->n_ranks + x->rank) * x->slot_bytes;
    pt.ids_off = pt.scores_off + x->ids_in_slot;
    pt.flags_off = x->flags_off + ((long long)parity * x->n_ranks + x->rank) * x->B_max * 4;
    pt.seq = (unsigned int)((x->step % 0x7FFFFFFFull) + 1);
    const char* slots = x->local_base + (size_t)parity * x->n_ranks * x->slot_bytes;
    pt.merge_scores = (const float*)slots;
    pt.merge_ids = (const int64_t*)(slots + x->ids_in_slot);
    pt.merge_stride_f = x->slot_bytes / 4; pt.merge_stride_i = x->slot_bytes / 8;
    pt.merge_flags = reinterpret_cast<const unsigned int*>(x->local_base + x->flags_off) + (size_t)parity * x->n_ranks * x->B_max;
    pt.merge_flag_stride = x->B_max;
    pt.merge_out_scores = out_scores_dev; pt.merge_out_ids = out_ids_dev;
    ++x->step;
    return ts::index_search_impl(h, q_dev, q_dtype, B, k, flags, path, nullptr, nullptr, stream, &pt);
  }
  int rc = ts_index_search_push(h, x, q_dev, q_dtype, B, k, flags, path, stream);
  if (rc) return rc;
  return ts_exchange_merge(x, B, k, out_scores_dev, out_ids_dev, stream);
}

int ts_index_search_sharded_host(ts_index* h, ts_exchange* x, const void* q_host, int q_dtype, int B, int k, unsigned flags,
                                 int path, float* out_scores_host, int64_t* out_ids_host, void* stream) {
  if (!h || !x || !q_host || !out_scores_host || !out_ids_host || B <= 0 || k <= 0 || k > TS_MAX_K) {
    set_error("ts_index_search_sharded_host: invalid argument");
    return TS_ERR_INVALID;
  }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  const size_t qb = (size_t)B * h->dim * dtype_size(q_dtype);
  int rc = ensure_bytes(&h->stage, &h->stage_b, qb);
  if (rc) return rc;
  const size_t sb = (size_t)B * k * sizeof(float), ib = (size_t)B * k * sizeof(int64_t);
  if ((rc = ensure_bytes(&h->hout, &h->hout_b, sb + ib + 256))) return rc;
  float* ds = (float*)h->hout;
  int64_t* di = (int64_t*)((char*)h->hout + ((sb + 255) / 256) * 256);
  TS_CUDA_OK(cudaMemcpyAsync(h->stage, q_host, qb, cudaMemcpyHostToDevice, st));
  if ((rc = ts_index_search_sharded(h, x, h->stage, q_dtype, B, k, flags, path, ds, di, stream))) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(out_scores_host, ds, sb, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaMemcpyAsync(out_ids_host, di, ib, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  return TS_OK;
}

int ts_index_search_host(ts_index* h, const void* q_host, int q_dtype, int B, int k, unsigned flags, int path,
                         float* out_scores_host, int64_t* out_ids_host, void* stream) {
  if (!h || !q_host || !out_scores_host || !out_ids_host || B <= 0 || k <= 0 || k > TS_MAX_K) {
    set_error("ts_index_search_host: invalid argument");
    return TS_ERR_INVALID;
  }
  if (h->n == 0) { set_error("No documents indexed. Call add_documents() first."); return TS_ERR_EMPTY; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  const size_t qb = (size_t)B * h->dim * dtype_size(q_dtype);
  int rc = ensure_bytes(&h->stage, &h->stage_b, qb);
  if (rc) return rc;
  const size_t sb = (size_t)B * k * sizeof(float), ib = (size_t)B * k * sizeof(int64_t);
  if ((rc = ensure_bytes(&h->hout, &h->hout_b, sb + ib + 256))) return rc;
  float* ds = (float*)h->hout;
  int64_t* di = (int64_t*)((char*)h->hout + ((sb + 255) / 256) * 256);
  TS_CUDA_OK(cudaMemcpyAsync(h->stage, q_host, qb, cudaMemcpyHostToDevice, st));
  rc = ts_index_search(h, h->stage, q_dtype, B, k, flags, path, ds, di, stream);
  if (rc) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(out_scores_host, ds, sb, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaMemcpyAsync(out_ids_host, di, ib, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  return TS_OK;
}

int ts_topk_merge(int device, const float* scores, const int64_t* ids, int n_lists, int B, int k, float* out_scores,
                  int64_t* out_ids, void* stream) {
  if (!scores || !ids || !out_scores || !out_ids) { set_error("ts_topk_merge: null pointer"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(device));
  return launch_merge_pairs(scores, ids, (long long)B * k, (long long)B * k, n_lists, B, k, out_scores, out_ids,
                            (cudaStream_t)stream);
}

int ts_topk_merge_packed(int device, const void* blob, int64_t list_stride_bytes, int64_t ids_offset_bytes, int n_lists,
                         int B, int k, float* out_scores, int64_t* out_ids, void* stream) {
  if (!blob || !out_scores || !out_ids || (list_stride_bytes % 8) || (ids_offset_bytes % 8) || ids_offset_bytes < (int64_t)B * k * 4) {
    set_error("ts_topk_merge_packed: bad layout");
    return TS_ERR_INVALID;
  }
  TS_CUDA_OK(cudaSetDevice(device));
  const float* scores = (const float*)blob;
  const int64_t* ids = (const int64_t*)((const char*)blob + ids_offset_bytes);
  return launch_merge_pairs(scores, ids, list_stride_bytes / 4, list_stride_bytes / 8, n_lists, B, k, out_scores, out_ids,
                            (cudaStream_t)stream);
}

int ts_exchange_push(int device, const void* blob_dev, int64_t nbytes, const int64_t* peer_bases_dev, int n_ranks, int rank,
                     int64_t slot_bytes, int64_t flags_offset, int parity, uint32_t seq, void* stream) {
  if (!blob_dev || !peer_bases_dev || n_ranks < 1 || rank < 0 || rank >= n_ranks || (parity != 0 && parity != 1) || nbytes > slot_bytes) {
    set_error("ts_exchange_push: invalid argument");
    return TS_ERR_INVALID;
  }
  TS_CUDA_OK(cudaSetDevice(device));
  const long long slot = ((long long)parity * n_ranks + rank) * slot_bytes;
  const long long flag = flags_offset + ((long long)parity * n_ranks + rank) * 4;
  return launch_exchange_push(blob_dev, (nbytes + 15) / 16 * 16, (const long long*)peer_bases_dev, n_ranks, slot, flag, seq,
                              (cudaStream_t)stream);
}

int ts_exchange_wait_merge(int device, const void* local_base_dev, int n_ranks, int B, int k, int64_t slot_bytes,
                           int64_t ids_offset, int64_t flags_offset, int parity, uint32_t seq, float* out_scores,
                           int64_t* out_ids, void* stream) {
  if (!local_base_dev || !out_scores || !out_ids || n_ranks < 1 || (parity != 0 && parity != 1) || (slot_bytes % 16) ||
      (ids_offset % 8) || ids_offset < (int64_t)B * k * 4 || ids_offset + (int64_t)B * k * 8 > slot_bytes) {
    set_error("ts_exchange_wait_merge: bad layout");
    return TS_ERR_INVALID;
  }
  TS_CUDA_OK(cudaSetDevice(device));
  const char* slots = (const char*)local_base_dev + (size_t)parity * n_ranks * slot_bytes;
  const unsigned int* flags = (const unsigned int*)((const char*)local_base_dev + flags_offset) + (size_t)parity * n_ranks;
  return launch_merge_pairs_wait((const float*)slots, (const int64_t*)(slots + ids_offset), slot_bytes / 4, slot_bytes / 8, n_ranks, B,
                                 k, flags, seq, out_scores, out_ids, (cudaStream_t)stream);
}

int ts_exchange_wait_sum(int device, const void* local_base_dev, int n_ranks, int64_t n_floats, int64_t slot_bytes,
                         int64_t flags_offset, int parity, uint32_t seq, float* out_dev, void* stream) {
  if (!local_base_dev || !out_dev || n_ranks < 1 || (parity != 0 && parity != 1) || (slot_bytes % 16) || n_floats * 4 > slot_bytes) {
    set_error("ts_exchange_wait_sum: bad layout");
    return TS_ERR_INVALID;
  }
  TS_CUDA_OK(cudaSetDevice(device));
  const char* slots = (const char*)local_base_dev + (size_t)parity * n_ranks * slot_bytes;
  const unsigned int* flags = (const unsigned int*)((const char*)local_base_dev + flags_offset) + (size_t)parity * n_ranks;
  return launch_exchange_wait_sum(slots, slot_bytes, flags, n_ranks, seq, n_floats, out_dev, (cudaStream_t)stream);
}

int ts_exchange_wait_take(int device, void* matrix_dev, const void* flags_dev, int n_ranks, uint32_t seq, int64_t n_floats,
                          float* out_dev, void* stream) {
  if (!matrix_dev || !flags_dev || !out_dev || n_ranks < 1 || n_floats <= 0 || seq == 0) { set_error("ts_exchange_wait_take: invalid argument"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(device));
  return launch_exchange_wait_take(matrix_dev, (const unsigned int*)flags_dev, n_ranks, seq, n_floats, out_dev, (cudaStream_t)stream);
}

int ts_index_debug_timeline(ts_index* h, uint64_t* out16) {
  if (!h || !out16 || !h->grid_bar) { set_error("ts_index_debug_timeline: invalid argument"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(h->device));
  TS_CUDA_OK(cudaDeviceSynchronize());
  TS_CUDA_OK(cudaMemcpy(out16, h->grid_bar + 16, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  return TS_OK;
}

int ts_index_get_rows(const ts_index* h, int64_t start, int64_t n, float* out_host) {
  if (!h || start < 0 || n < 0 || start + n > h->n || (n > 0 && !out_host)) { set_error("ts_index_get_rows: bad range"); return TS_ERR_INVALID; }
  if (n == 0) return TS_OK;
  TS_CUDA_OK(cudaSetDevice(h->device));
  float* tmp = nullptr;
  TS_CUDA_OK(cudaMalloc((void**)&tmp, (size_t)n * h->dim * sizeof(float)));
  const size_t row_b = (size_t)h->ld * dtype_size(h->dtype);
  int rc = launch_convert_rows((const char*)h->rows + (size_t)start * row_b, h->dtype, h->ld, tmp, TS_F32, h->dim, n,
                               h->dim, kNormNone, nullptr, 0);
  if (rc == TS_OK) {
    cudaError_t e = cudaMemcpy(out_host, tmp, (size_t)n * h->dim * sizeof(float), cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_error("copy back failed: %s", cudaGetErrorString(e)); rc = TS_ERR_CUDA; }
  }
  cudaFree(tmp);
  return rc;
}

// ------------------------------------------------------------------ Stage 2
int ts_tokstore_create(ts_tokstore** out, int device, int dim, int storage_dtype, int64_t reserve_docs,
                       int64_t reserve_tokens) {
  if (!out || dim <= 0 || dim > 4096 || !valid_storage(storage_dtype) || reserve_docs < 0 || reserve_tokens < 0) {
    set_error("ts_tokstore_create: invalid argument");
    return TS_ERR_INVALID;
  }
  DeviceInfo info;
  int rc = check_device(device, &info);
  if (rc) return rc;
  ts_tokstore* h = new ts_tokstore();
  memset(h, 0, sizeof(*h));
  h->device = device; h->dim = dim; h->dtype = storage_dtype; h->info = info;
  // tile layout (the Stage-2 flow kernel's operand image) whenever the shape allows it; TS_S2_FLOW=0 keeps
  // the row-major shard of the first tensor kernel (A/B timing, debugging)
  h->layout = (tok_tile_layout_ok(dim, storage_dtype) && env_flag("TS_S2_FLOW", kDefaultS2Flow)) ? kTokTile : kTokRowMajor;
  h->timer = new ScanTimer();
  h->hint_docs = reserve_docs;
  h->hint_rows = reserve_tokens + 8 * reserve_docs;  // every doc is padded to 8 rows
  *out = h;
  return TS_OK;
}

int ts_tokstore_destroy(ts_tokstore* h) {
  if (!h) return TS_OK;
  cudaSetDevice(h->device);
  void* ptrs[] = {h->tok, h->doc_off, h->doc_len, h->qbuf, h->stage, h->meta, h->hbuf, h->done_ctr, h->scratch_out};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (h->timer) { h->timer->destroy(); delete h->timer; }
  delete h;
  return TS_OK;
}

int ts_tokstore_set_profiling(ts_tokstore* h, int enable) { if (!h) return TS_ERR_INVALID; h->timer->on = enable != 0; h->timer->n = 0; return TS_OK; }
int ts_tokstore_scan_time(ts_tokstore* h, float* mean_ms, int* n) { if (!h) return TS_ERR_INVALID; cudaSetDevice(h->device); return h->timer->report(mean_ms, n); }

}  // extern "C"

namespace ts {
int tok_reserve(ts_tokstore* h, int64_t docs, int64_t rows, cudaStream_t st) {
  const int64_t hint_docs = h->hint_docs, hint_rows = h->hint_rows;
  const int64_t cur_docs = h->cap_docs, cur_rows = h->cap_rows;
  if (docs > cur_docs) {
    int64_t nc = cur_docs > 0 ? cur_docs : (hint_docs > 1024 ? hint_docs : 1024);
    while (nc < docs) nc *= 2;
    int64_t* noff = nullptr; int32_t* nlen = nullptr;
    if (cudaMalloc((void**)&noff, (size_t)nc * 8) != cudaSuccess || cudaMalloc((void**)&nlen, (size_t)nc * 4) != cudaSuccess) {
      cudaGetLastError();
      if (noff) cudaFree(noff);
      set_error("tokstore: cudaMalloc doc tables failed"); return TS_ERR_NOMEM;
    }
    if (h->ndocs > 0) {
      TS_CUDA_OK(cudaMemcpyAsync(noff, h->doc_off, (size_t)h->ndocs * 8, cudaMemcpyDeviceToDevice, st));
      TS_CUDA_OK(cudaMemcpyAsync(nlen, h->doc_len, (size_t)h->ndocs * 4, cudaMemcpyDeviceToDevice, st));
      TS_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (h->doc_off) cudaFree(h->doc_off);
    if (h->doc_len) cudaFree(h->doc_len);
    h->doc_off = noff; h->doc_len = nlen; h->cap_docs = nc;
  }
  if (rows > cur_rows) {
    int64_t nc = cur_rows > 0 ? cur_rows : (hint_rows > 8192 ? hint_rows : 8192);
    while (nc < rows) nc *= 2;
    const size_t row_b = (size_t)h->dim * dtype_size(h->dtype);
    void* nt = nullptr;
    cudaError_t e = cudaMalloc(&nt, (size_t)nc * row_b);
    if (e != cudaSuccess) {
      cudaGetLastError();
      nc = rows;
      e = cudaMalloc(&nt, (size_t)nc * row_b);
      if (e != cudaSuccess) { cudaGetLastError(); set_error("tokstore: cudaMalloc of %lld token rows failed: %s", (long long)nc, cudaGetErrorString(e)); return TS_ERR_NOMEM; }
    }
    if (h->nrows > 0) {
      TS_CUDA_OK(cudaMemcpyAsync(nt, h->tok, (size_t)h->nrows * row_b, cudaMemcpyDeviceToDevice, st));
      TS_CUDA_OK(cudaStreamSynchronize(st));
    }
    if (h->tok) cudaFree(h->tok);
    h->tok = nt; h->cap_rows = nc;
  }
  return TS_OK;
}


}

extern "C" {

int ts_tokstore_add(ts_tokstore* h, const void* tok, int src_dtype, int src_on_device, const int32_t* lens_host,
                    int n_docs, int normalize, void* stream) {
  if (!h || n_docs < 0 || (n_docs > 0 && (!tok || !lens_host))) { set_error("ts_tokstore_add: invalid argument"); return TS_ERR_INVALID; }
  if (src_dtype != TS_F32 && src_dtype != h->dtype) { set_error("ts_tokstore_add: src dtype must be f32 or the storage dtype"); return TS_ERR_INVALID; }
  if (n_docs == 0) return TS_OK;
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  std::vector<int64_t> src_off((size_t)n_docs), dst_off((size_t)n_docs);
  int64_t sum = 0, rows = h->nrows;
  for (int i = 0; i < n_docs; ++i) {
    const int L = lens_host[i];
    if (L < 1 || L > TS_S2_MAX_LD) { set_error("ts_tokstore_add: doc %d has %d tokens (allowed 1..%d)", i, L, TS_S2_MAX_LD); return TS_ERR_INVALID; }
    src_off[i] = sum; dst_off[i] = rows;
    sum += L; rows += (L + 7) & ~7;
  }
  if (rows > 0x7FFFFF00ll) { set_error("ts_tokstore_add: shard limited to 2^31 token rows"); return TS_ERR_UNSUPPORTED; }
  int rc = tok_reserve(h, h->ndocs + n_docs, rows, st);
  if (rc) return rc;
  // upload per-doc tables (src offsets are scratch; dst offsets + lens append to the store)
  if ((rc = ensure_bytes(&h->meta, &h->meta_b, (size_t)n_docs * 8))) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(h->meta, src_off.data(), (size_t)n_docs * 8, cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaMemcpyAsync(h->doc_off + h->ndocs, dst_off.data(), (size_t)n_docs * 8, cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaMemcpyAsync(h->doc_len + h->ndocs, lens_host, (size_t)n_docs * 4, cudaMemcpyHostToDevice, st));
  const void* dsrc = tok;
  if (!src_on_device) {
    const size_t nb = (size_t)sum * h->dim * dtype_size(src_dtype);
    if ((rc = ensure_bytes(&h->stage, &h->stage_b, nb))) return rc;
    TS_CUDA_OK(cudaMemcpyAsync(h->stage, tok, nb, cudaMemcpyHostToDevice, st));
    dsrc = h->stage;
  }
  rc = launch_tok_ingest(dsrc, src_dtype, (const int64_t*)h->meta, h->doc_off + h->ndocs, h->doc_len + h->ndocs, n_docs,
                         h->tok, h->dtype, h->dim, normalize, h->layout, st);
  if (rc) return rc;
  ++h->launches;
  TS_CUDA_OK(cudaStreamSynchronize(st));  // host vectors / staging are reused
  h->ndocs += n_docs; h->nrows = rows; h->ntokens += sum;
  return TS_OK;
}

int64_t ts_tokstore_ndocs(const ts_tokstore* h) { return h ? h->ndocs : -1; }
int64_t ts_tokstore_ntokens(const ts_tokstore* h) { return h ? h->ntokens : -1; }
int ts_tokstore_dim(const ts_tokstore* h) { return h ? h->dim : -1; }
int ts_tokstore_dtype(const ts_tokstore* h) { return h ? h->dtype : -1; }
int ts_tokstore_layout(const ts_tokstore* h) { return h ? h->layout : -1; }
int ts_tokstore_reset(ts_tokstore* h) { if (!h) return TS_ERR_INVALID; h->ndocs = 0; h->nrows = 0; h->ntokens = 0; return TS_OK; }
int ts_tokstore_set_id_base(ts_tokstore* h, int64_t b) { if (!h) return TS_ERR_INVALID; h->id_base = b; return TS_OK; }
int64_t ts_tokstore_launch_count(const ts_tokstore* h) { return h ? h->launches : -1; }

}  // extern "C"

namespace ts {
struct ScatterTarget { const long long* bases; int n_ranks, rank; long long off, flags_off; unsigned int seq; };
static int maxsim_impl(ts_tokstore* h, const void* q_tok, int q_dtype, const int32_t* q_len, int B, int lq_stride,
                       const int64_t* cand, const int32_t* n_cand, int C, int mode, unsigned flags, float* out, void* stream,
                       const ScatterTarget* sc);
}

extern "C" {

int ts_maxsim(ts_tokstore* h, const void* q_tok, int q_dtype, const int32_t* q_len, int B, int lq_stride,
              const int64_t* cand, const int32_t* n_cand, int C, int mode, unsigned flags, float* out, void* stream) {
  if (!out) { set_error("ts_maxsim: invalid argument"); return TS_ERR_INVALID; }
  return ts::maxsim_impl(h, q_tok, q_dtype, q_len, B, lq_stride, cand, n_cand, C, mode, flags, out, stream, nullptr);
}

int ts_maxsim_scatter(ts_tokstore* h, const void* q_tok, int q_dtype, const int32_t* q_len, int B, int lq_stride,
                      const int64_t* cand, const int32_t* n_cand, int C, int mode, unsigned flags,
                      const int64_t* peer_bases_dev, int n_ranks, int rank, int64_t matrix_offset, int64_t flags_offset,
                      uint32_t seq, void* stream) {
  if (!h || !peer_bases_dev || n_ranks < 1 || n_ranks > 256 || rank < 0 || rank >= n_ranks || (matrix_offset & 3) || (flags_offset & 3) || seq == 0) {
    set_error("ts_maxsim_scatter: invalid argument");
    return TS_ERR_INVALID;
  }
  TS_CUDA_OK(cudaSetDevice(h->device));
  if (!h->done_ctr) {
    if (cudaMalloc((void**)&h->done_ctr, 16) != cudaSuccess) { cudaGetLastError(); set_error("ts_maxsim_scatter: cudaMalloc failed"); return TS_ERR_NOMEM; }
    TS_CUDA_OK(cudaMemset(h->done_ctr, 0, 16));
  }
  int rc = ensure_bytes(&h->scratch_out, &h->scratch_out_b, (size_t)B * C * sizeof(float));
  if (rc) return rc;
  ts::ScatterTarget sc{(const long long*)peer_bases_dev, n_ranks, rank, matrix_offset, flags_offset, seq};
  return ts::maxsim_impl(h, q_tok, q_dtype, q_len, B, lq_stride, cand, n_cand, C, mode, flags, (float*)h->scratch_out, stream, &sc);
}

}  // extern "C"

namespace ts {
static int maxsim_impl(ts_tokstore* h, const void* q_tok, int q_dtype, const int32_t* q_len, int B, int lq_stride,
                       const int64_t* cand, const int32_t* n_cand, int C, int mode, unsigned flags, float* out, void* stream,
                       const ScatterTarget* sc) {
  if (!h || !q_tok || !cand || !out || B <= 0 || C <= 0 || lq_stride <= 0) { set_error("ts_maxsim: invalid argument"); return TS_ERR_INVALID; }
  if ((mode & 0xff) != TS_S2_MAXSIM && (mode & 0xff) != TS_S2_COLBERT) { set_error("ts_maxsim: bad mode %d", mode); return TS_ERR_INVALID; }
  if (q_dtype != TS_F32 && q_dtype != h->dtype) { set_error("ts_maxsim: query dtype must be f32 or the storage dtype"); return TS_ERR_INVALID; }
  if ((int64_t)B * C > 0x7FFFFFFFll) { set_error("ts_maxsim: B*C too large"); return TS_ERR_UNSUPPORTED; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  const int64_t qrows = (int64_t)B * lq_stride;
  // +128 rows of slack: a 128-row TMA box of the last query may run past the end
  int rc = ensure_bytes(&h->qbuf, &h->qbuf_b, (size_t)(qrows + 128) * h->dim * dtype_size(h->dtype));
  if (rc) return rc;
  rc = launch_convert_rows(q_tok, q_dtype, h->dim, h->qbuf, h->dtype, h->dim, qrows, h->dim,
                           (flags & TS_FLAG_NORMALIZE_Q) ? kNormStage2 : kNormNone, nullptr, st);
  if (rc) return rc;
  ++h->launches;
  MaxSimArgs a{};
  a.layout = h->layout;
  a.tok = h->tok; a.doc_off = h->doc_off; a.doc_len = h->doc_len; a.ndocs = h->ndocs; a.id_base = h->id_base;
  a.ntok_rows = h->nrows; a.dim = h->dim; a.dtype = h->dtype; a.q = h->qbuf; a.q_len = q_len; a.B = B;
  a.lq_stride = lq_stride; a.cand = cand; a.n_cand = n_cand; a.C = C; a.mode = mode; a.out = out;
  a.sm_count = h->info.sm_count;
  if (sc) {
    // the scatter lives in the flow kernel's finalize step: other kernels (row-major shards, fp32, Lq > 128) are
    // refused and the caller exchanges the matrix the ordinary way
    if (!maxsim_flow_takes(a) || (mode & 0x100)) { set_error("ts_maxsim_scatter: needs the tile-layout tensor kernel"); return TS_ERR_UNSUPPORTED; }
    a.scatter_bases = sc->bases; a.scatter_n = sc->n_ranks; a.scatter_rank = sc->rank; a.scatter_off = sc->off;
    a.scatter_flags_off = sc->flags_off; a.scatter_seq = sc->seq; a.scatter_done = h->done_ctr;
  }
  int launches = 0;
  h->timer->begin(st);
  rc = launch_maxsim(a, st, &launches);
  h->timer->end(st);
  h->launches += launches;
  return rc;
}
}  // namespace ts

extern "C" {

int ts_maxsim_host(ts_tokstore* h, const void* q_tok, int q_dtype, const int32_t* q_len, int B, int lq_stride,
                   const int64_t* cand, const int32_t* n_cand, int C, int mode, unsigned flags, float* out,
                   void* stream) {
  if (!h || !q_tok || !cand || !out || B <= 0 || C <= 0 || lq_stride <= 0) { set_error("ts_maxsim_host: invalid argument"); return TS_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->device));
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  const size_t qb = up((size_t)B * lq_stride * h->dim * dtype_size(q_dtype));
  const size_t cb = up((size_t)B * C * 8), lb = up((size_t)B * 4), ob = up((size_t)B * C * 4);
  int rc = ensure_bytes(&h->hbuf, &h->hbuf_b, qb + cb + 2 * lb + ob);
  if (rc) return rc;
  char* base = (char*)h->hbuf;
  void* dq = base; int64_t* dc = (int64_t*)(base + qb);
  int32_t* dql = (int32_t*)(base + qb + cb); int32_t* dnc = (int32_t*)(base + qb + cb + lb);
  float* dout = (float*)(base + qb + cb + 2 * lb);
  TS_CUDA_OK(cudaMemcpyAsync(dq, q_tok, (size_t)B * lq_stride * h->dim * dtype_size(q_dtype), cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaMemcpyAsync(dc, cand, (size_t)B * C * 8, cudaMemcpyHostToDevice, st));
  if (q_len) TS_CUDA_OK(cudaMemcpyAsync(dql, q_len, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  if (n_cand) TS_CUDA_OK(cudaMemcpyAsync(dnc, n_cand, (size_t)B * 4, cudaMemcpyHostToDevice, st));
  rc = ts_maxsim(h, dq, q_dtype, q_len ? dql : nullptr, B, lq_stride, dc, n_cand ? dnc : nullptr, C, mode, flags, dout, stream);
  if (rc) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(out, dout, (size_t)B * C * 4, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  return TS_OK;
}

int ts_rank_desc(int device, const float* scores, const int32_t* n_cand, int B, int C, int top_k, float* out_scores,
                 int32_t* out_pos, void* stream) {
  if (!scores || !out_scores || !out_pos) { set_error("ts_rank_desc: null pointer"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(device));
  return launch_rank_desc(scores, n_cand, B, C, top_k, out_scores, out_pos, (cudaStream_t)stream);
}

}  // extern "C"
