// Approximate Stage-1 mode: inverted lists over the resident row matrix.
//
// Replaces faiss.IndexIVFFlat(quantizer = IndexFlatIP(d), d, nlist, METRIC_INNER_PRODUCT) as the
// reference builds it when its first batch has more than 1000 rows
// (/root/reference/src/stage1_retriever.py:262-273: train, add, nprobe; later adds at :313;
// search at :380).
//
// Layout.  A ts_ivf does NOT copy the corpus: the rows stay where ts_index keeps them, in
// insertion order (so the exact scan and the approximate scan share one copy of the shard, and an
// add is an append).  The lists hold row NUMBERS, CSR style:
//     cent  [nlist][ldc] fp32      coarse centroids (ldc = dim rounded up to 8, pad columns zero)
//     off   [nlist + 1]  int64     list l = order[off[l] .. off[l+1])
//     order [n]          int32     row numbers grouped by list, ascending inside a list
// A row is one contiguous 16-byte-aligned run of ld*sizeof(T) bytes (1.5 - 2 KB at dim 768 - 1024),
// so gathering whole rows by number keeps every memory request coalesced.
//
// Kernels (all CUDA-core: the work is a gather-and-stream, bound by HBM -- one probe reads
// nprobe/nlist of the shard):
//   ivf_assign_kernel  add():    list of each new row = argmax_c <x, cent_c>, fp32
//   ivf_coarse_kernel  search(): per query, the nprobe lists with the largest <q, cent>
//   ivf_scan_kernel    search(): one CTA per (query, probed list, segment) streams the rows of its
//                      segment through 128-bit loads, keeps a per-warp top-k in shared memory
//                      (threshold + bitonic prune, as the exact stream scan does) and writes one
//                      sorted partial list; select_kernel (topk_select.cu) merges them.
// Algorithmic bytes per query: sum over its probed lists of len * ld * sizeof(T)
// (+ len * 4 for the row numbers, + nlist * ldc * 4 for the centroids).
//
// Scores are the exact scan's scores (same stored values, fp32 accumulation) and keys carry the
// row number, so ties resolve as everywhere else: score descending, then id ascending.  With
// nprobe == nlist the result equals the exact search.
#include <string.h>

#include <vector>

#include "ts_common.cuh"
#include "ts_handles.h"

using namespace ts;

struct ts_ivf {
  ts_index* base;
  int nlist, ldc, trained;
  int64_t n_assigned;
  int64_t base_gen;                 // base->reset_gen the lists were built under
  float* cent;                      // device [nlist][ldc]
  int64_t* off;                     // device [nlist + 1]
  int32_t* order; int64_t order_cap;
  std::vector<int32_t>* assign_host;   // list of every assigned row (host mirror: lists are rebuilt from it)
  std::vector<int64_t>* off_host;
  int64_t launches;
  // grow-only scratch
  void* q32; size_t q32_b;          // [B][ldc] fp32 queries for the coarse step
  void* qst; size_t qst_b;          // [B][ld] queries in the storage dtype for the scan
  void* probe; size_t probe_b;      // [B][nprobe] int32 lists + [B][nprobe] fp32 scores
  void* work; size_t work_b;        // [B * nprobe] (query, probe) pairs ordered by list (batches)
  void* partial; size_t partial_b;
  void* tmp0; size_t tmp0_b;
  void* tmp1; size_t tmp1_b;
  void* atmp; size_t atmp_b;        // assignments of the rows being added
  void* stage; size_t stage_b;      // host-variant staging
  void* hout; size_t hout_b;
};

namespace ts {
namespace {

constexpr int kIvfThreads = 256;
constexpr int kIvfWarps = kIvfThreads / 32;
constexpr int kIvfR = 4;            // rows per warp step
constexpr int kIvfJ = 4;            // centroids per assign step
constexpr int kIvfMergeCap = 4096;  // keys of the CTA merge buffer
constexpr int kIvfMaxSeg = 256;     // segments per probed list
constexpr int kIvfCtasPerSm = 4;    // grid target: CTAs per SM (TS_IVF_CTAS_PER_SM overrides, 1..32 -- tuning knob)
constexpr int kIvfMaxPairs = 4096;  // (query, probe) pairs the order kernel sorts (32 KB of keys)

// shared-memory twin of warp_prune_list (ts_common.cuh): the warp sorts list[0..cnt) and keeps the
// best k in list[0..k), descending, zero padded; returns the k-th key
template <int KPL>
__device__ __noinline__ uint64_t warp_prune_smem_t(uint64_t* list, int cnt, int k, int lane) {
  uint64_t v[KPL];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int e = j * 32 + lane;
    v[j] = (e < cnt) ? list[e] : 0ull;
  }
  warp_sort_desc<KPL>(v, lane);
  uint64_t kth_local = 0ull;
  const int kj = (k - 1) >> 5;
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int e = j * 32 + lane;
    if (e < k) list[e] = v[j];
    if (j == kj) kth_local = v[j];
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, kth_local, (k - 1) & 31);
}
__device__ __forceinline__ uint64_t warp_prune_smem(uint64_t* list, int cnt, int k, int lane, int cap) {
  return cap <= 256 ? warp_prune_smem_t<8>(list, cnt, k, lane) : warp_prune_smem_t<32>(list, cnt, k, lane);
}

__device__ __forceinline__ void load_f32_chunk(const float* p, float (&f)[4]) {
  const float4 t = *reinterpret_cast<const float4*>(p);
  f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
}
__device__ __forceinline__ void load_f32_chunk(const float* p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---- add(): list of each row ---------------------------------------------------------------
// One warp owns R rows at a time and walks the centroids J at a time; the rows are re-read from
// L1 for every centroid tile (R rows = 8 KB per warp), the centroids come from L1/L2
// (nlist * ldc * 4 bytes, 400 KB at nlist 100 x dim 1024).  2 * n * nlist * dim flop in total.
template <typename T>
__global__ void __launch_bounds__(kIvfThreads)
    ivf_assign_kernel(const T* __restrict__ X, int ld, int64_t row_lo, int64_t row_hi, const float* __restrict__ cent,
                      int ldc, int nlist, int32_t* __restrict__ assign_out) {
  constexpr int R = kIvfR, J = kIvfJ, EPC = Elem<T>::kPerChunk;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int C = ld / EPC;
  const uint4* Xc = reinterpret_cast<const uint4*>(X);
  const int64_t warps_total = (int64_t)gridDim.x * kIvfWarps;
  const int64_t gw = (int64_t)blockIdx.x * kIvfWarps + warp;
  const int64_t n_groups = (row_hi - row_lo + R - 1) / R;
  for (int64_t g = gw; g < n_groups; g += warps_total) {
    const int64_t row0 = row_lo + g * R;
    const uint4* xr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) xr[r] = Xc + ((row0 + r < row_hi) ? (row0 + r) : (row_hi - 1)) * C;
    float best[R];
    int bi[R];
#pragma unroll
    for (int r = 0; r < R; ++r) { best[r] = -INFINITY; bi[r] = 0; }
    for (int c0 = 0; c0 < nlist; c0 += J) {
      const float* cp[J];
#pragma unroll
      for (int j = 0; j < J; ++j) cp[j] = cent + (size_t)((c0 + j < nlist) ? (c0 + j) : (nlist - 1)) * ldc;
      float acc[R][J];
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < J; ++j) acc[r][j] = 0.f;
#pragma unroll 1
      for (int ch = lane; ch < C; ch += 32) {
        float xf[R][EPC];
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const uint4 v = __ldg(xr[r] + ch);
          Elem<T>::unpack(v, xf[r]);
        }
#pragma unroll
        for (int j = 0; j < J; ++j) {
          float cf[EPC];
          load_f32_chunk(cp[j] + (size_t)ch * EPC, cf);
#pragma unroll
          for (int r = 0; r < R; ++r)
#pragma unroll
            for (int e = 0; e < EPC; ++e) acc[r][j] = fmaf(xf[r][e], cf[e], acc[r][j]);
        }
      }
#pragma unroll
      for (int r = 0; r < R; ++r)
#pragma unroll
        for (int j = 0; j < J; ++j) acc[r][j] = warp_sum(acc[r][j]);
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (c0 + j < nlist) {
#pragma unroll
          for (int r = 0; r < R; ++r)
            if (acc[r][j] > best[r]) { best[r] = acc[r][j]; bi[r] = c0 + j; }   // strict: the lowest list wins a tie
        }
      }
    }
    if (lane == 0) {
#pragma unroll
      for (int r = 0; r < R; ++r)
        if (row0 + r < row_hi) assign_out[row0 + r - row_lo] = bi[r];
    }
  }
}

// ---- search(), step 1: quantizer.search(q, nprobe) -------------------------------------------
// One CTA per query: nlist fp32 dot products (a warp per centroid), one bitonic sort of the keys.
__global__ void __launch_bounds__(kIvfThreads)
    ivf_coarse_kernel(const float* __restrict__ Q32, int ldc, const float* __restrict__ cent, int nlist, int npow2, int nprobe,
                      int32_t* __restrict__ probe, float* __restrict__ pscore) {
  TS_DYN_SMEM(unsigned char, smem_raw);
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);            // [npow2]
  float* qs = reinterpret_cast<float*>(keys + npow2);                // [ldc]
  const int b = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ldc; i += blockDim.x) qs[i] = Q32[(size_t)b * ldc + i];
  for (int i = nlist + threadIdx.x; i < npow2; i += blockDim.x) keys[i] = 0ull;
  __syncthreads();
  const int C4 = ldc / 4;
  for (int c = warp; c < nlist; c += kIvfWarps) {
    const float4* cp = reinterpret_cast<const float4*>(cent + (size_t)c * ldc);
    const float4* qp = reinterpret_cast<const float4*>(qs);
    float acc = 0.f;
    for (int i = lane; i < C4; i += 32) {
      const float4 a = __ldg(cp + i), q = qp[i];
      acc = fmaf(a.x, q.x, acc); acc = fmaf(a.y, q.y, acc); acc = fmaf(a.z, q.z, acc); acc = fmaf(a.w, q.w, acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) keys[c] = make_key(acc, (uint32_t)c);
  }
  __syncthreads();
  block_sort_desc(keys, npow2);
  for (int j = threadIdx.x; j < nprobe; j += blockDim.x) {
    const uint64_t key = keys[j];
    probe[(size_t)b * nprobe + j] = key ? (int32_t)key_idx(key) : -1;
    pscore[(size_t)b * nprobe + j] = key ? key_score(key) : kLowestF32;
  }
}

// ---- search(), batches: co-schedule the probes of one list ---------------------------------------
// B queries x nprobe probes touch each list B*nprobe/nlist times on average.  Sorting the (query, probe)
// pairs by list makes the CTAs that scan the same list neighbours in launch order, so the list comes from
// HBM once and from the 126 MB L2 for the other queries.  One CTA, one bitonic sort of <= kIvfMaxPairs keys.
__global__ void __launch_bounds__(1024)
    ivf_order_kernel(const int32_t* __restrict__ probe, int n_pairs, int npow2, int32_t* __restrict__ work) {
  TS_DYN_SMEM(uint64_t, keys);
  for (int i = threadIdx.x; i < npow2; i += blockDim.x)
    keys[i] = (i < n_pairs) ? (((uint64_t)(uint32_t)(__ldg(probe + i) + 1) << 32) | (uint32_t)(n_pairs - 1 - i) | (1ull << 63)) : 0ull;
  __syncthreads();
  block_sort_desc(keys, npow2);
  for (int i = threadIdx.x; i < n_pairs; i += blockDim.x) work[i] = n_pairs - 1 - (int32_t)(uint32_t)keys[i];
}

// ---- search(), step 2: scan the probed lists --------------------------------------------------
// grid (S, nprobe, B): CTA (s, j, b) scans segment s of the j-th probed list of query b; with a work
// list (batches) the grid is (S, B * nprobe) and CTA (s, w) takes pair work[w] = b * nprobe + j.
template <typename T>
__global__ void __launch_bounds__(kIvfThreads, 3)
    ivf_scan_kernel(const T* __restrict__ X, int ld, const float* __restrict__ inv_norm, const T* __restrict__ Q,
                    const int32_t* __restrict__ probe, int nprobe, const int64_t* __restrict__ off,
                    const int32_t* __restrict__ order, int k, int CAP, uint64_t* __restrict__ partial, int B,
                    const int32_t* __restrict__ work) {
  constexpr int R = kIvfR, EPC = Elem<T>::kPerChunk;
  TS_DYN_SMEM(unsigned char, smem_raw);
  uint64_t* sbuf = reinterpret_cast<uint64_t*>(smem_raw);            // [kIvfMergeCap] CTA merge buffer
  uint64_t* wl = sbuf + kIvfMergeCap;                                // [warps][CAP] per-warp candidate lists
  float* Qs = reinterpret_cast<float*>(wl + (size_t)kIvfWarps * CAP);  // [ld]
  const int s = blockIdx.x, S = gridDim.x;
  int j = blockIdx.y, b = blockIdx.z;
  if (work) { const int pair = __ldg(work + blockIdx.y); b = pair / nprobe; j = pair % nprobe; }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < ld; i += blockDim.x) Qs[i] = Elem<T>::to_f32(Q[(size_t)b * ld + i]);
  __syncthreads();

  const int list = __ldg(probe + (size_t)b * nprobe + j);
  int64_t lo = 0, hi = 0;
  if (list >= 0) { lo = __ldg(off + list); hi = __ldg(off + list + 1); }
  int64_t seg = (hi - lo + S - 1) / S;
  seg = (seg + R - 1) / R * R;
  const int64_t p0 = lo + (int64_t)s * seg;
  const int64_t p1 = (p0 + seg < hi) ? (p0 + seg) : hi;

  uint64_t* my = wl + (size_t)warp * CAP;
  float tau = -INFINITY;
  int cnt = 0;
  const int C = ld / EPC;
  const uint4* Xc = reinterpret_cast<const uint4*>(X);
  for (int64_t g = p0 + (int64_t)warp * R; g < p1; g += (int64_t)kIvfWarps * R) {
    int rows[R];
#pragma unroll
    for (int r = 0; r < R; ++r) rows[r] = __ldg(order + ((g + r < p1) ? (g + r) : (p1 - 1)));
    float acc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = 0.f;
#pragma unroll 1
    for (int c = lane; c < C; c += 32) {
      uint4 xv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xv[r] = ldg_stream(Xc + (size_t)rows[r] * C + c);
      float qf[EPC];
      load_f32_chunk(Qs + (size_t)c * EPC, qf);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        float xf[EPC];
        Elem<T>::unpack(xv[r], xf);
#pragma unroll
        for (int e = 0; e < EPC; ++e) acc[r] = fmaf(xf[e], qf[e], acc[r]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
    // warp-uniform threshold filter; rows inside a list ascend, so a later tie never displaces an earlier one
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (g + r < p1) {
        const float sc = acc[r] * (inv_norm ? __ldg(inv_norm + rows[r]) : 1.0f);
        if (sc > tau) {
          if (lane == 0) my[cnt] = make_key(sc, (uint32_t)rows[r]);
          ++cnt;
        }
      }
    }
    if (cnt > CAP - R) {
      const uint64_t kth = warp_prune_smem(my, cnt, k, lane, CAP);
      cnt = k;
      tau = key_score(kth);
    }
  }
  warp_prune_smem(my, cnt, k, lane, CAP);   // my[0..k) sorted descending, zero padded
  __syncthreads();

  auto load = [&](int i) -> uint64_t { return wl[(size_t)(i / k) * CAP + (i % k)]; };
  block_select_topk(sbuf, kIvfMergeCap, k, kIvfWarps * k, load);
  uint64_t* out = partial + ((size_t)(j * S + s) * B + b) * k;
  for (int r = threadIdx.x; r < k; r += blockDim.x) out[r] = sbuf[r];
}

template <typename T>
int launch_assign_t(const ts_index* base, int64_t lo, int64_t hi, const float* cent, int ldc, int nlist, int32_t* out, cudaStream_t st) {
  const int64_t groups = (hi - lo + kIvfR - 1) / kIvfR;
  int64_t grid = (groups + kIvfWarps - 1) / kIvfWarps;
  const int64_t cap = (int64_t)base->info.sm_count * 8;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  TS_LAUNCH(ivf_assign_kernel<T>, (unsigned)grid, kIvfThreads, 0, st, reinterpret_cast<const T*>(base->rows), base->ld, lo, hi, cent, ldc,
            nlist, out);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

template <typename T>
int launch_scan_t(const ts_ivf* h, int B, int k, int nprobe, int S, const int32_t* work, cudaStream_t st) {
  const ts_index* base = h->base;
  const int CAP = cap_for_k(k);
  const size_t smem = (size_t)kIvfMergeCap * 8 + (size_t)kIvfWarps * CAP * 8 + (size_t)base->ld * sizeof(float);
  auto kern = ivf_scan_kernel<T>;
  if (smem > 48 * 1024) TS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const dim3 grid = work ? dim3(S, B * nprobe, 1) : dim3(S, nprobe, B);
  TS_LAUNCH(kern, grid, kIvfThreads, smem, st, reinterpret_cast<const T*>(base->rows), base->ld,
            (base->metric == TS_METRIC_COSINE) ? base->inv_norm : nullptr, reinterpret_cast<const T*>(h->qst),
            reinterpret_cast<const int32_t*>(h->probe), nprobe, h->off, h->order, k, CAP, reinterpret_cast<uint64_t*>(h->partial), B,
            work);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

// (re)build the CSR lists from the host mirror of the assignments and upload them
int rebuild_lists(ts_ivf* h, cudaStream_t st) {
  const std::vector<int32_t>& a = *h->assign_host;
  const int64_t n = (int64_t)a.size();
  std::vector<int64_t>& off = *h->off_host;
  off.assign((size_t)h->nlist + 1, 0);
  for (int64_t i = 0; i < n; ++i) ++off[(size_t)a[(size_t)i] + 1];
  for (int l = 0; l < h->nlist; ++l) off[(size_t)l + 1] += off[(size_t)l];
  std::vector<int32_t> order((size_t)n);
  std::vector<int64_t> cur(off.begin(), off.end() - 1);
  for (int64_t i = 0; i < n; ++i) order[(size_t)cur[(size_t)a[(size_t)i]]++] = (int32_t)i;   // stable: rows ascend inside a list
  if (n > h->order_cap) {
    if (h->order) { cudaFree(h->order); h->order = nullptr; h->order_cap = 0; }
    int64_t ncap = n + n / 2 + 1024;
    if (cudaMalloc((void**)&h->order, (size_t)ncap * 4) != cudaSuccess) { cudaGetLastError(); set_error("ivf: cudaMalloc of %lld list entries failed", (long long)ncap); return TS_ERR_NOMEM; }
    h->order_cap = ncap;
  }
  if (n > 0) TS_CUDA_OK(cudaMemcpyAsync(h->order, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaMemcpyAsync(h->off, off.data(), ((size_t)h->nlist + 1) * 8, cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));   // `order` (host) goes out of scope
  h->n_assigned = n;
  return TS_OK;
}

}  // namespace
}  // namespace ts

extern "C" {

int ts_ivf_create(ts_ivf** out, ts_index* base, int nlist) {
  if (!out || !base || nlist < 1 || nlist > TS_IVF_MAX_NLIST) { set_error("ts_ivf_create: invalid argument (1 <= nlist <= %d)", TS_IVF_MAX_NLIST); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(base->device));
  ts_ivf* h = new ts_ivf();
  memset(h, 0, sizeof(*h));
  h->base = base; h->nlist = nlist; h->ldc = (base->dim + 7) / 8 * 8; h->base_gen = base->reset_gen;
  h->assign_host = new std::vector<int32_t>();
  h->off_host = new std::vector<int64_t>((size_t)nlist + 1, 0);
  if (cudaMalloc((void**)&h->cent, (size_t)nlist * h->ldc * 4) != cudaSuccess || cudaMalloc((void**)&h->off, ((size_t)nlist + 1) * 8) != cudaSuccess) {
    cudaGetLastError();
    set_error("ts_ivf_create: cudaMalloc failed");
    ts_ivf_destroy(h);
    return TS_ERR_NOMEM;
  }
  cudaMemset(h->off, 0, ((size_t)nlist + 1) * 8);
  *out = h;
  return TS_OK;
}

int ts_ivf_destroy(ts_ivf* h) {
  if (!h) return TS_OK;
  cudaSetDevice(h->base->device);
  void* ptrs[] = {h->cent, h->off, h->order, h->q32, h->qst, h->probe, h->work, h->partial, h->tmp0, h->tmp1, h->atmp, h->stage, h->hout};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete h->assign_host;
  delete h->off_host;
  delete h;
  return TS_OK;
}

int ts_ivf_nlist(const ts_ivf* h) { return h ? h->nlist : -1; }
int ts_ivf_is_trained(const ts_ivf* h) { return h ? h->trained : -1; }
int64_t ts_ivf_nassigned(const ts_ivf* h) { return h ? h->n_assigned : -1; }
int64_t ts_ivf_launch_count(const ts_ivf* h) { return h ? h->launches : -1; }

int ts_ivf_set_centroids(ts_ivf* h, const float* centroids_host, void* stream) {
  if (!h || !centroids_host) { set_error("ts_ivf_set_centroids: invalid argument"); return TS_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->base->device));
  const int dim = h->base->dim;
  std::vector<float> padded((size_t)h->nlist * h->ldc, 0.f);
  for (int c = 0; c < h->nlist; ++c) memcpy(&padded[(size_t)c * h->ldc], centroids_host + (size_t)c * dim, (size_t)dim * 4);
  TS_CUDA_OK(cudaMemcpyAsync(h->cent, padded.data(), padded.size() * 4, cudaMemcpyHostToDevice, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  h->trained = 1;
  h->assign_host->clear();          // new centroids: every row has to be assigned again
  h->n_assigned = 0;
  return TS_OK;
}

int ts_ivf_get_centroids(const ts_ivf* h, float* out_host) {
  if (!h || !out_host || !h->trained) { set_error("ts_ivf_get_centroids: invalid argument or not trained"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(h->base->device));
  std::vector<float> padded((size_t)h->nlist * h->ldc);
  TS_CUDA_OK(cudaMemcpy(padded.data(), h->cent, padded.size() * 4, cudaMemcpyDeviceToHost));
  for (int c = 0; c < h->nlist; ++c) memcpy(out_host + (size_t)c * h->base->dim, &padded[(size_t)c * h->ldc], (size_t)h->base->dim * 4);
  return TS_OK;
}

int ts_ivf_sync(ts_ivf* h, void* stream) {
  if (!h) { set_error("ts_ivf_sync: invalid argument"); return TS_ERR_INVALID; }
  if (!h->trained) { set_error("ivf: no centroids yet (train / ts_ivf_set_centroids first)"); return TS_ERR_INVALID; }
  const ts_index* base = h->base;
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(base->device));
  if (base->reset_gen != h->base_gen || base->n < h->n_assigned) {   // the index was reset: every list is stale
    h->assign_host->clear(); h->n_assigned = 0; h->base_gen = base->reset_gen;
  }
  const int64_t lo = h->n_assigned, hi = base->n;
  if (lo == hi && (int64_t)h->assign_host->size() == hi) return TS_OK;
  if (hi > 0x7fffffffll) { set_error("ivf: shard limited to 2^31 rows"); return TS_ERR_UNSUPPORTED; }
  if (hi > lo) {
    int rc = ensure_bytes(&h->atmp, &h->atmp_b, (size_t)(hi - lo) * 4);
    if (rc) return rc;
    switch (base->dtype) {
      case TS_BF16: rc = launch_assign_t<__nv_bfloat16>(base, lo, hi, h->cent, h->ldc, h->nlist, (int32_t*)h->atmp, st); break;
      case TS_F16: rc = launch_assign_t<__half>(base, lo, hi, h->cent, h->ldc, h->nlist, (int32_t*)h->atmp, st); break;
      default: rc = launch_assign_t<float>(base, lo, hi, h->cent, h->ldc, h->nlist, (int32_t*)h->atmp, st); break;
    }
    if (rc) return rc;
    ++h->launches;
    h->assign_host->resize((size_t)hi);
    TS_CUDA_OK(cudaMemcpyAsync(h->assign_host->data() + lo, h->atmp, (size_t)(hi - lo) * 4, cudaMemcpyDeviceToHost, st));
    TS_CUDA_OK(cudaStreamSynchronize(st));
  }
  return rebuild_lists(h, st);
}

int ts_ivf_set_assignments(ts_ivf* h, const int32_t* assign_host, int64_t n, void* stream) {
  if (!h || n < 0 || (n > 0 && !assign_host)) { set_error("ts_ivf_set_assignments: invalid argument"); return TS_ERR_INVALID; }
  if (!h->trained) { set_error("ivf: no centroids yet (ts_ivf_set_centroids first)"); return TS_ERR_INVALID; }
  if (n != h->base->n) { set_error("ts_ivf_set_assignments: %lld assignments for %lld rows", (long long)n, (long long)h->base->n); return TS_ERR_INVALID; }
  for (int64_t i = 0; i < n; ++i)
    if (assign_host[i] < 0 || assign_host[i] >= h->nlist) { set_error("ts_ivf_set_assignments: row %lld names list %d of %d", (long long)i, assign_host[i], h->nlist); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaSetDevice(h->base->device));
  h->assign_host->assign(assign_host, assign_host + n);
  h->base_gen = h->base->reset_gen;
  return rebuild_lists(h, (cudaStream_t)stream);
}

int ts_ivf_get_assignments(const ts_ivf* h, int32_t* out_host, int64_t n) {
  if (!h || n < 0 || n > (int64_t)h->assign_host->size() || (n > 0 && !out_host)) { set_error("ts_ivf_get_assignments: bad range"); return TS_ERR_INVALID; }
  if (n) memcpy(out_host, h->assign_host->data(), (size_t)n * 4);
  return TS_OK;
}

int ts_ivf_list_sizes(const ts_ivf* h, int64_t* out_host) {
  if (!h || !out_host) { set_error("ts_ivf_list_sizes: invalid argument"); return TS_ERR_INVALID; }
  for (int l = 0; l < h->nlist; ++l) out_host[l] = (*h->off_host)[(size_t)l + 1] - (*h->off_host)[(size_t)l];
  return TS_OK;
}

// query prep + coarse step for queries [b0, b0 + Bc) already on the device
static int ivf_probe(ts_ivf* h, const void* q_dev, int q_dtype, int Bc, int nprobe, unsigned flags, cudaStream_t st) {
  const ts_index* base = h->base;
  const int norm = (flags & TS_FLAG_NORMALIZE_Q) ? kNormStage1 : kNormNone;
  int rc;
  if ((rc = ensure_bytes(&h->q32, &h->q32_b, (size_t)Bc * h->ldc * 4))) return rc;
  if ((rc = ensure_bytes(&h->qst, &h->qst_b, (size_t)Bc * base->ld * dtype_size(base->dtype)))) return rc;
  if ((rc = ensure_bytes(&h->probe, &h->probe_b, (size_t)Bc * nprobe * 8))) return rc;
  if ((rc = launch_convert_rows(q_dev, q_dtype, base->dim, h->q32, TS_F32, h->ldc, Bc, base->dim, norm, nullptr, st))) return rc;
  if ((rc = launch_convert_rows(q_dev, q_dtype, base->dim, h->qst, base->dtype, base->ld, Bc, base->dim, norm, nullptr, st))) return rc;
  const int npow2 = next_pow2(h->nlist < 2 ? 2 : h->nlist);
  const size_t smem = (size_t)npow2 * 8 + (size_t)h->ldc * 4;
  if (smem > 48 * 1024) TS_CUDA_OK(cudaFuncSetAttribute(ivf_coarse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int32_t* probe = (int32_t*)h->probe;
  float* pscore = (float*)((char*)h->probe + (size_t)Bc * nprobe * 4);
  TS_LAUNCH(ivf_coarse_kernel, Bc, kIvfThreads, smem, st, (const float*)h->q32, h->ldc, (const float*)h->cent, h->nlist, npow2, nprobe, probe, pscore);
  TS_CUDA_OK(cudaGetLastError());
  h->launches += 3;
  return TS_OK;
}

static int ivf_check_search(ts_ivf* h, const void* q, int q_dtype, int B, int k, int* nprobe, const void* o1, const void* o2, void* stream) {
  if (!h || !q || !o1 || !o2 || B <= 0) { set_error("ts_ivf_search: invalid argument"); return TS_ERR_INVALID; }
  if (k <= 0 || k > TS_MAX_K) { set_error("ts_ivf_search: k=%d outside 1..%d", k, TS_MAX_K); return TS_ERR_INVALID; }
  if (q_dtype != TS_F32 && q_dtype != h->base->dtype) { set_error("ts_ivf_search: query dtype must be f32 or the storage dtype"); return TS_ERR_INVALID; }
  if (!h->trained) { set_error("ivf: no centroids yet (train / ts_ivf_set_centroids first)"); return TS_ERR_INVALID; }
  if (h->base->n == 0) { set_error("No documents indexed. Call add_documents() first."); return TS_ERR_EMPTY; }
  if (*nprobe < 1) *nprobe = 1;
  if (*nprobe > h->nlist) *nprobe = h->nlist;
  if (h->n_assigned != h->base->n || h->base_gen != h->base->reset_gen) return ts_ivf_sync(h, stream);   // rows added (or index reset) since the last sync
  return TS_OK;
}

int ts_ivf_coarse_host(ts_ivf* h, const void* q_host, int B, int nprobe, unsigned flags, int32_t* out_lists_host,
                       float* out_scores_host, void* stream) {
  if (!h || !q_host || !out_lists_host || !out_scores_host || B <= 0 || B > 65535) { set_error("ts_ivf_coarse_host: invalid argument"); return TS_ERR_INVALID; }
  if (!h->trained) { set_error("ivf: no centroids yet (train / ts_ivf_set_centroids first)"); return TS_ERR_INVALID; }
  if (nprobe < 1 || nprobe > h->nlist) { set_error("ts_ivf_coarse_host: nprobe=%d outside 1..%d", nprobe, h->nlist); return TS_ERR_INVALID; }
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->base->device));
  const size_t qb = (size_t)B * h->base->dim * 4;
  int rc = ensure_bytes(&h->stage, &h->stage_b, qb);
  if (rc) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(h->stage, q_host, qb, cudaMemcpyHostToDevice, st));
  if ((rc = ivf_probe(h, h->stage, TS_F32, B, nprobe, flags, st))) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(out_lists_host, h->probe, (size_t)B * nprobe * 4, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaMemcpyAsync(out_scores_host, (char*)h->probe + (size_t)B * nprobe * 4, (size_t)B * nprobe * 4, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  return TS_OK;
}

int ts_ivf_search(ts_ivf* h, const void* q_dev, int q_dtype, int B, int k, int nprobe, unsigned flags, float* out_scores,
                  int64_t* out_ids, void* stream) {
  int rc = ivf_check_search(h, q_dev, q_dtype, B, k, &nprobe, out_scores, out_ids, stream);
  if (rc) return rc;
  const ts_index* base = h->base;
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(base->device));
  const int chunkB = 1024;
  for (int b0 = 0; b0 < B; b0 += chunkB) {
    const int Bc = (B - b0) < chunkB ? (B - b0) : chunkB;
    if ((rc = ivf_probe(h, (const char*)q_dev + (size_t)b0 * base->dim * dtype_size(q_dtype), q_dtype, Bc, nprobe, flags, st))) return rc;
    // enough CTAs for ~kIvfCtasPerSm per SM; a probed list is cut into at most kIvfMaxSeg segments
    int per_sm = kIvfCtasPerSm;
    if (const char* e = getenv("TS_IVF_CTAS_PER_SM")) { const int v = atoi(e); if (v >= 1 && v <= 32) per_sm = v; }
    int S = (per_sm * base->info.sm_count + Bc * nprobe - 1) / (Bc * nprobe);
    if (S < 1) S = 1;
    if (S > kIvfMaxSeg) S = kIvfMaxSeg;
    const int L = nprobe * S;
    if ((rc = ensure_bytes(&h->partial, &h->partial_b, (size_t)L * Bc * k * 8))) return rc;
    const size_t tmpk = merge_tmp_keys(L, Bc, k);
    if (tmpk) {
      if ((rc = ensure_bytes(&h->tmp0, &h->tmp0_b, tmpk * 8))) return rc;
      if ((rc = ensure_bytes(&h->tmp1, &h->tmp1_b, tmpk * 8))) return rc;
    }
    // batches: order the (query, probe) pairs by list so the CTAs of one list run together (L2 reuse)
    const int n_pairs = Bc * nprobe;
    const int32_t* work = nullptr;
    int launches = 1;
    int max_pairs = kIvfMaxPairs;
    if (const char* e = getenv("TS_IVF_MAXPAIRS")) { const int v = atoi(e); if (v >= 0 && v < max_pairs) max_pairs = v; }   // test knob
    if (Bc > 1 && n_pairs <= max_pairs && !env_on("TS_IVF_NOORDER")) {
      if ((rc = ensure_bytes(&h->work, &h->work_b, (size_t)n_pairs * 4))) return rc;
      const int np2 = next_pow2(n_pairs < 2 ? 2 : n_pairs);
      TS_LAUNCH(ivf_order_kernel, 1, 1024, (size_t)np2 * 8, st, (const int32_t*)h->probe, n_pairs, np2, (int32_t*)h->work);
      TS_CUDA_OK(cudaGetLastError());
      work = (const int32_t*)h->work;
      ++launches;
    }
    base->timer->begin(st);
    switch (base->dtype) {
      case TS_BF16: rc = launch_scan_t<__nv_bfloat16>(h, Bc, k, nprobe, S, work, st); break;
      case TS_F16: rc = launch_scan_t<__half>(h, Bc, k, nprobe, S, work, st); break;
      default: rc = launch_scan_t<float>(h, Bc, k, nprobe, S, work, st); break;
    }
    base->timer->end(st);
    if (rc) return rc;
    if ((rc = launch_merge_keys((const uint64_t*)h->partial, L, Bc, k, base->id_base, (uint64_t*)h->tmp0, (uint64_t*)h->tmp1,
                                out_scores + (size_t)b0 * k, out_ids + (size_t)b0 * k, st, &launches))) return rc;
    h->launches += launches;
  }
  return TS_OK;
}

int ts_ivf_search_host(ts_ivf* h, const void* q_host, int q_dtype, int B, int k, int nprobe, unsigned flags,
                       float* out_scores_host, int64_t* out_ids_host, void* stream) {
  int rc = ivf_check_search(h, q_host, q_dtype, B, k, &nprobe, out_scores_host, out_ids_host, stream);
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  TS_CUDA_OK(cudaSetDevice(h->base->device));
  const size_t qb = (size_t)B * h->base->dim * dtype_size(q_dtype);
  if ((rc = ensure_bytes(&h->stage, &h->stage_b, qb))) return rc;
  const size_t sb = (size_t)B * k * sizeof(float), ib = (size_t)B * k * sizeof(int64_t);
  if ((rc = ensure_bytes(&h->hout, &h->hout_b, sb + ib + 256))) return rc;
  float* ds = (float*)h->hout;
  int64_t* di = (int64_t*)((char*)h->hout + ((sb + 255) / 256) * 256);
  TS_CUDA_OK(cudaMemcpyAsync(h->stage, q_host, qb, cudaMemcpyHostToDevice, st));
  if ((rc = ts_ivf_search(h, h->stage, q_dtype, B, k, nprobe, flags, ds, di, stream))) return rc;
  TS_CUDA_OK(cudaMemcpyAsync(out_scores_host, ds, sb, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaMemcpyAsync(out_ids_host, di, ib, cudaMemcpyDeviceToHost, st));
  TS_CUDA_OK(cudaStreamSynchronize(st));
  return TS_OK;
}

}  // extern "C"
