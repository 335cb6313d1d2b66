// Exact top-k selection over candidate keys (one CTA per query and group).
//
// Replaces, on the device:
//   * FAISS's result heap/reservoir merge inside IndexFlatIP.search
//     (call site /root/reference/src/stage1_retriever.py:380) -- mode KEYS:
//     per-CTA partial top-k lists of the scan kernels -> final [B,k];
//   * the G*k -> k merge after the multi-GPU all-gather --
//     mode PAIRS;
//   * scored_candidates.sort(reverse=True)[:top_k]
//     (/root/reference/src/stage2_rescorer.py:294-297) -- mode RANK.
//
// Bound: latency (a few thousand keys per query in shared memory); HBM
// traffic is L*B*k*8 bytes read once.
#include <stdio.h>

#include "ts_common.cuh"
#include "ts_internal.h"

namespace ts {

namespace {

constexpr int kSelThreads = 512;
constexpr int kSelCap = 4096;  // keys of shared memory per CTA (32 KB)
constexpr int kSelCapLists = 2048;  // kLists after the umma scan: 16 KB, so a select CTA fits on an SM beside a scan CTA (198 KB) and
                                    // programmatic dependent launch can make it resident before the scan ends

enum SelMode { kKeys = 0, kPairs = 1, kRank = 2, kLists = 3 };
constexpr int kMaxLists = 256;

struct SelectParams {
  int mode;
  const uint64_t* keys;   // kKeys: [L][B][k_in]
  const float* scores;    // kPairs: [L][B][k_in]; kRank: [B][C]
  const int64_t* ids;     // kPairs: [L][B][k_in]
  const int32_t* n_cand;  // kRank: [B] or null
  int L, B, k_in, group, k_out, C;
  int final_pass;
  uint64_t* keys_out;     // !final: [n_groups][B][k_out]
  float* out_scores;      // final: [B][k_out]
  int64_t* out_ids;       // final kKeys / kPairs
  int32_t* out_pos;       // final kRank
  int64_t id_base;
  // kLists: unsorted candidate lists of the umma scan, lists[(cta*128 + row)*cap .. +counts[cta*128+row])
  const int* counts;
  int n_slices, spread, cap, rows_per_cta;
  long long pair_stride;  // kPairs: elements between consecutive lists (scores: floats, ids: int64s)
  long long pair_stride_ids;
  const float* pub;       // kLists: final per-slice J-th best scores [n_slices][bpad] (null = no filter)
  int kth_rule;           // kLists: one published best per slice (J == 1) and n_slices >= k: the bests are n_slices DISTINCT
                          // rows, so their k-th largest is a valid -- and far tighter -- bound than their minimum
  int bpad;
  int serial_prefix;      // kLists: 1 = first version of the count prefix / filter loop (TS_SELECT_V1)
  // kPairs after a peer-memory exchange: the lists are written by the other GPUs; flag l holds the sequence
  // number of the step whose list l is complete (null = lists are already there, e.g. after an NCCL all-gather)
  const unsigned int* wait_flags;
  unsigned int wait_seq;
  int wait_flag_stride;   // 0: flag l covers the whole list; > 0: flag of (list l, query b) at l * stride + b
  // final pass of a local search in a multi-GPU step: push the rows into every rank's receive buffer (PushTarget)
  PushTarget push;
  unsigned long long* tl;          // TS_DBG_TIMELINE: CTA 0 stamps [4] entry, [5] local rows sorted, [6] pushed + published, [7] peers' rows in, [8] exit
  int sel_cap;                     // keys of dynamic shared memory this launch has (kSelCap or kSelCapLists)
};

// system-scope flag accesses for the peer-memory exchange (NVLink): the producer's data stores are made
// visible by __threadfence_system() + st.release.sys, the consumer pairs them with ld.acquire.sys
constexpr long long kExchangeTimeoutCycles = 40000000000ll;  // ~20 s: a missing peer traps instead of hanging the GPU; ranks of an
                                                              // SPMD job may reach a step seconds apart (host-side work in between)
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
#ifdef TS_CUDASIM
  return *reinterpret_cast<const volatile unsigned int*>(p);
#else
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#endif
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
#ifdef TS_CUDASIM
  *reinterpret_cast<volatile unsigned int*>(p) = v;
#else
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
#endif
}

// Exchange, producer side: CTA d copies this GPU's packed [B,k] result into slot `rank` of GPU d's
// receive buffer (16-byte stores over NVLink; d == rank is the local copy) and then publishes the step.
__global__ void __launch_bounds__(256)
    exchange_push_kernel(const uint4* __restrict__ blob, int n16, const long long* __restrict__ peer_bases, long long slot_off,
                         long long flag_off, unsigned int seq) {
  char* base = reinterpret_cast<char*>(peer_bases[blockIdx.x]);
  uint4* dst = reinterpret_cast<uint4*>(base + slot_off);
  for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = blob[i];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) st_release_sys(reinterpret_cast<unsigned int*>(base + flag_off), seq);
}

// Exchange, consumer side for the Stage-2 score matrices: wait until all n_ranks slots of this step are
// published, then out[i] = sum over ranks of slot_r[i] (every candidate is owned by exactly one rank, the
// others contribute 0.0 -- the sum is what an all-reduce(SUM) of the per-rank outputs would deliver).
__global__ void __launch_bounds__(256)
    exchange_wait_sum_kernel(const char* __restrict__ slots, long long slot_bytes, const unsigned int* __restrict__ flags,
                             int n_ranks, unsigned int seq, long long n, float* __restrict__ out) {
  if ((int)threadIdx.x < n_ranks) {
    const unsigned int* f = flags + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) != seq) {
      TS_SPIN_YIELD();
      if (clock64() - t0 > kExchangeTimeoutCycles) {
        printf("[tristage] exchange timeout: rank %d never published step %u\n", (int)threadIdx.x, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int r = 0; r < n_ranks; ++r) acc += *reinterpret_cast<const float*>(slots + (size_t)r * slot_bytes + (size_t)i * 4);
    out[i] = acc;
  }
}

// Consumer of the Stage-2 scatter (ts_maxsim_scatter): wait until all n_ranks owners have published the step, then move
// this rank's score matrix out of the receive buffer and leave zeros behind -- entries nobody owns (ids outside every
// shard, positions beyond n_cand) read 0.0 like the local kernel's output, and the buffer is clean when its parity
// comes round again two steps later (no rank can be writing it before this rank's next push, see the parity argument
// in include/tristage.h).
__global__ void __launch_bounds__(256)
    exchange_wait_take_kernel(float* __restrict__ matrix, const unsigned int* __restrict__ flags, int n_ranks, unsigned int seq,
                              long long n, float* __restrict__ out) {
  grid_dep_launch();
  if ((int)threadIdx.x < n_ranks) {
    const unsigned int* f = flags + threadIdx.x;
    const long long t0 = clock64();
    while (ld_acquire_sys(f) != seq) {
      TS_SPIN_YIELD();
      if (clock64() - t0 > kExchangeTimeoutCycles) {
        printf("[tristage] exchange timeout: rank %d never published Stage-2 step %u\n", (int)threadIdx.x, seq);
        __trap();
      }
    }
  }
  __syncthreads();
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    out[i] = __ldcg(matrix + i);
    matrix[i] = 0.f;
  }
  grid_dep_wait();     // launched with programmatic stream serialization behind the scatter kernel: stay ordered behind it
}

// Merge of the exchange: G lists of k (score, id) pairs for query b, each SORTED the way the select kernel writes
// them (key = score, then position, strictly descending; id < 0 marks the unused tail).  A key's final rank is its
// position in its own list plus, for every other list, the number of greater keys there (binary search in shared
// memory): two block barriers instead of the ~55 of a 1024-key bitonic sort (12 -> ~3 us for 8 x 100 on the step
// time line).  Keys are unique (the position is part of the key), so the ranks are a permutation and the result is
// the one the sort gives.  G * k <= 4 * blockDim keys.
__device__ __forceinline__ void rank_merge_sorted(const float* __restrict__ scores, const int64_t* __restrict__ ids,
                                                  long long stride_f, long long stride_i, int G, int k, int b, uint64_t* sbuf,
                                                  float* __restrict__ out_scores, int64_t* __restrict__ out_ids) {
  const int total = G * k;
  uint64_t my_key[4];
  int64_t my_id[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = (int)threadIdx.x + u * (int)blockDim.x;
    my_key[u] = 0ull; my_id[u] = -1;
    if (i < total) {
      const int l = i / k, r = i - l * k;
      const size_t at = (size_t)b * k + r;
      my_id[u] = __ldcg(ids + (size_t)l * stride_i + at);
      if (my_id[u] >= 0) my_key[u] = make_key(__ldcg(scores + (size_t)l * stride_f + at), (uint32_t)i);
      sbuf[i] = my_key[u];
    }
  }
  for (int r = threadIdx.x; r < k; r += blockDim.x) {      // fewer than k valid rows in all lists together: the tail stays padded
    out_scores[(size_t)b * k + r] = kLowestF32;
    out_ids[(size_t)b * k + r] = -1;
  }
  __syncthreads();
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int i = (int)threadIdx.x + u * (int)blockDim.x;
    if (i >= total || my_key[u] == 0ull) continue;
    const int l = i / k;
    int rank = i - l * k;
    for (int l2 = 0; l2 < G; ++l2) {
      if (l2 == l) continue;
      const uint64_t* lst = sbuf + l2 * k;
      int lo = 0, hi = k;                                   // keys greater than mine in list l2 (its invalid tail is 0: never greater)
      while (lo < hi) { const int mid = (lo + hi) >> 1; if (lst[mid] > my_key[u]) lo = mid + 1; else hi = mid; }
      rank += lo;
    }
    if (rank < k) {
      out_scores[(size_t)b * k + rank] = key_score(my_key[u]);
      out_ids[(size_t)b * k + rank] = my_id[u];
    }
  }
}

__global__ void __launch_bounds__(kSelThreads) select_kernel(const SelectParams p) {
  TS_DYN_SMEM(uint64_t, sbuf);
  __shared__ int pre[kMaxLists + 1];
  __shared__ float s_min[kSelThreads / 32];
  __shared__ int s_wsum[kSelThreads / 32];
  __shared__ int s_cnt;
  const int b = blockIdx.x, g = blockIdx.y;
  // programmatic dependent launch: let the next kernel of the stream become resident now; our own inputs come from
  // the predecessor (scan / all-gather) -- except in the wait-merge of the peer exchange, whose inputs are guarded
  // by the peers' flags: it starts merging while this GPU's select kernel is still pushing and only orders itself
  // behind it at the very end (stream order stays transitive)
  grid_dep_launch();
  if (!p.wait_flags) grid_dep_wait();
  const bool tl_on = p.tl && b == 0 && g == 0 && threadIdx.x == 0;
  if (tl_on && !p.wait_flags) p.tl[4] = ts_globaltimer();
  if (p.wait_flags) {
    // one thread per list spins (system-scope acquire) until its producer GPU has published this step
    if ((int)threadIdx.x < p.L) {
      const unsigned int* f = p.wait_flags + (p.wait_flag_stride ? (size_t)threadIdx.x * p.wait_flag_stride + b : (size_t)threadIdx.x);
      const long long t0 = clock64();
      while (ld_acquire_sys(f) != p.wait_seq) {
        TS_SPIN_YIELD();
        if (clock64() - t0 > kExchangeTimeoutCycles) {
          printf("[tristage] exchange timeout: list %d never reached step %u (has %u)\n", (int)threadIdx.x, p.wait_seq, ld_acquire_sys(f));
          __trap();
        }
      }
    }
    __syncthreads();
    if (p.mode == kPairs && p.final_pass && gridDim.y == 1 && p.k_in == p.k_out && p.out_scores &&
        p.L * p.k_in <= p.sel_cap && p.L * p.k_in <= 4 * (int)blockDim.x) {
      // the wait-merge kernel of the two-kernel exchange: the lists are the peers' select outputs, i.e. sorted
      if (tl_on) p.tl[7] = ts_globaltimer();
      rank_merge_sorted(p.scores, p.ids, p.pair_stride, p.pair_stride_ids, p.L, p.k_in, b, sbuf, p.out_scores, p.out_ids);
      if (p.tl && b == 0) { __syncthreads(); if (threadIdx.x == 0) p.tl[8] = ts_globaltimer(); }
      grid_dep_wait();
      return;
    }
  }
  int total;
  const int l0 = g * p.group;
  size_t list_row0 = 0;   // kLists: (first CTA of this query's m-tile)*128 + this query's TMEM lane
  if (p.mode == kLists) {
    const int mt = b >> 7, qi = b & 127;
    int row = qi;
    if (p.spread) row = 64 + qi;       // small batches (B <= 64): the scan keeps query qi in TMEM lane 64 + qi
    list_row0 = (size_t)mt * p.n_slices * p.rows_per_cta + row;   // CTAs (mt, 0..n_slices) own this query tile
    if (p.serial_prefix) {
      // first version (TS_SELECT_V1=1, kept for A/B timing): thread 0 walks the counts -- n_slices
      // dependent global loads, ~20 us of the ~30 us this kernel took at B <= 32
      if (threadIdx.x == 0) {
        int acc = 0;
        for (int l = 0; l < p.n_slices; ++l) { pre[l] = acc; acc += min(max(p.counts[list_row0 + (size_t)l * p.rows_per_cta], 0), p.cap); }
        pre[p.n_slices] = acc;
      }
    } else {
      // one thread per list loads its count (all loads in flight at once), then a block-wide
      // exclusive scan: shuffle scan inside each warp + the totals of the warps before it
      const int t = threadIdx.x;
      const int c = (t < p.n_slices) ? min(max(__ldcg(p.counts + list_row0 + (size_t)t * p.rows_per_cta), 0), p.cap) : 0;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((t & 31) >= o) incl += v;
      }
      if ((t & 31) == 31) s_wsum[t >> 5] = incl;
      __syncthreads();
      if (t < p.n_slices) {                      // n_slices <= kMaxLists = 256: at most 8 warps carry counts
        int base = 0;
        for (int w = 0; w < (t >> 5); ++w) base += s_wsum[w];
        pre[t] = base + incl - c;
        if (t == p.n_slices - 1) pre[p.n_slices] = base + incl;
      }
    }
    __syncthreads();
    total = pre[p.n_slices];
  } else if (p.mode == kRank) {
    total = p.n_cand ? min(max(p.n_cand[b], 0), p.C) : p.C;
  } else {
    total = (min(p.L, l0 + p.group) - l0) * p.k_in;
  }
  auto load = [&](int i) -> uint64_t {
    if (p.mode == kLists) {
      int lo = 0, hi = p.n_slices;            // largest l with pre[l] <= i
      while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pre[mid] <= i) lo = mid; else hi = mid; }
      return __ldcg(p.keys + (list_row0 + (size_t)lo * p.rows_per_cta) * p.cap + (i - pre[lo]));
    } else if (p.mode == kKeys) {
      const int l = l0 + i / p.k_in, r = i % p.k_in;
      return p.keys[((size_t)l * p.B + b) * p.k_in + r];
    } else if (p.mode == kPairs) {
      const int l = l0 + i / p.k_in, r = i % p.k_in;
      const size_t at = ((size_t)b) * p.k_in + r;
      const int64_t id = p.ids[(size_t)l * p.pair_stride_ids + at];
      // position across lists keeps "id ascending" among equal scores when
      // lists are ordered by ascending id range
      return id < 0 ? 0ull : make_key(p.scores[(size_t)l * p.pair_stride + at], (uint32_t)(l * p.k_in + r));
    } else {
      return make_key(p.scores[(size_t)b * p.C + i], (uint32_t)i);
    }
  };
  bool done = false;
  if (p.mode == kLists && p.pub) {
    // Every slice's FINAL J-th best score is known now: at least k rows score
    // >= their minimum, so only keys at or above it can be in the top-k.
    // Filter + compact first (a few hundred survivors), then one small sort.
    float m = INFINITY;
    for (int c = threadIdx.x; c < p.n_slices; c += blockDim.x) m = fminf(m, __ldcg(p.pub + (size_t)c * p.bpad + b));
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_min[threadIdx.x >> 5] = m;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    m = s_min[0];
    for (int w = 1; w < kSelThreads / 32; ++w) m = fminf(m, s_min[w]);
    if (p.kth_rule) {
      // k-th largest of the <= 256 per-slice bests by bisection on their order-preserving integer image: warp 0,
      // 8 values per lane, 32 rounds of "how many are >= candidate" (a rolled loop: ~2 us).  Measured on a B200
      // (1.25 M-row shard, B = 32, k = 100): ~1700 -> ~200 survivors per query, the sort shrinks from 2048 to
      // 256 keys and the kernel from ~45 to ~22 us.  (The same bound inside the SCAN cut its appends 3.5x but
      // not its time -- the scan is not bound by its slow path -- so the scan keeps the plain minimum.)
      __syncthreads();                         // everybody has read s_min[] before thread 0 reuses s_min[0]
      if (threadIdx.x < 32) {
        uint32_t v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = j * 32 + (int)threadIdx.x;
          v[j] = (c < p.n_slices) ? f2ord(__ldcg(p.pub + (size_t)c * p.bpad + b)) : 0u;
        }
        uint32_t key = 0u;
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
          const uint32_t cand = key | (1u << bit);
          int n = 0;
#pragma unroll
          for (int j = 0; j < 8; ++j) n += (v[j] >= cand) ? 1 : 0;
#ifdef TS_CUDASIM
          n = warp_sum_int(n);
#else
          n = __reduce_add_sync(0xffffffffu, n);     // redux.sync: one instruction per round instead of five shuffle + add steps
#endif
          if (n >= p.k_out) key = cand;
        }
        if (threadIdx.x == 0) s_min[0] = key ? fmaxf(m, ord2f(key)) : m;
      }
      __syncthreads();
      m = s_min[0];
    }
    const uint64_t thr = (uint64_t)f2ord(m) << 32;      // smallest key with score m
    if (p.serial_prefix) {
      for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const uint64_t key = load(i);
        if (key >= thr && key != 0ull) {
          const int pos = atomicAdd(&s_cnt, 1);
          if (pos < p.sel_cap) sbuf[pos] = key;
        }
      }
    } else {
      // four independent key loads in flight per thread before any of them is tested
      for (int i0 = threadIdx.x; i0 < total; i0 += 4 * blockDim.x) {
        uint64_t key[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + u * (int)blockDim.x;
          key[u] = (i < total) ? load(i) : 0ull;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (key[u] >= thr && key[u] != 0ull) {
            const int pos = atomicAdd(&s_cnt, 1);
            if (pos < p.sel_cap) sbuf[pos] = key[u];
          }
        }
      }
    }
    __syncthreads();
    const int kept = s_cnt;
    if (kept <= p.sel_cap) {
      const int n = next_pow2(max(max(kept, p.k_out), 2));
      for (int i = kept + threadIdx.x; i < n; i += blockDim.x) sbuf[i] = 0ull;
      __syncthreads();
      block_sort_desc(sbuf, n);
      done = true;
    }
    __syncthreads();
  }
  if (!done) block_select_topk(sbuf, p.sel_cap, p.k_out, total, load);
  // sbuf[0..k_out) sorted descending, zero padded
  if (tl_on) p.tl[p.wait_flags ? 7 : 5] = ts_globaltimer();      // (wait-merge kernel: its peers' rows were in before the sort)
  for (int r = threadIdx.x; r < p.k_out; r += blockDim.x) {
    const uint64_t key = sbuf[r];
    if (!p.final_pass) {
      p.keys_out[((size_t)g * p.B + b) * p.k_out + r] = key;
      continue;
    }
    const size_t o = (size_t)b * p.k_out + r;
    float sc = kLowestF32;
    int64_t id = -1;
    if (key != 0ull) {
      sc = key_score(key);
      const uint32_t idx = key_idx(key);
      if (p.mode == kKeys || p.mode == kLists) {
        id = p.id_base + (int64_t)idx;
      } else if (p.mode == kPairs) {
        const int l = idx / p.k_in, r2 = idx % p.k_in;
        id = p.ids[(size_t)l * p.pair_stride_ids + (size_t)b * p.k_in + r2];
      } else {
        id = (int64_t)idx;
      }
    }
    if (p.out_scores) {
      p.out_scores[o] = sc;
      if (p.mode == kRank) p.out_pos[o] = (int32_t)id; else p.out_ids[o] = id;
    }
    if (p.push.peer_bases) {
      for (int d = 0; d < p.push.n_ranks; ++d) {
        char* base = reinterpret_cast<char*>(p.push.peer_bases[d]);
        reinterpret_cast<float*>(base + p.push.scores_off)[o] = sc;
        reinterpret_cast<int64_t*>(base + p.push.ids_off)[o] = id;
      }
    }
  }
  if (p.final_pass && p.push.peer_bases) {
    // publish: this query's row is complete in every rank's buffer.  The CTA barrier orders every thread's peer stores
    // before the flag threads' release, and a release is cumulative over what happens-before it (the construction a
    // cooperative-groups grid sync relies on): no system-scope fence per storing thread (it cost ~5 us per step here).
    __syncthreads();
    if ((int)threadIdx.x < p.push.n_ranks) {
      char* base = reinterpret_cast<char*>(p.push.peer_bases[threadIdx.x]);
      st_release_sys(reinterpret_cast<unsigned int*>(base + p.push.flags_off) + b, p.push.seq);
    }
    if (tl_on) p.tl[6] = ts_globaltimer();
  }
  if (p.final_pass && p.push.peer_bases && p.push.merge_scores) {
    // ---- fused wait + merge: the exchange of a multi-GPU step in ONE kernel.  This CTA has pushed query b's row to
    // every rank; now it waits for the G rows of query b that the other ranks' CTAs b push into THIS rank's buffer
    // (every rank pushes before it waits, and the launcher only fuses when all B CTAs are co-resident: no cycle), and
    // merges G * k -> k exactly like the separate wait-merge kernel (key = score, then position across lists).
    const int G = p.push.n_ranks, k = p.k_out;
    if ((int)threadIdx.x < G) {
      const unsigned int* f = p.push.merge_flags + (size_t)threadIdx.x * p.push.merge_flag_stride + b;
      const long long t0 = clock64();
      while (ld_acquire_sys(f) != p.push.seq) {
        TS_SPIN_YIELD();
        if (clock64() - t0 > kExchangeTimeoutCycles) {
          printf("[tristage] exchange timeout: rank %d never pushed query %d of step %u\n", (int)threadIdx.x, b, p.push.seq);
          __trap();
        }
      }
    }
    __syncthreads();                 // flags seen by the pollers, and every thread is done reading sbuf[0..k)
    if (tl_on) p.tl[7] = ts_globaltimer();
    rank_merge_sorted(p.push.merge_scores, p.push.merge_ids, p.push.merge_stride_f, p.push.merge_stride_i, G, k, b, sbuf,
                      p.push.merge_out_scores, p.push.merge_out_ids);
  }
  if (p.tl && b == 0 && g == 0) { __syncthreads(); if (threadIdx.x == 0) p.tl[8] = ts_globaltimer(); }
  if (p.wait_flags) grid_dep_wait();
}

int launch_select(SelectParams p, int n_groups, cudaStream_t st, bool pdl = false) {
  dim3 grid(p.B, n_groups);
  if (p.sel_cap == 0) p.sel_cap = kSelCap;
  if (pdl && env_flag("TS_PDL", kDefaultPdl)) TS_LAUNCH_PDL(select_kernel, grid, kSelThreads, p.sel_cap * sizeof(uint64_t), st, p);
  else TS_LAUNCH(select_kernel, grid, kSelThreads, p.sel_cap * sizeof(uint64_t), st, p);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

}  // namespace

size_t merge_tmp_keys(int L, int B, int k) {
  const int room = kSelCap - k;
  if ((long long)L * k <= room) return 0;
  const int group = max(2, room / k);
  return (size_t)((L + group - 1) / group) * B * k;
}

int launch_merge_keys(const uint64_t* keys, int L, int B, int k, int64_t id_base, uint64_t* tmp0, uint64_t* tmp1,
                      float* out_scores, int64_t* out_ids, cudaStream_t st, int* launches, const PushTarget* push) {
  if (k <= 0 || k > TS_MAX_K || B <= 0 || L <= 0) { set_error("merge: bad L/B/k"); return TS_ERR_INVALID; }
  const int room = kSelCap - k;
  const uint64_t* cur = keys;
  uint64_t* bufs[2] = {tmp0, tmp1};
  int which = 0;
  while ((long long)L * k > room) {
    const int group = max(2, room / k);
    const int ng = (L + group - 1) / group;
    if (!bufs[which]) { set_error("merge: missing tmp buffer"); return TS_ERR_INVALID; }
    SelectParams p{};
    p.mode = kKeys; p.keys = cur; p.L = L; p.B = B; p.k_in = k; p.group = group; p.k_out = k;
    p.final_pass = 0; p.keys_out = bufs[which];
    int rc = launch_select(p, ng, st);
    if (rc) return rc;
    if (launches) ++*launches;
    cur = bufs[which];
    which ^= 1;
    L = ng;
  }
  SelectParams p{};
  p.mode = kKeys; p.keys = cur; p.L = L; p.B = B; p.k_in = k; p.group = L; p.k_out = k;
  p.final_pass = 1; p.out_scores = out_scores; p.out_ids = out_ids; p.id_base = id_base;
  if (push) p.push = *push;
  int rc = launch_select(p, 1, st);
  if (rc) return rc;
  if (launches) ++*launches;
  return TS_OK;
}

int launch_merge_lists(const uint64_t* lists, const int* counts, const float* pub, const UmmaLayout& lay, int B, int k,
                       int64_t id_base, float* out_scores, int64_t* out_ids, cudaStream_t st, int* launches,
                       const PushTarget* push) {
  if (k <= 0 || k > TS_MAX_K || B <= 0 || lay.n_slices > kMaxLists) { set_error("merge_lists: bad arguments"); return TS_ERR_INVALID; }
  SelectParams p{};
  p.mode = kLists; p.keys = lists; p.counts = counts; p.n_slices = lay.n_slices; p.spread = lay.spread; p.cap = lay.cap;
  p.rows_per_cta = lay.rows_per_cta;
  p.pub = (lay.jrank > 0) ? pub : nullptr; p.bpad = lay.bpad;
  p.kth_rule = (lay.jrank == 1 && lay.kth_rule) ? 1 : 0;
  p.serial_prefix = env_on("TS_SELECT_V1") ? 1 : 0;
  p.tl = lay.tl;
  p.L = lay.n_slices; p.B = B; p.k_in = k; p.group = lay.n_slices; p.k_out = k;
  p.final_pass = 1; p.out_scores = out_scores; p.out_ids = out_ids; p.id_base = id_base;
  if (push) p.push = *push;
  p.sel_cap = (k <= 128) ? kSelCapLists : kSelCap;   // k = 500 (J = 4 bound): a few thousand keys pass the filter, keep the 32 KB buffer
  // programmatic dependent launch only with the small buffer: a 32 KB select CTA cannot sit beside a scan CTA, and at
  // k = 500 the pre-launched kernel measured +55 us on whatever follows it in the stream (graph replay, D2H of the host call)
  int rc = launch_select(p, 1, st, /*pdl=*/p.sel_cap == kSelCapLists);   // the scan executes griddepcontrol.launch_dependents at its start
  if (rc) return rc;
  if (launches) ++*launches;
  return TS_OK;
}

int launch_merge_pairs(const float* scores, const int64_t* ids, long long stride_scores, long long stride_ids, int L,
                       int B, int k, float* out_scores, int64_t* out_ids, cudaStream_t st) {
  if (k <= 0 || k > TS_MAX_K || B <= 0 || L <= 0) { set_error("merge: bad L/B/k"); return TS_ERR_INVALID; }
  if ((long long)L * k > 65536) { set_error("merge: n_lists*k too large (%d*%d)", L, k); return TS_ERR_UNSUPPORTED; }
  SelectParams p{};
  p.mode = kPairs; p.scores = scores; p.ids = ids; p.L = L; p.B = B; p.k_in = k; p.group = L; p.k_out = k;
  p.pair_stride = stride_scores; p.pair_stride_ids = stride_ids;
  p.final_pass = 1; p.out_scores = out_scores; p.out_ids = out_ids;
  return launch_select(p, 1, st);
}

int launch_exchange_push(const void* blob, long long nbytes, const long long* peer_bases_dev, int n_ranks, long long slot_off,
                         long long flag_off, unsigned int seq, cudaStream_t st) {
  if (!blob || !peer_bases_dev || n_ranks < 1 || nbytes <= 0 || (nbytes & 15) || (slot_off & 15) || (flag_off & 3)) {
    set_error("exchange_push: bad arguments (sizes and offsets must be 16-byte multiples)");
    return TS_ERR_INVALID;
  }
  TS_LAUNCH(exchange_push_kernel, n_ranks, 256, 0, st, (const uint4*)blob, (int)(nbytes / 16), peer_bases_dev, slot_off, flag_off, seq);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

int launch_exchange_wait_sum(const void* slots, long long slot_bytes, const unsigned int* flags, int n_ranks, unsigned int seq,
                             long long n, float* out, cudaStream_t st) {
  if (!slots || !flags || !out || n_ranks < 1 || n_ranks > 256 || n <= 0 || n * 4 > slot_bytes) { set_error("exchange_wait_sum: bad arguments"); return TS_ERR_INVALID; }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 4) blocks = 148 * 4;
  TS_LAUNCH(exchange_wait_sum_kernel, (unsigned)blocks, 256, 0, st, (const char*)slots, slot_bytes, flags, n_ranks, seq, n, out);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

int launch_exchange_wait_take(void* matrix, const unsigned int* flags, int n_ranks, unsigned int seq, long long n, float* out,
                              cudaStream_t st) {
  if (!matrix || !flags || !out || n_ranks < 1 || n_ranks > 256 || n <= 0) { set_error("exchange_wait_take: bad arguments"); return TS_ERR_INVALID; }
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (env_flag("TS_PDL", kDefaultPdl)) TS_LAUNCH_PDL(exchange_wait_take_kernel, (unsigned)blocks, 256, 0, st, (float*)matrix, flags, n_ranks, seq, n, out);
  else TS_LAUNCH(exchange_wait_take_kernel, (unsigned)blocks, 256, 0, st, (float*)matrix, flags, n_ranks, seq, n, out);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

int launch_merge_pairs_wait(const float* scores, const int64_t* ids, long long stride_scores, long long stride_ids, int L, int B,
                            int k, const unsigned int* wait_flags, unsigned int wait_seq, float* out_scores, int64_t* out_ids,
                            cudaStream_t st, int wait_flag_stride) {
  if (k <= 0 || k > TS_MAX_K || B <= 0 || L <= 0 || L > kSelThreads) { set_error("merge: bad L/B/k"); return TS_ERR_INVALID; }
  if ((long long)L * k > 65536) { set_error("merge: n_lists*k too large (%d*%d)", L, k); return TS_ERR_UNSUPPORTED; }
  SelectParams p{};
  p.mode = kPairs; p.scores = scores; p.ids = ids; p.L = L; p.B = B; p.k_in = k; p.group = L; p.k_out = k;
  p.pair_stride = stride_scores; p.pair_stride_ids = stride_ids;
  p.final_pass = 1; p.out_scores = out_scores; p.out_ids = out_ids;
  p.wait_flags = wait_flags; p.wait_seq = wait_seq; p.wait_flag_stride = wait_flag_stride;
  return launch_select(p, 1, st, /*pdl=*/true);      // overlaps this GPU's select + push kernel
}

int launch_rank_desc(const float* scores, const int32_t* n_cand, int B, int C, int top_k, float* out_scores,
                     int32_t* out_pos, cudaStream_t st) {
  if (top_k <= 0 || top_k > TS_MAX_K || B <= 0 || C <= 0) { set_error("rank: bad B/C/top_k"); return TS_ERR_INVALID; }
  SelectParams p{};
  p.mode = kRank; p.scores = scores; p.n_cand = n_cand; p.B = B; p.C = C; p.k_out = top_k; p.L = 1; p.group = 1;
  p.k_in = C; p.final_pass = 1; p.out_scores = out_scores; p.out_pos = out_pos;
  return launch_select(p, 1, st);
}

}  // namespace ts
