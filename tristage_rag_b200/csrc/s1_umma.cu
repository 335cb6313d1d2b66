// Stage-1 tensor path ("S1-umma"): exact inner-product top-k as a
// query-tile x corpus-tile contraction on tcgen05 / TMEM, fed by TMA, with the
// top-k fused into the TMEM epilogue.
//
// Replaces faiss.IndexFlatIP.search for batches
// (/root/reference/src/stage1_retriever.py:380) -- the S = Q X^T matrix is
// never written to HBM.
//
// Roofline: HBM for B <~ 128 (algorithmic bytes = N*ld*2 per call, read once),
// tensor pipe beyond (2*B*N*ld flop).
//
// Decomposition
//   grid = n_mt * n_slices CTAs, one per SM.  CTA (mt, slice) owns query tile mt (128 queries = the M rows of the MMA
//   = TMEM lanes) and corpus tiles of 256 rows (= the N columns of the MMA): its first tile is `slice`, every further
//   tile comes from an atomic counter (dynamic tile schedule, kSchedSlots below).  Per tile and 64-element K chunk, TMA
//   (SWIZZLE_128B) brings A = 128x64 of Q and B = 256x64 of X into a 4-stage shared-memory ring; one thread issues four
//   tcgen05.mma (M128 N256 K16, bf16/fp16 in, fp32 accumulate in TMEM).  Two 256-column TMEM accumulators are double
//   buffered against the epilogue.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA issuer (one lane), warps 2-5 = epilogue;
//   epilogue warp w reads TMEM lane quarter (w % 4) with tcgen05.ld 32x32b.x32, so thread t owns ONE query and walks
//   its 256 scores of the tile.
//
// Fused exact top-k (the [B, N] score matrix never exists):
//   * fast path: per 32-column chunk a thread takes the max of its 32 scores (a tree over eight groups of four) and
//     compares it ONCE with its bound tau; 96-99 % of a query's chunks end here.
//   * slow path: the groups whose maximum passes are visited, their survivors are appended straight from registers to
//     the thread's candidate list in global memory (L2 resident).  Within 32 entries of CAP the whole warp
//     bitonic-sorts the list in registers (warp_prune_list), keeps the best k and raises tau (never needed on the
//     benchmark configurations; exercised by adversarial score orders).
//   * shared bound: each slice c publishes pub[c][q] = the J-th best score it has seen for query q,
//     J = ceil(k / n_slices).  n_slices * J >= k rows score >= min_c pub[c][q], so every thread may drop anything
//     below that minimum.  With J = 1 and n_slices >= k the slices' bests are n_slices DISTINCT rows and their k-th
//     largest is a (much tighter) bound too: CTA q keeps it fresh for query q in tau_g[q] (kth_start / the refresh in
//     the main loop); otherwise slice 0 refreshes the minimum every other tile.
//   * one cooperative launch per search (mode 2, TS_FUSE, default): first tile -> pass 1 (per-slice best only) ->
//     publish -> grid barrier -> first bound -> pass 2 re-drains the same accumulator -> scan.  Modes 0 / 1 are the
//     two-launch form (pre-pass over every slice's first tile, then the scan).
//   * the lists leave the kernel UNSORTED with their counts; topk_select.cu filters them with the final bound, compacts
//     and sorts a few hundred survivors per query.
//
// fp32 storage (the reference's own dtype) runs the same kernel with kind::tf32: the TMA map converts fp32 -> tf32
//   (round to nearest) on the way into shared memory, a K chunk is 32 elements (the same 128-byte swizzle row), one MMA
//   covers 8 of them.  Default for B > 4 (TS_TF32); the CUDA-core scan keeps exact fp32 products for B <= 4.
//
// Small batches (B <= 64): the query tile is ONE right-sized TMA box (8 / 16 / 32 / 64 rows) placed at rows 64.. of the A
//   tile, i.e. the queries sit in TMEM lane quarters 2 and 3 -- epilogue warps 2 and 3, the two that do not share a warp
//   scheduler with the TMA producer (warp 0) or the MMA issuer (warp 1).  The first layout spread the queries over all four
//   quarters in 8-row boxes: every extra small box per K chunk cost ~27 ns of the producer / TMA path (B = 1 3.01 ms,
//   B = 32 (4 boxes) 3.28 ms, B = 64 (8 boxes) 3.82 ms on 10 M x 1024), and one 128-row box with >= 96 out-of-bounds rows
//   cost more still (4.5 ms).
#include <stdlib.h>

#include "ts_common.cuh"
#include "ts_internal.h"
#include "ts_ptx.cuh"

namespace ts {
namespace {

using namespace ts::ptx;

constexpr int kThreads = 192;
constexpr int kTileM = 128, kTileN = 256, kChunkK = 64;   // kChunkK: 16-bit elements per K chunk (one 128-B swizzle row)
constexpr int kChunkBytes = kChunkK * 2;
constexpr int kABytes = kTileM * kChunkBytes;  // 16 KB
constexpr int kBBytes = kTileN * kChunkBytes;  // 32 KB
// operand kinds of the single-CTA kernel (the values are the UMMA A/B format codes)
constexpr int kOpF16 = 0, kOpBF16 = 1, kOpTF32 = 2;
constexpr int kRingBytes = 192 * 1024;         // 4 x (A + B); the CTA-pair kernel: 6 x (A + B/2)
constexpr int kBarBytes = 256;
constexpr int kSmemBytes = kRingBytes + kBarBytes + 1024;  // + alignment slack
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 4;
// Dynamic tile schedule of the scan: a CTA's first tile is `slice` (the fused pre-pass needs one tile per CTA), every
// further tile comes from an atomic counter, so an SM that streams a few percent slower than the others (the
// TS_DBG_TRACE timeline shows +-3.5 % per-SM tile times: 87 us between the median and the last CTA of a 10 M-row
// scan) takes fewer tiles instead of finishing late.  The producer thread draws the tile and hands its number to the
// MMA and epilogue warps through an 8-slot shared-memory ring; it can be at most ~4 tiles ahead of the epilogue
// (4 ring stages + 2 accumulators), so 8 slots never wrap onto an unread entry.
constexpr int kSchedSlots = 8;
static_assert((16 + kSchedSlots) * 8 + kSchedSlots * 4 <= kBarBytes, "barrier area too small");

struct UmmaParams {
  int64_t N;
  int nK, B, k;
  int n_slices, n_tiles;
  int spread, q_box_rows, cap;   // spread: B <= 64, the queries are rows 64.. of the A tile (one q_box_rows-row TMA box)
  int dbg_notopk;
  int mode;        // 0 = threshold pre-pass (first tile of every slice, publishes pub), 1 = scan,
                   // 2 = both in one cooperative launch (grid barrier after the first tile)
  unsigned int* grid_bar;  // arrival counter of the grid barrier (mode 2), zero at launch
  unsigned int* tile_ctr;  // scan modes: next-tile counters [n_mt], zero at launch (null = static round-robin tiles)
  int jrank;       // j = ceil(k / n_slices) if <= 8, else 0 (threshold sharing off)
  int kth_rule;    // jrank == 1 and k <= n_slices <= 256: the slices' bests are n_slices distinct rows, their k-th largest is a bound
  int bpad;        // row pitch of pub
  uint64_t* lists; // [grid][rows_per_cta][cap] candidate keys
  int* counts;     // [grid][rows_per_cta] entries per list (out)
  float* pub;      // [n_slices][bpad] per-slice j-th best score per query
  float* tau_g;    // [bpad] min over slices of pub, refreshed by slice 0
  const float* inv_norm;
  unsigned long long* stats;  // debug: [0] appends [1] prunes [2] slow-path chunks (null = off)
  unsigned long long* trace;  // debug (TS_DBG_TRACE): [grid][kTraceSlots] globaltimer stamps of epilogue warp 2 (null = off)
  unsigned long long* tl;     // debug (TS_DBG_TIMELINE): step time line, [2] = min over CTAs of the entry, [3] = max of the exit
};
constexpr int kTraceSlots = 64;   // [0] entry, [1] pass-1 done, [2] grid barrier passed, [3] exit, [4] tiles, [5] first bound in hand,
                                  // [6] first tile drained, [8 + i] accumulator i ready
__device__ __forceinline__ void trace_stamp(const UmmaParams& p, int slot) {
#ifndef TS_CUDASIM
  if (p.trace && slot < kTraceSlots) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    p.trace[(size_t)blockIdx.x * kTraceSlots + slot] = t;
  }
#endif
}

// per-thread state of one query (one TMEM lane of one accumulator).  Scalars only: an array member indexed by the
// run-time J put the whole struct into local memory (208-byte stack frame, LDL/STL on every access of the epilogue).
struct QState {
  float tau, tjJ, pub_last;
  float t0, t1, t2, t3, t4, t5, t6, t7;   // best 8 scores appended so far, descending (J > 1 only)
  int cnt, q;
  int n_app, n_prune, n_slow;   // debug counters
  bool active;
  uint64_t* lst;
};

__device__ __forceinline__ float min_over_slices(const UmmaParams& p, int q) {
  // 32 independent L2 loads in flight per round: ~5 rounds for 148 slices (the first version issued 8 per round,
  // 19 dependent rounds = ~7 us -- longer than the slack the fused start-up has, see the TS_DBG_TRACE timeline)
  float m = INFINITY;
  int c = 0;
  for (; c + 32 <= p.n_slices; c += 32) {
    float v[32];
#pragma unroll
    for (int u = 0; u < 32; ++u) v[u] = __ldcg(p.pub + (size_t)(c + u) * p.bpad + q);
#pragma unroll
    for (int u = 0; u < 32; ++u) m = fminf(m, v[u]);
  }
  for (; c + 4 <= p.n_slices; c += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcg(p.pub + (size_t)(c + u) * p.bpad + q);
#pragma unroll
    for (int u = 0; u < 4; ++u) m = fminf(m, v[u]);
  }
  for (; c < p.n_slices; ++c) m = fminf(m, __ldcg(p.pub + (size_t)c * p.bpad + q));
  return m;
}

__device__ __forceinline__ void topj_reset(QState& s) {
  s.tjJ = -INFINITY;
  s.t0 = s.t1 = s.t2 = s.t3 = s.t4 = s.t5 = s.t6 = s.t7 = -INFINITY;
}
__device__ __forceinline__ void topj_insert(QState& s, float v, int J) {
#define TS_INS(t) { const float hi = fmaxf(t, v); v = fminf(t, v); t = hi; }
  TS_INS(s.t0) TS_INS(s.t1) TS_INS(s.t2) TS_INS(s.t3) TS_INS(s.t4) TS_INS(s.t5) TS_INS(s.t6) TS_INS(s.t7)
#undef TS_INS
  s.tjJ = J == 1 ? s.t0 : J == 2 ? s.t1 : J == 3 ? s.t2 : J == 4 ? s.t3 : J == 5 ? s.t4 : J == 6 ? s.t5 : J == 7 ? s.t6 : s.t7;
}

// one accumulator (this warp's 32 lanes x ncols columns) -> candidates of this thread's query
__device__ __forceinline__ void drain_acc(const UmmaParams& p, QState& s, uint32_t t_addr, int64_t n0, int ncols,
                                          bool warp_active, bool prepass, int J, int CAP, int lane) {
  if (!warp_active) return;
  for (int c0 = 0; c0 < ncols; c0 += 32) {
    uint32_t r[32];
    tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
    tmem_ld_wait();
    if (s.active) {
      if (p.inv_norm) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
          r[j] = __float_as_uint(__uint_as_float(r[j]) * __ldg(p.inv_norm + n0 + ((c0 + j < ncols) ? c0 + j : 0)));
      }
      if (c0 + 32 > ncols) {        // last, partial chunk of the corpus: columns past the end never win
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if (c0 + j >= ncols) r[j] = __float_as_uint(-INFINITY);
      }
      // fast path: the maximum of the 32 scores, as a tree over eight groups of four (depth 5: this warp is alone
      // on its scheduler, a 32-long dependent chain would cost ~130 cycles per chunk), compared once
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i)
        g[i] = fmaxf(fmaxf(__uint_as_float(r[4 * i]), __uint_as_float(r[4 * i + 1])),
                     fmaxf(__uint_as_float(r[4 * i + 2]), __uint_as_float(r[4 * i + 3])));
      const float m = fmaxf(fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3])), fmaxf(fmaxf(g[4], g[5]), fmaxf(g[6], g[7])));
      const float thr = prepass ? s.tjJ : s.tau;
      if (m > thr) {
        // slow path (a few percent of the chunks per query, but most chunks of a warp while the bound warms up):
        // only the groups whose maximum passes are looked at, their survivors are appended straight from the
        // registers.  (The first version copied the 32 scores to local memory and walked a bit mask: ~1-2 us per
        // chunk, 10 us per pass over the first tile -- TS_DBG_TRACE timeline, profiles/README.md.)
        ++s.n_slow;
        if (!prepass) {
          int cnt = s.cnt;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (g[i] > thr) {
#pragma unroll
              for (int u = 0; u < 4; ++u) {
                const float v = __uint_as_float(r[4 * i + u]);
                if (v > thr) s.lst[cnt++] = make_key(v, (uint32_t)(n0 + c0 + 4 * i + u));
              }
            }
          }
          s.n_app += cnt - s.cnt;
          s.cnt = cnt;
        }
        if (J == 1) {
          s.tjJ = fmaxf(s.tjJ, m);   // the slice's best score is the running maximum: no insertion
        } else if (J > 1 && m > s.tjJ) {
          // scores that enter the slice's own J best (~J * ln(rows / J) times per scan): ONE copy of the insertion
          // code, reached through a bit mask and a local copy of the chunk (32 inlined insertions cost the k = 500
          // scan 12 % -- instruction-cache misses on a path that is taken a few times per tile)
          uint32_t mask = 0;
          float loc[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float v = __uint_as_float(r[j]);
            loc[j] = v;
            mask |= (v > s.tjJ) ? (1u << j) : 0u;
          }
#pragma unroll 1
          while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1;
            const float v = loc[j];
            if (v > s.tjJ) topj_insert(s, v, J);
          }
        }
      }
    }
    if (!prepass) {
      unsigned full = __ballot_sync(0xffffffffu, s.active && s.cnt > CAP - 32);
      while (full) {
        const int L = __ffs(full) - 1;
        full &= full - 1;
        uint64_t* lp = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(s.lst), L));
        const int lc = __shfl_sync(0xffffffffu, s.cnt, L);
        const uint64_t kth = warp_prune_list(lp, lc, p.k, lane, CAP);
        if (lane == L) { s.cnt = p.k; s.tau = fmaxf(s.tau, key_score(kth)); ++s.n_prune; }
      }
    }
  }
}

__device__ __forceinline__ void init_state(const UmmaParams& p, QState& s, int a, int mt0, int quarter, int lane,
                                           int lane_row, int rows_per_cta, int CAP, bool present, int cta_id) {
  const int qi = p.spread ? lane_row - 64 : lane_row;      // small batches live in lane quarters 2 and 3
  s.q = (mt0 + a) * kTileM + qi;
  s.active = present && qi >= 0 && s.q < p.B;
  s.lst = p.lists + ((size_t)cta_id * rows_per_cta + a * kTileM + lane_row) * CAP;
  s.cnt = 0;
  s.n_app = s.n_prune = s.n_slow = 0;
  topj_reset(s);
  s.tau = p.dbg_notopk ? INFINITY : -INFINITY;
  s.pub_last = -INFINITY;
}

// threshold sharing, start of kernel (J > 0)
__device__ __forceinline__ void start_state(const UmmaParams& p, QState& s, int slice, bool prepass) {
  if (!s.active) return;
  if (prepass) {
    if (slice == 0) p.tau_g[s.q] = -INFINITY;   // reset before the scan kernel reads it
  } else {
    // never publish below what the pre-pass already published for this slot
    s.pub_last = __ldcg(p.pub + (size_t)slice * p.bpad + s.q);
    const float m = min_over_slices(p, s.q);
    if (slice == 0) p.tau_g[s.q] = m;
    if (!p.dbg_notopk) s.tau = nextafterf(m, -INFINITY);   // keep rows that tie with the bound
  }
}

// threshold sharing, after every tile of the scan (J > 0)
__device__ __forceinline__ void share_state(const UmmaParams& p, QState& s, int slice, int iter) {
  if (!s.active) return;
  if (s.tjJ > s.pub_last) { p.pub[(size_t)slice * p.bpad + s.q] = s.tjJ; s.pub_last = s.tjJ; }
  float m;
  if (!p.kth_rule && slice == 0 && (iter & 1)) {   // designated refresher of the shared bound (kth_rule: see kth_refresh)
    m = min_over_slices(p, s.q);
    p.tau_g[s.q] = m;
  } else {
    m = __ldcg(p.tau_g + s.q);
  }
  if (!p.dbg_notopk) s.tau = fmaxf(s.tau, nextafterf(m, -INFINITY));
}

__device__ __forceinline__ void finish_state(const UmmaParams& p, QState& s, int a, int slice, bool prepass, int J,
                                             int rows_per_cta, int lane_row, bool present, int cta_id) {
  if (!present) return;
  if (prepass) {
    if (s.active && J > 0) p.pub[(size_t)slice * p.bpad + s.q] = s.tjJ;
  } else {
    // lists stay unsorted: the select kernel merges them (no in-kernel final sort)
    p.counts[(size_t)cta_id * rows_per_cta + a * kTileM + lane_row] = s.active ? s.cnt : 0;
    if (p.stats && s.active) {
      atomicAdd(p.stats + 0, (unsigned long long)s.n_app);
      atomicAdd(p.stats + 1, (unsigned long long)s.n_prune);
      atomicAdd(p.stats + 2, (unsigned long long)s.n_slow);
    }
  }
}

__device__ __forceinline__ void named_bar_sync128(int id) { named_bar_sync(id, 128); }

// Grid-wide barrier for the epilogue threads of a cooperative launch (all CTAs co-resident).  One arrival counter,
// zeroed by the query-prep kernel that precedes every scan (convert_rows.cu), so there is no reset / generation
// protocol: one release-add per CTA, then acquire-polls until the count reaches the grid size.  The first version
// (__threadfence = fence.sc.gpu around an add + exchange + generation word) completed 4-13 us after the last
// arrival while the TMA stream was saturating the memory system (TS_DBG_TRACE) -- longer than the one tile time
// (11 us) the first accumulator can be held without stalling the pipeline.  Bounded spin: a protocol bug traps.
__device__ __forceinline__ void grid_barrier_epilogue(unsigned int* bar, int warp, int lane) {
  named_bar_sync128(1);        // the CTA's pub stores are ordered before the release below (CTA-scope barrier + cumulativity)
  if (warp == 2) {
    // the whole warp polls (one request per load: same address), so it stays converged up to the barrier below
#ifdef TS_CUDASIM
    if (lane == 0) atomicAdd(bar, 1u);
    __syncwarp();
    while (*reinterpret_cast<volatile unsigned int*>(bar) < gridDim.x) TS_SPIN_YIELD();
#else
    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    __syncwarp();
    const long long t0 = clock64();
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
      if (seen < gridDim.x && clock64() - t0 > TS_WAIT_TIMEOUT_CYCLES) {
        if (lane == 0) printf("[tristage] grid barrier timeout: block %d saw %u of %u\n", (int)blockIdx.x, seen, gridDim.x);
        __trap();
      }
    } while (__any_sync(0xffffffffu, seen < gridDim.x));
#endif
  }
  named_bar_sync128(1);
}

// k-th largest of the n_slices (<= 256) published per-slice bests of query q, by one warp: 8 values per lane,
// bisection over their order-preserving integer image (32 rounds of "how many are >= candidate").  The slices are
// disjoint row sets, so k rows score >= this value: a valid bound, and a far tighter one than the minimum of the
// bests (0.4 % of the rows pass it instead of 2 % right after the first tile of a 148-slice scan at k = 100).
__device__ __forceinline__ float warp_kth_of_slices(const UmmaParams& p, int q, int lane) {
  uint32_t v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = j * 32 + lane;
    v[j] = (c < p.n_slices) ? f2ord(__ldcg(p.pub + (size_t)c * p.bpad + q)) : 0u;
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
    n = __reduce_add_sync(0xffffffffu, n);   // redux.sync: one instruction instead of five shuffle + add steps per round
#endif
    if (n >= p.k) key = cand;
  }
  return fmaxf(ord2f(key), -3.0e38f);    // finite: -inf is the "not published yet" value of tau_g
}

// fused start-up with the k-th-of-slices rule: CTA q (q < B) owns the first bound of query q -- its warp 2 computes
// it right after the grid barrier and publishes it in tau_g[q]; every thread then waits for the bound of its own
// query (all CTAs are co-resident, and every owner publishes before it waits, so the waits cannot form a cycle).
__device__ __forceinline__ void kth_start(const UmmaParams& p, QState& s, float published, int warp, int lane) {
  if (warp == 2 && (int)blockIdx.x < p.B) {
    const float kth = warp_kth_of_slices(p, (int)blockIdx.x, lane);
    if (lane == 0) {
#ifdef TS_CUDASIM
      *reinterpret_cast<volatile float*>(p.tau_g + blockIdx.x) = kth;
#else
      asm volatile("st.release.gpu.global.f32 [%0], %1;" ::"l"(p.tau_g + blockIdx.x), "f"(kth) : "memory");
#endif
    }
  }
  if (!s.active) return;
  s.pub_last = published;
  float m;
#ifdef TS_CUDASIM
  while ((m = *reinterpret_cast<volatile float*>(p.tau_g + s.q)) == -INFINITY) TS_SPIN_YIELD();
#else
  const long long t0 = clock64();
  for (;;) {
    asm volatile("ld.acquire.gpu.global.f32 %0, [%1];" : "=f"(m) : "l"(p.tau_g + s.q) : "memory");
    if (m != -INFINITY) break;
    if (clock64() - t0 > TS_WAIT_TIMEOUT_CYCLES) {
      printf("[tristage] first-bound wait timeout: block %d query %d\n", (int)blockIdx.x, s.q);
      __trap();
    }
  }
#endif
  if (!p.dbg_notopk) s.tau = nextafterf(m, -INFINITY);   // keep rows that tie with the bound
}

template <int OP>
__global__ void __launch_bounds__(kThreads, 1)
    s1_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ8,
                   const __grid_constant__ CUtensorMap tmX, const UmmaParams p) {
  constexpr int CK = (OP == kOpTF32) ? kChunkBytes / 4 : kChunkK;   // elements per K chunk
  const int CAP = p.cap;
  constexpr int n_stages = kMaxStages;
  constexpr int stage_bytes = kABytes + kBBytes;
  constexpr int b_off = kABytes;                      // B chunk offset inside a stage
  TS_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes);
  uint64_t* full_bar = bars;                        // [kMaxStages] TMA -> MMA
  uint64_t* empty_bar = bars + kMaxStages;          // [kMaxStages] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kMaxStages;      // [2] MMA -> epilogue (per accumulator)
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2; // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);
  uint64_t* sched_bar = bars + 16;                  // [kSchedSlots] producer -> MMA / epilogue: tile number of iteration i is in sched[i % 8]
  volatile int* sched = reinterpret_cast<volatile int*>(bars + 16 + kSchedSlots);

  grid_dep_launch();   // the select kernel (launched with programmatic stream serialization) may become resident now; it waits for this grid's completion
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt0 = blockIdx.x / p.n_slices, slice = blockIdx.x % p.n_slices;   // mt0: query tile
  // the pre-pass looks at the first tile of the slice only
  const int t_end = (p.mode == 0) ? ((slice + 1 < p.n_tiles) ? slice + 1 : p.n_tiles) : p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    for (int i = 0; i < kSchedSlots; ++i) mbar_init(&sched_bar[i], 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const bool dynamic = (p.mode != 0) && (p.tile_ctr != nullptr);
  // tile of iteration `iter` as the MMA / epilogue warps learn it (-1 = no more tiles)
  auto tile_of = [&](int iter) -> int {
    if (dynamic) {
      mbar_wait(&sched_bar[iter & (kSchedSlots - 1)], (uint32_t)((iter / kSchedSlots) & 1), 7);
      return sched[iter & (kSchedSlots - 1)];
    }
    const int t = slice + iter * p.n_slices;
    return t < t_end ? t : -1;
  };

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer ------
      prefetch_tmap(&tmQ); prefetch_tmap(&tmQ8); prefetch_tmap(&tmX);
      const uint32_t tx = (p.spread ? (uint32_t)p.q_box_rows * 128u : (uint32_t)kABytes) + (uint32_t)kBBytes;
      const uint64_t x_policy = (gridDim.x > (unsigned)p.n_slices) ? kEvictNormal : kEvictFirst;
      int stage = 0; uint32_t phase = 0;
      for (int iter = 0;; ++iter) {
        int t = slice + iter * p.n_slices;
        if (dynamic) {
          if (iter > 0) t = p.n_slices + (int)atomicAdd(p.tile_ctr + mt0, 1u);
          if (t >= t_end) t = -1;
          sched[iter & (kSchedSlots - 1)] = t;
          mbar_arrive(&sched_bar[iter & (kSchedSlots - 1)]);     // release: the tile number is visible to whoever completes the wait
        } else if (t >= t_end) {
          t = -1;
        }
        if (t < 0) break;
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          unsigned char* sA = smem + stage * stage_bytes;
          unsigned char* sB = sA + b_off;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          if (p.spread) {
            tma_load_2d(sA + 64 * 128, &tmQ8, &full_bar[stage], kc * CK, 0, kEvictLast);   // tmQ8: the q_box_rows-row box
          } else {
            tma_load_2d(sA, &tmQ, &full_bar[stage], kc * CK, mt0 * kTileM, kEvictLast);
          }
          tma_load_2d(sB, &tmX, &full_bar[stage], kc * CK, t * kTileN, x_policy);
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer --------
      constexpr uint32_t idesc = (OP == kOpTF32) ? make_idesc_tf32(kTileM, kTileN) : make_idesc_f16(kTileM, kTileN, OP == kOpBF16);
      constexpr int kSteps = kChunkBytes / 32;          // one MMA consumes 32 bytes of K: 16 x 16-bit or 8 x tf32
      auto mma = [](uint32_t d, uint64_t ad, uint64_t bd, uint32_t id, uint32_t accumulate) {
        if constexpr (OP == kOpTF32) umma_tf32_ss(d, ad, bd, id, accumulate); else umma_f16_ss(d, ad, bd, id, accumulate);
      };
      int stage = 0; uint32_t phase = 0;
      for (int iter = 0; tile_of(iter) >= 0; ++iter) {
        const int acc = iter & 1;                       // the two accumulators alternate per tile
        const uint32_t par = (uint32_t)((iter >> 1) & 1);
        mbar_wait(&tempty_bar[acc], par ^ 1u, 2);
        tc_fence_after();
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * stage_bytes);
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = make_desc_kmajor_sw128(a_addr + b_off);
#pragma unroll
          for (int ks = 0; ks < kSteps; ++ks)
            mma(tmem_base + (uint32_t)(acc * kTileN), adesc + ks * kDescKStep, bdesc + ks * kDescKStep, idesc,
                (kc | ks) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == n_stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);      // accumulator complete
      }
    }
  } else {
    // -------------------------------------------------- epilogue ----------
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    constexpr int rows_per_cta = kTileM;
    const bool prepass = (p.mode == 0);
    const bool fused = (p.mode == 2);          // pre-pass + scan in one cooperative launch 
    const int J = p.jrank;
    QState s0;
    init_state(p, s0, 0, mt0, quarter, lane, lane_row, rows_per_cta, CAP, true, (int)blockIdx.x);
    if (J > 0 && !fused) start_state(p, s0, slice, prepass);

    const bool wact0 = __any_sync(0xffffffffu, s0.active);
    const bool tracer = (p.trace != nullptr) && warp == (p.spread ? 2 : 4) && lane == 0;   // the warp that holds query 0 (lane 64 for small batches, else lane 0)
    if (tracer) trace_stamp(p, 0);
    if (p.tl && warp == 4 && lane == 0) atomicMin(p.tl + 2, ts_globaltimer());
    int iter = 0;
    for (;; ++iter) {
      const int t = tile_of(iter);
      if (t < 0) break;
      const int64_t n0 = (int64_t)t * kTileN;
      const int ncols = (int)((p.N - n0) < (int64_t)kTileN ? (p.N - n0) : (int64_t)kTileN);
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int acc = iter & 1;
      mbar_wait(&tfull_bar[acc], (uint32_t)((iter >> 1) & 1), 4);
      tc_fence_after();
      if (tracer) trace_stamp(p, 8 + iter);
      if (fused && iter == 0) {
        // first tile, pass 1: only the thread's J best scores -> publish -> wait for every slice
        drain_acc(p, s0, lane_addr + (uint32_t)(acc * kTileN), n0, ncols, wact0, true, J, CAP, lane);
        const float published = s0.tjJ;
        if (s0.active) {
          p.pub[(size_t)slice * p.bpad + s0.q] = published;
          if (slice == 0) p.tau_g[s0.q] = -INFINITY;   // never let a stale bound of an earlier call be read
        }
        if (tracer) trace_stamp(p, 1);
        grid_barrier_epilogue(p.grid_bar, warp, lane);
        if (tracer) trace_stamp(p, 2);
        // pass 2 re-reads the same accumulator with the shared bound in place; the J-best
        // registers restart from scratch so no row is counted twice
        topj_reset(s0);
        if (p.kth_rule) kth_start(p, s0, published, warp, lane);
        else start_state(p, s0, slice, false);
        if (tracer) trace_stamp(p, 5);
      }
      drain_acc(p, s0, lane_addr + (uint32_t)(acc * kTileN), n0, ncols, wact0, prepass, J, CAP, lane);
      if (tracer && iter == 0) trace_stamp(p, 6);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if (!prepass && J > 0) {
        // k-th-of-slices rule: CTA q keeps the bound of query q fresh (every other tile, one warp, ~1.5 us), so the
        // refresh work is spread over B CTAs instead of slice 0 recomputing B minima
        if (p.kth_rule && (iter & 1) && warp == 2 && (int)blockIdx.x < p.B) {
          const float kth = warp_kth_of_slices(p, (int)blockIdx.x, lane);
          if (lane == 0) p.tau_g[blockIdx.x] = kth;
        }
        share_state(p, s0, slice, iter);
      }
    }
    finish_state(p, s0, 0, slice, prepass, J, rows_per_cta, lane_row, true, (int)blockIdx.x);
    if (tracer) { trace_stamp(p, 3); p.trace[(size_t)blockIdx.x * kTraceSlots + 4] = (unsigned long long)iter; }
    if (p.tl && warp == 4 && lane == 0 && p.mode != 0) atomicMax(p.tl + 3, ts_globaltimer());
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ---------------------------------------------------------------------------
// CTA-pair variant (TS_PAIR, default for B >= 129; bit-equal to the single-CTA scan on hardware, x1.12 at B = 256,
// x1.03-1.09 at B = 1024 -- profiles/README.md).
//
// Why: for B >= 256 the scan is bound by L2->SM throughput (each CTA pulls 16 KB of Q and 32 KB of X
// per K chunk: ~12.5 TB/s chip-wide at the measured rate, the LTS cap).  A cluster of two CTAs on one
// TPC takes two query tiles (mt = 2*pair + rank) of the SAME slice and runs ONE tcgen05.mma
// cta_group::2 (M 256 x N 256 x K 16) per step: every CTA loads its own 128 x 64 of Q and only HALF
// of the corpus chunk (rows t*256 + 128*rank ..), 32 KB per chunk instead of 48 KB; the tensor cores
// read the other half from the partner's shared memory.  Each CTA's TMEM still holds 128 queries x
// 256 corpus rows per accumulator, double buffered, so the epilogue and the fused top-k are the
// single-CTA ones.  Stages shrink to 32 KB: a 6-deep ring.
//
// Protocol (per stage s, accumulator a):
//   full[s]   leader's barrier, count 2: the leader's producer arrives with expect_tx(2 x 32 KB), the
//             partner's producer arrives remotely; BOTH CTAs' TMA loads credit the leader's barrier;
//   empty[s]  one per CTA, count 1: the leader's tcgen05.commit multicasts to both;
//   tfull[a]  one per CTA, count 1: commit multicast after the last K chunk;
//   tempty[a] leader's barrier, count 8: the four epilogue warps of both CTAs (partner: remote arrive).
constexpr int kPairStages = 6;
constexpr int kPairStageBytes = kABytes + kBBytes / 2;   // 32 KB
static_assert(kPairStages * kPairStageBytes <= kRingBytes, "pair ring must fit the ring area");

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
    s1_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmXh, const UmmaParams p) {
  const int CAP = p.cap;
  TS_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kRingBytes);
  uint64_t* full_bar = bars;                           // [kPairStages]
  uint64_t* empty_bar = bars + kPairStages;            // [kPairStages]
  uint64_t* tfull_bar = bars + 2 * kPairStages;        // [2]
  uint64_t* tempty_bar = bars + 2 * kPairStages + 2;   // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kPairStages + 4);

  grid_dep_launch();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1) / p.n_slices, slice = (int)(blockIdx.x >> 1) % p.n_slices;
  const int mt = 2 * pair + (int)rank;
  const int cta_id = mt * p.n_slices + slice;          // logical id: the list layout select_kernel expects
  const int t_end = (p.mode == 0) ? ((slice + 1 < p.n_tiles) ? slice + 1 : p.n_tiles) : p.n_tiles;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kPairStages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 8); }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_2sm(tmem_slot, kTmemCols);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync();                                      // both CTAs' barriers exist before any remote arrive / TMA credit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      // ------------------------------------------------ TMA producer (both CTAs) ------
      prefetch_tmap(&tmQ); prefetch_tmap(&tmXh);
      int stage = 0; uint32_t phase = 0;
      for (int t = slice; t < t_end; t += p.n_slices) {
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          unsigned char* sA = smem + stage * kPairStageBytes;
          unsigned char* sB = sA + kABytes;
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2u * (uint32_t)kPairStageBytes);
          else mbar_arrive_cluster(&full_bar[stage], 0);
          tma_load_2d_2sm(sA, &tmQ, &full_bar[stage], kc * kChunkK, mt * kTileM, kEvictLast);
          tma_load_2d_2sm(sB, &tmXh, &full_bar[stage], kc * kChunkK, t * kTileN + (int)rank * (kTileN / 2), kEvictNormal);
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 0) {
      // ------------------------------------------------ MMA issuer (leader only) ------
      constexpr uint32_t idesc = make_idesc_f16(2 * kTileM, kTileN, BF16);
      int stage = 0; uint32_t phase = 0;
      int iter = 0;
      for (int t = slice; t < t_end; t += p.n_slices, ++iter) {
        const int acc = iter & 1;
        mbar_wait(&tempty_bar[acc], (uint32_t)((iter >> 1) & 1) ^ 1u, 2);
        tc_fence_after();
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kPairStageBytes);
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = make_desc_kmajor_sw128(a_addr + kABytes);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks)
            umma_f16_ss_2sm(tmem_base + (uint32_t)(acc * kTileN), adesc + ks * kDescKStep, bdesc + ks * kDescKStep, idesc,
                            (kc | ks) ? 1u : 0u);
          umma_commit_2sm(&empty_bar[stage], 3);       // both CTAs' stage is reusable once these MMAs retire
          if (++stage == kPairStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&tfull_bar[acc], 3);           // both CTAs' accumulator halves are complete
      }
    }
  } else {
    // -------------------------------------------------- epilogue (both CTAs) ----------
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    const bool prepass = (p.mode == 0);
    const int J = p.jrank;
    QState s0;
    init_state(p, s0, 0, mt, quarter, lane, lane_row, kTileM, CAP, true, cta_id);
    if (J > 0) start_state(p, s0, slice, prepass);
    const bool wact0 = __any_sync(0xffffffffu, s0.active);
    int iter = 0;
    for (int t = slice; t < t_end; t += p.n_slices, ++iter) {
      const int64_t n0 = (int64_t)t * kTileN;
      const int ncols = (int)((p.N - n0) < (int64_t)kTileN ? (p.N - n0) : (int64_t)kTileN);
      const uint32_t lane_addr = tmem_base + ((uint32_t)(quarter * 32) << 16);
      const int acc = iter & 1;
      mbar_wait(&tfull_bar[acc], (uint32_t)((iter >> 1) & 1), 4);
      tc_fence_after();
      drain_acc(p, s0, lane_addr + (uint32_t)(acc * kTileN), n0, ncols, wact0, prepass, J, CAP, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(&tempty_bar[acc]); else mbar_arrive_cluster(&tempty_bar[acc], 0);
      }
      if (!prepass && J > 0) share_state(p, s0, slice, iter);
    }
    finish_state(p, s0, 0, slice, prepass, J, kTileM, lane_row, true, cta_id);
  }

  tc_fence_before();
  cluster_sync();                                      // nobody leaves while the partner may still read its smem / TMEM
  if (warp == 2) tmem_dealloc_2sm(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ host ---
#ifndef TS_CUDASIM
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess || !p)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}
#endif  // !TS_CUDASIM

}  // namespace

// 2-D row-major [rows][dim] (pitch ld elements) tensor, box = one 128-byte K chunk x box_rows, SWIZZLE_128B.
// fp32 tensors are mapped as TFLOAT32: the copy engine rounds to tf32 on the way into shared memory.
int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int dim, int ld, int box_rows) {
  const int esz = dtype_size(dtype);
#ifdef TS_CUDASIM
  return ptx::sim_make_tmap_2d(out, base, dtype, rows, dim, ld, kChunkBytes / esz, box_rows);
#else
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return TS_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * esz};
  cuuint32_t box[2] = {(cuuint32_t)(kChunkBytes / esz), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType mdt = dtype == TS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16
                                  : dtype == TS_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_TFLOAT32;
  CUresult r = enc(out, mdt, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows %lld dim %d ld %d box %d)", (int)r, (long long)rows, dim, ld, box_rows); return TS_ERR_CUDA; }
  return TS_OK;
#endif
}

namespace {
struct UmmaPlan { int n_mt, n_slices, n_tiles, grid, pair; };
UmmaPlan umma_plan(const ScanArgs& a) {
  UmmaPlan pl;
  pl.n_mt = (a.B + kTileM - 1) / kTileM;
  // (A single-CTA variant with two query tiles per CTA -- TS_DUAL -- halved the L2->SM traffic too, but its two
  // accumulators could not be double buffered against the epilogue: 21.5 vs 19.2 ms at B = 1024, 10 M x 1024.  The
  // CTA-pair kernel below is the design that keeps the double buffering; the variant was removed in round 2.)
  // CTA pairs (cta_group::2): two query tiles per cluster; an odd tile count is padded with an idle tile
  pl.pair = (pl.n_mt >= 2 && a.sm_count >= 2 && a.dtype != TS_F32 && env_flag("TS_PAIR", kDefaultPair)) ? 1 : 0;
  if (pl.pair) pl.n_mt = (pl.n_mt + 1) & ~1;
  pl.n_tiles = (int)((a.n + kTileN - 1) / kTileN);
  int s = a.sm_count / pl.n_mt;
  if (s < 1) s = 1;
  if (s > pl.n_tiles) s = pl.n_tiles;
  pl.n_slices = s;
  pl.grid = pl.n_mt * pl.n_slices;
  return pl;
}
}  // namespace

int s1_umma_plan(const ScanArgs& a, UmmaLayout* lay) {
  if (a.dtype != TS_BF16 && a.dtype != TS_F16 && a.dtype != TS_F32) { set_error("umma path: bad storage dtype %d", a.dtype); return TS_ERR_INVALID; }
  if (a.B > 1024) { set_error("umma path: at most 1024 queries per launch"); return TS_ERR_INVALID; }
  const UmmaPlan pl = umma_plan(a);
  lay->n_slices = pl.n_slices;
  lay->n_mt = pl.n_mt;
  lay->rows_per_cta = kTileM;
  lay->grid = pl.grid;
  lay->cap = cap_for_k(a.k);
  lay->pair = pl.pair;
  lay->spread = (a.B <= 64 && !env_on("TS_DBG_NOSPREAD")) ? 1 : 0;
  lay->bpad = pl.n_mt * kTileM;
  lay->lists_keys = (size_t)pl.grid * lay->rows_per_cta * lay->cap;
  lay->counts_n = (size_t)pl.grid * lay->rows_per_cta;
  lay->pub_n = (size_t)(pl.n_slices + 1) * lay->bpad;   // + one row for tau_g
  const int j = (a.k + pl.n_slices - 1) / pl.n_slices;
  lay->jrank = (j <= 8 && !env_on("TS_DBG_NOSHARE")) ? j : 0;
  // One cooperative launch (pre-pass + grid barrier + scan) instead of two launches: bit-equal to the
  // two-launch sequence on a B200 and x1.02-1.03 on 1.25 M-row shards (profiles/README.md); TS_FUSE=0 for A/B.
  lay->fused = (lay->jrank > 0 && !pl.pair && env_flag("TS_FUSE", kDefaultFuse)) ? 1 : 0;
  lay->kth_rule = (lay->jrank == 1 && lay->n_slices >= a.k && lay->n_slices <= 256 && !env_on("TS_DBG_NOKTH")) ? 1 : 0;
  return TS_OK;
}

int launch_s1_umma(const ScanArgs& a, const UmmaLayout& lay, cudaStream_t st, int* launches) {
  int rc;
  CUtensorMap tmQ, tmQ8, tmX;
  if ((rc = make_tmap_2d(&tmQ, a.q, a.dtype, a.B, a.dim, a.ld, kTileM))) return rc;
  const int q_box_rows = a.B <= 8 ? 8 : a.B <= 16 ? 16 : a.B <= 32 ? 32 : 64;
  if ((rc = make_tmap_2d(&tmQ8, a.q, a.dtype, a.B, a.dim, a.ld, q_box_rows))) return rc;
  if ((rc = make_tmap_2d(&tmX, a.rows, a.dtype, a.n, a.dim, a.ld, kTileN))) return rc;
  UmmaParams p{};
  const int chunk_elems = kChunkBytes / dtype_size(a.dtype);
  p.N = a.n; p.nK = (a.dim + chunk_elems - 1) / chunk_elems; p.B = a.B; p.k = a.k;
  p.n_slices = lay.n_slices; p.n_tiles = (int)((a.n + kTileN - 1) / kTileN);
  p.spread = lay.spread;
  p.q_box_rows = q_box_rows;
  p.cap = lay.cap;
  p.dbg_notopk = env_on("TS_DBG_NOTOPK") ? 1 : 0;
  p.jrank = lay.jrank; p.bpad = lay.bpad; 
  // in the scan every query needs an owner CTA (blockIdx.x == q) for its bound: grid >= B (the select kernel's use of the rule has no such need)
  p.kth_rule = (lay.kth_rule && lay.grid >= a.B && !lay.pair && !env_on("TS_DBG_NOKTHSTART")) ? 1 : 0;
  p.lists = a.lists; p.counts = a.counts; p.pub = a.pub; p.tau_g = a.pub + (size_t)lay.n_slices * lay.bpad;
  p.inv_norm = a.inv_norm; p.tl = a.tl;
  p.tile_ctr = (a.grid_bar && !env_on("TS_DBG_STATIC")) ? a.grid_bar + 1 : nullptr;   // TS_DBG_STATIC=1: round-robin tiles (A/B)
  static unsigned long long* d_stats = nullptr;
  if (env_on("TS_DBG_STATS")) {
    if (!d_stats) cudaMalloc((void**)&d_stats, 32);
    cudaMemsetAsync(d_stats, 0, 32, st);
    p.stats = d_stats;
  }
  static unsigned long long* d_trace = nullptr;
  if (env_on("TS_DBG_TRACE") && !lay.pair) {
    if (!d_trace) cudaMalloc((void**)&d_trace, (size_t)1024 * kTraceSlots * 8);
    cudaMemsetAsync(d_trace, 0, (size_t)1024 * kTraceSlots * 8, st);
    p.trace = d_trace;
  }
  const bool bf16 = a.dtype == TS_BF16;
  if (lay.pair) {
    CUtensorMap tmXh;
    if ((rc = make_tmap_2d(&tmXh, a.rows, a.dtype, a.n, a.dim, a.ld, kTileN / 2))) return rc;
    auto pk = bf16 ? s1_pair_kernel<true> : s1_pair_kernel<false>;
    TS_CUDA_OK(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    for (int mode = (p.jrank > 0 ? 0 : 1); mode <= 1; ++mode) {
      p.mode = mode;
#ifdef TS_CUDASIM
      cudasim::launch_cluster(lay.grid, kThreads, kSmemBytes, 2, [&]() { pk(tmQ, tmXh, p); });
#else
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(lay.grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      cfg.attrs = attr; cfg.numAttrs = 1;
      TS_CUDA_OK(cudaLaunchKernelEx(&cfg, pk, tmQ, tmXh, p));
#endif
      if (launches) ++*launches;
    }
    return TS_OK;
  }
  auto kern = bf16 ? s1_umma_kernel<kOpBF16> : (a.dtype == TS_F16 ? s1_umma_kernel<kOpF16> : s1_umma_kernel<kOpTF32>);
  TS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  if (p.jrank > 0 && lay.fused && a.grid_bar && a.coop) {
    // one cooperative launch (all CTAs co-resident): first tile -> publish -> grid barrier -> scan
    p.mode = 2;
    p.grid_bar = a.grid_bar;
#ifdef TS_CUDASIM
    cudasim::launch_cooperative(lay.grid, kThreads, kSmemBytes, [&]() { kern(tmQ, tmQ8, tmX, p); });
#else
    if (env_on("TS_DBG_NOCOOP")) {
      // measurement only: a plain launch of the same grid (one CTA per SM is co-resident on an idle GPU, but nothing
      // guarantees it) -- shows what the cooperative launch itself costs
      TS_LAUNCH(kern, lay.grid, kThreads, kSmemBytes, st, tmQ, tmQ8, tmX, p);
      TS_CUDA_OK(cudaGetLastError());
    } else {
      void* args[] = {(void*)&tmQ, (void*)&tmQ8, (void*)&tmX, (void*)&p};
      TS_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(lay.grid), dim3(kThreads), args, kSmemBytes, st));
    }
#endif
    if (launches) ++*launches;
  } else {
    if (p.jrank > 0) {
      p.mode = 0;   // threshold pre-pass over the first tile of every slice
      TS_LAUNCH(kern, lay.grid, kThreads, kSmemBytes, st, tmQ, tmQ8, tmX, p);
      TS_CUDA_OK(cudaGetLastError());
      if (launches) ++*launches;
    }
    p.mode = 1;
    TS_LAUNCH(kern, lay.grid, kThreads, kSmemBytes, st, tmQ, tmQ8, tmX, p);
    TS_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
  }
  if (p.trace) {
    // one JSON line per launch on stderr: per-CTA stamps relative to the earliest entry (ns)
    const size_t n = (size_t)lay.grid * kTraceSlots;
    unsigned long long* h = (unsigned long long*)malloc(n * 8);
    cudaMemcpyAsync(h, d_trace, n * 8, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    unsigned long long t0 = ~0ull;
    for (int c = 0; c < lay.grid; ++c) if (h[(size_t)c * kTraceSlots] && h[(size_t)c * kTraceSlots] < t0) t0 = h[(size_t)c * kTraceSlots];
    fprintf(stderr, "[ts trace] {\"B\": %d, \"grid\": %d, \"t0\": %llu, \"ctas\": [", a.B, lay.grid, t0);
    for (int c = 0; c < lay.grid; ++c) {
      const unsigned long long* r = h + (size_t)c * kTraceSlots;
      const int tiles = (int)r[4];
      fprintf(stderr, "%s{\"entry\": %lld, \"pass1\": %lld, \"bar\": %lld, \"bound\": %lld, \"tile0\": %lld, \"exit\": %lld, \"acc\": [", c ? ", " : "",
              (long long)(r[0] - t0), r[1] ? (long long)(r[1] - t0) : -1ll, r[2] ? (long long)(r[2] - t0) : -1ll,
              r[5] ? (long long)(r[5] - t0) : -1ll, r[6] ? (long long)(r[6] - t0) : -1ll, (long long)(r[3] - t0));
      for (int i = 0; i < tiles && 8 + i < kTraceSlots; ++i) fprintf(stderr, "%s%lld", i ? ", " : "", (long long)(r[8 + i] - t0));
      fprintf(stderr, "]}");
    }
    fprintf(stderr, "]}\n");
    free(h);
  }
  if (p.stats) {
    unsigned long long h[4];
    cudaMemcpyAsync(h, d_stats, 32, cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    fprintf(stderr, "[ts stats] B=%d k=%d grid=%d slices=%d pair=%d J=%d: appends=%llu (%.1f/list) prunes=%llu slow_chunks=%llu\n", a.B, a.k, lay.grid,
            lay.n_slices, lay.pair, lay.jrank, h[0], (double)h[0] / ((double)lay.n_slices * a.B), h[1], h[2]);
  }
  return TS_OK;
}

}  // namespace ts
