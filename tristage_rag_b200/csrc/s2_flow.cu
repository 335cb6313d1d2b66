// Stage-2 late-interaction scoring, tensor kernel for tile-layout token shards ("S2-flow").
//
// Replaces the per-candidate loop of ColBERTScorer.rescore_candidates
// (/root/reference/src/stage2_rescorer.py:268-291) around _maxsim_score (:167-183: cosine sim matrix, max
// over doc tokens, MEAN over query tokens) and _colbert_score (:185-201: softmax-weighted sum of the maxima).
//
// Why this kernel looks the way it does (round-2 measurements of the first kernel, profiles/r02_s2_*):
// 0.53 of HBM with the memory system idle half of the time -- the ONE producer warp executed ~700
// instructions per 256-column tile (per-doc shuffles, up to 5 TMA boxes per doc and K chunk, each issued
// through an elect/R2UR loop because UTMALDG takes uniform registers) and the epilogue ~600 per warp and
// tile (masked column maxima, maxima through shared memory, a 5-shuffle sum per doc).  Ring depth, the
// epilogue variants and a resident query tile changed nothing: the kernel was bound by instruction issue of
// its serial roles, not by HBM latency.  So the work per byte had to go down, starting with the layout:
//   * token shards are stored in the TILE layout (TokLayout::kTokTile, tok_ingest.cu): every 8-row group
//     is the un-swizzled K-major core-matrix image tcgen05.mma reads, so a doc -- or any 8-row-aligned part
//     of it -- lands in its columns of a tile with ONE contiguous cp.async.bulk (~25 KB, one HBM burst
//     run) instead of ~8 strided TMA boxes;
//   * pad rows repeat the doc's last token: a pad column can never win a maximum, the epilogue does not mask;
//   * docs flow through full tiles (a doc may split across two tiles), placement is a warp prefix sum over
//     32 candidates, not a per-doc loop;
//   * per-(doc, query token) maxima are combined with shared-memory atomicMax on order-preserving integers
//     into one slot per doc -- across lane quarters AND across tiles, so a split doc needs no carry -- and
//     docs are finalised 32 at a time, one doc per lane, by one of the four epilogue warps in turn;
//   * the query tile (128 rows x dim, SWIZZLE_128B via TMA) is loaded once per (CTA, query).
// Work is split by flat candidate number, [c*T/G, (c+1)*T/G) for CTA c of G.
// Roofline: HBM, algorithmic bytes per candidate = pad8(Ld) * dim * 2.
//
// Roles (320 threads): warp 0 producer, warp 1 one MMA thread (tcgen05.mma M128 x N(used) x K16, two
// 256-column TMEM accumulators), warps 2-9 epilogue (two per TMEM lane quarter: tcgen05.ld, column maxima,
// slot atomics, finalize).
#include <stdlib.h>

#include "ts_common.cuh"
#include "ts_internal.h"
#include "ts_ptx.cuh"

namespace ts {

int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int dim, int ld, int box_rows);

namespace {

using namespace ts::ptx;

constexpr int kEpiWarps = 8;                        // two per TMEM lane quarter: each takes half of the quarter's columns
constexpr int kThreads = 64 + kEpiWarps * 32;
constexpr int kTileM = 128, kChunkK = 64;
constexpr int kAChunkBytes = kTileM * kChunkK * 2;   // 16 KB: one 64-wide K chunk of the query tile (SWIZZLE_128B)
constexpr int kMaxStages = 6;
constexpr int kMetaSlots = 8;
constexpr int kMaxSegs = 32;                         // 256 columns / 8
constexpr int kTmemCols = 512;
constexpr int kAccCols = 256;                        // columns per TMEM accumulator
constexpr int kBatch = 32;                           // docs finalised together (one per lane)
constexpr int kBatches = 3;                          // slot generations in flight (see the deadlock note at gfree)
constexpr int kSlots = kBatch * kBatches;
constexpr int kBarBytes = 512;
constexpr int kTraceSlots = 16;

struct FlowMeta {
  int used;          // columns in use (multiple of 8); 0 = end of work
  int nseg;          // doc segments in this tile
  int flags;         // bit 0: first tile of a query (wait for its query tile)
  int a_buf, a_par;  // query-tile buffer and the parity of its "loaded" barrier
  int fin_upto;      // docs (in stream order) that are complete once this tile is drained
  int lq;            // query tokens of this tile's query
  int pad_;
  uint32_t seg[kMaxSegs];   // first column | doc slot << 16, ascending columns
};
constexpr int kMetaBytes = kMetaSlots * (int)sizeof(FlowMeta);

struct FlowParams {
  const unsigned char* tok;    // tile-layout shard
  const int64_t* doc_off;
  const int32_t* doc_len;
  int64_t ndocs, id_base;
  const int32_t* q_len;
  int B, lq_stride, nK, dim;
  const int64_t* cand;
  const int32_t* n_cand;
  int C, mode;
  int n_stages, a_bufs;
  int tile_shift;              // log2(columns per tile): 8 or 7
  int slot_stride;             // bytes per doc slot: lq_cap * 4 + 16
  int lq_cap;                  // 32 or 128
  float* out;
  unsigned long long* trace;   // [grid][kTraceSlots] cycle counters (TRACE instantiation only)
  // multi-GPU scatter (MaxSimArgs::scatter_*): 0 ranks = off
  const long long* sc_bases;
  int sc_n, sc_rank;
  long long sc_off, sc_flags_off;
  unsigned int sc_seq;
  unsigned int* sc_done;
};

__host__ __device__ constexpr int flow_fixed_bytes(int slot_stride) {
  return kSlots * slot_stride + kSlots * 8 + kMetaBytes + kBarBytes + 1024;
}

#define TS_TR_BEGIN() long long _tr0 = 0; if constexpr (TRACE) _tr0 = clock64()
#define TS_TR_END(acc) do { if constexpr (TRACE) (acc) += (unsigned long long)(clock64() - _tr0); } while (0)

template <bool BF16, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
    maxsim_flow_kernel(const __grid_constant__ CUtensorMap tmQ8, const __grid_constant__ CUtensorMap tmQ32,
                       const __grid_constant__ CUtensorMap tmQ128, const FlowParams p) {
  grid_dep_launch();   // the consumer of a multi-GPU scatter (exchange_wait_take_kernel) may become resident and poll the peers' flags
  TS_DYN_SMEM(unsigned char, smem_raw);
  // offset arithmetic on the __shared__ array keeps the address space known to the compiler (LDS/STS)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kStages = p.n_stages;
  const int row_bytes = p.dim * 2;
  const int tile_cols = 1 << p.tile_shift;
  const int stage_bytes = tile_cols * row_bytes;
  const int a_buf_bytes = p.nK * kAChunkBytes;
  unsigned char* sA = smem;                                           // [a_bufs][nK][128 x 64] SWIZZLE_128B
  unsigned char* sB = smem + p.a_bufs * a_buf_bytes;                  // [stages][tile_cols / 8][dim / 8][8][16 B]
  unsigned char* slots = sB + kStages * stage_bytes;                  // [kSlots][slot_stride]: ord(max) per query token
  int* doc_out = reinterpret_cast<int*>(slots + kSlots * p.slot_stride);   // [kSlots] b*C + j
  int* doc_lq = doc_out + kSlots;                                          // [kSlots]
  FlowMeta* metas = reinterpret_cast<FlowMeta*>(doc_lq + kSlots);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(metas) + kMetaBytes);
  uint64_t* full_bar = bars;                           // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;             // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint64_t* afull_bar = tempty_bar + 2;                // [2]
  uint64_t* gfree_bar = afull_bar + 2;                 // [kBatches]
  uint64_t* mfull_bar = gfree_bar + 4;                 // [kMetaSlots]
  uint64_t* mempty_bar = mfull_bar + kMetaSlots;       // [kMetaSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mempty_bar + kMetaSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t_start = 0;
  if constexpr (TRACE) t_start = clock64();

  for (int i = threadIdx.x; i < (kSlots * p.slot_stride) / 4; i += kThreads) reinterpret_cast<uint32_t*>(slots)[i] = 0u;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], kEpiWarps); mbar_init(&afull_bar[a], 1); }
    for (int g = 0; g < kBatches; ++g) mbar_init(&gfree_bar[g], 1);
    for (int s = 0; s < kMetaSlots; ++s) { mbar_init(&mfull_bar[s], 1); mbar_init(&mempty_bar[s], kEpiWarps); }
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

  if (warp == 0) {
    // ------------------------------------------ producer (whole warp) -----
    if (lane == 0) { prefetch_tmap(&tmQ8); prefetch_tmap(&tmQ32); prefetch_tmap(&tmQ128); }
    unsigned long long tr_wait_empty = 0, tr_wait_meta = 0, tr_wait_a = 0, tr_wait_g = 0;
    // this CTA's candidates: flat numbers [f0, f1) of b*C + j
    const long long T = (long long)p.B * p.C;
    const long long f0 = T * blockIdx.x / gridDim.x, f1 = T * (blockIdx.x + 1) / gridDim.x;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int sh = p.tile_shift;

    uint32_t seq = 0;                       // tiles emitted
    int stage_i = 0; uint32_t stage_ph = 0; // ring position of tile `seq`
    uint32_t n_queries = 0;                 // query tiles loaded
    // last tile that read query buffer 0 / 1: its ordinal + 1, ring stage and fill phase
    uint32_t a_last_seq0 = 0, a_last_seq1 = 0; int a_last_st0 = 0, a_last_st1 = 0; uint32_t a_last_ph0 = 0, a_last_ph1 = 0;
    int b_cur = -1, lq_cur = 1, a_buf = 0, a_par = 0;
    bool a_new = false;
    int ds = 0;                             // docs placed so far (stream order); doc d uses slot d % kSlots
    int cur = 0;                            // column cursor inside the current query's stream
    bool open = false;                      // a tile is open: meta slot and ring stage acquired, copies may land
    FlowMeta* m = nullptr;
    int nseg_open = 0, tile_base = 0;

    auto open_tile = [&]() {
      const uint32_t slot = seq % kMetaSlots;
      {
        TS_TR_BEGIN();
        mbar_wait(&mempty_bar[slot], ((seq / kMetaSlots) & 1u) ^ 1u, 12);
        TS_TR_END(tr_wait_meta);
      }
      {
        TS_TR_BEGIN();
        mbar_wait(&empty_bar[stage_i], stage_ph ^ 1u, 11);
        TS_TR_END(tr_wait_empty);
      }
      m = &metas[slot];
      nseg_open = 0;
      tile_base = (cur >> sh) << sh;
      open = true;
    };
    // publish the open tile: its copies were issued while docs were placed, each with its own expect_tx
    auto emit_tile = [&](int used, int fin_upto) {
      if (lane == 0) {
        m->used = used; m->nseg = nseg_open; m->flags = a_new ? 1 : 0; m->a_buf = a_buf; m->a_par = a_par;
        m->fin_upto = fin_upto; m->lq = lq_cur;
      }
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full_bar[stage_i]);                       // the producer's one arrival: phase ends when the bytes are in
        mbar_arrive(&mfull_bar[seq % kMetaSlots]);             // release the layout to MMA + epilogue
      }
      if (a_buf == 0) { a_last_seq0 = seq + 1; a_last_st0 = stage_i; a_last_ph0 = stage_ph; }
      else { a_last_seq1 = seq + 1; a_last_st1 = stage_i; a_last_ph1 = stage_ph; }
      ++seq;
      if (++stage_i == kStages) { stage_i = 0; stage_ph ^= 1u; }
      a_new = false;
      open = false;
      __syncwarp();
    };
    // a new query: its token rows go into a query-tile buffer once the MMAs that read the buffer's previous
    // occupant have retired (the ring stage of their last tile has been handed back)
    auto start_query = [&](int b, int lq) {
      a_buf = (int)(n_queries % (uint32_t)p.a_bufs);
      a_par = (int)((n_queries / (uint32_t)p.a_bufs) & 1u);
      if (n_queries >= (uint32_t)p.a_bufs) {
        const uint32_t lf = a_buf == 0 ? a_last_seq0 : a_last_seq1;
        if (seq < lf + (uint32_t)kStages) {        // otherwise a later tile already waited for that stage
          TS_TR_BEGIN();
          mbar_wait(&empty_bar[a_buf == 0 ? a_last_st0 : a_last_st1], a_buf == 0 ? a_last_ph0 : a_last_ph1, 13);
          TS_TR_END(tr_wait_a);
        }
      }
      const int q_row = b * p.lq_stride;
      const int a_groups = (lq + 7) >> 3;
      const bool rep4 = lq <= 32;            // query tokens replicated into all four TMEM lane quarters
      const bool a_boxes8 = lq <= 64;
      const uint32_t per_chunk = rep4 ? (uint32_t)kAChunkBytes : a_boxes8 ? (uint32_t)a_groups * 1024u : (uint32_t)kAChunkBytes;
      uint64_t* ab = &afull_bar[a_buf];
      if (lane == 0) mbar_arrive_expect_tx(ab, per_chunk * (uint32_t)p.nK);
      __syncwarp();
      for (int kc = 0; kc < p.nK; ++kc) {
        unsigned char* dA = sA + a_buf * a_buf_bytes + kc * kAChunkBytes;
        const int kx = kc * kChunkK;
        if (rep4) {
          if (lane < 4) tma_load_2d(dA + lane * 4096, &tmQ32, ab, kx, q_row, kEvictLast);
        } else if (a_boxes8) {
          if (lane < a_groups) tma_load_2d(dA + lane * 1024, &tmQ8, ab, kx, q_row + lane * 8, kEvictLast);
        } else {
          if (lane == 0) tma_load_2d(dA, &tmQ128, ab, kx, q_row, kEvictLast);
        }
      }
      ++n_queries;
      b_cur = b; lq_cur = lq; a_new = true; cur = 0;
      __syncwarp();
    };

    // candidate lookup, one per lane; the loads of chunk i+1 are issued before chunk i is laid out
    struct Look { bool valid; long long off; int len; int b; int lq; };
    auto lookup = [&](long long fbase) -> Look {
      Look L{false, 0, 0, -1, 1};
      const long long f = fbase + lane;
      if (f >= f1) return L;
      const int b = (int)(f / p.C), j = (int)(f - (long long)b * p.C);
      const int nc = p.n_cand ? min(max(p.n_cand[b], 0), p.C) : p.C;
      L.b = b;
      L.lq = p.q_len ? min(max(p.q_len[b], 1), min(p.lq_stride, kTileM)) : min(p.lq_stride, kTileM);
      if (j < nc) {
        const int64_t id = p.cand[f] - p.id_base;
        if (id >= 0 && id < p.ndocs) { L.off = p.doc_off[id]; L.len = p.doc_len[id]; L.valid = L.len > 0; }
      }
      return L;
    };
    Look look = lookup(f0);
    for (long long fb = f0; fb < f1; fb += 32) {
      const Look nxt = lookup(fb + 32);
      unsigned vmask = __ballot_sync(0xffffffffu, look.valid);
      while (vmask) {
        // the lanes of ONE query (consecutive flat numbers): a run
        const int src = __ffs(vmask) - 1;
        const int b = __shfl_sync(0xffffffffu, look.b, src);
        const int lq = __shfl_sync(0xffffffffu, look.lq, src);
        const unsigned run = __ballot_sync(0xffffffffu, look.valid && look.b == b) & vmask;
        vmask &= ~run;
        const bool mine = (run >> lane) & 1u;
        if (b != b_cur) {
          if (open) emit_tile(cur - tile_base, ds);
          start_query(b, lq);
        }
        // placement: prefix sum of the padded lengths -> every doc's columns in the query's stream
        const int pad = mine ? ((look.len + 7) & ~7) : 0;
        int incl = pad;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        const int pos = cur + incl - pad, end = cur + incl;
        const int n_run = __popc(run);
        const int my_ds = ds + __popc(run & lt_mask);
        const int my_slot = my_ds % kSlots;
        // slot generations: doc d reuses the slot of doc d - kSlots, which must have been finalised.  A run
        // touches at most one new batch k; batch k - kBatches ended >= 33 docs (>= 264 columns) earlier, so
        // the tile holding its last column is complete and already emitted: the wait cannot deadlock.
        const int kb = (ds + n_run - 1) / kBatch;
        if (kb >= kBatches && (ds == 0 || kb != (ds - 1) / kBatch)) {
          TS_TR_BEGIN();
          mbar_wait(&gfree_bar[kb % kBatches], (uint32_t)((kb / kBatches - 1) & 1), 14);
          TS_TR_END(tr_wait_g);
        }
        if (mine) { doc_out[my_slot] = (int)(fb + lane); doc_lq[my_slot] = lq; }
        const int new_cur = cur + total;
        const int t_first = pos >> sh, t_last = (end - 1) >> sh;
        for (int t = cur >> sh; t <= ((new_cur - 1) >> sh); ++t) {
          if (!open) open_tile();
          const int tile_lo = t << sh, tile_hi = tile_lo + tile_cols;
          const bool has = mine && t_first <= t && t <= t_last;
          const unsigned hmask = __ballot_sync(0xffffffffu, has);
          if (has) {
            const int a = pos > tile_lo ? pos : tile_lo, e = end < tile_hi ? end : tile_hi;
            const int col = a - tile_lo;
            const uint32_t bytes = (uint32_t)(e - a) * (uint32_t)row_bytes;
            m->seg[nseg_open + __popc(hmask & lt_mask)] = (uint32_t)col | ((uint32_t)my_slot << 16);
            uint64_t* fbar = &full_bar[stage_i];
            mbar_expect_tx(fbar, bytes);
            bulk_load(sB + stage_i * stage_bytes + col * row_bytes,
                      p.tok + (size_t)(look.off + (a - pos)) * (size_t)row_bytes, bytes, fbar, kEvictFirst);
          }
          nseg_open += __popc(hmask);
          if (new_cur >= tile_hi) {
            // docs of this run ending inside (or before) this tile are complete once it is drained
            const int done = ds + __popc(run & __ballot_sync(0xffffffffu, mine && end <= tile_hi));
            cur = tile_hi;
            emit_tile(tile_cols, done);
          }
        }
        cur = new_cur;
        ds += n_run;
      }
      look = nxt;
    }
    if (open) emit_tile(cur - tile_base, ds);
    // end-of-work sentinel: carries the number of docs placed, the epilogue finalises what is left
    {
      const uint32_t slot = seq % kMetaSlots;
      mbar_wait(&mempty_bar[slot], ((seq / kMetaSlots) & 1u) ^ 1u, 15);
      if (lane == 0) { metas[slot].used = 0; metas[slot].nseg = 0; metas[slot].fin_upto = ds; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mfull_bar[slot]);
    }
    if constexpr (TRACE) {
      if (lane == 0) {
        unsigned long long* t = p.trace + (size_t)blockIdx.x * kTraceSlots;
        t[0] = tr_wait_empty; t[1] = tr_wait_meta; t[2] = tr_wait_a + tr_wait_g; t[3] = seq; t[4] = (unsigned long long)ds;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer --------
      unsigned long long tr_meta = 0, tr_tempty = 0, tr_full = 0, tr_afull = 0;
      int stage_i = 0; uint32_t stage_ph = 0;
      int acc = 0; uint32_t acc_phase = 0;
      const int n_ks = p.dim >> 4;                         // K = 16 per instruction
      const uint32_t sbo = (uint32_t)p.dim * 16u;          // bytes between 8-row groups of the tile layout
      for (uint32_t seq = 0;; ++seq) {
        const uint32_t slot = seq % kMetaSlots;
        {
          TS_TR_BEGIN();
          mbar_wait(&mfull_bar[slot], (seq / kMetaSlots) & 1u, 21);
          TS_TR_END(tr_meta);
        }
        const FlowMeta* mm = &metas[slot];
        const int used = mm->used;
        if (used == 0) break;
        const int abuf = mm->a_buf;
        if (mm->flags & 1) {
          TS_TR_BEGIN();
          mbar_wait(&afull_bar[abuf], (uint32_t)mm->a_par, 24);
          TS_TR_END(tr_afull);
        }
        const int n_mma = used < 16 ? 16 : ((used + 15) & ~15);
        const uint32_t idesc = make_idesc_f16(kTileM, n_mma, BF16);
        {
          TS_TR_BEGIN();
          mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 22);
          TS_TR_END(tr_tempty);
        }
        {
          TS_TR_BEGIN();
          mbar_wait(&full_bar[stage_i], stage_ph, 23);
          TS_TR_END(tr_full);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kAccCols);
        const uint32_t a_addr0 = smem_u32(sA + abuf * a_buf_bytes);
        const uint32_t b_addr0 = smem_u32(sB + stage_i * stage_bytes);
        for (int ks = 0; ks < n_ks; ++ks) {
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr0 + (uint32_t)((ks >> 2) * kAChunkBytes)) + (uint64_t)(ks & 3) * kDescKStep;
          const uint64_t bdesc = make_desc_kmajor_nosw(b_addr0 + (uint32_t)ks * 256u, 128u, sbo);
          umma_f16_ss(d_tmem, adesc, bdesc, idesc, ks ? 1u : 0u);
        }
        umma_commit(&empty_bar[stage_i]);
        umma_commit(&tfull_bar[acc]);
        if (++stage_i == kStages) { stage_i = 0; stage_ph ^= 1u; }
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
      if constexpr (TRACE) {
        unsigned long long* t = p.trace + (size_t)blockIdx.x * kTraceSlots;
        t[5] = tr_meta; t[6] = tr_tempty; t[7] = tr_full; t[8] = tr_afull;
      }
    }
  } else {
    // -------------------------------------------------- epilogue ----------
    // Two warps per TMEM lane quarter (a warp may only read lanes 32 * (warp % 4) ..): `half` picks the warp's
    // share of the quarter's columns.  lq <= 32: the query tokens sit in ALL four lane quarters, the warps of
    // quarter q drain columns [64q, 64q+64), 32 each.  lq > 32: tokens span the quarters, the two warps of a
    // quarter walk one half of the tile's columns each for its 32 tokens.  A thread's maximum over a doc's
    // columns goes into the doc's slot with atomicMax on an order-preserving integer.
    unsigned long long tr_meta = 0, tr_tfull = 0, tr_drain = 0, tr_fin = 0;
    const int quarter = warp & 3;
    const int ew = warp - 2;                              // 0 .. kEpiWarps-1
    const int half = ew >> 2;
    int acc = 0; uint32_t acc_phase = 0;
    int fin_done = 0;                                     // docs finalised so far (all four warps agree)
    const bool colbert = p.mode != TS_S2_MAXSIM;

    // finalise docs [fin_done, fin_done + n): one doc per lane, by the warp whose turn it is
    auto finalize = [&](int n) {
      named_bar_sync(1, kEpiWarps * 32);                  // every warp's atomics for these docs are done
      const int batch = fin_done / kBatch;
      if ((batch % kEpiWarps) == ew) {
        if (lane < n) {
          const int slot = (fin_done + lane) % kSlots;
          uint32_t* sp = reinterpret_cast<uint32_t*>(slots + slot * p.slot_stride);
          const int lq = doc_lq[slot];
          float res;
          if (!colbert) {
            float sacc = 0.f;
            for (int i = 0; i < p.lq_cap; i += 4) {
              const uint4 v = *reinterpret_cast<const uint4*>(sp + i);
              if (i < lq) sacc += ord2f(v.x);
              if (i + 1 < lq) sacc += ord2f(v.y);
              if (i + 2 < lq) sacc += ord2f(v.z);
              if (i + 3 < lq) sacc += ord2f(v.w);
            }
            res = sacc / (float)lq;
          } else {
            float mx = -INFINITY;
            for (int i = 0; i < lq; ++i) mx = fmaxf(mx, ord2f(sp[i]));
            float z = 0.f, sacc = 0.f;
            for (int i = 0; i < lq; ++i) { const float mv = ord2f(sp[i]); const float e = __expf(mv - mx); z += e; sacc += e * mv; }
            res = sacc / z;
          }
          p.out[doc_out[slot]] = res;
          // multi-GPU: the owner of a candidate stores its score straight into every rank's score matrix over NVLink
          // (consecutive lanes hold consecutive candidates: one 128-byte store per peer and batch).  Every entry has
          // exactly one owner, so there is nothing to reduce -- this replaces the all-reduce(SUM) of the shards' outputs.
          for (int d = 0; d < p.sc_n; ++d)
            reinterpret_cast<float*>(reinterpret_cast<char*>(p.sc_bases[d]) + p.sc_off)[doc_out[slot]] = res;
          uint4 zero; zero.x = zero.y = zero.z = zero.w = 0u;
          for (int i = 0; i < p.lq_cap; i += 4) *reinterpret_cast<uint4*>(sp + i) = zero;
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&gfree_bar[batch % kBatches]);   // the producer may hand these slots out again
      }
      fin_done += n;
    };

    for (uint32_t seq = 0;; ++seq) {
      const uint32_t slot = seq % kMetaSlots;
      {
        TS_TR_BEGIN();
        mbar_wait(&mfull_bar[slot], (seq / kMetaSlots) & 1u, 31);
        TS_TR_END(tr_meta);
      }
      const FlowMeta* m = &metas[slot];
      const int used = m->used;
      const int fin_upto = m->fin_upto;
      if (used == 0) {
        while (fin_done < fin_upto) finalize(min(kBatch, fin_upto - fin_done));
        break;
      }
      const int nseg = m->nseg, lq = m->lq;
      const bool rep4 = lq <= 32;
      const uint32_t my_seg = lane < nseg ? m->seg[lane] : 0xFFFFu;      // col | slot << 16
      const int my_col = (int)(my_seg & 0xFFFFu);
      {
        TS_TR_BEGIN();
        mbar_wait(&tfull_bar[acc], acc_phase, 32);
        TS_TR_END(tr_tfull);
      }
      tc_fence_after();
      long long t_dr = 0;
      if constexpr (TRACE) t_dr = clock64();
      const bool warp_active = rep4 ? true : (quarter * 32 < lq);
      if (warp_active) {
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kAccCols);
        unsigned char* my_tok = slots + (rep4 ? lane : quarter * 32 + lane) * 4;
        const int half_cols = (1 << p.tile_shift) >> 1;
        const int b_lo = rep4 ? quarter * 64 + half * 32 : half * half_cols;
        const int b_hi = rep4 ? min(used, b_lo + 32) : min(used, b_lo + half_cols);
        for (int c_lo = b_lo; c_lo < b_hi; c_lo += 64) {
          const int c_hi = min(b_hi, c_lo + 64);
          // segment holding column c_lo, and the 8-column units of this block that start a new segment
          int s = __popc(__ballot_sync(0xffffffffu, my_col <= c_lo)) - 1;
          const uint32_t starts = warp_or((my_col > c_lo && my_col < c_hi) ? (1u << ((my_col - c_lo) >> 3)) : 0u);
          uint32_t r0[32], r1[32];
          const bool two = c_lo + 32 < c_hi;           // warp-uniform
          tmem_ld_32x32b_x32(t_addr + (uint32_t)c_lo, r0);
          if (two) tmem_ld_32x32b_x32(t_addr + (uint32_t)(c_lo + 32), r1);
          tmem_ld_wait();
          float best = -INFINITY;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            if (c_lo + u * 8 < c_hi) {                  // warp-uniform
              if ((starts >> u) & 1u) {                 // warp-uniform: the next segment starts at this unit
                const int dslot = (int)(__shfl_sync(0xffffffffu, my_seg, s) >> 16);
                atomicMax(reinterpret_cast<uint32_t*>(my_tok + dslot * p.slot_stride), f2ord(best));
                ++s;
                best = -INFINITY;
              }
              float v[8];
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) v[j2] = __uint_as_float(u < 4 ? r0[u * 8 + j2] : r1[(u - 4) * 8 + j2]);
              const float m01 = fmaxf(v[0], v[1]), m23 = fmaxf(v[2], v[3]);
              const float m45 = fmaxf(v[4], v[5]), m67 = fmaxf(v[6], v[7]);
              best = fmaxf(best, fmaxf(fmaxf(m01, m23), fmaxf(m45, m67)));
            }
          }
          const int dslot = (int)(__shfl_sync(0xffffffffu, my_seg, s) >> 16);
          atomicMax(reinterpret_cast<uint32_t*>(my_tok + dslot * p.slot_stride), f2ord(best));
        }
      }
      // accumulator fully read -> hand TMEM back to the MMA warp; meta slot reusable
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(&tempty_bar[acc]); mbar_arrive(&mempty_bar[slot]); }
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      if constexpr (TRACE) { const long long t1 = clock64(); tr_drain += (unsigned long long)(t1 - t_dr); t_dr = t1; }
      while (fin_done + kBatch <= fin_upto) finalize(kBatch);
      if constexpr (TRACE) tr_fin += (unsigned long long)(clock64() - t_dr);
    }
    if constexpr (TRACE) {
      if (warp == 2 && lane == 0) {
        unsigned long long* t = p.trace + (size_t)blockIdx.x * kTraceSlots;
        t[9] = tr_meta; t[10] = tr_tfull; t[11] = tr_drain; t[12] = 0; t[13] = tr_fin;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if (p.sc_n && threadIdx.x == 0) {
    // last CTA of the grid publishes the step in every rank's buffer: the CTA barrier above orders every thread's peer
    // stores before this thread's system-scope fence, each CTA's arrival is ordered behind its fence, and the last
    // CTA's release is cumulative over all of it (the construction a cooperative-groups grid sync relies on)
    __threadfence_system();
    const unsigned int old = atomicAdd(p.sc_done, 1u);
    if (old == gridDim.x - 1) {
      *p.sc_done = 0u;                    // the next launch on this handle is stream-ordered behind this kernel
      __threadfence_system();
      for (int d = 0; d < p.sc_n; ++d) {
        unsigned int* f = reinterpret_cast<unsigned int*>(reinterpret_cast<char*>(p.sc_bases[d]) + p.sc_flags_off) + p.sc_rank;
#ifdef TS_CUDASIM
        *reinterpret_cast<volatile unsigned int*>(f) = p.sc_seq;
#else
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(p.sc_seq) : "memory");
#endif
      }
    }
  }
  if constexpr (TRACE) {
    if (threadIdx.x == 0) p.trace[(size_t)blockIdx.x * kTraceSlots + 14] = (unsigned long long)(clock64() - t_start);
  }
}

}  // namespace

bool maxsim_flow_takes(const MaxSimArgs& a) {
  return a.layout == kTokTile && tok_tile_layout_ok(a.dim, a.dtype) && a.lq_stride >= 1 && a.lq_stride <= TS_S2_MAX_LQ &&
         (long long)a.B * a.C < (1ll << 31);
}

// last trace of a TS_S2_TRACE=1 launch: per-role cycle counters, mean and max over the CTAs
static unsigned long long* g_trace_dev = nullptr;
static int g_trace_grid = 0;

int launch_maxsim_flow(const MaxSimArgs& a, cudaStream_t st, int* launches) {
  CUtensorMap tq8, tq32, tq128;
  int rc;
  const int64_t qrows = (int64_t)a.B * a.lq_stride;
  if ((rc = make_tmap_2d(&tq8, a.q, a.dtype, qrows, a.dim, a.dim, 8))) return rc;
  if ((rc = make_tmap_2d(&tq32, a.q, a.dtype, qrows, a.dim, a.dim, 32))) return rc;
  if ((rc = make_tmap_2d(&tq128, a.q, a.dtype, qrows, a.dim, a.dim, 128))) return rc;
  FlowParams p{};
  p.tok = (const unsigned char*)a.tok;
  p.doc_off = a.doc_off; p.doc_len = a.doc_len; p.ndocs = a.ndocs; p.id_base = a.id_base;
  p.q_len = a.q_len; p.B = a.B; p.lq_stride = a.lq_stride; p.nK = (a.dim + kChunkK - 1) / kChunkK; p.dim = a.dim;
  p.cand = a.cand; p.n_cand = a.n_cand; p.C = a.C; p.mode = a.mode & 0xff;
  p.out = a.out;
  p.sc_bases = a.scatter_bases; p.sc_n = a.scatter_bases ? a.scatter_n : 0; p.sc_rank = a.scatter_rank;
  p.sc_off = a.scatter_off; p.sc_flags_off = a.scatter_flags_off; p.sc_seq = a.scatter_seq; p.sc_done = a.scatter_done;
  p.lq_cap = a.lq_stride <= 32 ? 32 : 128;
  p.slot_stride = p.lq_cap * 4 + 16;        // + 16: consecutive slots start in different 16-byte bank groups
  // shared memory: query tile(s) + ring + slots within the 227 KB opt-in limit
  const int limit = 232448;
  const int fixed = flow_fixed_bytes(p.slot_stride);
  const int row_bytes = a.dim * 2;
  // a CTA meets a new query every C candidates: with short candidate lists a second query-tile buffer hides
  // the reload behind the previous query's last tiles (when the ring can spare the room)
  int a_bufs = (a.C < 128 && p.nK <= 2) ? 2 : 1;
  { const char* e = getenv("TS_S2_ABUFS"); if (e && (atoi(e) == 1 || atoi(e) == 2)) a_bufs = atoi(e); }
  int shift = 8;
  { const char* e = getenv("TS_S2_TILE"); if (e && atoi(e) == 128) shift = 7; }
  auto stages_for = [&](int sh, int ab) { return (limit - fixed - ab * p.nK * kAChunkBytes) / ((1 << sh) * row_bytes); };
  if (stages_for(shift, a_bufs) < 2 && a_bufs == 2) a_bufs = 1;
  if (stages_for(shift, a_bufs) < 2) shift = 7;
  int n_stages = stages_for(shift, a_bufs);
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  { const char* e = getenv("TS_S2_STAGES"); if (e && atoi(e) >= 1 && atoi(e) <= n_stages) n_stages = atoi(e); }
  if (n_stages < 1) { set_error("maxsim: no room for the token ring (dim %d)", a.dim); return TS_ERR_UNSUPPORTED; }
  p.n_stages = n_stages; p.a_bufs = a_bufs; p.tile_shift = shift;
  const int smem = fixed + a_bufs * p.nK * kAChunkBytes + n_stages * (1 << shift) * row_bytes;
  const long long T = (long long)a.B * a.C;
  // at least ~8 candidates per CTA: tiny batches do not pay for 148 pipelines
  long long want = (T + 7) / 8;
  int grid = (int)(want < a.sm_count ? (want < 1 ? 1 : want) : a.sm_count);
  const bool trace = env_on("TS_S2_TRACE");
  if (trace) {
    if (g_trace_grid < grid) {
      if (g_trace_dev) cudaFree(g_trace_dev);
      TS_CUDA_OK(cudaMalloc((void**)&g_trace_dev, (size_t)grid * kTraceSlots * sizeof(unsigned long long)));
      g_trace_grid = grid;
    }
    TS_CUDA_OK(cudaMemsetAsync(g_trace_dev, 0, (size_t)grid * kTraceSlots * sizeof(unsigned long long), st));
    p.trace = g_trace_dev;
  }
  auto launch = [&](auto kern) -> int {
    TS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    TS_LAUNCH(kern, grid, kThreads, smem, st, tq8, tq32, tq128, p);
    return TS_OK;
  };
  if (a.dtype == TS_BF16) rc = trace ? launch(maxsim_flow_kernel<true, true>) : launch(maxsim_flow_kernel<true, false>);
  else rc = trace ? launch(maxsim_flow_kernel<false, true>) : launch(maxsim_flow_kernel<false, false>);
  if (rc) return rc;
  TS_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  if (trace) {
    // debugging aid: synchronise and print the per-role wait cycles (mean / max over the CTAs)
    TS_CUDA_OK(cudaStreamSynchronize(st));
    unsigned long long* h = (unsigned long long*)malloc((size_t)grid * kTraceSlots * sizeof(unsigned long long));
    if (h) {
      TS_CUDA_OK(cudaMemcpy(h, g_trace_dev, (size_t)grid * kTraceSlots * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      static const char* names[kTraceSlots] = {"prod_wait_empty", "prod_wait_meta", "prod_wait_a_or_slots", "tiles", "docs",
                                               "mma_wait_meta", "mma_wait_tempty", "mma_wait_full", "mma_wait_afull",
                                               "epi_wait_meta", "epi_wait_tfull", "epi_drain", "-", "epi_finalize",
                                               "cta_cycles", "-"};
      fprintf(stderr, "[tristage s2 trace] {\"grid\": %d, \"stages\": %d, \"a_bufs\": %d, \"tile_cols\": %d", grid, n_stages, a_bufs, 1 << shift);
      for (int s = 0; s < 15; ++s) {
        if (names[s][0] == '-') continue;
        double sum = 0; unsigned long long mx = 0;
        for (int c = 0; c < grid; ++c) { const unsigned long long v = h[(size_t)c * kTraceSlots + s]; sum += (double)v; if (v > mx) mx = v; }
        fprintf(stderr, ", \"%s\": [%.0f, %llu]", names[s], sum / grid, mx);
      }
      fprintf(stderr, "}\n");
      free(h);
    }
  }
  return TS_OK;
}

}  // namespace ts
