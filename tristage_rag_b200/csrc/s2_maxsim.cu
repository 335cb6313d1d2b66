// Stage-2 late-interaction scoring ("S2-maxsim"): one launch scores every
// (query, candidate) pair of a batch.
//
// Replaces the per-candidate Python loop of ColBERTScorer.rescore_candidates
// (/root/reference/src/stage2_rescorer.py:268-273) around _maxsim_score
// (:167-183: cosine sim matrix, max over doc tokens, MEAN over query tokens)
// and _colbert_score (:185-201: softmax-weighted sum of the row maxima).
// Tokens are L2-normalised once at ingest (convert_rows.cu), so the kernel is
// a plain contraction + fused reductions.
//
// Roofline: HBM -- algorithmic bytes per candidate = pad8(Ld)*dim*2 (doc
// tokens read once; the query tile is re-read from L2), 2*Lq*Ld*dim flop,
// i.e. ~Lq flop/byte: far below the ridge, but only the tensor pipe sustains
// it (~210 TFLOP/s at 6.5 TB/s for Lq = 32).
//
// Tensor path (maxsim_umma_kernel)
//   Token store: docs concatenated row-wise, each padded to a multiple of 8
//   rows, [rows][dim] bf16/fp16.  A work item is (query b, 32 consecutive
//   candidates).  The producer warp looks the candidates up (offset, length),
//   greedily packs their padded token rows into B tiles of <= 256 rows and,
//   per 64-wide K chunk, gathers them with TMA boxes of 128/64/32/16/8 rows
//   (binary decomposition of the padded length -- the ragged docs are staged
//   side by side in SWIZZLE_128B shared memory).  A = the query's token rows
//   (8-row TMA boxes).  One thread issues tcgen05.mma M128 x N(used) x K16
//   into one of two TMEM accumulators: query tokens on lanes, doc tokens on
//   columns.  Epilogue: each thread (one query token) takes the max over each
//   doc's valid columns (pad columns masked), the per-doc maxima go through
//   shared memory and a warp reduces them to mean (maxsim) or softmax-weighted
//   sum (colbert) and writes out[b][j].  Tile layout travels producer ->
//   MMA/epilogue through an 8-slot shared-memory ring guarded by mbarriers.
//
// CUDA-core path (maxsim_simt_kernel): one CTA per (query, candidate); used
//   for shapes the tensor path does not take (dim % 8 != 0, Lq > 128) and as
//   an independent on-device cross-check in the tests.
#include <stdlib.h>

#include "ts_common.cuh"
#include "ts_internal.h"
#include "ts_ptx.cuh"

namespace ts {

int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int dim, int ld, int box_rows);


namespace {

using namespace ts::ptx;

// ============================================================ SIMT path ====
template <typename T>
__global__ void __launch_bounds__(128)
    maxsim_simt_kernel(const T* __restrict__ tok, const int64_t* __restrict__ doc_off,
                       const int32_t* __restrict__ doc_len, int64_t ndocs, int64_t id_base, int dim,
                       const T* __restrict__ q, const int32_t* __restrict__ q_len, int lq_stride,
                       const int64_t* __restrict__ cand, const int32_t* __restrict__ n_cand, int C, int mode,
                       float* __restrict__ out, int layout) {
  TS_DYN_SMEM(float, sm);  // m[lq_stride]
  const int b = blockIdx.y, j = blockIdx.x;
  const int nc = n_cand ? n_cand[b] : C;
  if (j >= nc) return;
  const int64_t id = cand[(size_t)b * C + j] - id_base;
  if (id < 0 || id >= ndocs) return;
  const int Lq = q_len ? min(max(q_len[b], 1), lq_stride) : lq_stride;   // as the tensor path clamps it
  const int Ld = doc_len[id];
  const long long row0 = doc_off[id];
  const T* qb = q + (size_t)b * lq_stride * dim;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = warp; i < Lq; i += nw) {
    float best = -INFINITY;
    for (int t = 0; t < Ld; ++t) {
      float acc = 0.f;
      for (int c = lane; c < dim; c += 32)
        acc = fmaf(Elem<T>::to_f32(qb[(size_t)i * dim + c]), Elem<T>::to_f32(tok[tok_elem(layout, row0 + t, c, dim)]), acc);
      acc = warp_sum(acc);
      best = fmaxf(best, acc);
    }
    if (lane == 0) sm[i] = best;
  }
  __syncthreads();
  if (warp == 0) {
    float res;
    if (mode == TS_S2_MAXSIM) {
      float s = 0.f;
      for (int i = lane; i < Lq; i += 32) s += sm[i];
      res = warp_sum(s) / (float)Lq;
    } else {
      float mx = -INFINITY;
      for (int i = lane; i < Lq; i += 32) mx = fmaxf(mx, sm[i]);
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
      float z = 0.f, s = 0.f;
      for (int i = lane; i < Lq; i += 32) { const float e = __expf(sm[i] - mx); z += e; s += e * sm[i]; }
      z = warp_sum(z); s = warp_sum(s);
      res = s / z;
    }
    if (lane == 0) out[(size_t)b * C + j] = res;
  }
}

// ========================================================== tensor path ====
constexpr int kThreads = 192;
constexpr int kThreadsEpi2 = 320;   // producer + MMA + two epilogue warpgroups (TS_S2_EPI2)
constexpr int kMaxStages = 4;   // 3 stages + double-buffered maxima, or 4 stages + single buffer (p.n_stages)
constexpr int kTileM = 128, kTileN = 256, kChunkK = 64;
constexpr int kABytes = kTileM * kChunkK * 2;
constexpr int kBBytes = kTileN * kChunkK * 2;
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kMetaSlots = 8;
constexpr int kMaxDocsPerTile = 32;   // 256 columns / 8
constexpr int kItemCands = 32;        // candidates per work item (one per producer lane)
constexpr int kTmemCols = 512;

struct TileMeta {
  int used;        // columns in use (multiple of 8); 0 = end of work
  int ndocs;
  int lq;          // real query tokens of this item's query
  int pad_;
  int out_idx[kMaxDocsPerTile];          // b*C + j per doc
  int seg_col[kMaxDocsPerTile];          // first column of doc d
  int seg_len[kMaxDocsPerTile];          // real token count of doc d
  long long seg_row[kMaxDocsPerTile];    // first store row of doc d
  int q_row;       // first row of the query's tokens in the Q tensor
  int pad2_[3];
};

constexpr int kMvalsOne = kMaxDocsPerTile * kTileM;         // floats per buffer
constexpr int kMvalsBytes = 2 * kMvalsOne * 4;             // two buffers, 32 KB
constexpr int kMetaBytes = kMetaSlots * (int)sizeof(TileMeta);
constexpr int kBarBytes = 512;
constexpr int kRingMvalsBytes = 3 * kStageBytes + kMvalsBytes;   // == 4 * kStageBytes + kMvalsBytes / 2 - 32 KB
static_assert(4 * kStageBytes + kMvalsBytes / 2 <= kRingMvalsBytes + 32 * 1024, "layout");
constexpr int kSmemBytes = 4 * kStageBytes + kMvalsBytes / 2 + kMetaBytes + kBarBytes + 1024;

struct MaxSimParams {
  const int64_t* doc_off;
  const int32_t* doc_len;
  int64_t ndocs, id_base;
  const int32_t* q_len;
  int B, lq_stride, nK;
  const int64_t* cand;
  const int32_t* n_cand;
  int C, mode, n_chunks, n_items;
  int n_stages;   // 3: maxima double buffered (one barrier per tile); 4: single buffer (two barriers)
  float* out;
};

// V2 (opt-in, TS_S2_V2=1; written from the round-1 ncu source view, not yet validated on
// hardware): the same pipeline with a shorter epilogue critical path --
//   * shared-memory pointers keep their address space (LDS/STS instead of generic LD/ST for the
//     tile meta and the per-doc maxima),
//   * no per-tile "-inf" initialisation of the maxima: the producer records the first doc of
//     every 64-column quarter in the tile meta and the finalize step only reads the quarters a
//     doc really overlaps,
//   * the drain keeps two tcgen05.ld in flight and reduces each 8-column unit with a max tree
//     (unmasked when the unit lies inside a doc) instead of a 32-long dependent FMNMX chain.
// EPI2 (opt-in, TS_S2_EPI2=1; written without a GPU): TWO epilogue warpgroups (320 threads).  Group g
// (warps 2+4g .. 5+4g) owns the tiles of accumulator g -- the tiles with seq % 2 == g -- so the drain +
// finalize of one tile overlaps the drain + finalize of the next instead of following it; the round-1
// profile showed the MMA warp waiting for a free accumulator on 99 % of the tiles.  Each group has its
// own maxima buffer and named barriers; the producer ends the work with one sentinel per group.
template <bool BF16, bool V2, bool EPI2>
__global__ void __launch_bounds__(EPI2 ? kThreadsEpi2 : kThreads, 1)
    maxsim_umma_kernel(const __grid_constant__ CUtensorMap tmQ8, const __grid_constant__ CUtensorMap tmQ32,
                       const __grid_constant__ CUtensorMap tmQ128,
                       const __grid_constant__ CUtensorMap tmT8, const __grid_constant__ CUtensorMap tmT16,
                       const __grid_constant__ CUtensorMap tmT32, const __grid_constant__ CUtensorMap tmT64,
                       const __grid_constant__ CUtensorMap tmT128, const MaxSimParams p) {
  TS_DYN_SMEM(unsigned char, smem_raw);
  unsigned char* smem;
  if constexpr (V2) {
    // offset arithmetic on the __shared__ array keeps the address space known to the compiler
    smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  } else {
    smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  }
  const int kStages = p.n_stages;
  const int mvals_bufs = (kStages == 3) ? 2 : 1;
  float* mvals_base = reinterpret_cast<float*>(smem + kStages * kStageBytes);           // [bufs][32][128]
  unsigned char* after = smem + kStages * kStageBytes + mvals_bufs * kMvalsOne * 4;
  TileMeta* metas = reinterpret_cast<TileMeta*>(after);
  uint64_t* bars = reinterpret_cast<uint64_t*>(after + kMetaBytes);
  uint64_t* full_bar = bars;                           // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;             // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;         // [2]
  uint64_t* tempty_bar = bars + 2 * kMaxStages + 2;    // [2]
  uint64_t* mfull_bar = bars + 2 * kMaxStages + 4;     // [kMetaSlots]
  uint64_t* mempty_bar = mfull_bar + kMetaSlots;       // [kMetaSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mempty_bar + kMetaSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
    for (int s = 0; s < kMetaSlots; ++s) { mbar_init(&mfull_bar[s], 1); mbar_init(&mempty_bar[s], 4); }
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
    if (lane == 0) {
      prefetch_tmap(&tmQ8); prefetch_tmap(&tmQ32); prefetch_tmap(&tmQ128); prefetch_tmap(&tmT8); prefetch_tmap(&tmT16);
      prefetch_tmap(&tmT32); prefetch_tmap(&tmT64); prefetch_tmap(&tmT128);
    }
    int stage = 0; uint32_t phase = 0;
    uint32_t seq = 0;  // tiles emitted by this CTA

    // emit the tile described by metas[seq % slots] (already filled in).
    // The whole warp issues the TMA traffic: lane d gathers doc d of the tile
    // (its padded rows as boxes of 128/64/32/16/8), lanes 28-31 (or 24-31 /
    // 31) bring the query rows -- one warp-wide instruction per box size
    // instead of a per-doc loop on a single thread.
    auto emit_tile = [&](TileMeta* m) {
      __syncwarp();
      if (lane == 0) mbar_arrive(&mfull_bar[seq % kMetaSlots]);   // release meta to MMA + epilogue
      const int lq = m->lq;
      const int nd = m->ndocs;
      const int q_row = m->q_row;
      const int a_groups = (lq + 7) >> 3;
      const bool rep4 = lq <= 32;          // query tokens replicated into all four TMEM lane quarters
      const bool a_boxes8 = lq <= 64;
      const uint32_t tx = (rep4 ? (uint32_t)kABytes : a_boxes8 ? (uint32_t)a_groups * 1024u : (uint32_t)kABytes) +
                          (uint32_t)m->used * 128u;
      const bool has_doc = lane < nd;
      const int my_rows = has_doc ? ((m->seg_len[lane] + 7) & ~7) : 0;
      const int my_row0 = has_doc ? (int)m->seg_row[lane] : 0;
      const int my_col = has_doc ? m->seg_col[lane] : 0;
      for (int kc = 0; kc < p.nK; ++kc) {
        mbar_wait(&empty_bar[stage], phase ^ 1u, 11);
        unsigned char* sA = smem + stage * kStageBytes;
        unsigned char* sB = sA + kABytes;
        uint64_t* fb = &full_bar[stage];
        if (lane == 0) mbar_arrive_expect_tx(fb, tx);
        __syncwarp();
        const int kx = kc * kChunkK;
        if (rep4) {
          if (lane >= 28) tma_load_2d(sA + (lane - 28) * 4096, &tmQ32, fb, kx, q_row, kEvictLast);
        } else if (a_boxes8) {
          if (lane >= 24 && lane - 24 < a_groups) tma_load_2d(sA + (lane - 24) * 1024, &tmQ8, fb, kx, q_row + (lane - 24) * 8, kEvictLast);
        } else {
          if (lane == 31) tma_load_2d(sA, &tmQ128, fb, kx, q_row, kEvictLast);
        }
        int rows = my_rows, r = my_row0;
        unsigned char* dst = sB + my_col * 128;
        if (rows >= 256) { tma_load_2d(dst, &tmT128, fb, kx, r, kEvictFirst); rows -= 128; r += 128; dst += 128 * 128; }
        if (rows >= 128) { tma_load_2d(dst, &tmT128, fb, kx, r, kEvictFirst); rows -= 128; r += 128; dst += 128 * 128; }
        if (rows >= 64) { tma_load_2d(dst, &tmT64, fb, kx, r, kEvictFirst); rows -= 64; r += 64; dst += 64 * 128; }
        if (rows >= 32) { tma_load_2d(dst, &tmT32, fb, kx, r, kEvictFirst); rows -= 32; r += 32; dst += 32 * 128; }
        if (rows >= 16) { tma_load_2d(dst, &tmT16, fb, kx, r, kEvictFirst); rows -= 16; r += 16; dst += 16 * 128; }
        if (rows >= 8) { tma_load_2d(dst, &tmT8, fb, kx, r, kEvictFirst); }
        if (++stage == kStages) { stage = 0; phase ^= 1u; }
      }
      ++seq;
      __syncwarp();
    };
    auto begin_tile = [&]() -> TileMeta* {
      const uint32_t slot = seq % kMetaSlots;
      mbar_wait(&mempty_bar[slot], ((seq / kMetaSlots) & 1u) ^ 1u, 12);
      return &metas[slot];
    };

    // candidate lookup (one per lane) for a work item; the loads of item i+1
    // are issued before item i is packed so their latency hides behind the TMA issue
    struct Look { bool valid; long long off; int len; int lq; };
    auto lookup = [&](int item) -> Look {
      Look L{false, 0, 0, 1};
      if (item >= p.n_items) return L;
      const int b = item / p.n_chunks, j = (item % p.n_chunks) * kItemCands + lane;
      const int nc = p.n_cand ? min(max(p.n_cand[b], 0), p.C) : p.C;
      L.lq = p.q_len ? min(max(p.q_len[b], 1), min(p.lq_stride, kTileM)) : min(p.lq_stride, kTileM);
      if (j < nc) {
        const int64_t id = p.cand[(size_t)b * p.C + j] - p.id_base;
        if (id >= 0 && id < p.ndocs) { L.off = p.doc_off[id]; L.len = p.doc_len[id]; L.valid = L.len > 0; }
      }
      return L;
    };
    Look cur = lookup(blockIdx.x);
    for (int item = blockIdx.x; item < p.n_items; item += gridDim.x) {
      const Look nxt = lookup(item + gridDim.x);
      const int b = item / p.n_chunks, j0 = (item % p.n_chunks) * kItemCands;
      const int lq = cur.lq;
      const bool valid = cur.valid; const long long off = cur.off; const int len = cur.len;
      unsigned vmask = __ballot_sync(0xffffffffu, valid);
      if (!vmask) { cur = nxt; continue; }
      TileMeta* m = begin_tile();
      int cols = 0, nd = 0;
      int qf1 = -1, qf2 = -1, qf3 = -1;   // V2: first doc reaching into columns >= 64 / 128 / 192
      while (vmask) {
        const int src = __ffs(vmask) - 1;
        vmask &= vmask - 1;
        const int L = __shfl_sync(0xffffffffu, len, src);
        const long long o = __shfl_sync(0xffffffffu, off, src);
        const int pad = (L + 7) & ~7;
        if (cols + pad > kTileN) {
          if (lane == 0) {
            m->used = cols; m->ndocs = nd; m->lq = lq; m->q_row = b * p.lq_stride;
            if constexpr (V2) { m->pad2_[0] = qf1; m->pad2_[1] = qf2; m->pad2_[2] = qf3; }
          }
          emit_tile(m);
          m = begin_tile();
          cols = 0; nd = 0;
          qf1 = qf2 = qf3 = -1;
        }
        if constexpr (V2) {
          if (qf1 < 0 && cols + pad > 64) qf1 = nd;
          if (qf2 < 0 && cols + pad > 128) qf2 = nd;
          if (qf3 < 0 && cols + pad > 192) qf3 = nd;
        }
        if (lane == 0) {
          m->out_idx[nd] = b * p.C + j0 + src;
          m->seg_col[nd] = cols;
          m->seg_len[nd] = L;
          m->seg_row[nd] = o;
        }
        cols += pad; ++nd;
      }
      if (lane == 0) {
        m->used = cols; m->ndocs = nd; m->lq = lq; m->q_row = b * p.lq_stride;
        if constexpr (V2) { m->pad2_[0] = qf1; m->pad2_[1] = qf2; m->pad2_[2] = qf3; }
      }
      emit_tile(m);
      cur = nxt;
    }
    // end-of-work sentinel (EPI2: one per epilogue group -- consecutive seq numbers belong to different groups)
    for (int rep = 0; rep < (EPI2 ? 2 : 1); ++rep) {
      TileMeta* m = begin_tile();
      if (lane == 0) { m->used = 0; m->ndocs = 0; }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mfull_bar[seq % kMetaSlots]);
      ++seq;
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer --------
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (uint32_t seq = 0;; ++seq) {
        const uint32_t slot = seq % kMetaSlots;
        mbar_wait(&mfull_bar[slot], (seq / kMetaSlots) & 1u, 21);
        const int used = *reinterpret_cast<volatile int*>(&metas[slot].used);
        if (used == 0) break;
        const int n_mma = used < 16 ? 16 : ((used + 15) & ~15);
        const uint32_t idesc = make_idesc_f16(kTileM, n_mma, BF16);
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 22);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kTileN);
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&full_bar[stage], phase, 23);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = make_desc_kmajor_sw128(a_addr + kABytes);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks)
            umma_f16_ss(d_tmem, adesc + ks * kDescKStep, bdesc + ks * kDescKStep, idesc, (kc | ks) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // -------------------------------------------------- epilogue ----------
    // lq <= 32: the query tokens sit in ALL four lane quarters (rep4), so warp
    // (quarter q) owns columns [64q, 64q+64) of the tile and the four warps
    // split the drain.  lq > 32: tokens span the quarters, every warp walks
    // all columns of its 32 tokens.  Per-doc partial maxima meet in mvals
    // (double buffered: one named barrier per tile).
    const int quarter = warp & 3;
    const int grp = EPI2 ? ((warp - 2) >> 2) : 0;          // epilogue warpgroup
    const int ew = EPI2 ? ((warp - 2) & 3) : (warp - 2);   // 0..3: finalize docs d with d % 4 == ew
    const int bar_a = EPI2 ? 1 + 2 * grp : 1, bar_b = EPI2 ? 2 + 2 * grp : 2;   // named barriers of this group
    const int lane_row = quarter * 32 + lane;
    int acc = EPI2 ? grp : 0; uint32_t acc_phase = 0;
    for (uint32_t seq = EPI2 ? (uint32_t)grp : 0u;; seq += EPI2 ? 2u : 1u) {
      const uint32_t slot = seq % kMetaSlots;
      mbar_wait(&mfull_bar[slot], (seq / kMetaSlots) & 1u, 31);
      const TileMeta* m = &metas[slot];
      const int used = m->used;
      if (used == 0) break;
      const int nd = m->ndocs, lq = m->lq;
      const bool rep4 = lq <= 32;
      float* mvals = mvals_base + ((mvals_bufs == 2) ? (seq & 1u) : 0u) * kMvalsOne;
      mbar_wait(&tfull_bar[acc], acc_phase, 32);
      tc_fence_after();
      const int c_lo = rep4 ? quarter * 64 : 0;
      const int c_hi = rep4 ? ((used < c_lo + 64) ? used : c_lo + 64) : used;
      const bool warp_active = rep4 ? true : (quarter * 32 < lq);
      if constexpr (V2) {
        // first doc of this warp's column range comes from the tile meta (no -inf initialisation)
        int d = rep4 ? (quarter == 0 ? 0 : m->pad2_[quarter - 1]) : 0;
        if (warp_active && d >= 0 && c_lo < c_hi) {
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kTileN);
          int seg_end = m->seg_col[d] + m->seg_len[d];
          int seg_next = m->seg_col[d] + ((m->seg_len[d] + 7) & ~7);
          float best = -INFINITY;
          for (int g0 = c_lo; g0 < c_hi; g0 += 64) {
            uint32_t r0[32], r1[32];
            const bool two = g0 + 32 < c_hi;           // warp-uniform
            tmem_ld_32x32b_x32(t_addr + (uint32_t)g0, r0);
            if (two) tmem_ld_32x32b_x32(t_addr + (uint32_t)(g0 + 32), r1);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              const int cu = g0 + u * 8;
              if (cu < c_hi) {                         // warp-uniform
                if (cu >= seg_next) {                  // warp-uniform: the next doc starts here
                  mvals[d * kTileM + lane_row] = best;
                  ++d;
                  best = -INFINITY;
                  seg_end = m->seg_col[d] + m->seg_len[d];
                  seg_next = m->seg_col[d] + ((m->seg_len[d] + 7) & ~7);
                }
                float v[8];
#pragma unroll
                for (int j2 = 0; j2 < 8; ++j2) v[j2] = __uint_as_float(u < 4 ? r0[u * 8 + j2] : r1[(u - 4) * 8 + j2]);
                if (cu + 8 > seg_end) {                // warp-uniform: the doc's last, partly padded unit
#pragma unroll
                  for (int j2 = 0; j2 < 8; ++j2) v[j2] = (cu + j2 < seg_end) ? v[j2] : -INFINITY;
                }
                const float m01 = fmaxf(v[0], v[1]), m23 = fmaxf(v[2], v[3]);
                const float m45 = fmaxf(v[4], v[5]), m67 = fmaxf(v[6], v[7]);
                best = fmaxf(best, fmaxf(fmaxf(m01, m23), fmaxf(m45, m67)));
              }
            }
          }
          mvals[d * kTileM + lane_row] = best;
        }
      } else
      if (warp_active) {
        // docs that do not touch this warp's column range contribute -inf
        int d = -1;
        for (int dd = 0; dd < nd; ++dd) {
          const int col = m->seg_col[dd];
          const int nxt = col + ((m->seg_len[dd] + 7) & ~7);
          if (nxt <= c_lo || col >= c_hi) mvals[dd * kTileM + lane_row] = -INFINITY;
          else if (d < 0) d = dd;                         // first doc overlapping the range
        }
        if (d >= 0) {
          const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kTileN);
          int seg_end = m->seg_col[d] + m->seg_len[d];                 // first masked column of doc d
          int seg_next = m->seg_col[d] + ((m->seg_len[d] + 7) & ~7);   // first column of doc d+1
          float best = -INFINITY;
          for (int c0 = c_lo; c0 < c_hi; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int cu = c0 + u * 8;
              if (cu < c_hi) {                   // warp-uniform
                if (cu >= seg_next) {            // warp-uniform: next doc starts at this 8-column unit
                  mvals[d * kTileM + lane_row] = best;
                  ++d;
                  best = -INFINITY;
                  seg_end = m->seg_col[d] + m->seg_len[d];
                  seg_next = m->seg_col[d] + ((m->seg_len[d] + 7) & ~7);
                }
#pragma unroll
                for (int j2 = 0; j2 < 8; ++j2) {
                  const float v = __uint_as_float(r[u * 8 + j2]);
                  best = (cu + j2 < seg_end) ? fmaxf(best, v) : best;
                }
              }
            }
          }
          mvals[d * kTileM + lane_row] = best;
        }
      }
      // accumulator fully read -> hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      if constexpr (EPI2) { acc_phase ^= 1u; }             // the group keeps its accumulator: one use per own tile
      else { acc ^= 1; if (acc == 0) acc_phase ^= 1u; }
      named_bar_sync(bar_a, 128);   // all per-doc maxima of this tile are in mvals
      for (int d = ew; d < nd; d += 4) {
        // m_i = max over doc tokens for query token i (lane i, i + 32, ...)
        float res;
        float mv[4];
        int nmv = 0;
        if constexpr (V2) {
          // only the lane quarters whose 64-column range the doc overlaps hold a value for it
          const int col = m->seg_col[d];
          const int q_lo = rep4 ? (col >> 6) : 0;
          const int q_hi = rep4 ? ((col + ((m->seg_len[d] + 7) & ~7) - 1) >> 6) : 0;
          for (int i = lane; i < lq; i += 32) {
            float v = mvals[d * kTileM + (rep4 ? q_lo * 32 : 0) + i];
            for (int qq = q_lo + 1; qq <= q_hi; ++qq) v = fmaxf(v, mvals[d * kTileM + qq * 32 + i]);
            mv[nmv++] = v;
          }
        } else
        for (int i = lane; i < lq; i += 32) {
          float v = mvals[d * kTileM + i];
          if (rep4) v = fmaxf(fmaxf(v, mvals[d * kTileM + 32 + i]), fmaxf(mvals[d * kTileM + 64 + i], mvals[d * kTileM + 96 + i]));
          mv[nmv++] = v;
        }
        if (p.mode == TS_S2_MAXSIM) {
          float sacc = 0.f;
          for (int t = 0; t < nmv; ++t) sacc += mv[t];
          res = warp_sum(sacc) / (float)lq;
        } else {
          float mx = -INFINITY;
          for (int t = 0; t < nmv; ++t) mx = fmaxf(mx, mv[t]);
#pragma unroll
          for (int o = 16; o >= 1; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
          float z = 0.f, sacc = 0.f;
          for (int t = 0; t < nmv; ++t) {
            const float e = __expf(mv[t] - mx);
            z += e; sacc += e * mv[t];
          }
          z = warp_sum(z); sacc = warp_sum(sacc);
          res = sacc / z;
        }
        if (lane == 0) p.out[m->out_idx[d]] = res;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mempty_bar[slot]);   // meta slot reusable
      // two buffers: mvals[(seq & 1)] is rewritten at tile seq + 2, after the barrier of tile
      // seq + 1, which every warp reaches only after this finalize.  one buffer: wait here.
      // EPI2: buffer (seq & 1) belongs to this group alone and is rewritten by its NEXT tile: wait as well.
      if (EPI2 || mvals_bufs == 1) named_bar_sync(bar_b, 128);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

}  // namespace

int launch_maxsim(const MaxSimArgs& a, cudaStream_t st, int* launches) {
  if (a.B <= 0 || a.C <= 0) { set_error("maxsim: empty batch"); return TS_ERR_INVALID; }
  TS_CUDA_OK(cudaMemsetAsync(a.out, 0, (size_t)a.B * a.C * sizeof(float), st));
  if (a.ndocs == 0 && !a.scatter_bases) return TS_OK;   // (a scatter call still has to publish its step: the flow kernel runs over no work)
  // tensor kernels: the flow kernel reads tile-layout shards, the first kernel row-major ones; everything
  // else (fp32, odd dims, Lq > 128, the test switch) takes the CUDA-core kernel, which reads either layout
  bool tensor_ok = (a.dtype == TS_BF16 || a.dtype == TS_F16) && (a.dim % 8 == 0) && a.lq_stride >= 1 &&
                   a.lq_stride <= TS_S2_MAX_LQ && !(a.mode & 0x100);
  if (tensor_ok && a.layout == kTokTile && !maxsim_flow_takes(a)) tensor_ok = false;
  const int mode = a.mode & 0xff;
  if (!tensor_ok) {
    if (a.lq_stride > 4096) { set_error("maxsim: lq_stride too large"); return TS_ERR_UNSUPPORTED; }
    dim3 grid(a.C, a.B);
    const size_t smem = (size_t)a.lq_stride * sizeof(float);
#define TS_SIMT(T)                                                                                    \
  do {                                                                                                \
    auto kern = maxsim_simt_kernel<T>;                                                                \
    TS_LAUNCH(kern, grid, 128, smem, st, (const T*)a.tok, a.doc_off, a.doc_len, a.ndocs, a.id_base,   \
              a.dim, (const T*)a.q, a.q_len, a.lq_stride, a.cand, a.n_cand, a.C, mode, a.out, a.layout); \
  } while (0)
    if (a.dtype == TS_BF16) TS_SIMT(__nv_bfloat16);
    else if (a.dtype == TS_F16) TS_SIMT(__half);
    else TS_SIMT(float);
#undef TS_SIMT
    TS_CUDA_OK(cudaGetLastError());
    if (launches) ++*launches;
    return TS_OK;
  }
  if (a.layout == kTokTile) {
    MaxSimArgs a2 = a;
    a2.mode = mode;
    return launch_maxsim_flow(a2, st, launches);
  }
  CUtensorMap tq8, tq32, tq128, t8, t16, t32, t64, t128;
  int rc;
  const int64_t qrows = (int64_t)a.B * a.lq_stride;
  if ((rc = make_tmap_2d(&tq8, a.q, a.dtype, qrows, a.dim, a.dim, 8))) return rc;
  if ((rc = make_tmap_2d(&tq32, a.q, a.dtype, qrows, a.dim, a.dim, 32))) return rc;
  if ((rc = make_tmap_2d(&tq128, a.q, a.dtype, qrows, a.dim, a.dim, 128))) return rc;
  if ((rc = make_tmap_2d(&t8, a.tok, a.dtype, a.ntok_rows, a.dim, a.dim, 8))) return rc;
  if ((rc = make_tmap_2d(&t16, a.tok, a.dtype, a.ntok_rows, a.dim, a.dim, 16))) return rc;
  if ((rc = make_tmap_2d(&t32, a.tok, a.dtype, a.ntok_rows, a.dim, a.dim, 32))) return rc;
  if ((rc = make_tmap_2d(&t64, a.tok, a.dtype, a.ntok_rows, a.dim, a.dim, 64))) return rc;
  if ((rc = make_tmap_2d(&t128, a.tok, a.dtype, a.ntok_rows, a.dim, a.dim, 128))) return rc;
  MaxSimParams p{};
  p.doc_off = a.doc_off; p.doc_len = a.doc_len; p.ndocs = a.ndocs; p.id_base = a.id_base;
  p.q_len = a.q_len; p.B = a.B; p.lq_stride = a.lq_stride; p.nK = (a.dim + kChunkK - 1) / kChunkK;
  p.cand = a.cand; p.n_cand = a.n_cand; p.C = a.C; p.mode = mode;
  p.n_chunks = (a.C + kItemCands - 1) / kItemCands;
  p.n_items = a.B * p.n_chunks;
  p.out = a.out;
  { const char* e = getenv("TS_S2_STAGES"); p.n_stages = (e && atoi(e) == 4) ? 4 : 3; }
  int grid = a.sm_count < p.n_items ? a.sm_count : p.n_items;
  const bool v2 = env_flag("TS_S2_V2", kDefaultS2V2);   // opt-in until validated on hardware
  const bool epi2 = env_flag("TS_S2_EPI2", kDefaultS2Epi2);   // two epilogue warpgroups; opt-in until validated on hardware
  if (epi2) p.n_stages = 3;                               // one maxima buffer per group
  const int threads = epi2 ? kThreadsEpi2 : kThreads;
  auto launch = [&](auto kern) -> int {
    TS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    TS_LAUNCH(kern, grid, threads, kSmemBytes, st, tq8, tq32, tq128, t8, t16, t32, t64, t128, p);
    return TS_OK;
  };
  if (a.dtype == TS_BF16) {
    if (epi2) rc = v2 ? launch(maxsim_umma_kernel<true, true, true>) : launch(maxsim_umma_kernel<true, false, true>);
    else rc = v2 ? launch(maxsim_umma_kernel<true, true, false>) : launch(maxsim_umma_kernel<true, false, false>);
  } else {
    if (epi2) rc = v2 ? launch(maxsim_umma_kernel<false, true, true>) : launch(maxsim_umma_kernel<false, false, true>);
    else rc = v2 ? launch(maxsim_umma_kernel<false, true, false>) : launch(maxsim_umma_kernel<false, false, false>);
  }
  if (rc) return rc;
  TS_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return TS_OK;
}

}  // namespace ts
