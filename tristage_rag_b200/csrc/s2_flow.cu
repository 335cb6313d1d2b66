// Stage-2 late-interaction scoring, second tensor kernel ("S2-flow"): the candidates' token rows
// FLOW through full 256-column tiles against a query tile that stays resident in shared memory.
//
// Same arithmetic and the same token-store layout as maxsim_umma_kernel (s2_maxsim.cu); replaces
// the per-candidate loop of ColBERTScorer.rescore_candidates
// (/root/reference/src/stage2_rescorer.py:268-291) around _maxsim_score (:167-183) and
// _colbert_score (:185-201).  What changed, and why (round-1 profile: 0.53 of HBM, DRAM 42 %,
// tensor pipe 23 %, one CTA per SM with 1.5 tiles of look-ahead, tiles 75 % full):
//   * the query tile A (128 rows x dim, <= 64 KB) is loaded ONCE per (CTA, query) and stays in
//     shared memory -- the first kernel re-fetched it through TMA for every tile and K chunk
//     (32 KB of L2->SM traffic per <= 64 KB of doc tokens) and paid a third of every ring stage
//     for it.  The ring now holds doc tokens only: 5 stages of 32 KB at dim 128 = 2.5 tiles
//     of look-ahead per SM.
//   * documents may SPLIT across tiles (8-row granularity): every tile but the last of a query
//     is completely full (next-fit packing filled 75 %), so all ring bytes are HBM bytes.  The
//     running maximum of the split document crosses the tile boundary through a 1 KB carry.
//   * work is split by flat candidate number, [c*T/G, (c+1)*T/G) for CTA c of G: balanced to one
//     candidate, consecutive candidates of one query share tiles and the resident query tile.
// Roofline: HBM, algorithmic bytes per candidate = pad8(Ld)*dim*2.
//
// Roles (192 threads): warp 0 producer (candidate lookup one chunk ahead, tile layout, TMA
// gather of 128/64/32/16/8-row boxes, query tile loads), warp 1 one MMA thread
// (tcgen05.mma M128 x N(used) x K16, two 256-column TMEM accumulators), warps 2-5 epilogue
// (tcgen05.ld, per-(query token, doc) max, mean / softmax-weighted sum).
#include <stdlib.h>

#include "ts_common.cuh"
#include "ts_internal.h"
#include "ts_ptx.cuh"

namespace ts {

int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int dim, int ld, int box_rows);

namespace {

using namespace ts::ptx;

constexpr int kThreads = 192;
constexpr int kTileM = 128, kTileN = 256, kChunkK = 64;
constexpr int kAChunkBytes = kTileM * kChunkK * 2;   // 16 KB: one K chunk of the query tile
constexpr int kStageBytes = kTileN * kChunkK * 2;    // 32 KB: one K chunk of a doc-token tile
constexpr int kMaxStages = 6;
constexpr int kMaxNK = 4;                            // dim <= 256: the query tile fits beside the ring
constexpr int kMetaSlots = 8;
constexpr int kMaxSegs = 16;                         // doc segments per tile
constexpr int kTmemCols = 512;
constexpr int kMvalsOne = kMaxSegs * kTileM;         // floats per maxima buffer
constexpr int kMvalsBytes = 2 * kMvalsOne * 4;       // double buffered: 16 KB
constexpr int kCarryBytes = 2 * kTileM * 4;          // running maxima of the doc split across a tile boundary
constexpr int kBarBytes = 512;
constexpr int kTraceSlots = 16;

struct FlowMeta {
  int used;          // columns in use (multiple of 8); 0 = end of work
  int nseg;          // doc segments in this tile
  int lq;            // real query tokens
  int flags;         // bit 0: segment 0 continues a doc of the previous tile; bit 1: the last segment continues
                     // in the next tile; bit 2: first tile of a query (wait for its query tile)
  int a_buf, a_par;  // query-tile buffer and the parity of its "loaded" barrier
  int qf[3];         // first segment reaching into columns >= 64 / 128 / 192 (-1: none)
  int pad_[3];
  int out_idx[kMaxSegs];        // b*C + j of the segment's doc
  int seg_col[kMaxSegs];        // first column
  int seg_len[kMaxSegs];        // real tokens in this segment
  long long seg_row[kMaxSegs];  // first store row
};
constexpr int kMetaBytes = kMetaSlots * (int)sizeof(FlowMeta);

struct FlowParams {
  const int64_t* doc_off;
  const int32_t* doc_len;
  int64_t ndocs, id_base;
  const int32_t* q_len;
  int B, lq_stride, nK;
  const int64_t* cand;
  const int32_t* n_cand;
  int C, mode;
  int n_stages, a_bufs;
  float* out;
  unsigned long long* trace;   // [grid][kTraceSlots] cycle counters (TRACE instantiation only)
};

__host__ __device__ constexpr int flow_fixed_bytes() { return kMvalsBytes + kCarryBytes + kMetaBytes + kBarBytes + 1024; }

#define TS_TR_BEGIN() long long _tr0 = 0; if constexpr (TRACE) _tr0 = clock64()
#define TS_TR_END(acc) do { if constexpr (TRACE) (acc) += (unsigned long long)(clock64() - _tr0); } while (0)

template <bool BF16, bool TRACE>
__global__ void __launch_bounds__(kThreads, 1)
    maxsim_flow_kernel(const __grid_constant__ CUtensorMap tmQ8, const __grid_constant__ CUtensorMap tmQ32,
                       const __grid_constant__ CUtensorMap tmQ128,
                       const __grid_constant__ CUtensorMap tmT8, const __grid_constant__ CUtensorMap tmT16,
                       const __grid_constant__ CUtensorMap tmT32, const __grid_constant__ CUtensorMap tmT64,
                       const __grid_constant__ CUtensorMap tmT128, const FlowParams p) {
  TS_DYN_SMEM(unsigned char, smem_raw);
  // offset arithmetic on the __shared__ array keeps the address space known to the compiler (LDS/STS)
  unsigned char* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int kStages = p.n_stages;
  const int a_buf_bytes = p.nK * kAChunkBytes;
  unsigned char* sA = smem;                                           // [a_bufs][nK][128 x 64]
  unsigned char* sB = smem + p.a_bufs * a_buf_bytes;                  // [stages][256 x 64]
  float* mvals_base = reinterpret_cast<float*>(sB + kStages * kStageBytes);   // [2][kMaxSegs][128]
  float* carry = mvals_base + 2 * kMvalsOne;                          // [2][128]
  FlowMeta* metas = reinterpret_cast<FlowMeta*>(reinterpret_cast<unsigned char*>(carry) + kCarryBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(metas) + kMetaBytes);
  uint64_t* full_bar = bars;                           // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;             // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;         // [2]
  uint64_t* tempty_bar = tfull_bar + 2;                // [2]
  uint64_t* afull_bar = tempty_bar + 2;                // [2]
  uint64_t* mfull_bar = afull_bar + 2;                 // [kMetaSlots]
  uint64_t* mempty_bar = mfull_bar + kMetaSlots;       // [kMetaSlots]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mempty_bar + kMetaSlots);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long long t_start = 0;
  if constexpr (TRACE) t_start = clock64();

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); mbar_init(&afull_bar[a], 1); }
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
    unsigned long long tr_wait_empty = 0, tr_wait_meta = 0, tr_wait_a = 0;
    // this CTA's candidates: flat numbers [f0, f1) of b*C + j
    const long long T = (long long)p.B * p.C;
    const long long f0 = T * blockIdx.x / gridDim.x, f1 = T * (blockIdx.x + 1) / gridDim.x;

    uint32_t fills = 0;          // ring stages filled so far: stage = fills % kStages, phase = (fills / kStages) & 1
    uint32_t seq = 0;            // tiles emitted
    uint32_t n_queries = 0;      // query tiles loaded
    uint32_t a_last_fill0 = 0, a_last_fill1 = 0;   // fills counter after the last chunk that read query buffer 0 / 1
    int b_cur = -1, lq_cur = 1, a_buf = 0, a_par = 0;
    bool a_new = false;
    // tile under construction
    FlowMeta* m = nullptr;
    int cols = 0, nseg = 0, qf1 = -1, qf2 = -1, qf3 = -1;
    bool first_cont = false;

    auto begin_tile = [&]() {
      const uint32_t slot = seq % kMetaSlots;
      TS_TR_BEGIN();
      mbar_wait(&mempty_bar[slot], ((seq / kMetaSlots) & 1u) ^ 1u, 12);
      TS_TR_END(tr_wait_meta);
      m = &metas[slot];
      cols = 0; nseg = 0; qf1 = qf2 = qf3 = -1;
    };
    // publish the tile under construction and gather its token rows: lane s brings segment s
    auto emit_tile = [&](bool last_cont) {
      if (lane == 0) {
        m->used = cols; m->nseg = nseg; m->lq = lq_cur;
        m->flags = (first_cont ? 1 : 0) | (last_cont ? 2 : 0) | (a_new ? 4 : 0);
        m->a_buf = a_buf; m->a_par = a_par;
        m->qf[0] = qf1; m->qf[1] = qf2; m->qf[2] = qf3;
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&mfull_bar[seq % kMetaSlots]);   // release the layout to MMA + epilogue
      const bool has_seg = lane < nseg;
      const int my_rows = has_seg ? ((m->seg_len[lane] + 7) & ~7) : 0;
      const int my_row0 = has_seg ? (int)m->seg_row[lane] : 0;
      const int my_col = has_seg ? m->seg_col[lane] : 0;
      const uint32_t tx = (uint32_t)cols * 128u;
      for (int kc = 0; kc < p.nK; ++kc) {
        const int stage = (int)(fills % (uint32_t)kStages);
        const uint32_t phase = (fills / (uint32_t)kStages) & 1u;
        {
          TS_TR_BEGIN();
          mbar_wait(&empty_bar[stage], phase ^ 1u, 11);
          TS_TR_END(tr_wait_empty);
        }
        uint64_t* fb = &full_bar[stage];
        if (lane == 0) mbar_arrive_expect_tx(fb, tx);
        __syncwarp();
        const int kx = kc * kChunkK;
        int rows = my_rows, r = my_row0;
        unsigned char* dst = sB + stage * kStageBytes + my_col * 128;
        if (rows >= 256) { tma_load_2d(dst, &tmT128, fb, kx, r, kEvictFirst); rows -= 128; r += 128; dst += 128 * 128; }
        if (rows >= 128) { tma_load_2d(dst, &tmT128, fb, kx, r, kEvictFirst); rows -= 128; r += 128; dst += 128 * 128; }
        if (rows >= 64) { tma_load_2d(dst, &tmT64, fb, kx, r, kEvictFirst); rows -= 64; r += 64; dst += 64 * 128; }
        if (rows >= 32) { tma_load_2d(dst, &tmT32, fb, kx, r, kEvictFirst); rows -= 32; r += 32; dst += 32 * 128; }
        if (rows >= 16) { tma_load_2d(dst, &tmT16, fb, kx, r, kEvictFirst); rows -= 16; r += 16; dst += 16 * 128; }
        if (rows >= 8) { tma_load_2d(dst, &tmT8, fb, kx, r, kEvictFirst); }
        ++fills;
      }
      if (a_buf == 0) a_last_fill0 = fills; else a_last_fill1 = fills;
      ++seq;
      a_new = false;
      m = nullptr;
      __syncwarp();
    };
    // a new query: its token rows go into a query-tile buffer once the MMAs that read the buffer's
    // previous occupant have retired (their last ring stage has been handed back)
    auto start_query = [&](int b, int lq) {
      a_buf = (int)(n_queries % (uint32_t)p.a_bufs);
      a_par = (int)((n_queries / (uint32_t)p.a_bufs) & 1u);
      if (n_queries >= (uint32_t)p.a_bufs) {
        const uint32_t lf = a_buf == 0 ? a_last_fill0 : a_last_fill1;   // >= 1: every query emits a tile
        if (fills - lf < (uint32_t)kStages) {      // otherwise the refill of that stage already waited for it
          const uint32_t f = lf - 1u;
          TS_TR_BEGIN();
          mbar_wait(&empty_bar[f % (uint32_t)kStages], (f / (uint32_t)kStages) & 1u, 13);
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
      b_cur = b; lq_cur = lq; a_new = true;
      __syncwarp();
    };

    // candidate lookup, one per lane; the loads of chunk i+1 are issued before chunk i is laid out
    struct Look { bool valid; long long off; int len; int b; int lq; };
    auto lookup = [&](long long fbase) -> Look {
      Look L{false, 0, 0, 0, 1};
      const long long f = fbase + lane;
      if (f >= f1) return L;
      const int b = (int)(f / p.C), j = (int)(f % p.C);
      const int nc = p.n_cand ? min(max(p.n_cand[b], 0), p.C) : p.C;
      L.b = b;
      L.lq = p.q_len ? min(max(p.q_len[b], 1), min(p.lq_stride, kTileM)) : min(p.lq_stride, kTileM);
      if (j < nc) {
        const int64_t id = p.cand[f] - p.id_base;
        if (id >= 0 && id < p.ndocs) { L.off = p.doc_off[id]; L.len = p.doc_len[id]; L.valid = L.len > 0; }
      }
      return L;
    };
    Look cur = lookup(f0);
    for (long long fb = f0; fb < f1; fb += 32) {
      const Look nxt = lookup(fb + 32);
      unsigned vmask = __ballot_sync(0xffffffffu, cur.valid);
      while (vmask) {
        const int src = __ffs(vmask) - 1;
        vmask &= vmask - 1;
        const int L = __shfl_sync(0xffffffffu, cur.len, src);
        const long long o = __shfl_sync(0xffffffffu, cur.off, src);
        const int b = __shfl_sync(0xffffffffu, cur.b, src);
        const int lq = __shfl_sync(0xffffffffu, cur.lq, src);
        if (b != b_cur) {
          if (m) emit_tile(false);
          start_query(b, lq);
          first_cont = false;
        }
        int rem_pad = (L + 7) & ~7, rem_real = L;
        long long row = o;
        const int oidx = (int)(fb + src);
        while (rem_pad > 0) {
          if (!m) begin_tile();
          const int room = kTileN - cols;
          const int piece = rem_pad < room ? rem_pad : room;
          const int real = rem_real < piece ? rem_real : piece;
          if (qf1 < 0 && cols + piece > 64) qf1 = nseg;
          if (qf2 < 0 && cols + piece > 128) qf2 = nseg;
          if (qf3 < 0 && cols + piece > 192) qf3 = nseg;
          if (lane == 0) {
            m->out_idx[nseg] = oidx;
            m->seg_col[nseg] = cols;
            m->seg_len[nseg] = real;
            m->seg_row[nseg] = row;
          }
          cols += piece; ++nseg; rem_pad -= piece; rem_real -= real; row += piece;
          if (cols == kTileN || nseg == kMaxSegs) {
            const bool cont = rem_pad > 0;
            emit_tile(cont);
            first_cont = cont;
          }
        }
      }
      cur = nxt;
    }
    if (m) emit_tile(false);
    // end-of-work sentinel
    begin_tile();
    if (lane == 0) { m->used = 0; m->nseg = 0; }
    __syncwarp();
    if (lane == 0) mbar_arrive(&mfull_bar[seq % kMetaSlots]);
    ++seq;
    if constexpr (TRACE) {
      if (lane == 0) {
        unsigned long long* t = p.trace + (size_t)blockIdx.x * kTraceSlots;
        t[0] = tr_wait_empty; t[1] = tr_wait_meta; t[2] = tr_wait_a; t[3] = seq - 1; t[4] = fills;
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer --------
      unsigned long long tr_meta = 0, tr_tempty = 0, tr_full = 0, tr_afull = 0;
      uint32_t fills = 0;
      int acc = 0; uint32_t acc_phase = 0;
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
        const int flags = mm->flags, abuf = mm->a_buf;
        if (flags & 4) {
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
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kTileN);
        const uint32_t a_addr0 = smem_u32(sA + abuf * a_buf_bytes);
        for (int kc = 0; kc < p.nK; ++kc) {
          const int stage = (int)(fills % (uint32_t)kStages);
          {
            TS_TR_BEGIN();
            mbar_wait(&full_bar[stage], (fills / (uint32_t)kStages) & 1u, 23);
            TS_TR_END(tr_full);
          }
          tc_fence_after();
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr0 + (uint32_t)(kc * kAChunkBytes));
          const uint64_t bdesc = make_desc_kmajor_sw128(smem_u32(sB + stage * kStageBytes));
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks)
            umma_f16_ss(d_tmem, adesc + ks * kDescKStep, bdesc + ks * kDescKStep, idesc, (kc | ks) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);
          ++fills;
        }
        umma_commit(&tfull_bar[acc]);
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
      if constexpr (TRACE) {
        unsigned long long* t = p.trace + (size_t)blockIdx.x * kTraceSlots;
        t[5] = tr_meta; t[6] = tr_tempty; t[7] = tr_full; t[8] = tr_afull;
      }
    }
  } else {
    // -------------------------------------------------- epilogue ----------
    // lq <= 32: the query tokens sit in ALL four lane quarters, warp (quarter q) drains columns
    // [64q, 64q+64).  lq > 32: tokens span the quarters, every warp walks all columns of its 32
    // tokens.  Per-segment maxima meet in mvals (double buffered: one named barrier per tile).
    unsigned long long tr_meta = 0, tr_tfull = 0, tr_drain = 0, tr_bar = 0, tr_fin = 0;
    const int quarter = warp & 3;
    const int ew = warp - 2;                              // finalize segments s with s % 4 == ew
    const int lane_row = quarter * 32 + lane;
    int acc = 0; uint32_t acc_phase = 0;
    for (uint32_t seq = 0;; ++seq) {
      const uint32_t slot = seq % kMetaSlots;
      {
        TS_TR_BEGIN();
        mbar_wait(&mfull_bar[slot], (seq / kMetaSlots) & 1u, 31);
        TS_TR_END(tr_meta);
      }
      const FlowMeta* m = &metas[slot];
      const int used = m->used;
      if (used == 0) break;
      const int nseg = m->nseg, lq = m->lq, flags = m->flags;
      const bool rep4 = lq <= 32;
      float* mvals = mvals_base + (seq & 1u) * kMvalsOne;
      {
        TS_TR_BEGIN();
        mbar_wait(&tfull_bar[acc], acc_phase, 32);
        TS_TR_END(tr_tfull);
      }
      tc_fence_after();
      long long t_dr = 0;
      if constexpr (TRACE) t_dr = clock64();
      const int c_lo = rep4 ? quarter * 64 : 0;
      const int c_hi = rep4 ? ((used < c_lo + 64) ? used : c_lo + 64) : used;
      const bool warp_active = rep4 ? true : (quarter * 32 < lq);
      // first segment of this warp's column range comes from the tile meta (no -inf initialisation)
      int s = rep4 ? (quarter == 0 ? 0 : m->qf[quarter - 1]) : 0;
      if (warp_active && s >= 0 && c_lo < c_hi) {
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kTileN);
        int seg_end = m->seg_col[s] + m->seg_len[s];                 // first masked column of segment s
        int seg_next = m->seg_col[s] + ((m->seg_len[s] + 7) & ~7);   // first column of segment s+1
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
              if (cu >= seg_next) {                  // warp-uniform: the next segment starts here
                mvals[s * kTileM + lane_row] = best;
                ++s;
                best = -INFINITY;
                seg_end = m->seg_col[s] + m->seg_len[s];
                seg_next = m->seg_col[s] + ((m->seg_len[s] + 7) & ~7);
              }
              float v[8];
#pragma unroll
              for (int j2 = 0; j2 < 8; ++j2) v[j2] = __uint_as_float(u < 4 ? r0[u * 8 + j2] : r1[(u - 4) * 8 + j2]);
              if (cu + 8 > seg_end) {                // warp-uniform: the segment's last, partly padded unit
#pragma unroll
                for (int j2 = 0; j2 < 8; ++j2) v[j2] = (cu + j2 < seg_end) ? v[j2] : -INFINITY;
              }
              const float m01 = fmaxf(v[0], v[1]), m23 = fmaxf(v[2], v[3]);
              const float m45 = fmaxf(v[4], v[5]), m67 = fmaxf(v[6], v[7]);
              best = fmaxf(best, fmaxf(fmaxf(m01, m23), fmaxf(m45, m67)));
            }
          }
        }
        mvals[s * kTileM + lane_row] = best;
      }
      // accumulator fully read -> hand TMEM back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      if constexpr (TRACE) { const long long t1 = clock64(); tr_drain += (unsigned long long)(t1 - t_dr); t_dr = t1; }
      named_bar_sync(1, 128);   // all per-segment maxima of this tile are in mvals
      if constexpr (TRACE) { const long long t1 = clock64(); tr_bar += (unsigned long long)(t1 - t_dr); t_dr = t1; }
      float* carry_w = carry + (seq & 1u) * kTileM;            // written by this tile's continuing segment
      const float* carry_r = carry + ((seq & 1u) ^ 1u) * kTileM;   // left by the previous tile
      for (int d = ew; d < nseg; d += 4) {
        // m_i = max over the segment's tokens for query token i (lane i, i + 32, ...); only the lane
        // quarters whose 64-column range the segment overlaps hold a value for it
        const int col = m->seg_col[d];
        const int q_lo = rep4 ? (col >> 6) : 0;
        const int q_hi = rep4 ? ((col + ((m->seg_len[d] + 7) & ~7) - 1) >> 6) : 0;
        const bool from_prev = (d == 0) && (flags & 1);
        const bool to_next = (d == nseg - 1) && (flags & 2);
        float mv[4];
        int nmv = 0;
        for (int i = lane; i < lq; i += 32) {
          float v = mvals[d * kTileM + (rep4 ? q_lo * 32 : 0) + i];
          for (int qq = q_lo + 1; qq <= q_hi; ++qq) v = fmaxf(v, mvals[d * kTileM + qq * 32 + i]);
          if (from_prev) v = fmaxf(v, carry_r[i]);
          if (to_next) carry_w[i] = v;
          mv[nmv++] = v;
        }
        if (to_next) continue;                      // the doc's score is written by the tile it ends in
        float res;
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
      if constexpr (TRACE) tr_fin += (unsigned long long)(clock64() - t_dr);
      // mvals[(seq & 1)] and carry[(seq & 1)] are rewritten at tile seq + 2, after the barrier of tile
      // seq + 1, which every warp reaches only after this finalize.
    }
    if constexpr (TRACE) {
      if (warp == 2 && lane == 0) {
        unsigned long long* t = p.trace + (size_t)blockIdx.x * kTraceSlots;
        t[9] = tr_meta; t[10] = tr_tfull; t[11] = tr_drain; t[12] = tr_bar; t[13] = tr_fin;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
  if constexpr (TRACE) {
    if (threadIdx.x == 0) p.trace[(size_t)blockIdx.x * kTraceSlots + 14] = (unsigned long long)(clock64() - t_start);
  }
}

}  // namespace

bool maxsim_flow_takes(const MaxSimArgs& a) {
  return (a.dtype == TS_BF16 || a.dtype == TS_F16) && (a.dim % 8 == 0) && a.lq_stride >= 1 &&
         a.lq_stride <= TS_S2_MAX_LQ && (a.dim + kChunkK - 1) / kChunkK <= kMaxNK &&
         (long long)a.B * a.C < (1ll << 31);
}

// last trace of a TS_S2_TRACE=1 launch: per-role cycle counters, mean and max over the CTAs
static unsigned long long* g_trace_dev = nullptr;
static int g_trace_grid = 0;

int launch_maxsim_flow(const MaxSimArgs& a, cudaStream_t st, int* launches) {
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
  FlowParams p{};
  p.doc_off = a.doc_off; p.doc_len = a.doc_len; p.ndocs = a.ndocs; p.id_base = a.id_base;
  p.q_len = a.q_len; p.B = a.B; p.lq_stride = a.lq_stride; p.nK = (a.dim + kChunkK - 1) / kChunkK;
  p.cand = a.cand; p.n_cand = a.n_cand; p.C = a.C; p.mode = a.mode & 0xff;
  p.out = a.out;
  // shared memory: query tile(s) + ring + fixed part within the 227 KB opt-in limit
  const int limit = 232448;
  // a CTA meets a new query every C candidates: with short candidate lists a second query-tile buffer
  // hides the reload behind the previous query's last tiles (at the price of ring depth)
  int a_bufs = (a.C < 128 && p.nK <= 2) ? 2 : 1;
  { const char* e = getenv("TS_S2_ABUFS"); if (e && (atoi(e) == 1 || atoi(e) == 2)) a_bufs = atoi(e); }
  int n_stages = (limit - flow_fixed_bytes() - a_bufs * p.nK * kAChunkBytes) / kStageBytes;
  if (n_stages < 2 && a_bufs == 2) { a_bufs = 1; n_stages = (limit - flow_fixed_bytes() - p.nK * kAChunkBytes) / kStageBytes; }
  if (n_stages > kMaxStages) n_stages = kMaxStages;
  { const char* e = getenv("TS_S2_STAGES"); if (e && atoi(e) >= 2 && atoi(e) <= n_stages) n_stages = atoi(e); }
  if (n_stages < 2) { set_error("maxsim: no room for the token ring (dim %d)", a.dim); return TS_ERR_UNSUPPORTED; }
  p.n_stages = n_stages; p.a_bufs = a_bufs;
  const int smem = flow_fixed_bytes() + a_bufs * p.nK * kAChunkBytes + n_stages * kStageBytes;
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
    TS_LAUNCH(kern, grid, kThreads, smem, st, tq8, tq32, tq128, t8, t16, t32, t64, t128, p);
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
      static const char* names[kTraceSlots] = {"prod_wait_empty", "prod_wait_meta", "prod_wait_a", "tiles", "fills",
                                               "mma_wait_meta", "mma_wait_tempty", "mma_wait_full", "mma_wait_afull",
                                               "epi_wait_meta", "epi_wait_tfull", "epi_drain", "epi_bar", "epi_finalize",
                                               "cta_cycles", "-"};
      fprintf(stderr, "[tristage s2 trace] {\"grid\": %d, \"stages\": %d, \"a_bufs\": %d", grid, n_stages, a_bufs);
      for (int s = 0; s < 15; ++s) {
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
