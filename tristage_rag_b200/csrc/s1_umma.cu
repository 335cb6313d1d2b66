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
//   grid = n_mt * n_slices CTAs, one per SM.  CTA (mt, slice) owns query tile
//   mt (128 queries = the M rows of the MMA = TMEM lanes) and corpus tiles
//   slice, slice+n_slices, ... (256 rows = the N columns of the MMA).
//   Per tile and 64-element K chunk, TMA (SWIZZLE_128B) brings A = 128x64 of Q
//   and B = 256x64 of X into a 4-stage shared-memory ring; one thread issues
//   four tcgen05.mma (M128 N256 K16, bf16/fp16 in, fp32 accumulate in TMEM).
//   Two 256-column TMEM accumulators are double buffered against the epilogue.
//
// Warp roles (192 threads): warp 0 = TMA producer (one lane), warp 1 = MMA
//   issuer (one lane), warps 2-5 = epilogue; epilogue warp w reads TMEM lane
//   quarter (w % 4) with tcgen05.ld 32x32b.x32, so thread t owns ONE query and
//   walks its 256 scores of the tile.
//
// Fused top-k: thread-private running threshold tau (register) and candidate
//   count; scores > tau are appended to the thread's candidate list in global
//   memory (L2 resident).  When a list is within 32 entries of CAP the whole
//   warp bitonic-sorts it in registers (warp_prune_list), keeps the best k
//   and raises tau.  At the end every list is sorted once more and written as
//   partial[slice][query][k]; topk_select.cu merges the n_slices lists.
//
// Small batches (B <= 64): the queries are spread over the four TMEM lane
//   quarters in groups of 8 rows (8-row TMA boxes), so all four epilogue warps
//   share the list maintenance instead of one.
#include "ts_common.cuh"
#include "ts_internal.h"
#include "ts_ptx.cuh"

namespace ts {
namespace {

using namespace ts::ptx;

constexpr int kThreads = 192;
constexpr int kStages = 4;
constexpr int kTileM = 128, kTileN = 256, kChunkK = 64;
constexpr int kABytes = kTileM * kChunkK * 2;  // 16 KB
constexpr int kBBytes = kTileN * kChunkK * 2;  // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kBarBytes = 256;
constexpr int kSmemBytes = kStages * kStageBytes + kBarBytes + 1024;  // + alignment slack
constexpr int kTmemCols = 512;

struct UmmaParams {
  int64_t N;
  int nK, B, k;
  int n_slices, n_tiles;
  int spread, n_qgroups, cap;
  uint64_t* lists;
  uint64_t* partial;
  const float* inv_norm;
};

template <bool BF16>
__global__ void __launch_bounds__(kThreads, 1)
    s1_umma_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmQ8,
                   const __grid_constant__ CUtensorMap tmX, const UmmaParams p) {
  const int CAP = p.cap;
  extern __shared__ unsigned char smem_raw[];
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
  uint64_t* full_bar = bars;                  // [kStages] TMA -> MMA
  uint64_t* empty_bar = bars + kStages;       // [kStages] MMA -> TMA
  uint64_t* tfull_bar = bars + 2 * kStages;   // [2] MMA -> epilogue
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = blockIdx.x / p.n_slices, slice = blockIdx.x % p.n_slices;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull_bar[a], 1); mbar_init(&tempty_bar[a], 4); }
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
    if (lane == 0) {
      // ------------------------------------------------ TMA producer ------
      prefetch_tmap(&tmQ); prefetch_tmap(&tmQ8); prefetch_tmap(&tmX);
      const uint32_t tx = (p.spread ? (uint32_t)p.n_qgroups * 1024u : (uint32_t)kABytes) + (uint32_t)kBBytes;
      const uint64_t x_policy = (gridDim.x > (unsigned)p.n_slices) ? kEvictNormal : kEvictFirst;
      int stage = 0; uint32_t phase = 0;
      for (int t = slice; t < p.n_tiles; t += p.n_slices) {
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&empty_bar[stage], phase ^ 1u, 1);
          unsigned char* sA = smem + stage * kStageBytes;
          unsigned char* sB = sA + kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], tx);
          if (p.spread) {
            for (int g = 0; g < p.n_qgroups; ++g) {
              const int row = ((g & 3) * 32) + ((g >> 2) * 8);
              tma_load_2d(sA + row * 128, &tmQ8, &full_bar[stage], kc * kChunkK, g * 8, kEvictLast);
            }
          } else {
            tma_load_2d(sA, &tmQ, &full_bar[stage], kc * kChunkK, mt * kTileM, kEvictLast);
          }
          tma_load_2d(sB, &tmX, &full_bar[stage], kc * kChunkK, t * kTileN, x_policy);
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // ------------------------------------------------ MMA issuer --------
      constexpr uint32_t idesc = make_idesc_f16(kTileM, kTileN, BF16);
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      for (int t = slice; t < p.n_tiles; t += p.n_slices) {
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u, 2);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * kTileN);
        for (int kc = 0; kc < p.nK; ++kc) {
          mbar_wait(&full_bar[stage], phase, 3);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * kStageBytes);
          const uint64_t adesc = make_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = make_desc_kmajor_sw128(a_addr + kABytes);
#pragma unroll
          for (int ks = 0; ks < kChunkK / 16; ++ks)
            umma_f16_ss(d_tmem, adesc + ks * kDescKStep, bdesc + ks * kDescKStep, idesc, (kc | ks) ? 1u : 0u);
          umma_commit(&empty_bar[stage]);  // smem stage reusable once these MMAs retire
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&tfull_bar[acc]);      // accumulator complete
        acc ^= 1; if (acc == 0) acc_phase ^= 1u;
      }
    }
  } else {
    // -------------------------------------------------- epilogue ----------
    const int quarter = warp & 3;
    const int lane_row = quarter * 32 + lane;
    int qi;
    if (p.spread) qi = (((lane >> 3) * 4 + quarter) * 8) + (lane & 7);
    else qi = lane_row;
    const int q_global = mt * kTileM + qi;
    const bool active = q_global < p.B;
    const bool warp_active = __any_sync(0xffffffffu, active);
    uint64_t* lst = p.lists + ((size_t)blockIdx.x * kTileM + lane_row) * CAP;
    float tau = -INFINITY;
    int cnt = 0;
    int acc = 0; uint32_t acc_phase = 0;
    for (int t = slice; t < p.n_tiles; t += p.n_slices) {
      mbar_wait(&tfull_bar[acc], acc_phase, 4);
      tc_fence_after();
      if (warp_active) {
        const int64_t n0 = (int64_t)t * kTileN;
        const int ncols = (int)((p.N - n0) < (int64_t)kTileN ? (p.N - n0) : (int64_t)kTileN);
        const uint32_t t_addr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * kTileN);
        for (int c0 = 0; c0 < ncols; c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(t_addr + (uint32_t)c0, r);
          tmem_ld_wait();
          if (active) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              float s = __uint_as_float(r[j]);
              const int col = c0 + j;
              if (p.inv_norm) s *= __ldg(p.inv_norm + n0 + (col < ncols ? col : 0));
              if (col < ncols && s > tau) lst[cnt++] = make_key(s, (uint32_t)(n0 + col));
            }
          }
          unsigned full = __ballot_sync(0xffffffffu, active && cnt > CAP - 32);
          while (full) {
            const int L = __ffs(full) - 1;
            full &= full - 1;
            uint64_t* lp = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(lst), L));
            const int lc = __shfl_sync(0xffffffffu, cnt, L);
            const uint64_t kth = warp_prune_list(lp, lc, p.k, lane, CAP);
            if (lane == L) { cnt = p.k; tau = key_score(kth); }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1; if (acc == 0) acc_phase ^= 1u;
    }
    // final flush: sort every list, emit partial[slice][q][0..k)
    unsigned act = __ballot_sync(0xffffffffu, active);
    while (act) {
      const int L = __ffs(act) - 1;
      act &= act - 1;
      uint64_t* lp = reinterpret_cast<uint64_t*>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(lst), L));
      const int lc = __shfl_sync(0xffffffffu, cnt, L);
      const int qL = __shfl_sync(0xffffffffu, q_global, L);
      warp_prune_list(lp, lc, p.k, lane, CAP, p.partial + ((size_t)slice * p.B + qL) * p.k);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

// ------------------------------------------------------------------ host ---
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

}  // namespace

// 2-D row-major [rows][dim] (pitch ld elements) 16-bit tensor, box = 64 x box_rows, SWIZZLE_128B
int make_tmap_2d(CUtensorMap* out, const void* base, int dtype, int64_t rows, int dim, int ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) { set_error("cuTensorMapEncodeTiled unavailable (no CUDA driver?)"); return TS_ERR_CUDA; }
  cuuint64_t gdim[2] = {(cuuint64_t)dim, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kChunkK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, dtype == TS_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                   const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: CUresult %d (rows %lld dim %d ld %d box %d)", (int)r, (long long)rows, dim, ld, box_rows); return TS_ERR_CUDA; }
  return TS_OK;
}

namespace {
struct UmmaPlan { int n_mt, n_slices, n_tiles, grid; };
UmmaPlan umma_plan(const ScanArgs& a) {
  UmmaPlan pl;
  pl.n_mt = (a.B + kTileM - 1) / kTileM;
  pl.n_tiles = (int)((a.n + kTileN - 1) / kTileN);
  int s = a.sm_count / pl.n_mt;
  if (s < 1) s = 1;
  if (s > pl.n_tiles) s = pl.n_tiles;
  pl.n_slices = s;
  pl.grid = pl.n_mt * pl.n_slices;
  return pl;
}
}  // namespace

int s1_umma_plan(const ScanArgs& a, int* L, size_t* lists_keys) {
  if (a.dtype != TS_BF16 && a.dtype != TS_F16) { set_error("umma path needs bf16/fp16 storage"); return TS_ERR_UNSUPPORTED; }
  if (a.B > 1024) { set_error("umma path: at most 1024 queries per launch"); return TS_ERR_INVALID; }
  const UmmaPlan pl = umma_plan(a);
  *L = pl.n_slices;
  *lists_keys = (size_t)pl.grid * kTileM * cap_for_k(a.k);
  return TS_OK;
}

int launch_s1_umma(const ScanArgs& a, cudaStream_t st, int* launches) {
  int L; size_t lk;
  int rc = s1_umma_plan(a, &L, &lk);
  if (rc) return rc;
  const UmmaPlan pl = umma_plan(a);
  CUtensorMap tmQ, tmQ8, tmX;
  if ((rc = make_tmap_2d(&tmQ, a.q, a.dtype, a.B, a.dim, a.ld, kTileM))) return rc;
  if ((rc = make_tmap_2d(&tmQ8, a.q, a.dtype, a.B, a.dim, a.ld, 8))) return rc;
  if ((rc = make_tmap_2d(&tmX, a.rows, a.dtype, a.n, a.dim, a.ld, kTileN))) return rc;
  UmmaParams p{};
  p.N = a.n; p.nK = (a.dim + kChunkK - 1) / kChunkK; p.B = a.B; p.k = a.k;
  p.n_slices = pl.n_slices; p.n_tiles = pl.n_tiles;
  p.spread = (a.B <= 64) ? 1 : 0;
  p.n_qgroups = (a.B + 7) / 8;
  p.cap = cap_for_k(a.k);
  p.lists = a.lists; p.partial = a.partial; p.inv_norm = a.inv_norm;
  const bool bf16 = a.dtype == TS_BF16;
#define TS_LAUNCH(BF)                                                                                 \
  do {                                                                                                     \
    auto kern = s1_umma_kernel<BF>;                                                                   \
    TS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));       \
    kern<<<pl.grid, kThreads, kSmemBytes, st>>>(tmQ, tmQ8, tmX, p);                                        \
  } while (0)
  if (bf16) TS_LAUNCH(true); else TS_LAUNCH(false);
#undef TS_LAUNCH
  TS_CUDA_OK(cudaGetLastError());
  if (launches) ++*launches;
  return TS_OK;
}

}  // namespace ts
