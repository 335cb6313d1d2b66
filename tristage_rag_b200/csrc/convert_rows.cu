// Ingest / query-prep kernel: cast rows to the storage dtype, optionally
// L2-normalising them first.
//
// Replaces _normalize_embeddings + faiss_index.add for documents
// (/root/reference/src/stage1_retriever.py:285-288, :307, :270/277/313) and the
// query normalisation at :377; for Stage 2 it applies F.normalize
// (/root/reference/src/stage2_rescorer.py:173-174) once at ingest instead of
// once per (query, candidate).
//
// Bound: HBM, one read + one write of the rows (n*dim*(src+dst bytes)); one
// warp per row, fp32 sum of squares, true division like numpy.
#include "ts_common.cuh"
#include "ts_internal.h"

namespace ts {
namespace {

template <typename TS, typename TD>
__global__ void __launch_bounds__(256) convert_rows_kernel(const TS* __restrict__ src, int64_t src_ld,
                                                           TD* __restrict__ dst, int64_t dst_ld, int64_t n, int dim,
                                                           int norm_mode, float* __restrict__ inv_norm_out,
                                                           unsigned int* __restrict__ zero_word,
                                                           unsigned long long* __restrict__ tl) {
  const int lane = threadIdx.x & 31;
  // TS_DBG_TIMELINE: GPU time line of one search step (ts_index_debug_timeline): this kernel's start, and the reset of
  // the scan's first-entry / last-exit slots
  if (tl && blockIdx.x == 0 && threadIdx.x == 0) { tl[0] = ts_globaltimer(); tl[2] = ~0ull; tl[3] = 0ull; }
  // query prep of a search: reset the 16 scheduling words of the scan (grid-barrier arrivals, next-tile counters:
  // s1_umma.cu) -- the kernel boundary orders these stores before the scan, so neither needs a reset protocol
  if (zero_word && blockIdx.x == 0 && threadIdx.x < 16) zero_word[threadIdx.x] = 0u;
  const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < n; row += warps) {
    const TS* s = src + row * src_ld;
    TD* d = dst + row * dst_ld;
    float denom = 1.0f;
    bool scale = false;
    if (norm_mode != kNormNone) {
      float ss = 0.f;
#pragma unroll 8
      for (int c = lane; c < dim; c += 32) {
        const float v = Elem<TS>::to_f32(s[c]);
        ss = fmaf(v, v, ss);
      }
      ss = warp_sum(ss);
      const float nrm = sqrtf(ss);
      denom = (norm_mode == kNormStage1) ? (nrm + 1e-8f) : fmaxf(nrm, 1e-12f);
      if (inv_norm_out) {
        if (lane == 0) inv_norm_out[row] = __fdiv_rn(1.0f, denom);
      } else {
        scale = true;
      }
    }
#pragma unroll 8
    for (int c = lane; c < (int)dst_ld; c += 32) {     // eight independent loads in flight (the query prep of a search is latency-bound)
      float v = (c < dim) ? Elem<TS>::to_f32(s[c]) : 0.f;
      if (scale) v = __fdiv_rn(v, denom);
      d[c] = Elem<TD>::from_f32(v);
    }
  }
  if (tl && blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x == 0) tl[1] = ts_globaltimer();
  }
}

template <typename TS, typename TD>
int launch_t(const void* src, int64_t src_ld, void* dst, int64_t dst_ld, int64_t n, int dim, int norm_mode,
             float* inv, cudaStream_t st, unsigned int* zero_word, unsigned long long* tl) {
  if (n == 0) return TS_OK;
  const int warps_per_block = 8;
  int64_t blocks = (n + warps_per_block - 1) / warps_per_block;
  if (blocks > 148 * 16) blocks = 148 * 16;
  auto kern = convert_rows_kernel<TS, TD>;
  TS_LAUNCH(kern, (unsigned)blocks, 256, 0, st, (const TS*)src, src_ld, (TD*)dst, dst_ld, n, dim, norm_mode, inv, zero_word, tl);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

}  // namespace

int launch_convert_rows(const void* src, int src_dtype, int64_t src_ld, void* dst, int dst_dtype, int64_t dst_ld,
                        int64_t n, int dim, int norm_mode, float* inv, cudaStream_t st, unsigned int* zero_word,
                        unsigned long long* tl) {
#define TS_CASE(SD, ST_, DD, DT_)                                                                 \
  if (src_dtype == SD && dst_dtype == DD)                                                         \
    return launch_t<ST_, DT_>(src, src_ld, dst, dst_ld, n, dim, norm_mode, inv, st, zero_word, tl);
  TS_CASE(TS_F32, float, TS_F32, float)
  TS_CASE(TS_F32, float, TS_BF16, __nv_bfloat16)
  TS_CASE(TS_F32, float, TS_F16, __half)
  TS_CASE(TS_BF16, __nv_bfloat16, TS_BF16, __nv_bfloat16)
  TS_CASE(TS_F16, __half, TS_F16, __half)
  TS_CASE(TS_BF16, __nv_bfloat16, TS_F32, float)
  TS_CASE(TS_F16, __half, TS_F32, float)
#undef TS_CASE
  set_error("convert_rows: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  return TS_ERR_UNSUPPORTED;
}

}  // namespace ts
