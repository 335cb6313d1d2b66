// Token-store ingest: scatter ragged per-doc token matrices into the 8-row
// padded store layout, L2-normalising each token (F.normalize,
// /root/reference/src/stage2_rescorer.py:174) and casting to the storage
// dtype.  The store is what encode_documents_batch (:207-242) would return
// for every document, kept resident instead of being recomputed per query.
//
// Bound: HBM, one read of the source rows + one write of the padded rows.
// One CTA per doc, one warp per token row.
#include "ts_common.cuh"
#include "ts_internal.h"

namespace ts {
namespace {

template <typename TS, typename TD>
__global__ void __launch_bounds__(128)
    tok_ingest_kernel(const TS* __restrict__ src, const int64_t* __restrict__ src_off,
                      const int64_t* __restrict__ dst_off, const int32_t* __restrict__ lens, int n_docs,
                      TD* __restrict__ dst, int dim, int normalize) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int doc = blockIdx.x; doc < n_docs; doc += gridDim.x) {
    const int L = lens[doc];
    const int Lp = (L + 7) & ~7;
    const TS* s0 = src + src_off[doc] * dim;
    TD* d0 = dst + dst_off[doc] * dim;
    for (int r = warp; r < Lp; r += nw) {
      TD* d = d0 + (size_t)r * dim;
      if (r >= L) {
        for (int c = lane; c < dim; c += 32) d[c] = Elem<TD>::from_f32(0.f);
        continue;
      }
      const TS* s = s0 + (size_t)r * dim;
      float denom = 1.f;
      if (normalize) {
        float ss = 0.f;
        for (int c = lane; c < dim; c += 32) { const float v = Elem<TS>::to_f32(s[c]); ss = fmaf(v, v, ss); }
        ss = warp_sum(ss);
        denom = fmaxf(sqrtf(ss), 1e-12f);
      }
      for (int c = lane; c < dim; c += 32) {
        float v = Elem<TS>::to_f32(s[c]);
        if (normalize) v = __fdiv_rn(v, denom);
        d[c] = Elem<TD>::from_f32(v);
      }
    }
  }
}

template <typename TS, typename TD>
int launch_t(const void* src, const int64_t* so, const int64_t* dof, const int32_t* len, int n_docs, void* dst, int dim,
             int normalize, cudaStream_t st) {
  int grid = n_docs < 148 * 64 ? n_docs : 148 * 64;
  auto kern = tok_ingest_kernel<TS, TD>;
  TS_LAUNCH(kern, grid, 128, 0, st, (const TS*)src, so, dof, len, n_docs, (TD*)dst, dim, normalize);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

}  // namespace

int launch_tok_ingest(const void* src, int src_dtype, const int64_t* so, const int64_t* dof, const int32_t* len,
                      int n_docs, void* dst, int dst_dtype, int dim, int normalize, cudaStream_t st) {
#define TS_CASE(SD, ST_, DD, DT_) \
  if (src_dtype == SD && dst_dtype == DD) return launch_t<ST_, DT_>(src, so, dof, len, n_docs, dst, dim, normalize, st);
  TS_CASE(TS_F32, float, TS_F32, float)
  TS_CASE(TS_F32, float, TS_BF16, __nv_bfloat16)
  TS_CASE(TS_F32, float, TS_F16, __half)
  TS_CASE(TS_BF16, __nv_bfloat16, TS_BF16, __nv_bfloat16)
  TS_CASE(TS_F16, __half, TS_F16, __half)
#undef TS_CASE
  set_error("tok_ingest: unsupported dtype pair %d -> %d", src_dtype, dst_dtype);
  return TS_ERR_UNSUPPORTED;
}

}  // namespace ts
