// Token-store ingest: scatter ragged per-doc token matrices into the 8-row
// padded store layout, L2-normalising each token (F.normalize,
// /root/reference/src/stage2_rescorer.py:174) and casting to the storage
// dtype.  The store is what encode_documents_batch (:207-242) would return
// for every document, kept resident instead of being recomputed per query.
//
// Bound: HBM, one read of the source rows + one write of the padded rows.
// One CTA per doc, one warp per token row.
//
// Two shard layouts (TokLayout, ts_internal.h): row-major with zero pad rows, or the tile layout --
// tok[row/8][dim/8][row%8][8] with the doc's last token repeated in the pad rows -- which is the
// shared-memory image the Stage-2 tensor kernel (s2_flow.cu) feeds to tcgen05.mma, so that a doc
// moves into a tile with one contiguous bulk copy.  tok_relayout_kernel converts a range of docs
// in place (shard files keep the row-major image; ts_tokstore_save / append_file convert).
#include "ts_common.cuh"
#include "ts_internal.h"

namespace ts {
namespace {

template <typename TS, typename TD>
__global__ void __launch_bounds__(128)
    tok_ingest_kernel(const TS* __restrict__ src, const int64_t* __restrict__ src_off,
                      const int64_t* __restrict__ dst_off, const int32_t* __restrict__ lens, int n_docs,
                      TD* __restrict__ dst, int dim, int normalize, int layout) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int doc = blockIdx.x; doc < n_docs; doc += gridDim.x) {
    const int L = lens[doc];
    const int Lp = (L + 7) & ~7;
    const TS* s0 = src + src_off[doc] * dim;
    const long long row0 = dst_off[doc];
    for (int r = warp; r < Lp; r += nw) {
      if (r >= L && layout == kTokRowMajor) {
        for (int c = lane; c < dim; c += 32) dst[tok_elem(layout, row0 + r, c, dim)] = Elem<TD>::from_f32(0.f);
        continue;
      }
      const TS* s = s0 + (size_t)(r < L ? r : L - 1) * dim;     // tile layout: pad rows repeat the last token
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
        dst[tok_elem(layout, row0 + r, c, dim)] = Elem<TD>::from_f32(v);
      }
    }
  }
}

// In-place layout change of docs [doc_lo, doc_lo + n): one warp per doc walks its 8-row groups; a group
// (16 * dim bytes) is staged in shared memory as 16-byte chunks and written back permuted.  Chunk (r, j) =
// row r of the group, columns [8j, 8j + 8): row-major index r * (dim/8) + j, tile index j * 8 + r.  Going to
// the tile layout the pad rows of the last group take the doc's last token, going back they are zeroed.
__global__ void __launch_bounds__(128)
    tok_relayout_kernel(uint4* __restrict__ tok, const int64_t* __restrict__ doc_off, const int32_t* __restrict__ doc_len,
                        long long doc_lo, long long n_docs, int dim, int to_layout) {
  TS_DYN_SMEM(uint4, sm);                       // [warps][dim] chunks
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int kc = dim >> 3;                      // chunks per row
  uint4* buf = sm + (size_t)warp * dim;
  for (long long d = (long long)blockIdx.x * nw + warp; d < n_docs; d += (long long)gridDim.x * nw) {
    const int L = doc_len[doc_lo + d];
    const int groups = (L + 7) >> 3;
    uint4* g0 = tok + (size_t)(doc_off[doc_lo + d] >> 3) * dim;
    for (int g = 0; g < groups; ++g) {
      uint4* gp = g0 + (size_t)g * dim;
      for (int i = lane; i < dim; i += 32) buf[i] = gp[i];
      __syncwarp();
      const int valid = min(8, L - g * 8);      // real rows in this group
      for (int i = lane; i < dim; i += 32) {
        // i = destination chunk index in the target layout
        const int r = to_layout == kTokTile ? (i & 7) : (i / kc);
        const int j = to_layout == kTokTile ? (i >> 3) : (i % kc);
        uint4 v;
        if (r < valid) {
          v = buf[to_layout == kTokTile ? r * kc + j : j * 8 + r];
        } else if (to_layout == kTokTile) {
          v = buf[(valid - 1) * kc + j];        // pad row: the doc's last token
        } else {
          v.x = v.y = v.z = v.w = 0u;
        }
        gp[i] = v;
      }
      __syncwarp();
    }
  }
}

template <typename TS, typename TD>
int launch_t(const void* src, const int64_t* so, const int64_t* dof, const int32_t* len, int n_docs, void* dst, int dim,
             int normalize, int layout, cudaStream_t st) {
  int grid = n_docs < 148 * 64 ? n_docs : 148 * 64;
  auto kern = tok_ingest_kernel<TS, TD>;
  TS_LAUNCH(kern, grid, 128, 0, st, (const TS*)src, so, dof, len, n_docs, (TD*)dst, dim, normalize, layout);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

}  // namespace

int launch_tok_relayout(void* tok, int dtype, const int64_t* doc_off, const int32_t* doc_len, int64_t doc_lo, int64_t n_docs,
                        int dim, int to_layout, cudaStream_t st) {
  if (n_docs <= 0) return TS_OK;
  if (dtype == TS_F32 || dim % 8 != 0 || dim > 1024) { set_error("tok_relayout: 2-byte dtypes with dim %% 8 == 0 only"); return TS_ERR_UNSUPPORTED; }
  const int warps = 4;
  long long blocks = (n_docs + warps - 1) / warps;
  const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);
  TS_LAUNCH(tok_relayout_kernel, grid, warps * 32, (size_t)warps * dim * sizeof(uint4), st, (uint4*)tok, doc_off, doc_len,
            (long long)doc_lo, (long long)n_docs, dim, to_layout);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

int launch_tok_ingest(const void* src, int src_dtype, const int64_t* so, const int64_t* dof, const int32_t* len,
                      int n_docs, void* dst, int dst_dtype, int dim, int normalize, int layout, cudaStream_t st) {
  if (layout == kTokTile && !tok_tile_layout_ok(dim, dst_dtype)) { set_error("tok_ingest: tile layout needs a 2-byte dtype and dim %% 16 == 0"); return TS_ERR_INVALID; }
#define TS_CASE(SD, ST_, DD, DT_) \
  if (src_dtype == SD && dst_dtype == DD) return launch_t<ST_, DT_>(src, so, dof, len, n_docs, dst, dim, normalize, layout, st);
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
