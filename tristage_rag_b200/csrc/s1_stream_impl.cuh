// Stage-1 bandwidth path ("S1-stream"): exact inner-product top-k for small
// query batches (B <= 4 per pass) on the CUDA cores.
//
// Replaces faiss.IndexFlatIP.search for a handful of queries
// (/root/reference/src/stage1_retriever.py:380; the reference API is batch-1).
//
// Roofline: HBM.  Each pass reads the corpus shard exactly once --
// algorithmic bytes = N*ld*sizeof(T) (+ B*ld*sizeof(T) queries, + L*B*k*8
// partial keys) -- at 2*B flop per bf16 element, far below the ridge.
//
// Layout / mapping
//   * rows are [N][ld] with ld a multiple of 16 bytes, so a row is C = ld*s/16
//     128-bit chunks; lane l of a warp loads chunks l, l+32, ... of R = 4 rows
//     at once (coalesced 512-B warp requests, L1 no-allocate streaming loads,
//     >= 8 independent 128-bit loads in flight per lane);
//   * the NB queries live in shared memory as fp32; each lane keeps R*NB fp32
//     partial sums, reduced with warp shuffles once per row group;
//   * warps own interleaved row groups, so concurrently running warps touch
//     adjacent DRAM pages.
//
// Fused top-k: every reduced score is compared with the warp's running
// threshold tau[b] (the k-th best it has kept so far); survivors are appended
// to a per-(warp, query) candidate list in global memory (L2 resident) and the
// warp bitonic-sorts that list in registers whenever it fills (CAP entries),
// keeping the best k and raising tau.  At the end the CTA merges its warps'
// lists in shared memory and writes one sorted partial list per query; the
// select kernel (topk_select.cu) merges the per-CTA lists.  The [B, N] score
// matrix never exists.
#pragma once
#include "ts_common.cuh"
#include "ts_internal.h"

namespace ts {
namespace stream_impl {

constexpr int kStreamThreads = 256;
constexpr int kStreamWarps = kStreamThreads / 32;
constexpr int kRowsPerGroup = 4;
constexpr int kCtaMergeCap = 4096;  // keys (>= 8 warps * 512 / ... see plan)

template <typename T, int NB>
__global__ void __launch_bounds__(kStreamThreads, 2)
    s1_stream_kernel(const T* __restrict__ X, int64_t N, int ld, const T* __restrict__ Q,
                     const float* __restrict__ inv_norm, int k, int CAP, uint64_t* __restrict__ lists,
                     uint64_t* __restrict__ partial, int B_total, int b0) {
  constexpr int R = kRowsPerGroup;
  constexpr int EPC = Elem<T>::kPerChunk;
  TS_DYN_SMEM(unsigned char, smem_raw);
  uint64_t* sbuf = reinterpret_cast<uint64_t*>(smem_raw);                       // kCtaMergeCap keys
  float* Qs = reinterpret_cast<float*>(smem_raw + kCtaMergeCap * sizeof(uint64_t));  // [NB][ld]

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < NB * ld; i += blockDim.x) Qs[i] = Elem<T>::to_f32(Q[i]);
  __syncthreads();

  const int C = ld / EPC;
  const uint4* Xc = reinterpret_cast<const uint4*>(X);
  const int64_t warps_total = (int64_t)gridDim.x * kStreamWarps;
  const int64_t gw = (int64_t)blockIdx.x * kStreamWarps + warp;
  uint64_t* my_lists = lists + (size_t)gw * NB * CAP;

  float tau[NB];
  int cnt[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) { tau[b] = -INFINITY; cnt[b] = 0; }

  const int64_t n_groups = (N + R - 1) / R;
  for (int64_t g = gw; g < n_groups; g += warps_total) {
    const int64_t row0 = g * R;
    const uint4* xr[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int64_t row = (row0 + r < N) ? (row0 + r) : (N - 1);
      xr[r] = Xc + row * C;
    }
    float acc[R][NB];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = 0.f;

#pragma unroll 1
    for (int c = lane; c < C; c += 32) {
      uint4 xv[R];
#pragma unroll
      for (int r = 0; r < R; ++r) xv[r] = ldg_stream(xr[r] + c);
      float xf[R][EPC];
#pragma unroll
      for (int r = 0; r < R; ++r) Elem<T>::unpack(xv[r], xf[r]);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float4* qp = reinterpret_cast<const float4*>(Qs + (size_t)b * ld + (size_t)c * EPC);
        float qf[EPC];
#pragma unroll
        for (int v = 0; v < EPC / 4; ++v) {
          const float4 t = qp[v];
          qf[4 * v] = t.x; qf[4 * v + 1] = t.y; qf[4 * v + 2] = t.z; qf[4 * v + 3] = t.w;
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
          for (int e = 0; e < EPC; ++e) acc[r][b] = fmaf(xf[r][e], qf[e], acc[r][b]);
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[r][b] = warp_sum(acc[r][b]);

    // fused threshold filter (warp-uniform: every lane holds the same sums)
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (row0 + r < N) {
        const float scale = inv_norm ? __ldg(inv_norm + row0 + r) : 1.0f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const float s = acc[r][b] * scale;
          if (s > tau[b]) {
            if (lane == 0) my_lists[(size_t)b * CAP + cnt[b]] = make_key(s, (uint32_t)(row0 + r));
            ++cnt[b];
          }
        }
      }
    }
    // a row group adds at most R entries per query: prune with that margin
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      if (cnt[b] > CAP - R) {
        const uint64_t kth = warp_prune_list(my_lists + (size_t)b * CAP, cnt[b], k, lane, CAP);
        cnt[b] = k;
        tau[b] = key_score(kth);
      }
    }
  }

  // final per-warp sort: list[0..k) sorted descending, zero padded
#pragma unroll
  for (int b = 0; b < NB; ++b) warp_prune_list(my_lists + (size_t)b * CAP, cnt[b], k, lane, CAP);
  __threadfence_block();
  __syncthreads();

  // CTA merge of the 8 warp lists -> one partial list per query
  const uint64_t* cta_lists = lists + (size_t)blockIdx.x * kStreamWarps * NB * CAP;
  for (int b = 0; b < NB; ++b) {
    auto load = [&](int i) -> uint64_t {
      const int w = i / k, r = i % k;
      return __ldcg(cta_lists + ((size_t)w * NB + b) * CAP + r);
    };
    block_select_topk(sbuf, kCtaMergeCap, k, kStreamWarps * k, load);
    uint64_t* out = partial + ((size_t)blockIdx.x * B_total + (b0 + b)) * k;
    for (int r = threadIdx.x; r < k; r += blockDim.x) out[r] = sbuf[r];
    __syncthreads();
  }
}

inline int stream_grid(const ScanArgs& a) {
  const int64_t n_groups = (a.n + kRowsPerGroup - 1) / kRowsPerGroup;
  int64_t want = (n_groups + kStreamWarps - 1) / kStreamWarps;
  const int64_t cap = (int64_t)a.sm_count * 2;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}

template <typename T, int NB>
int launch_nb(const ScanArgs& a, int b0, cudaStream_t st) {
  const int grid = stream_grid(a);
  const size_t smem = kCtaMergeCap * sizeof(uint64_t) + (size_t)NB * a.ld * sizeof(float);
  auto kern = s1_stream_kernel<T, NB>;
  if (smem > 48 * 1024) TS_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const T* q = reinterpret_cast<const T*>(a.q) + (size_t)b0 * a.ld;
  TS_LAUNCH(kern, grid, kStreamThreads, smem, st, reinterpret_cast<const T*>(a.rows), a.n, a.ld, q, a.inv_norm, a.k,
            cap_for_k(a.k), a.lists, a.partial, a.B, b0);
  TS_CUDA_OK(cudaGetLastError());
  return TS_OK;
}

template <typename T>
int launch_t(const ScanArgs& a, cudaStream_t st, int* launches) {
  for (int b0 = 0; b0 < a.B; b0 += 4) {
    const int nb = (a.B - b0) < 4 ? (a.B - b0) : 4;
    int rc;
    switch (nb) {
      case 1: rc = launch_nb<T, 1>(a, b0, st); break;
      case 2: rc = launch_nb<T, 2>(a, b0, st); break;
      case 3: rc = launch_nb<T, 3>(a, b0, st); break;
      default: rc = launch_nb<T, 4>(a, b0, st); break;
    }
    if (rc) return rc;
    if (launches) ++*launches;
  }
  return TS_OK;
}

}  // namespace stream_impl
}  // namespace ts
