// Shared device/host helpers for libtristage (sm_100a only).
//
// Candidate keys.  Every (score, row) candidate travels as one u64:
//     key = ord(score) << 32 | (0xFFFFFFFF - idx)
// ord() maps fp32 to an unsigned int with the same total order, so sorting
// keys DESCENDING yields "score descending, then idx ascending" -- the
// deterministic tie rule the oracle uses (oracle/flat_ip.py).  key 0 (the
// image of a negative NaN, never produced) is the empty-slot sentinel.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ts_launch.h"

namespace ts {

constexpr float kLowestF32 = -3.4028234663852886e38f;  // FAISS pad score

__host__ __device__ __forceinline__ uint32_t f2ord(float f) {
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(f);
#else
  union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float s, uint32_t idx) {
  return ((uint64_t)f2ord(s) << 32) | (uint64_t)(0xFFFFFFFFu - idx);
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return ord2f((uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_idx(uint64_t k) { return 0xFFFFFFFFu - (uint32_t)k; }

// ---------------------------------------------------------------------------
// Warp-wide bitonic sort, DESCENDING, of KPL*32 keys held KPL per lane.
// Element e lives in register v[e >> 5] of lane (e & 31): exchanges with
// stride >= 32 are register-register, strides < 32 are one shuffle.
// ---------------------------------------------------------------------------
template <int KPL>
__device__ __forceinline__ void warp_sort_desc(uint64_t (&v)[KPL], int lane) {
  constexpr int N = KPL * 32;
#pragma unroll
  for (int size = 2; size <= N; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
      if (stride >= 32) {
        const int js = stride >> 5;
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          if ((j & js) == 0) {
            const int p = j | js;
            const bool desc = (((j * 32) & size) == 0);
            const uint64_t a = v[j], b = v[p];
            const bool sw = desc ? (a < b) : (a > b);
            v[j] = sw ? b : a;
            v[p] = sw ? a : b;
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < KPL; ++j) {
          const uint64_t o = __shfl_xor_sync(0xffffffffu, v[j], stride);
          const int e = j * 32 + lane;
          const bool desc = ((e & size) == 0);
          const bool lower = ((lane & stride) == 0);
          const bool keep_max = (desc == lower);
          const uint64_t mx = v[j] > o ? v[j] : o;
          const uint64_t mn = v[j] > o ? o : v[j];
          v[j] = keep_max ? mx : mn;
        }
      }
    }
  }
}

// A candidate list in global memory owned by one warp (stream kernel) or one
// epilogue thread (umma kernel).  The whole warp sorts list[0..cnt) and keeps
// the best k in list[0..k), sorted descending, zero padded.  Returns the k-th
// key (0 if fewer than k entries).  CAP = KPL*32 >= cnt.
template <int KPL>
__device__ __noinline__ uint64_t warp_prune_list_t(uint64_t* list, int cnt, int k, int lane, uint64_t* also_out) {
  uint64_t v[KPL];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int e = j * 32 + lane;
    v[j] = (e < cnt) ? __ldcg(list + e) : 0ull;
  }
  warp_sort_desc<KPL>(v, lane);
  uint64_t kth_local = 0ull;
  const int kj = (k - 1) >> 5;
#pragma unroll
  for (int j = 0; j < KPL; ++j) {
    const int e = j * 32 + lane;
    if (e < k) {
      list[e] = v[j];
      if (also_out) also_out[e] = v[j];
    }
    if (j == kj) kth_local = v[j];
  }
  __syncwarp();
  return __shfl_sync(0xffffffffu, kth_local, (k - 1) & 31);
}
// cap = 256 (k <= 128) or 1024 (k <= 512); one out-of-line copy of each sort
// per kernel keeps code size and register pressure of the scan loops small.
__device__ __forceinline__ uint64_t warp_prune_list(uint64_t* list, int cnt, int k, int lane, int cap,
                                                    uint64_t* also_out = nullptr) {
  return cap <= 256 ? warp_prune_list_t<8>(list, cnt, k, lane, also_out)
                    : warp_prune_list_t<32>(list, cnt, k, lane, also_out);
}

// ---------------------------------------------------------------------------
// Block-wide bitonic sort (descending) of n = 2^m keys in shared memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void block_sort_desc(uint64_t* s, int n) {
  for (int size = 2; size <= n; size <<= 1) {
    for (int stride = size >> 1, lg = 31 - __clz(size >> 1); stride >= 1; stride >>= 1, --lg) {
      for (int i = threadIdx.x; i < (n >> 1); i += blockDim.x) {
        const int e = ((i >> lg) << (lg + 1)) | (i & (stride - 1));   // strides are powers of two
        const int p = e | stride;
        const bool desc = ((e & size) == 0);
        const uint64_t a = s[e], b = s[p];
        const bool sw = desc ? (a < b) : (a > b);
        if (sw) { s[e] = b; s[p] = a; }
      }
      __syncthreads();
    }
  }
}

__host__ __device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// Select the best k of `total` keys produced by load(i) into sbuf[0..k),
// sorted descending (zero padded).  sbuf holds capb (power of two, > k) keys.
// All threads of the block must call; ends with a barrier.
template <class Load>
__device__ __forceinline__ void block_select_topk(uint64_t* sbuf, int capb, int k, int total, Load load) {
  const int room = capb - k;
  for (int i = threadIdx.x; i < k; i += blockDim.x) sbuf[i] = 0ull;
  __syncthreads();
  for (int off = 0; off < total || off == 0; off += room) {
    const int take = min(room, total - off);
    const int n = next_pow2(max(k + take, 2));
    for (int i = threadIdx.x; i < n - k; i += blockDim.x) sbuf[k + i] = (i < take) ? load(off + i) : 0ull;
    __syncthreads();
    block_sort_desc(sbuf, n);
    if (total == 0) break;
  }
}

// ----------------------------------------------------------------- dtypes ---
template <typename T> struct Elem;
template <> struct Elem<float> {
  static constexpr int kPerChunk = 4;  // elements per 16-byte chunk
  static __device__ __forceinline__ void unpack(const uint4& c, float (&f)[4]) {
    f[0] = __uint_as_float(c.x); f[1] = __uint_as_float(c.y);
    f[2] = __uint_as_float(c.z); f[3] = __uint_as_float(c.w);
  }
  static __device__ __forceinline__ float to_f32(float v) { return v; }
  static __device__ __forceinline__ float from_f32(float v) { return v; }
};
template <> struct Elem<__nv_bfloat16> {
  static constexpr int kPerChunk = 8;
  static __device__ __forceinline__ void unpack(const uint4& c, float (&f)[8]) {
    f[0] = __uint_as_float(c.x << 16); f[1] = __uint_as_float(c.x & 0xffff0000u);
    f[2] = __uint_as_float(c.y << 16); f[3] = __uint_as_float(c.y & 0xffff0000u);
    f[4] = __uint_as_float(c.z << 16); f[5] = __uint_as_float(c.z & 0xffff0000u);
    f[6] = __uint_as_float(c.w << 16); f[7] = __uint_as_float(c.w & 0xffff0000u);
  }
  static __device__ __forceinline__ float to_f32(__nv_bfloat16 v) { return __bfloat162float(v); }
  static __device__ __forceinline__ __nv_bfloat16 from_f32(float v) { return __float2bfloat16_rn(v); }
};
template <> struct Elem<__half> {
  static constexpr int kPerChunk = 8;
  static __device__ __forceinline__ void unpack(const uint4& c, float (&f)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&c);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(h[i]);
      f[2 * i] = t.x; f[2 * i + 1] = t.y;
    }
  }
  static __device__ __forceinline__ float to_f32(__half v) { return __half2float(v); }
  static __device__ __forceinline__ __half from_f32(float v) { return __float2half_rn(v); }
};

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
#ifdef TS_CUDASIM
  return *p;
#else
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
#endif
}

// element index of (store row, column) in a token shard (TokLayout in ts_internal.h: 0 row-major, 1 tile)
__host__ __device__ __forceinline__ size_t tok_elem(int layout, long long row, int c, int dim) {
  if (layout == 0) return (size_t)row * dim + c;
  return (size_t)(row >> 3) * 8 * dim + (size_t)(c >> 3) * 64 + (size_t)(row & 7) * 8 + (c & 7);
}

__device__ __forceinline__ unsigned long long ts_globaltimer() {
#ifdef TS_CUDASIM
  return 0ull;
#else
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
#endif
}
// programmatic dependent launch (see TS_LAUNCH_PDL): no-ops when the kernel was launched the ordinary way
__device__ __forceinline__ void grid_dep_wait() {
#ifndef TS_CUDASIM
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void grid_dep_launch() {
#ifndef TS_CUDASIM
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}

__device__ __forceinline__ int warp_sum_int(int v) {
#if defined(TS_CUDASIM)
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
#else
  return __reduce_add_sync(0xffffffffu, v);      // redux.sync: one instruction
#endif
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace ts
