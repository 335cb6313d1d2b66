// TEST INFRASTRUCTURE ONLY -- a small SIMT emulator that lets the CUDA-core kernels of
// libtristage (csrc/convert_rows.cu, tok_ingest.cu, topk_select.cu, s1_stream_impl.cuh, the SIMT
// kernel of s2_maxsim.cu, ...) be compiled with g++ and EXECUTED on the CPU, thread by thread,
// with CUDA's block / warp semantics:
//   * every simulated thread is a fiber (ucontext); blocks run one after another on one OS
//     thread, so __shared__ variables are plain statics and atomics are plain operations;
//   * __syncthreads() and the *_sync warp primitives (shfl, shfl_xor, shfl_up, shfl_down,
//     ballot, any, syncwarp) are rendezvous points: a fiber yields until every thread named
//     in the barrier / mask has arrived; a barrier that can never complete (divergent
//     __syncthreads, a lane missing from a mask) aborts with a diagnostic instead of hanging;
//   * "device" memory is host memory (tests/hostsim/cuda_runtime.h).
// It checks LOGIC (indexing, reductions, selection, tie rules), not memory-model races and not
// performance.  The tensor-path kernels run on top of it through ts_ptx_sim.cuh, a functional model
// of the TMA / mbarrier / tcgen05 wrappers.  Nothing under tristage_rag_b200/ includes this header
// unless TS_CUDASIM is defined by tests/cudasim/Makefile.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <functional>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3 { unsigned x, y, z; };
struct uint4 { unsigned x, y, z, w; };
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };

namespace cudasim {
struct ThreadCtx {
  uint3 threadIdx_, blockIdx_;
  dim3 blockDim_, gridDim_;
};
extern ThreadCtx* g_cur;   // the running simulated thread
void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body);
void launch_cooperative(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()>& body);
void launch_cluster(dim3 grid, dim3 block, size_t dyn_smem_bytes, int cluster_size, const std::function<void()>& body);
void yield_spin();   // a spinning thread lets the others run (global-memory flags, mbarrier waits)
void* dyn_smem();
void sync_threads();
enum Collective { kShflIdx, kShflXor, kShflUp, kShflDown, kBallot, kSyncWarp };
uint64_t warp_collective(Collective kind, unsigned mask, uint64_t value, int aux);
extern unsigned long long g_launches, g_blocks_run, g_switches;   // statistics the tests read
}  // namespace cudasim

#define threadIdx (cudasim::g_cur->threadIdx_)
#define blockIdx (cudasim::g_cur->blockIdx_)
#define blockDim (cudasim::g_cur->blockDim_)
#define gridDim (cudasim::g_cur->gridDim_)

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__ __attribute__((noinline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __grid_constant__

inline void __syncthreads() { cudasim::sync_threads(); }
inline void __syncwarp(unsigned mask = 0xffffffffu) { cudasim::warp_collective(cudasim::kSyncWarp, mask, 0, 0); }
inline void __threadfence() {}
inline void __threadfence_block() {}
inline void __threadfence_system() {}
[[noreturn]] inline void __trap() { fprintf(stderr, "[cudasim] __trap()\n"); abort(); }
inline long long clock64() { return 0; }

namespace cudasim {
template <typename T> inline uint64_t to_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle payload"); memcpy(&b, &v, sizeof(T)); return b; }
template <typename T> inline T from_bits(uint64_t b) { T v; memcpy(&v, &b, sizeof(T)); return v; }
}  // namespace cudasim
template <typename T> inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  if (width != 32) { fprintf(stderr, "[cudasim] shuffle width %d unsupported\n", width); abort(); }
  return cudasim::from_bits<T>(cudasim::warp_collective(cudasim::kShflIdx, mask, cudasim::to_bits(v), src & 31));
}
template <typename T> inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
  if (width != 32) abort();
  return cudasim::from_bits<T>(cudasim::warp_collective(cudasim::kShflXor, mask, cudasim::to_bits(v), lane_mask));
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  if (width != 32) abort();
  return cudasim::from_bits<T>(cudasim::warp_collective(cudasim::kShflUp, mask, cudasim::to_bits(v), (int)delta));
}
template <typename T> inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
  if (width != 32) abort();
  return cudasim::from_bits<T>(cudasim::warp_collective(cudasim::kShflDown, mask, cudasim::to_bits(v), (int)delta));
}
inline unsigned __ballot_sync(unsigned mask, int pred) { return (unsigned)cudasim::warp_collective(cudasim::kBallot, mask, pred ? 1 : 0, 0); }
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == mask; }

// single OS thread: atomics are plain read-modify-writes
template <typename T, typename U> inline T atomicAdd(T* p, U v) { T old = *p; *p = (T)(old + (T)v); return old; }
template <typename T, typename U> inline T atomicExch(T* p, U v) { T old = *p; *p = (T)v; return old; }
template <typename T, typename U> inline T atomicMax(T* p, U v) { T old = *p; if ((T)v > old) *p = (T)v; return old; }
template <typename T, typename U> inline T atomicMin(T* p, U v) { T old = *p; if ((T)v < old) *p = (T)v; return old; }
template <typename T> inline T atomicCAS(T* p, T cmp, T v) { T old = *p; if (old == cmp) *p = v; return old; }

template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }

inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __expf(float x) { return expf(x); }
inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }

// CUDA's integer / float min-max overload set
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline long min(long a, long b) { return a < b ? a : b; }
inline long max(long a, long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }

// launch attributes are meaningless here
enum cudaFuncAttribute { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename F> inline int cudaFuncSetAttribute(F, cudaFuncAttribute, int) { return 0; }
