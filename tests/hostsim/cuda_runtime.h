// TEST INFRASTRUCTURE ONLY -- a stand-in for <cuda_runtime.h> that lets the HOST side of
// libtristage (csrc/api.cu, csrc/shard_file.cu: argument checks, growth, staging, shard files)
// be compiled with g++ and exercised on a machine without a GPU.  "Device" memory is host
// memory, streams are synchronous, events are no-ops.  Nothing under tristage_rag_b200/ ever
// includes this header or links the library built from it (tests/hostsim/Makefile); the
// product library always builds against the real CUDA runtime.
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorInvalidValue = 1 };
typedef struct hostsim_stream* cudaStream_t;
typedef struct hostsim_event* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToHost = 0, cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaHostAllocDefault = 0, cudaEventDisableTiming = 2, cudaEnableDefault = 0 };
enum cudaDeviceAttr { cudaDevAttrCooperativeLaunch = 95 };
struct cudaDeviceProp {
  int multiProcessorCount, major, minor;
  size_t sharedMemPerBlockOptin;
};

// bookkeeping the tests read back (tests/hostsim/sim_launchers.cc)
extern "C" {
extern long long hostsim_live_allocs;      // cudaMalloc'ed blocks not yet freed
extern long long hostsim_live_pinned;      // cudaHostAlloc'ed blocks not yet freed
extern long long hostsim_fail_malloc_over; // cudaMalloc of more bytes than this fails (0 = never)
}

static inline const char* cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "hostsim error"; }
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
static inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
static inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) {
  const char* sms = getenv("HOSTSIM_SM_COUNT");          // tests shrink the "GPU" to keep emulated grids small
  p->multiProcessorCount = (sms && atoi(sms) > 0) ? atoi(sms) : 148; p->major = 10; p->minor = 0; p->sharedMemPerBlockOptin = 232448;
  return cudaSuccess;
}
static inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr, int) { *v = 1; return cudaSuccess; }   // cooperative launch: yes
static inline cudaError_t cudaMalloc(void** p, size_t n) {
  if (hostsim_fail_malloc_over > 0 && (long long)n > hostsim_fail_malloc_over) { *p = nullptr; return cudaErrorMemoryAllocation; }
  // 0xA5 fill: reading memory the library never wrote shows up in comparisons
  *p = malloc(n ? n : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  memset(*p, 0xA5, n);
  ++hostsim_live_allocs;
  return cudaSuccess;
}
static inline cudaError_t cudaFree(void* p) { if (p) { free(p); --hostsim_live_allocs; } return cudaSuccess; }
static inline cudaError_t cudaHostAlloc(void** p, size_t n, unsigned) {
  *p = malloc(n ? n : 1);
  if (!*p) return cudaErrorMemoryAllocation;
  ++hostsim_live_pinned;
  return cudaSuccess;
}
static inline cudaError_t cudaFreeHost(void* p) { if (p) { free(p); --hostsim_live_pinned; } return cudaSuccess; }
static inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
static inline cudaError_t cudaMemset(void* d, int v, size_t n) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { memset(d, v, n); return cudaSuccess; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
static inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
