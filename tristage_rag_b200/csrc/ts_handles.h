// Definitions of the opaque ABI handles (ts_index, ts_tokstore) and the helpers
// shared by api.cu and shard_file.cu.  Internal: never included by callers of
// include/tristage.h.
#pragma once
#include "ts_internal.h"

namespace ts {
// ring of CUDA event pairs around the dominant kernel (measurement aid)
struct ScanTimer {
  static constexpr int kMax = 256;
  bool on = false;
  int n = 0;
  cudaEvent_t e0[kMax], e1[kMax];
  bool made = false;
  void begin(cudaStream_t st) {
    if (!on || n >= kMax) return;
    if (!made) { for (int i = 0; i < kMax; ++i) { cudaEventCreate(&e0[i]); cudaEventCreate(&e1[i]); } made = true; }
    cudaEventRecord(e0[n], st);
  }
  void end(cudaStream_t st) {
    if (!on || n >= kMax) return;
    cudaEventRecord(e1[n], st);
    ++n;
  }
  int report(float* mean_ms, int* count) {
    float tot = 0.f;
    for (int i = 0; i < n; ++i) {
      if (cudaEventSynchronize(e1[i]) != cudaSuccess) { set_error("event sync failed"); return TS_ERR_CUDA; }
      float ms = 0.f;
      cudaEventElapsedTime(&ms, e0[i], e1[i]);
      tot += ms;
    }
    if (mean_ms) *mean_ms = n ? tot / n : 0.f;
    if (count) *count = n;
    n = 0;
    return TS_OK;
  }
  void destroy() { if (made) { for (int i = 0; i < kMax; ++i) { cudaEventDestroy(e0[i]); cudaEventDestroy(e1[i]); } made = false; } }
};

// grow-only device scratch
int ensure_bytes(void** p, size_t* cur, size_t need);
int check_device(int device, DeviceInfo* info);
}  // namespace ts

struct ts_index {
  int device, dim, ld, dtype, metric;
  int64_t n, cap, id_base;
  void* rows;
  float* inv_norm;
  ts::DeviceInfo info;
  int64_t launches;
  // scratch (grow-only)
  void* qbuf; size_t qbuf_b;
  void* lists; size_t lists_b;
  void* partial; size_t partial_b;
  void* tmp0; size_t tmp0_b;
  void* tmp1; size_t tmp1_b;
  void* counts; size_t counts_b;
  void* pub; size_t pub_b;
  unsigned int* grid_bar;            // 16 words zeroed by every query-prep launch: grid-barrier arrivals + next-tile counters of the scan
  int coop;                          // device supports cooperative launch
  void* stage; size_t stage_b;       // staging for host inputs (add / search_host)
  void* hout; size_t hout_b;         // device result buffers for search_host
  ts::ScanTimer* timer;
  int64_t reset_gen;                 // bumped by ts_index_reset: views over the rows (ts_ivf) drop their state
};

struct ts_tokstore {
  int device, dim, dtype;
  int layout;                       // ts::TokLayout of tok (fixed at creation)
  int64_t ndocs, nrows, cap_docs, cap_rows, id_base, ntokens;
  int64_t hint_docs, hint_rows;     // reservation hints honoured by the first add
  void* tok;
  int64_t* doc_off;
  int32_t* doc_len;
  ts::DeviceInfo info;
  int64_t launches;
  void* qbuf; size_t qbuf_b;
  void* stage; size_t stage_b;
  void* meta; size_t meta_b;         // per-add src offsets scratch / host-variant buffers
  void* hbuf; size_t hbuf_b;
  unsigned int* done_ctr;           // last-CTA election of the multi-GPU scatter (ts_maxsim_scatter), zero between launches
  void* scratch_out; size_t scratch_out_b;   // local [B][C] matrix of a scatter call (the result lives in the receive buffers)
  ts::ScanTimer* timer;
};


namespace ts {
// make room for `rows` corpus rows (doubling growth; existing rows are copied on `st`)
int index_reserve(ts_index* h, int64_t rows, cudaStream_t st);
// make room for `docs` doc-table entries and `rows` padded token rows
int tok_reserve(ts_tokstore* h, int64_t docs, int64_t rows, cudaStream_t st);
}  // namespace ts
