// TEST INFRASTRUCTURE ONLY (see tests/hostsim/cuda_runtime.h).  Stand-ins for the kernel
// launchers that csrc/api.cu calls: the two INGEST launchers are restated on the CPU so the
// host-side bookkeeping around them (growth, staging, padding, shard files) can be checked
// end to end; the scan / scoring / selection launchers are NOT simulated -- they fail with
// TS_ERR_UNSUPPORTED, there is no CPU compute path for them anywhere in this repository.
#include <math.h>

#include <vector>

#include "ts_internal.h"

extern "C" {
long long hostsim_live_allocs = 0;
long long hostsim_live_pinned = 0;
long long hostsim_fail_malloc_over = 0;
int hostsim_is_simulation(void) { return 1; }
}

namespace {
uint16_t f32_to_bf16(float f) {          // round to nearest even, NaN kept quiet
  uint32_t u; memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
float bf16_to_f32(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }
uint16_t f32_to_f16(float f) { _Float16 h = (_Float16)f; uint16_t u; memcpy(&u, &h, 2); return u; }
float f16_to_f32(uint16_t u) { _Float16 h; memcpy(&h, &u, 2); return (float)h; }
float load(const void* p, int dt, size_t i) {
  if (dt == TS_F32) return ((const float*)p)[i];
  return dt == TS_BF16 ? bf16_to_f32(((const uint16_t*)p)[i]) : f16_to_f32(((const uint16_t*)p)[i]);
}
void store(void* p, int dt, size_t i, float v) {
  if (dt == TS_F32) ((float*)p)[i] = v;
  else ((uint16_t*)p)[i] = dt == TS_BF16 ? f32_to_bf16(v) : f32_to_f16(v);
}
}  // namespace

namespace ts {

int launch_convert_rows(const void* src, int sdt, int64_t src_ld, void* dst, int ddt, int64_t dst_ld, int64_t n, int dim,
                        int norm_mode, float* inv_norm_out, cudaStream_t, unsigned int* zero_word, unsigned long long*) {
  if (zero_word) for (int i = 0; i < 16; ++i) zero_word[i] = 0u;
  for (int64_t r = 0; r < n; ++r) {
    float denom = 1.f; bool scale = false;
    if (norm_mode != kNormNone) {
      float ss = 0.f;
      for (int c = 0; c < dim; ++c) { const float v = load(src, sdt, (size_t)(r * src_ld + c)); ss += v * v; }
      const float nrm = sqrtf(ss);
      denom = norm_mode == kNormStage1 ? nrm + 1e-8f : fmaxf(nrm, 1e-12f);
      if (inv_norm_out) inv_norm_out[r] = 1.0f / denom; else scale = true;
    }
    for (int64_t c = 0; c < dst_ld; ++c) {
      float v = c < dim ? load(src, sdt, (size_t)(r * src_ld + c)) : 0.f;
      if (scale) v /= denom;
      store(dst, ddt, (size_t)(r * dst_ld + c), v);
    }
  }
  return TS_OK;
}

static size_t tok_elem(int layout, long long row, int c, int dim) {
  if (layout == kTokRowMajor) return (size_t)row * dim + c;
  return (size_t)(row >> 3) * 8 * dim + (size_t)(c >> 3) * 64 + (size_t)(row & 7) * 8 + (c & 7);
}

int launch_tok_ingest(const void* src, int sdt, const int64_t* so, const int64_t* dof, const int32_t* len, int n_docs,
                      void* dst, int ddt, int dim, int normalize, int layout, cudaStream_t) {
  for (int d = 0; d < n_docs; ++d) {
    const int L = len[d], Lp = (L + 7) & ~7;
    for (int r = 0; r < Lp; ++r) {
      if (r >= L && layout == kTokRowMajor) { for (int c = 0; c < dim; ++c) store(dst, ddt, tok_elem(layout, dof[d] + r, c, dim), 0.f); continue; }
      const size_t s = (size_t)(so[d] + (r < L ? r : L - 1)) * dim;   // tile layout: pad rows repeat the last token
      float denom = 1.f;
      if (normalize) {
        float ss = 0.f;
        for (int c = 0; c < dim; ++c) { const float v = load(src, sdt, s + c); ss += v * v; }
        denom = fmaxf(sqrtf(ss), 1e-12f);
      }
      for (int c = 0; c < dim; ++c) { float v = load(src, sdt, s + c); if (normalize) v /= denom; store(dst, ddt, tok_elem(layout, dof[d] + r, c, dim), v); }
    }
  }
  return TS_OK;
}

int launch_tok_relayout(void* tok, int dtype, const int64_t* doc_off, const int32_t* doc_len, int64_t doc_lo, int64_t n_docs,
                        int dim, int to_layout, cudaStream_t) {
  if (dtype == TS_F32 || dim % 8) return TS_ERR_UNSUPPORTED;
  uint16_t* t = (uint16_t*)tok;
  std::vector<uint16_t> g((size_t)8 * dim);
  const int from = to_layout == kTokTile ? kTokRowMajor : kTokTile;
  for (int64_t d = doc_lo; d < doc_lo + n_docs; ++d) {
    const int L = doc_len[d];
    for (int r0 = 0; r0 < L; r0 += 8) {
      const long long row0 = doc_off[d] + r0;
      const int valid = L - r0 < 8 ? L - r0 : 8;
      for (int r = 0; r < 8; ++r) for (int c = 0; c < dim; ++c) g[(size_t)r * dim + c] = t[tok_elem(from, row0 + r, c, dim)];
      for (int r = 0; r < 8; ++r) for (int c = 0; c < dim; ++c) {
        uint16_t v = 0;
        if (r < valid) v = g[(size_t)r * dim + c];
        else if (to_layout == kTokTile) v = g[(size_t)(valid - 1) * dim + c];
        t[tok_elem(to_layout, row0 + r, c, dim)] = v;
      }
    }
  }
  return TS_OK;
}

static int not_simulated(const char* what) {
  set_error("%s: not simulated -- the scan/scoring kernels only exist for sm_100a (hostsim is test infrastructure)", what);
  return TS_ERR_UNSUPPORTED;
}
int s1_stream_plan(const ScanArgs&, int*, size_t*) { return not_simulated("s1_stream"); }
int s1_umma_plan(const ScanArgs&, UmmaLayout*) { return not_simulated("s1_umma"); }
int launch_s1_stream(const ScanArgs&, cudaStream_t, int*) { return not_simulated("s1_stream"); }
int launch_s1_umma(const ScanArgs&, const UmmaLayout&, cudaStream_t, int*) { return not_simulated("s1_umma"); }
int launch_merge_keys(const uint64_t*, int, int, int, int64_t, uint64_t*, uint64_t*, float*, int64_t*, cudaStream_t, int*, const PushTarget*) { return not_simulated("merge_keys"); }
size_t merge_tmp_keys(int, int, int) { return 0; }
int launch_merge_lists(const uint64_t*, const int*, const float*, const UmmaLayout&, int, int, int64_t, float*, int64_t*, cudaStream_t, int*, const PushTarget*) { return not_simulated("merge_lists"); }
int launch_merge_pairs(const float*, const int64_t*, long long, long long, int, int, int, float*, int64_t*, cudaStream_t) { return not_simulated("merge_pairs"); }
int launch_rank_desc(const float*, const int32_t*, int, int, int, float*, int32_t*, cudaStream_t) { return not_simulated("rank_desc"); }
int launch_exchange_push(const void*, long long, const long long*, int, long long, long long, unsigned int, cudaStream_t) { return not_simulated("exchange_push"); }
int launch_exchange_wait_sum(const void*, long long, const unsigned int*, int, unsigned int, long long, float*, cudaStream_t) { return not_simulated("exchange_wait_sum"); }
int launch_merge_pairs_wait(const float*, const int64_t*, long long, long long, int, int, int, const unsigned int*, unsigned int, float*, int64_t*, cudaStream_t, int) { return not_simulated("merge_pairs_wait"); }
int launch_maxsim(const MaxSimArgs&, cudaStream_t, int*) { return not_simulated("maxsim"); }
bool maxsim_flow_takes(const MaxSimArgs&) { return false; }
int launch_exchange_wait_take(void*, const unsigned int*, int, unsigned int, long long, float*, cudaStream_t) { return not_simulated("exchange_wait_take"); }

}  // namespace ts
