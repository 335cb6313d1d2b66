// Stage-1 bandwidth path: planning + dtype dispatch (kernel in s1_stream_impl.cuh).
#include "ts_internal.h"

namespace ts {
namespace stream_impl {
constexpr int kStreamWarps = 8;
constexpr int kRowsPerGroup = 4;
constexpr int kCtaMergeCap = 4096;
inline int stream_grid(const ScanArgs& a) {
  const int64_t n_groups = (a.n + kRowsPerGroup - 1) / kRowsPerGroup;
  int64_t want = (n_groups + kStreamWarps - 1) / kStreamWarps;
  const int64_t cap = (int64_t)a.sm_count * 2;
  if (want > cap) want = cap;
  if (want < 1) want = 1;
  return (int)want;
}
}  // namespace stream_impl
int launch_s1_stream_bf16(const ScanArgs& a, cudaStream_t st, int* launches);
int launch_s1_stream_f16(const ScanArgs& a, cudaStream_t st, int* launches);
int launch_s1_stream_f32(const ScanArgs& a, cudaStream_t st, int* launches);

int s1_stream_plan(const ScanArgs& a, int* L, size_t* lists_keys) {
  const int grid = stream_impl::stream_grid(a);
  *L = grid;
  const int nb = a.B < 4 ? a.B : 4;
  *lists_keys = (size_t)grid * stream_impl::kStreamWarps * nb * cap_for_k(a.k);
  if ((size_t)a.ld * 4 * 4 + stream_impl::kCtaMergeCap * 8 > 200 * 1024) {
    set_error("stream path: dim %d too large for the shared-memory query tile", a.dim);
    return TS_ERR_UNSUPPORTED;
  }
  return TS_OK;
}

int launch_s1_stream(const ScanArgs& a, cudaStream_t st, int* launches) {
  switch (a.dtype) {
    case TS_BF16: return launch_s1_stream_bf16(a, st, launches);
    case TS_F16: return launch_s1_stream_f16(a, st, launches);
    case TS_F32: return launch_s1_stream_f32(a, st, launches);
  }
  set_error("stream path: bad dtype %d", a.dtype);
  return TS_ERR_INVALID;
}

}  // namespace ts
