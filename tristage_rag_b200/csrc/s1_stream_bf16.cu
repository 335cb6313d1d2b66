// S1-stream instantiation for __nv_bfloat16 storage (see s1_stream_impl.cuh).
#include "s1_stream_impl.cuh"
namespace ts {
int launch_s1_stream_bf16(const ScanArgs& a, cudaStream_t st, int* launches) {
  return stream_impl::launch_t<__nv_bfloat16>(a, st, launches);
}
}  // namespace ts
