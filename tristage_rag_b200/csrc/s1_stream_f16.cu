// S1-stream instantiation for __half storage (see s1_stream_impl.cuh).
#include "s1_stream_impl.cuh"
namespace ts {
int launch_s1_stream_f16(const ScanArgs& a, cudaStream_t st, int* launches) {
  return stream_impl::launch_t<__half>(a, st, launches);
}
}  // namespace ts
