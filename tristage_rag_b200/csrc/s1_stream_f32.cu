// S1-stream instantiation for float storage (see s1_stream_impl.cuh).
#include "s1_stream_impl.cuh"
namespace ts {
int launch_s1_stream_f32(const ScanArgs& a, cudaStream_t st, int* launches) {
  return stream_impl::launch_t<float>(a, st, launches);
}
}  // namespace ts
