// Kernel-launch and dynamic-shared-memory spelling.  The product build (nvcc) expands these to the
// ordinary CUDA syntax.  tests/cudasim/ defines TS_CUDASIM and compiles the SAME kernel sources
// with g++ against a fiber-based SIMT emulator, so the CUDA-core kernels (ingest, stream scan,
// selection, SIMT MaxSim) can be executed and checked on a machine without a GPU; that build is
// test infrastructure and is never loaded by the package.
#pragma once
#ifdef TS_CUDASIM
#define TS_LAUNCH(kern, grid, block, smem, stream, ...) \
  cudasim::launch((grid), (block), (size_t)(smem), [&]() { kern(__VA_ARGS__); })
#define TS_LAUNCH_PDL(kern, grid, block, smem, stream, ...) TS_LAUNCH(kern, grid, block, smem, stream, __VA_ARGS__)
#define TS_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(cudasim::dyn_smem())
#define TS_SPIN_YIELD() cudasim::yield_spin()
#define ts_nanosleep(ns) ((void)0)
#else
#define TS_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
// Programmatic dependent launch: the kernel may be scheduled while its predecessor in the stream is still running
// (once every CTA of the predecessor has executed griddepcontrol.launch_dependents or exited); it must execute
// griddepcontrol.wait (ts::ptx::grid_dep_wait) before it touches anything the predecessor writes.
#define TS_LAUNCH_PDL(kern, grid_, block_, smem_, stream_, ...)                                          \
  do {                                                                                                   \
    cudaLaunchConfig_t ts_cfg_ = {};                                                                     \
    ts_cfg_.gridDim = dim3(grid_); ts_cfg_.blockDim = dim3(block_);                                      \
    ts_cfg_.dynamicSmemBytes = (size_t)(smem_); ts_cfg_.stream = (stream_);                              \
    cudaLaunchAttribute ts_attr_[1];                                                                     \
    ts_attr_[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;                                 \
    ts_attr_[0].val.programmaticStreamSerializationAllowed = 1;                                          \
    ts_cfg_.attrs = ts_attr_; ts_cfg_.numAttrs = 1;                                                      \
    cudaLaunchKernelEx(&ts_cfg_, kern, __VA_ARGS__);                                                     \
  } while (0)
#define TS_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define TS_SPIN_YIELD() ((void)0)   /* a spinning thread of the emulator lets the others run */
#define ts_nanosleep(ns) __nanosleep(ns)
#endif
