// TEST INFRASTRUCTURE ONLY (tests/cudasim): the tensor-path launchers (TMA / tcgen05) cannot be
// emulated; they refuse, so a test that reaches them fails loudly instead of computing anything.
#include "ts_internal.h"

extern "C" {
long long hostsim_live_allocs = 0;
long long hostsim_live_pinned = 0;
long long hostsim_fail_malloc_over = 0;
int hostsim_is_simulation(void) { return 2; }   // 2 = SIMT emulation build
}

namespace ts {
static int not_emulated(const char* what) {
  set_error("%s: the tcgen05/TMA kernels only exist for sm_100a and cannot be emulated (cudasim is test infrastructure)", what);
  return TS_ERR_UNSUPPORTED;
}
int s1_umma_plan(const ScanArgs&, UmmaLayout*) { return not_emulated("s1_umma"); }
int launch_s1_umma(const ScanArgs&, const UmmaLayout&, cudaStream_t, int*) { return not_emulated("s1_umma"); }
}  // namespace ts
