// TEST INFRASTRUCTURE ONLY (tests/cudasim): bookkeeping symbols shared with tests/hostsim.
#include "ts_internal.h"

extern "C" {
long long hostsim_live_allocs = 0;
long long hostsim_live_pinned = 0;
long long hostsim_fail_malloc_over = 0;
int hostsim_is_simulation(void) { return 2; }   // 2 = SIMT emulation build
}

