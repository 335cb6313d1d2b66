// TEST INFRASTRUCTURE ONLY (tests/cudasim): C entry points that let the tests drive internal
// launchers which the public ABI only reaches through the (non-emulable) tensor-path scan.
#include "ts_internal.h"

extern "C" int cudasim_merge_lists(const uint64_t* lists, const int* counts, const float* pub, int n_slices, int n_mt,
                                   int cap, int spread, int bpad, int jrank, int B, int k, int64_t id_base,
                                   float* out_scores, int64_t* out_ids) {
  ts::UmmaLayout lay{};
  lay.n_slices = n_slices; lay.n_mt = n_mt; lay.grid = n_mt * n_slices; lay.cap = cap; lay.spread = spread;
  lay.bpad = bpad; lay.jrank = jrank; lay.rows_per_cta = 128; lay.fused = 0;
  int launches = 0;
  return ts::launch_merge_lists(lists, counts, pub, lay, B, k, id_base, out_scores, out_ids, nullptr, &launches);
}
