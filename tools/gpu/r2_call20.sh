#!/bin/bash
# Round 2, call 20 (1 GPU, ~3 min): same-box A/B of the k = 500 path (C5: Stage 1 k = 500, d = 768 -> Stage 2 -> top-100) between
# commit 29916a4 (worktree under build/) and HEAD: the bench line showed 0.458 -> 0.543 ms per 64-query step at N = 2.
mkdir -p gpurun_out
OLD=build/old_29916a4
for rep in 1 2; do
for tree in $OLD .; do
  echo "== $tree C5"; timeout 300 python $tree/tools/e2e_c5.py --docs 625000 --steps 30 2>/dev/null | tail -1 | cut -c1-260
  echo "== $tree stage1 k=500"; timeout 300 python $tree/tools/step_probe.py --rows 625000 --dim 768 --k 500 --batches 64 --steps 100 --variants TS_FUSE=1 2>/dev/null | tail -1 | cut -c1-260
done; done
echo "== HEAD variants"; timeout 300 python tools/step_probe.py --rows 625000 --dim 768 --k 500 --batches 64 --steps 100 --variants TS_FUSE=1,TS_DBG_STATIC=1,TS_PDL=0,TS_FUSE=0,TS_FUSE=1 2>/dev/null | cut -c1-200
TS_DBG_STATS=1 timeout 120 python tools/step_probe.py --rows 625000 --dim 768 --k 500 --batches 64 --steps 2 --variants TS_FUSE=1 2>&1 | grep "ts stats" | tail -1
TS_DBG_STATS=1 timeout 120 python $OLD/tools/step_probe.py --rows 625000 --dim 768 --k 500 --batches 64 --steps 2 --variants TS_FUSE=1 2>&1 | grep "ts stats" | tail -1
