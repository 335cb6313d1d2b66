#!/bin/bash
# Round 2, call 7 (1 GPU, ~3 min): k-th-of-slices rule kept fresh by a dedicated bound warp -- parity + step probe A/B.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1 tests/test_gpu_stage1.py
timeout 300 python tools/step_probe.py --rows 1250000 --variants TS_FUSE=1,TS_DBG_NOKTH=1,TS_FUSE=1,TS_DBG_NOKTH=1 > gpurun_out/step_probe.jsonl 2> gpurun_out/step_probe.err; echo "rc=$?"; cat gpurun_out/step_probe.jsonl; tail -3 gpurun_out/step_probe.err
timeout 300 python tools/step_probe.py --rows 10000000 --steps 30 --batches 1,32,128 --variants TS_FUSE=1,TS_DBG_NOKTH=1,TS_FUSE=1,TS_DBG_NOKTH=1 > gpurun_out/step_probe10.jsonl 2>> gpurun_out/step_probe.err; cat gpurun_out/step_probe10.jsonl
TS_DBG_STATS=1 timeout 120 python tools/step_probe.py --rows 1250000 --steps 2 --batches 32 --variants TS_FUSE=1 2>&1 | grep "ts stats" | tail -2
TS_DBG_STATS=1 timeout 120 python tools/step_probe.py --rows 1250000 --steps 2 --batches 32 --variants TS_DBG_NOKTH=1 2>&1 | grep "ts stats" | tail -2
run zfull tests/test_gpu_zzz_fullsize.py
