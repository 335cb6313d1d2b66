#!/bin/bash
# Round 2, call 10 (1 GPU, ~4 min): the fused scan's start-up after the trace of call 9 -- append-from-registers slow path,
# J = 1 pass 1 without a slow path, 32-wide min over the slices, release/acquire grid barrier zeroed by the query-prep
# kernel: parity suites, timeline, step probe (compare with call 9: 463 us scan at 1.25 M rows x B = 32).
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1 tests/test_gpu_stage1.py
for cfg in "1250000 1" "1250000 32" "1250000 128"; do
  set -- $cfg
  TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows $1 --steps 3 --batches $2 --variants TS_FUSE=1 2> gpurun_out/trace_$1_b$2.err | tail -1 | cut -c1-120
  grep "ts trace" gpurun_out/trace_$1_b$2.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace2_$1_b$2.json
  rm -f gpurun_out/trace_$1_b$2.err
done
timeout 300 python tools/step_probe.py --rows 1250000 --batches 1,32,128 --variants TS_FUSE=1,TS_FUSE=0,TS_FUSE=1,TS_FUSE=0 | tee gpurun_out/step_probe.jsonl
timeout 300 python tools/step_probe.py --rows 10000000 --steps 30 --batches 1,32,128 --variants TS_FUSE=1,TS_FUSE=0 | tee gpurun_out/step_probe10.jsonl
TS_DBG_STATS=1 timeout 120 python tools/step_probe.py --rows 1250000 --steps 2 --batches 32 --variants TS_FUSE=1 2>&1 | grep "ts stats" | tail -1
run zfull tests/test_gpu_zzz_fullsize.py
run rest tests/test_gpu_pipeline.py tests/test_gpu_zz_tf32.py tests/test_gpu_zz_ivf.py tests/test_gpu_z_hybrid.py
