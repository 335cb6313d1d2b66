#!/bin/bash
# Round 2, call 14 (1 GPU, ~5 min): k-th-of-slices first bound + distributed refresh in the fused scan, programmatic
# dependent launch of the select kernel (16 KB of shared memory: resident beside the scan), unrolled query prep:
# parity suites, interleaved A/B on the 8-GPU per-rank shard and on the full corpus, timeline.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1 tests/test_gpu_stage1.py
V="TS_FUSE=1,TS_PDL=0,TS_DBG_NOKTHSTART=1,TS_FUSE=1,TS_PDL=0,TS_DBG_NOKTHSTART=1"
timeout 300 python tools/step_probe.py --rows 1250000 --batches 1,32,128 --variants $V | tee gpurun_out/step_probe.jsonl
timeout 300 python tools/step_probe.py --rows 10000000 --steps 30 --batches 32 --variants TS_FUSE=1,TS_DBG_NOKTHSTART=1,TS_FUSE=1,TS_DBG_NOKTHSTART=1 | tee gpurun_out/step_probe10.jsonl
TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows 1250000 --steps 3 --batches 32 --variants TS_FUSE=1 2> gpurun_out/trace.err | tail -1 | cut -c1-100
grep "ts trace\]" gpurun_out/trace.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace5_1250000_b32.json; rm -f gpurun_out/trace.err
TS_DBG_STATS=1 timeout 120 python tools/step_probe.py --rows 1250000 --steps 2 --batches 32 --variants TS_FUSE=1 2>&1 | grep "ts stats" | tail -1
TS_DBG_STATS=1 timeout 120 python tools/step_probe.py --rows 1250000 --steps 2 --batches 32 --variants TS_DBG_NOKTHSTART=1 2>&1 | grep "ts stats" | tail -1
run zfull tests/test_gpu_zzz_fullsize.py
run rest tests/test_gpu_pipeline.py tests/test_gpu_zz_tf32.py tests/test_gpu_zz_ivf.py tests/test_gpu_z_hybrid.py tests/test_gpu_stage2.py tests/test_gpu_z_shards.py
