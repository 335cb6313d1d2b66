#!/bin/bash
# Round 2, call 9 (1 GPU, ~2 min): per-CTA timeline of the fused scan (TS_DBG_TRACE) on the per-rank shard of the 8-GPU job
# and on the full corpus: where the fixed ~30-50 us of a scan go (start-up, grid barrier, warm-up of the bound, tail skew).
mkdir -p gpurun_out
for cfg in "1250000 1" "1250000 32" "10000000 32"; do
  set -- $cfg
  TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows $1 --steps 3 --batches $2 --variants TS_FUSE=1 2> gpurun_out/trace_$1_b$2.err | tail -1
  grep "ts trace" gpurun_out/trace_$1_b$2.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace_$1_b$2.json
  rm -f gpurun_out/trace_$1_b$2.err
done
TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows 1250000 --steps 3 --batches 32 --variants TS_FUSE=0 2> gpurun_out/trace_nofuse.err | tail -1
grep "ts trace" gpurun_out/trace_nofuse.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace_1250000_b32_nofuse.json; rm -f gpurun_out/trace_nofuse.err
timeout 300 python tools/step_probe.py --rows 1250000 --variants TS_FUSE=1,TS_FUSE=0,TS_FUSE=1,TS_FUSE=0 | tee gpurun_out/step_probe.jsonl
ls -la gpurun_out/trace_*
