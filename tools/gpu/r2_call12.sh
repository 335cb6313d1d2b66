#!/bin/bash
# Round 2, call 12 (1 GPU, ~5 min): dynamic tile schedule of the scan (A/B against round-robin tiles, TS_DBG_STATIC) and the
# cost of the cooperative launch (TS_DBG_NOCOOP): parity suites, step probes on the 8-GPU per-rank shard and on the full
# corpus, timeline, then the bench line.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1 tests/test_gpu_stage1.py
V="TS_FUSE=1,TS_DBG_STATIC=1,TS_DBG_NOCOOP=1,TS_FUSE=1,TS_DBG_STATIC=1,TS_DBG_NOCOOP=1"
timeout 300 python tools/step_probe.py --rows 1250000 --batches 1,32,128 --variants $V | tee gpurun_out/step_probe.jsonl
timeout 300 python tools/step_probe.py --rows 10000000 --steps 30 --batches 1,32,128 --variants TS_FUSE=1,TS_DBG_STATIC=1,TS_FUSE=1,TS_DBG_STATIC=1 | tee gpurun_out/step_probe10.jsonl
for cfg in "1250000 32" "10000000 32"; do
  set -- $cfg
  TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows $1 --steps 3 --batches $2 --variants TS_FUSE=1 2> gpurun_out/trace_$1_b$2.err | tail -1 | cut -c1-120
  grep "ts trace" gpurun_out/trace_$1_b$2.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace3_$1_b$2.json
  rm -f gpurun_out/trace_$1_b$2.err
done
run zfull tests/test_gpu_zzz_fullsize.py
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-3000
