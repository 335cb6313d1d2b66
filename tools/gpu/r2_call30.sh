#!/bin/bash
# Round 2, call 30 (1 GPU, ~3 min): compact small-batch layout (one right-sized TMA box in lane quarters 2-3): Stage-1 parity,
# full-size tests, batch sweep, bench line.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1 tests/test_gpu_stage1.py tests/test_gpu_zz_tf32.py tests/test_gpu_z_exchange.py tests/test_gpu_pipeline.py
timeout 400 python tools/step_probe.py --rows 10000000 --steps 30 --batches 1,8,16,20,32,64 --variants TS_FUSE=1,TS_FUSE=1 | tee gpurun_out/step_probe_compact.jsonl | python -c "
import sys, json
for l in sys.stdin:
    r=json.loads(l); print(r['B'], 'scan', r['scan_us'], 'eager', r['eager_us'])
"
timeout 300 python tools/step_probe.py --rows 1250000 --batches 32,64 --variants TS_FUSE=1,TS_FUSE=1 | cut -c1-200
run zfull tests/test_gpu_zzz_fullsize.py
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-1500
