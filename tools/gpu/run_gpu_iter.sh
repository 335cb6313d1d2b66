#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1_rest tests/test_gpu_stage1.py -k "not stream_path and not umma_path"
python bench.py --steps 50 --warmup 5 --no-cpu --no-extra > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.err | tail -3
python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --batch 1 > gpurun_out/bench_n1_b1.json 2>> gpurun_out/bench_n1.err
python bench.py --steps 50 --warmup 5 --no-cpu --no-extra --rows 1000000 --dim 768 > gpurun_out/bench_1M_b32.json 2>> gpurun_out/bench_n1.err
python - <<'PY'
import json
for f in ("bench_n1","bench_n1_b1","bench_1M_b32"):
    d=json.load(open(f"gpurun_out/{f}.json"))
    print(f, "q/s", round(d['value']), "ms/step", round(d['ms_per_step'],4), "eager", round(d['config']['ms_per_step_eager'],4), d['config']['launch'], "scan_ms", round(d['roofline']['kernel_ms'],4), "frac", round(d['roofline']['frac'],3), "e2e", round(d['e2e']['value']))
PY
