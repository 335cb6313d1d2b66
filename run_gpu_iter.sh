#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s2 tests/test_gpu_stage2.py
run s1_umma tests/test_gpu_stage1.py -k "umma_path"
run s1_rest tests/test_gpu_stage1.py -k "not stream_path and not umma_path"
run pipe tests/test_gpu_pipeline.py
python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
P="python tools/perf_probe.py --paths umma"
$P --rows 1000000 --dim 768 --batches 1,32,128,1024 --tag v5 > gpurun_out/exp_1M.jsonl 2>> gpurun_out/exp.err
echo done
