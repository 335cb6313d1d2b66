#!/bin/bash
# Runs the GPU test groups in separate processes (a trapped kernel poisons its
# CUDA context) and keeps each log under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout 1200 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/$name.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1_stream tests/test_gpu_stage1.py -k "stream_path"
run s1_umma tests/test_gpu_stage1.py -k "umma_path"
run s1_rest tests/test_gpu_stage1.py -k "not stream_path and not umma_path"
run s2 tests/test_gpu_stage2.py
run pipe tests/test_gpu_pipeline.py
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
