#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1_umma tests/test_gpu_stage1.py -k "umma_path"
run s1_rest tests/test_gpu_stage1.py -k "not stream_path and not umma_path"
P="python tools/perf_probe.py --paths umma"
$P --batches 1,8,32,128,256,1024 --tag v4 > gpurun_out/exp_10M.jsonl 2> gpurun_out/exp.err
TS_DBG_NODUAL=1 $P --batches 256,1024 --tag v4_nodual >> gpurun_out/exp_10M.jsonl 2>> gpurun_out/exp.err
$P --rows 1000000 --dim 768 --batches 1,32,128,1024 --tag v4 > gpurun_out/exp_1M.jsonl 2>> gpurun_out/exp.err
TS_DBG_STATS=1 $P --steps 1 --batches 32,1024 --tag st > /dev/null 2> gpurun_out/st.err
grep "ts stats" gpurun_out/st.err | sort | uniq -c | tail -4
echo done
