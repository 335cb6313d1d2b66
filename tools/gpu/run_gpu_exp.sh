#!/bin/bash
mkdir -p gpurun_out
P="python tools/perf_probe.py --paths umma"
$P --batches 1,8,32,128,1024 --tag base > gpurun_out/exp_10M.jsonl 2> gpurun_out/exp.err
TS_DBG_NOTOPK=1 $P --batches 1,8,32,128,1024 --tag notopk >> gpurun_out/exp_10M.jsonl 2>> gpurun_out/exp.err
TS_DBG_NOSPREAD=1 $P --batches 8,32 --tag nospread >> gpurun_out/exp_10M.jsonl 2>> gpurun_out/exp.err
TS_DBG_NOSPREAD=1 TS_DBG_NOTOPK=1 $P --batches 8,32 --tag nospread_notopk >> gpurun_out/exp_10M.jsonl 2>> gpurun_out/exp.err
$P --rows 1000000 --dim 768 --batches 32,128,1024 --tag base > gpurun_out/exp_1M.jsonl 2>> gpurun_out/exp.err
TS_DBG_NOTOPK=1 $P --rows 1000000 --dim 768 --batches 32,128,1024 --tag notopk >> gpurun_out/exp_1M.jsonl 2>> gpurun_out/exp.err
echo done
