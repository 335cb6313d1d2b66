#!/bin/bash
# Round 2, call 34 (1 GPU, ~2 min): ncu of the scan kernel at HEAD (compact small-batch layout) + launch list of the bench step.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-parity"
timeout 300 $CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"s1_umma|select_kernel|convert_rows" -c 30 --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 400 $NCU -k regex:s1_umma -s 4 -c 1 -o gpurun_out/prof_final_s1_b32 $CMD > gpurun_out/ncu_final_s1_b32.log 2>&1; echo "s1 b32 rc=$?"
ls -la gpurun_out/prof_final_s1_b32.ncu-rep
