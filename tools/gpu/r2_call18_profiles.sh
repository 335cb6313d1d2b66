#!/bin/bash
# Round 2, call 18 (1 GPU, ~8 min): ncu evidence for HEAD.  Every command runs plain first and is profiled only if that run
# exited 0.  Reports are read on the CPU box with tools/ncu_summary.py and summarised under profiles/.
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# 1. launch list of the bench step (kernel shares), own kernels only
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-parity"
timeout 300 $CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"s1_umma|select_kernel|convert_rows" -c 60 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# 2. dominant kernel at the headline config (10 M x 1024, B = 32), fused launch
timeout 900 $NCU -k regex:s1_umma -s 4 -c 1 -o gpurun_out/prof_h_s1_b32 $CMD > gpurun_out/ncu_h_s1_b32.log 2>&1; echo "s1 b32 rc=$?"
# 3. the select kernel of the same step
timeout 900 $NCU -k regex:select_kernel -s 4 -c 1 -o gpurun_out/prof_h_select $CMD > gpurun_out/ncu_h_select.log 2>&1; echo "select rc=$?"
# 4. per-rank shard of the 8-GPU job (1.25 M rows, B = 32)
CMD="python tools/step_probe.py --rows 1250000 --steps 3 --batches 32 --variants TS_FUSE=1"
timeout 300 $CMD > gpurun_out/plain_step.log 2>&1 && \
timeout 900 $NCU -k regex:s1_umma -s 4 -c 1 -o gpurun_out/prof_h_s1_b32_shard $CMD > gpurun_out/ncu_h_s1_shard.log 2>&1; echo "s1 shard rc=$?"
# 5. tensor-bound regime: CTA-pair kernel at B = 1024 (4 M x 1024)
CMD="python tools/perf_probe.py --paths umma --rows 4000000 --dim 1024 --batches 1024 --steps 2"
timeout 600 $CMD > gpurun_out/plain_s1_b1024.log 2>&1 && \
timeout 900 $NCU -k regex:s1_pair -s 3 -c 1 -o gpurun_out/prof_h_s1_b1024_pair $CMD > gpurun_out/ncu_h_s1_b1024.log 2>&1; echo "pair rc=$?"
# 6. Stage 2, config #4 (flow kernel, eight epilogue warps)
CMD="python tools/s2_probe.py --steps 2"
timeout 600 $CMD > gpurun_out/plain_s2.log 2>&1 && \
timeout 900 $NCU -k regex:maxsim -s 2 -c 1 -o gpurun_out/prof_h_s2_flow $CMD > gpurun_out/ncu_h_s2.log 2>&1; echo "s2 rc=$?"
# 7. approximate mode list scan at batch 1
CMD="python tools/ivf_probe.py --rows 10000000 --dim 1024 --batches 1 --steps 2"
timeout 600 $CMD > gpurun_out/plain_ivf.log 2>&1 && \
timeout 900 $NCU -k regex:ivf_scan -s 2 -c 1 -o gpurun_out/prof_h_ivf_scan $CMD > gpurun_out/ncu_h_ivf.log 2>&1; echo "ivf rc=$?"
# 8. side numbers for the README: Stage-2 shape sweep and the fp32 / C2 probes
P="timeout 300 python tools/s2_probe.py"
O=gpurun_out/s2_probe_head.jsonl; : > $O
for cfg in "--ndocs 1000000" "--Lq 128" "--lo 16 --hi 40" "--lo 180 --hi 180" "--dim 64" "--B 8 --C 500"; do $P $cfg --tag "head $cfg" >> $O 2>> gpurun_out/s2_probe_head.err; done
cat $O | cut -c1-260
timeout 300 python tools/perf_probe.py --paths umma --rows 1000000 --dim 768 --batches 1,32,1024 --steps 10 --tag "C2 bf16" > gpurun_out/c2_head.jsonl 2> gpurun_out/c2_head.err
timeout 300 python tools/perf_probe.py --paths stream,umma --rows 1000000 --dim 768 --batches 1,32,1024 --steps 5 --dtype fp32 --tag "C2 fp32" >> gpurun_out/c2_head.jsonl 2>> gpurun_out/c2_head.err
cat gpurun_out/c2_head.jsonl | cut -c1-260
timeout 300 python tools/ivf_probe.py --rows 10000000 --dim 1024 --batches 1,32 --steps 10 > gpurun_out/ivf_head.jsonl 2> gpurun_out/ivf_head.err; cat gpurun_out/ivf_head.jsonl | cut -c1-300
ls -la gpurun_out/prof_h_*.ncu-rep
