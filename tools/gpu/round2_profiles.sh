#!/bin/bash
# ncu evidence for round 2 (1 GPU, one gpurun call; run AFTER round2_first.sh is green).  Every command is run
# plain first and profiled only if that run exited 0 (`&&`, no pipe), as the profiling recipe requires.
# Read the reports on the CPU box:  python tools/ncu_summary.py gpurun_out/prof_X.ncu-rep profiles/r02_X
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
# 1. launch list of the bench command (kernel SHARES of the step)
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-variants"
timeout 600 $CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
# 2. dominant kernel, headline config (10M x 1024, B = 32): pre-pass + scan launches
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra"
timeout 600 $CMD > gpurun_out/plain_s1.log 2>&1 && \
timeout 900 $NCU -k regex:s1_umma -s 4 -c 2 -o gpurun_out/prof_s1_b32 $CMD > gpurun_out/ncu_s1_b32.log 2>&1
echo "s1 B=32 rc=$?"
# 3. tensor-bound regime (B = 1024): is L2->SM traffic the bound?  (lts__t_bytes, l1tex__m_xbar2l1tex_read_bytes)
CMD="python tools/perf_probe.py --paths umma --rows 4000000 --dim 1024 --batches 1024 --steps 2"
timeout 600 $CMD > gpurun_out/plain_s1_b1024.log 2>&1 && \
timeout 900 $NCU -k regex:s1_umma -s 3 -c 1 -o gpurun_out/prof_s1_b1024 $CMD > gpurun_out/ncu_s1_b1024.log 2>&1
echo "s1 B=1024 rc=$?"
# 4. Stage 2, config #4, both epilogues
for v in 0 1; do
  CMD="python tools/s2_probe.py --steps 2"
  TS_S2_V2=$v timeout 600 $CMD > gpurun_out/plain_s2_v$v.log 2>&1 && \
  TS_S2_V2=$v timeout 900 $NCU -k regex:maxsim_umma -s 2 -c 1 -o gpurun_out/prof_s2_v$v $CMD > gpurun_out/ncu_s2_v$v.log 2>&1
  echo "s2 v2=$v rc=$?"
done
# 5. approximate mode: the list scan at the reference's batch 1 (gathered rows; is it HBM-bound?)
CMD="python tools/ivf_probe.py --rows 10000000 --dim 1024 --batches 1 --steps 2"
timeout 600 $CMD > gpurun_out/plain_ivf.log 2>&1 && \
timeout 900 $NCU -k regex:ivf_scan -s 2 -c 1 -o gpurun_out/prof_ivf_scan $CMD > gpurun_out/ncu_ivf.log 2>&1
echo "ivf rc=$?"
ls -la gpurun_out/*.ncu-rep
