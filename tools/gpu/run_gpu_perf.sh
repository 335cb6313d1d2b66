#!/bin/bash
mkdir -p gpurun_out
tools/gpu/run_gpu_tests.sh
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
CMD="python bench.py --steps 2 --warmup 1 --no-cpu --no-variants"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launchlist rc=$?"
CMD2="python bench.py --steps 2 --warmup 1 --no-cpu --no-extra"
$CMD2 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:s1_umma -s 2 -c 2 -o gpurun_out/prof_s1_umma $CMD2 > gpurun_out/ncu_umma.log 2>&1
echo "ncu umma rc=$?"
