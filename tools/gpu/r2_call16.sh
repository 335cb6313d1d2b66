#!/bin/bash
# Round 2, call 16 (1 GPU, ~6 min): full GPU suite + bench after the barrier fix (barrier.sync + warp-uniform polling),
# k-th-of-slices start, dynamic tiles, programmatic dependent launch; launch list of a bench step.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_suite.log 2>&1; echo "suite rc=$? $(tail -1 gpurun_out/gpu_suite.log)"
timeout 300 python tools/step_probe.py --rows 1250000 --batches 1,32,128 --variants TS_FUSE=1,TS_DBG_NOKTHSTART=1,TS_FUSE=1,TS_DBG_NOKTHSTART=1 | tee gpurun_out/step_probe.jsonl
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-2600
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-cpu --no-parity"
timeout 300 $CMD > gpurun_out/plain_bench.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
