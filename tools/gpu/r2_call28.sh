#!/bin/bash
# Round 2, call 28 (1 GPU, ~3 min): full GPU suite, smoke and the bench line at HEAD.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_suite.log 2>&1; echo "suite rc=$? $(tail -1 gpurun_out/gpu_suite.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-1500
