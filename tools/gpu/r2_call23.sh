#!/bin/bash
# Round 2, call 23 (1 GPU, ~5 min): programmatic dependent launch off by default -- clean A/B (no events between the kernels),
# full GPU suite, smoke, bench line.
mkdir -p gpurun_out
timeout 300 python tools/step_probe.py --rows 1250000 --batches 1,32,128 --variants TS_PDL=0,TS_PDL=1,TS_PDL=0,TS_PDL=1 | tee gpurun_out/step_probe_pdl.jsonl | cut -c1-250
timeout 300 python tools/step_probe.py --rows 625000 --dim 768 --k 500 --batches 64 --steps 100 --variants TS_PDL=0,TS_PDL=1 | cut -c1-250
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_suite.log 2>&1; echo "suite rc=$? $(tail -1 gpurun_out/gpu_suite.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json | cut -c1-2400
