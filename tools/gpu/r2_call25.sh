#!/bin/bash
# Round 2, call 25 (1 GPU, ~3 min): full GPU suite at HEAD incl. the single-GPU tests of the exchange kernels; smoke.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_suite.log 2>&1; echo "suite rc=$? $(tail -1 gpurun_out/gpu_suite.log)"; grep -E "FAILED|Error" gpurun_out/gpu_suite.log | head -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
