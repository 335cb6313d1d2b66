#!/bin/bash
# Round 2, call 33 (1 GPU, <1 min): the single-GPU exchange tests and Stage-1 parity after the rank merge; smoke.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_z_exchange.py tests/test_gpu_stage1.py tests/test_gpu_pipeline.py -q -m gpu -p no:cacheprovider -x > gpurun_out/gpu_last.log 2>&1; echo "rc=$? $(tail -1 gpurun_out/gpu_last.log)"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
