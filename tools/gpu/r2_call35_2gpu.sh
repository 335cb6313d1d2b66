#!/bin/bash
# Round 2, call 35 (2 GPUs, <1 min): dist_check incl. the sharded drop-in classes on hardware.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29611 tools/dist_check.py > gpurun_out/dist_default.log 2>&1; echo "dist_check rc=$? $(grep 'dist_check ok' gpurun_out/dist_default.log)"; tail -4 gpurun_out/dist_default.log | cut -c1-300
