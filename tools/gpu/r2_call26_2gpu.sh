#!/bin/bash
# Round 2, call 26 (2 GPUs, ~1.5 min): step time line tool -- one GPU (1.25 M-row shard, 10 M rows) and two GPUs.
mkdir -p gpurun_out
timeout 200 python tools/step_timeline.py --rows 1250000 2>&1 | tail -1 | tee gpurun_out/timeline_n1_1p25M.json | cut -c1-700
timeout 200 python tools/step_timeline.py --rows 10000000 2>&1 | tail -1 | tee gpurun_out/timeline_n1_10M.json | cut -c1-700
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tools/step_timeline.py --rows 2500000 2>&1 | grep '"world"' | tee gpurun_out/timeline_n2.json | cut -c1-900
