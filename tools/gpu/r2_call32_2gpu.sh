#!/bin/bash
# Round 2, call 32 (2 GPUs, ~1 min): rank merge in the exchange -- dist_check (one kernel = two kernels = oracle, NCCL plane), time line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tools/dist_check.py > gpurun_out/dist_default.log 2>&1; echo "dist_check default rc=$? $(grep 'dist_check ok' gpurun_out/dist_default.log)"; tail -2 gpurun_out/dist_default.log | cut -c1-200
timeout 200 $TR --master-port 29613 tools/step_timeline.py --rows 2500000 2>&1 | grep '"world"' | tee gpurun_out/timeline_n2_rankmerge.json | cut -c1-600
timeout 200 $TR --master-port 29614 bench.py --gpus 2 --steps 50 --warmup 5 --rows 2500000 --no-extra --no-cpu 2>/dev/null | python -c "
import sys, json
r=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench small: ms', r['ms_per_step'], 'parity', r.get('parity'))"
