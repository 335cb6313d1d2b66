#!/bin/bash
# Round 2, call 29 (1 GPU, ~1.5 min): why is B = 32 3 - 8 % slower than B = 1 on the same 128-row MMA tiles?  Batch sweep with the
# queries spread over the four TMEM lane quarters (8-row TMA boxes, four epilogue warps) and packed into one (one box, one warp).
mkdir -p gpurun_out
timeout 400 python tools/step_probe.py --rows 10000000 --steps 30 --batches 1,8,16,32,64 --variants TS_FUSE=1,TS_DBG_NOSPREAD=1,TS_FUSE=1,TS_DBG_NOSPREAD=1 | tee gpurun_out/step_probe_spread.jsonl | cut -c1-230
