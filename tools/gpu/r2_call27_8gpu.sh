#!/bin/bash
# Round 2, call 27 (8 GPUs, ~2 min): dist_check after the fence changes; GPU time line of the 8-GPU step; bench line.
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29612 tools/dist_check.py > gpurun_out/dist_default_$N.log 2>&1; echo "dist_check rc=$? $(grep 'dist_check ok' gpurun_out/dist_default_$N.log)"
timeout 200 $TR --master-port 29613 tools/step_timeline.py --rows 10000000 > gpurun_out/timeline_n8.jsonl 2> gpurun_out/timeline_n8.err; echo "timeline rc=$?"; grep '"world"' gpurun_out/timeline_n8.jsonl | cut -c1-900
TS_P2P=0 timeout 200 $TR --master-port 29614 tools/step_timeline.py --rows 10000000 > gpurun_out/timeline_n8_nccl.jsonl 2>> gpurun_out/timeline_n8.err; grep '"world"' gpurun_out/timeline_n8_nccl.jsonl | cut -c1-900
timeout 400 $TR --master-port 29617 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/b${N}_final.json 2> gpurun_out/b${N}_final.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/b8_final.json').read().strip().splitlines()[-1]); ro=r['roofline']
print(f"N=8: value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} scan={ro['kernel_ms']} exch={ro['exchange']}")
print('   also:', json.dumps(ro['also'])); print('   parity:', json.dumps(r['parity']))
PY
