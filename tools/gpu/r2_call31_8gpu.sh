#!/bin/bash
# Round 2, call 31 (8 GPUs, ~1.5 min): the bench line at N = 8 with the compact small-batch layout.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 400 $TR --master-port 29617 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/b8_final2.json 2> gpurun_out/b8_final2.err; echo "bench rc=$?"
python - <<'PY'
import json
r=json.loads(open('gpurun_out/b8_final2.json').read().strip().splitlines()[-1]); ro=r['roofline']
print(f"N=8: value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} ({r['e2e']['ms_per_step']}) scan={ro['kernel_ms']} frac={ro['frac']} exch={ro['exchange']} clocks={r['clocks']}")
print('   also:', json.dumps(ro['also'])); print('   parity:', json.dumps(r['parity']))
PY
tail -n 2 gpurun_out/b8_final2.err
