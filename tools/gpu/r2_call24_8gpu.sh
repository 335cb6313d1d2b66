#!/bin/bash
# Round 2, call 24 (8-GPU box, ~3 min): N = 1 and N = 8 back to back on ONE box, the way the driver measures scaling.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 python bench.py --gpus 1 --steps 50 --warmup 5 --no-cpu > gpurun_out/s_n1.json 2> gpurun_out/s_n1.err; echo "n1 rc=$?"
timeout 400 $TR --master-port 29617 bench.py --gpus 8 --steps 50 --warmup 5 > gpurun_out/s_n8.json 2> gpurun_out/s_n8.err; echo "n8 rc=$?"
TS_P2P=0 timeout 200 $TR --master-port 29614 bench.py --gpus 8 --steps 50 --warmup 5 --no-extra --no-cpu --no-parity > gpurun_out/s_n8_nccl.json 2> gpurun_out/s_n8_nccl.err; echo "n8 nccl rc=$?"
python - <<'PY'
import json
v={}
for n,f in ((1,'gpurun_out/s_n1.json'),(8,'gpurun_out/s_n8.json'),('8nccl','gpurun_out/s_n8_nccl.json')):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1]); ro=r['roofline']; v[n]=r['value']
        print(f"N={n}: value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} scan={ro['kernel_ms']} frac={ro['frac']} exch={ro['exchange']} clocks={r['clocks']}")
        if 'also' in ro: print('   also:', json.dumps(ro['also']))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-300:])
if 1 in v and 8 in v: print('efficiency N=8:', v[8]/(8*v[1]), ' nccl:', v.get('8nccl',0)/(8*v[1]))
PY
