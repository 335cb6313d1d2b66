#!/bin/bash
# Round 2, call 22 (8 GPUs, ~3 min): defaults at N = 8 -- dist_check (one-kernel Stage-1 exchange, Stage-2 scatter), the full bench
# line, and the Stage-1 step with NCCL for comparison.
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29612 tools/dist_check.py > gpurun_out/dist_default_$N.log 2>&1; echo "dist_check default rc=$? $(grep 'dist_check ok' gpurun_out/dist_default_$N.log)"; tail -2 gpurun_out/dist_default_$N.log | cut -c1-300
timeout 500 $TR --master-port 29617 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/b${N}_default_full.json 2> gpurun_out/b${N}_default_full.err; echo "full default rc=$?"
B="bench.py --gpus $N --steps 100 --warmup 5 --no-extra --no-cpu --no-parity"
timeout 200 $TR --master-port 29613 $B > gpurun_out/b${N}_default.json 2> gpurun_out/b${N}_default.err; echo "default rc=$?"
TS_P2P=0 timeout 200 $TR --master-port 29614 $B > gpurun_out/b${N}_nccl.json 2> gpurun_out/b${N}_nccl.err; echo "nccl rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b8_default*.json'))+['gpurun_out/b8_nccl.json']:
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        ro=r['roofline']
        print(f"{f:40s} value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} (with events {ro.get('ms_with_kernel_events')}) e2e={r['e2e']['value']:.0f} ({r['e2e']['ms_per_step']} ms) scan={ro['kernel_ms']} launch={ro['launch']} exch={ro['exchange']}")
        if 'also' in ro: print('   also:', json.dumps(ro['also']))
        if 'parity' in r: print('   parity:', json.dumps(r['parity']), json.dumps(r.get('comm')))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-300:])
PY
for f in gpurun_out/b8_default_full.err; do echo "== $f"; tail -n 3 $f; done
