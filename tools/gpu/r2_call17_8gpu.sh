#!/bin/bash
# Round 2, call 17 (8 GPUs, ~4 min): the multi-GPU data planes at N = 8 -- dist_check on the peer-memory exchange, the
# Stage-1 step with NCCL and with the fused exchange (no side measurements), then the full bench line on the exchange.
mkdir -p gpurun_out
N=8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
TS_P2P=1 timeout 200 $TR --master-port 29612 tools/dist_check.py > gpurun_out/dist_p2p_$N.log 2>&1; echo "dist_check p2p rc=$? $(grep 'dist_check ok' gpurun_out/dist_p2p_$N.log)"
B="bench.py --gpus $N --steps 50 --warmup 5 --no-extra --no-cpu"
TS_P2P=0 timeout 200 $TR --master-port 29613 $B > gpurun_out/b${N}_nccl.json 2> gpurun_out/b${N}_nccl.err; echo "nccl rc=$?"
TS_P2P=1 timeout 200 $TR --master-port 29614 $B > gpurun_out/b${N}_p2p.json 2> gpurun_out/b${N}_p2p.err; echo "p2p rc=$?"
TS_P2P=1 timeout 500 $TR --master-port 29617 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/b${N}_p2p_full.json 2> gpurun_out/b${N}_p2p_full.err; echo "full p2p rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b8_*.json')):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        ro=r['roofline']
        print(f"{f:40s} value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} ({r['e2e']['ms_per_step']} ms) scan={ro['kernel_ms']} launch={ro['launch']} eager={ro['ms_eager']} graph={ro['ms_graph']} exch={ro['exchange']}")
        if 'also' in ro: print('   also:', json.dumps(ro['also']))
        if 'parity' in r: print('   parity:', json.dumps(r['parity']))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-300:])
PY
for f in gpurun_out/b8_*.err; do echo "== $f"; tail -3 $f; done
