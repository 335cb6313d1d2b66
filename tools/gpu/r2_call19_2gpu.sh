#!/bin/bash
# Round 2, call 19 (2 GPUs, ~4 min): Stage-2 scatter fused into the scoring kernel -- dist_check on the default planes
# (peer-memory Stage-1 exchange + Stage-2 scatter) and on NCCL, then the full bench line with both.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tools/dist_check.py > gpurun_out/dist_default.log 2>&1; echo "dist_check default rc=$? $(grep 'dist_check ok' gpurun_out/dist_default.log)"; tail -3 gpurun_out/dist_default.log | cut -c1-300
TS_P2P=0 timeout 200 $TR --master-port 29612 tools/dist_check.py > gpurun_out/dist_nccl.log 2>&1; echo "dist_check nccl rc=$? $(grep 'dist_check ok' gpurun_out/dist_nccl.log)"
timeout 420 $TR --master-port 29616 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_default.json 2> gpurun_out/b2_default.err; echo "full default rc=$?"
TS_P2P=0 timeout 420 $TR --master-port 29617 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_nccl.json 2> gpurun_out/b2_nccl.err; echo "full nccl rc=$?"
python - <<'PY'
import json,glob
for f in ['gpurun_out/b2_default.json','gpurun_out/b2_nccl.json']:
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        ro=r['roofline']
        print(f"{f:40s} value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} scan={ro['kernel_ms']} exch={ro['exchange']}")
        print('   s2_c4:', json.dumps(ro['also'].get('s2_c4')), ' c5:', json.dumps(ro['also'].get('c5')))
        print('   parity:', json.dumps(r.get('parity')), json.dumps(r.get('comm')))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-300:])
PY
tail -n 4 gpurun_out/b2_default.err
