#!/bin/bash
# Round 2, call 8 (2 GPUs, ~6 min): re-validate the multi-GPU bench after the collective-alignment fixes
# (every rank makes the Stage-2 parity call, barriers after rank-0-only checks): dist_check + full bench.py on
# both data planes, then the per-rank load of the 8-GPU job (2.5 M rows over 2 GPUs) with and without the peer exchange.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tools/dist_check.py > gpurun_out/dist_nccl.log 2>&1; echo "dist_check nccl rc=$? $(grep 'dist_check ok' gpurun_out/dist_nccl.log)"
TS_P2P=1 timeout 200 $TR --master-port 29612 tools/dist_check.py > gpurun_out/dist_p2p.log 2>&1; echo "dist_check p2p rc=$? $(grep 'dist_check ok' gpurun_out/dist_p2p.log)"
TS_P2P=0 timeout 420 $TR --master-port 29616 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_nccl.json 2> gpurun_out/b2_nccl.err; echo "full nccl rc=$?"
TS_P2P=1 timeout 420 $TR --master-port 29617 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_p2p.json 2> gpurun_out/b2_p2p.err; echo "full p2p rc=$?"
B="bench.py --gpus 2 --steps 50 --warmup 5 --rows 2500000 --no-extra --no-cpu --no-parity"
TS_P2P=0 timeout 200 $TR --master-port 29613 $B > gpurun_out/b2_small_nccl.json 2> gpurun_out/b2_small_nccl.err; echo "small nccl rc=$?"
TS_P2P=1 timeout 200 $TR --master-port 29614 $B > gpurun_out/b2_small_p2p.json 2> gpurun_out/b2_small_p2p.err; echo "small p2p rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2_*.json')):
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        ro=r['roofline']
        print(f"{f:40s} value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} ({r['e2e']['ms_per_step']} ms) scan={ro['kernel_ms']} launch={ro['launch']} eager={ro['ms_eager']} graph={ro['ms_graph']} exch={ro['exchange']}")
        if 'also' in ro: print('   also:', json.dumps(ro['also']))
        if 'parity' in r: print('   parity:', json.dumps(r['parity']))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-300:])
PY
tail -4 gpurun_out/b2_nccl.err gpurun_out/b2_p2p.err
