#!/bin/bash
# Round 2, call 5 (2 GPUs, ~4 min): the multi-GPU data planes.  (1) parity of the sharded search / Stage 2 / shard files
# under NCCL and under the fused peer-memory exchange (TS_P2P=1); (2) bench.py at the per-rank load of the 8-GPU
# job (2.5 M rows over 2 GPUs = 1.25 M rows per rank): NCCL eager vs CUDA graph vs fused exchange; (3) the full
# bench at N = 2 with both planes.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29611 tools/dist_check.py > gpurun_out/dist_nccl.log 2>&1; echo "dist_check nccl rc=$? $(grep 'dist_check ok' gpurun_out/dist_nccl.log)"
TS_P2P=1 timeout 300 $TR --master-port 29612 tools/dist_check.py > gpurun_out/dist_p2p.log 2>&1; echo "dist_check p2p rc=$? $(grep 'dist_check ok' gpurun_out/dist_p2p.log)"; tail -5 gpurun_out/dist_p2p.log
B="bench.py --gpus 2 --steps 50 --warmup 5 --rows 2500000 --no-extra --no-cpu --no-parity"
TS_P2P=0 timeout 300 $TR --master-port 29613 $B > gpurun_out/b2_small_nccl.json 2> gpurun_out/b2_small_nccl.err; echo "small nccl rc=$?"
TS_P2P=1 timeout 300 $TR --master-port 29614 $B > gpurun_out/b2_small_p2p.json 2> gpurun_out/b2_small_p2p.err; echo "small p2p rc=$?"; tail -3 gpurun_out/b2_small_p2p.err
TS_P2P=1 TS_FUSE=0 timeout 300 $TR --master-port 29615 $B > gpurun_out/b2_small_p2p_nofuse.json 2> gpurun_out/b2_small_p2p_nofuse.err; echo "small p2p nofuse rc=$?"
TS_P2P=0 timeout 600 $TR --master-port 29616 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_nccl.json 2> gpurun_out/b2_nccl.err; echo "full nccl rc=$?"
TS_P2P=1 timeout 600 $TR --master-port 29617 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_p2p.json 2> gpurun_out/b2_p2p.err; echo "full p2p rc=$?"
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
