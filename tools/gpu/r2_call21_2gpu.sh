#!/bin/bash
# Round 2, call 21 (2 GPUs, ~3 min): the whole Stage-1 exchange in ONE kernel (select + push + wait + merge): dist_check,
# per-rank-load A/B against the two-kernel form and NCCL, full bench line.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 200 $TR --master-port 29611 tools/dist_check.py > gpurun_out/dist_default.log 2>&1; echo "dist_check default rc=$? $(grep 'dist_check ok' gpurun_out/dist_default.log)"; tail -2 gpurun_out/dist_default.log | cut -c1-300
B="bench.py --gpus 2 --steps 100 --warmup 5 --rows 2500000 --no-extra --no-cpu --no-parity"
for rep in 1 2; do
timeout 200 $TR --master-port 29613 $B > gpurun_out/b2s_xfuse_$rep.json 2> gpurun_out/b2s.err; echo "xfuse rc=$?"
TS_XFUSE=0 timeout 200 $TR --master-port 29614 $B > gpurun_out/b2s_twok_$rep.json 2>> gpurun_out/b2s.err; echo "two-kernel rc=$?"
TS_P2P=0 timeout 200 $TR --master-port 29615 $B > gpurun_out/b2s_nccl_$rep.json 2>> gpurun_out/b2s.err; echo "nccl rc=$?"
done
timeout 420 $TR --master-port 29616 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_default.json 2> gpurun_out/b2_default.err; echo "full default rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/b2s_*.json'))+['gpurun_out/b2_default.json']:
    try:
        r=json.loads(open(f).read().strip().splitlines()[-1])
        ro=r['roofline']
        print(f"{f:36s} value={r['value']:.0f} q/s ms={r['ms_per_step']:.4f} e2e={r['e2e']['value']:.0f} ({r['e2e']['ms_per_step']}) scan={ro['kernel_ms']} launches={r['gpu_launches']} exch={ro['exchange']}")
        if 'also' in ro: print('   s2_c4:', json.dumps(ro['also'].get('s2_c4')), ' c5:', json.dumps(ro['also'].get('c5')))
        if 'parity' in r: print('   parity:', json.dumps(r.get('parity')))
    except Exception as e: print(f, 'ERR', e, open(f).read()[-300:])
PY
tail -n 3 gpurun_out/b2s.err
