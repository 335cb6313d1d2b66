#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1200 python -m pytest "$@" -q -m gpu -p no:cacheprovider --timeout 900 > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s2 tests/test_gpu_stage2.py
P="python tools/s2_probe.py"
$P --tag c4 > gpurun_out/s2.jsonl 2> gpurun_out/s2.err
$P --tag all180 --lo 180 --hi 180 >> gpurun_out/s2.jsonl 2>> gpurun_out/s2.err
$P --tag all248 --lo 248 --hi 248 >> gpurun_out/s2.jsonl 2>> gpurun_out/s2.err
$P --tag short --lo 16 --hi 40 >> gpurun_out/s2.jsonl 2>> gpurun_out/s2.err
$P --tag lq128 --Lq 128 >> gpurun_out/s2.jsonl 2>> gpurun_out/s2.err
$P --tag dim768 --dim 768 --ndocs 50000 >> gpurun_out/s2.jsonl 2>> gpurun_out/s2.err
python - <<'PY'
import json
for l in open('gpurun_out/s2.jsonl'):
    r=json.loads(l); print(f"{r['tag']:8s} kernel={r['kernel_ms']:.3f}ms cand/s={r['cand_per_s']/1e6:.1f}M GB/s={r['GBps']:.0f} frac={r['hbm_frac']:.2f}")
PY
CMD="python tools/s2_probe.py --steps 3"
$CMD > gpurun_out/plain_s2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:maxsim_umma -s 2 -c 1 -o gpurun_out/prof_s2 $CMD > gpurun_out/ncu_s2.log 2>&1
echo "ncu s2 rc=$?"
