#!/bin/bash
# Round 2, call 6 (1 GPU, ~6 min): the k-th-of-slices threshold rule and the 8-warp Stage-2 epilogue on hardware:
# parity suites, per-rank step probe with the rule on / off and without the fused top-k (TS_DBG_NOTOPK: scan only),
# launch list of the step, Stage-2 probes.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s1 tests/test_gpu_stage1.py
run s2 tests/test_gpu_stage2.py
run zfull tests/test_gpu_zzz_fullsize.py tests/test_gpu_zz_tf32.py
V="TS_FUSE=1,TS_DBG_NOKTH=1,TS_DBG_NOTOPK=1,TS_FUSE=0"
timeout 300 python tools/step_probe.py --rows 1250000 --variants $V > gpurun_out/step_probe.jsonl 2> gpurun_out/step_probe.err; echo "rc=$?"; cat gpurun_out/step_probe.jsonl; tail -3 gpurun_out/step_probe.err
timeout 300 python tools/step_probe.py --rows 10000000 --steps 30 --batches 32 --variants TS_FUSE=1,TS_DBG_NOKTH=1,TS_DBG_NOTOPK=1 >> gpurun_out/step_probe.jsonl 2>> gpurun_out/step_probe.err; tail -3 gpurun_out/step_probe.jsonl
P="timeout 300 python tools/s2_probe.py"
O=gpurun_out/s2_probe8.jsonl; : > $O
for cfg in "--ndocs 1000000" "--Lq 128" "--Lq 64" "--lo 16 --hi 40" "--lo 180 --hi 180" "--dim 64" "--B 8 --C 500"; do $P $cfg --tag "flow8 $cfg" >> $O 2>> gpurun_out/s2_probe8.err; done
python - <<'PY'
import json
for l in open('gpurun_out/s2_probe8.jsonl'):
    r=json.loads(l); print(f"{r['tag']:45s} kernel={r['kernel_ms']:.3f} ms  {r['cand_per_s']/1e6:.1f} Mcand/s  {r['GBps']:.0f} GB/s hbm={r['hbm_frac']:.2f}")
PY
CMD="python tools/step_probe.py --rows 1250000 --steps 3 --batches 32 --variants TS_FUSE=1"
timeout 300 $CMD > gpurun_out/plain_step.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"s1_umma|select_kernel|convert_rows" -c 60 --csv --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_step.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_step.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID']
if hi:
    h=rows[hi[0]]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
    for r in rows[hi[0]+1:][-21:]:
        if len(r)>mv: print(r[kn][:50], r[mv])
PY
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json
