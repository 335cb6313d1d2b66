#!/bin/bash
# Round 2, call 4 (1 GPU, ~3 min): the per-rank Stage-1 step of an 8-GPU job on one GPU (1.25 M-row shard):
# eager / CUDA graph / host call, TS_FUSE on and off, and the launch list of the product's own kernels.
mkdir -p gpurun_out
timeout 300 python tools/step_probe.py --rows 1250000 > gpurun_out/step_probe.jsonl 2> gpurun_out/step_probe.err; echo "rc=$?"; cat gpurun_out/step_probe.jsonl; tail -3 gpurun_out/step_probe.err
timeout 300 python tools/step_probe.py --rows 10000000 --steps 30 --batches 32 >> gpurun_out/step_probe.jsonl 2>> gpurun_out/step_probe.err; tail -2 gpurun_out/step_probe.jsonl
CMD="python tools/step_probe.py --rows 1250000 --steps 3 --batches 32"
timeout 300 $CMD > gpurun_out/plain_step.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:ts:: -c 200 --csv --log-file gpurun_out/launches_step.csv $CMD > gpurun_out/ncu_step.log 2>&1
echo "launch list rc=$?"
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/launches_step.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
for r in rows[hi+1:][-40:]:
    if len(r)>mv: print(r[kn][:70], r[mv])
PY
