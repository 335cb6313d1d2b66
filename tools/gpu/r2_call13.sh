#!/bin/bash
# Round 2, call 13 (1 GPU, ~1 min): GPU time line of one search step on the 8-GPU per-rank shard -- end of the query prep,
# first / last CTA of the scan, start of the select kernel (launch gaps that programmatic dependent launch could hide).
mkdir -p gpurun_out
for B in 1 32; do
  TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows 1250000 --steps 4 --batches $B --variants TS_FUSE=1 2> gpurun_out/trace.err | tail -1 | cut -c1-100
  grep "ts trace" gpurun_out/trace.err | tail -6 | cut -c1-90
  grep "ts trace\]" gpurun_out/trace.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace4_b$B.json
  grep "ts trace select" gpurun_out/trace.err | tail -1 > gpurun_out/trace4_b$B.select
done
python - <<'PY'
import json
for B in (1, 32):
    d = json.load(open(f'gpurun_out/trace4_b{B}.json'))
    sel = int(open(f'gpurun_out/trace4_b{B}.select').read().split()[-1])
    ex = max(c['exit'] for c in d['ctas']); en = min(c['entry'] for c in d['ctas'])
    print(f"B={B}: prep exit {d['prep_exit']} ns before/after first scan stamp (t0); first entry {en}; last exit {ex}; select entry {sel - d['t0']} -> gap scan->select {sel - d['t0'] - ex} ns, gap prep->scan {-d['prep_exit']} ns")
PY
