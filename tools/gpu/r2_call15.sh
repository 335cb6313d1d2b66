#!/bin/bash
# Round 2, call 15 (1 GPU, <1 min): where the 15 us between the grid barrier and the release of the first accumulator go
mkdir -p gpurun_out
for v in TS_FUSE=1 TS_DBG_NOKTHSTART=1; do
for B in 1 32; do
  TS_DBG_TRACE=1 timeout 200 python tools/step_probe.py --rows 1250000 --steps 3 --batches $B --variants $v 2> gpurun_out/trace.err | tail -1 | cut -c1-80
  grep "ts trace\]" gpurun_out/trace.err | tail -1 | sed 's/^\[ts trace\] //' > gpurun_out/trace6_${v}_b$B.json; rm -f gpurun_out/trace.err
done; done
python - <<'PY'
import json, glob, statistics as st
for f in sorted(glob.glob('gpurun_out/trace6_*.json')):
    c = json.load(open(f))['ctas']
    q = lambda k: tuple(round(x/1000,2) for x in (min(x[k] for x in c), st.median(x[k] for x in c), max(x[k] for x in c)))
    print(f.split('/')[-1], 'acc0', tuple(round(v/1000,2) for v in (min(x['acc'][0] for x in c), st.median(x['acc'][0] for x in c), max(x['acc'][0] for x in c))),
          'pass1', q('pass1'), 'bar', q('bar'), 'bound', q('bound'), 'tile0', q('tile0'), 'acc1', round(st.median(x['acc'][1] for x in c)/1000,2),
          'exit-lastacc', round(st.median(x['exit']-x['acc'][-1] for x in c)/1000,2))
PY
