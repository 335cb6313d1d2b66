#!/bin/bash
# Round 2, Stage-2 call (1 GPU, ~6 min): flow kernel on tile-layout shards -- parity, A/B against the first kernel
# (TS_S2_FLOW=0: row-major shards), role trace, ncu.
# Usage: gpurun --timeout 1500 -- tools/gpu/r2_call1.sh
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run s2 tests/test_gpu_stage2.py
TS_S2_FLOW=0 run s2_first tests/test_gpu_stage2.py
P="timeout 300 python tools/s2_probe.py"
O=gpurun_out/s2_probe.jsonl; E=gpurun_out/s2_probe.err; : > $O; : > $E
$P --ndocs 1000000 --tag "flow c4_1M" >> $O 2>> $E
TS_S2_FLOW=0 $P --ndocs 1000000 --tag "first c4_1M" >> $O 2>> $E
for st in 1 2; do TS_S2_STAGES=$st $P --tag "flow stages=$st" >> $O 2>> $E; done
TS_S2_TILE=128 $P --tag "flow tile=128" >> $O 2>> $E
TS_S2_TILE=128 TS_S2_STAGES=3 $P --tag "flow tile=128 stages=3" >> $O 2>> $E
TS_S2_ABUFS=2 $P --tag "flow abufs=2" >> $O 2>> $E
for cfg in "--lo 16 --hi 180" "--lo 180 --hi 180" "--lo 16 --hi 40" "--Lq 128" "--dim 64" "--dim 256 --ndocs 100000" "--B 2000 --C 32" "--B 1 --C 1000" "--B 8 --C 500"; do
  $P $cfg --tag "flow $cfg" >> $O 2>> $E
  TS_S2_FLOW=0 $P $cfg --tag "first $cfg" >> $O 2>> $E
done
TS_S2_ABUFS=1 $P --B 2000 --C 32 --tag "flow abufs=1 --B 2000 --C 32" >> $O 2>> $E
TS_S2_ABUFS=2 $P --B 2000 --C 32 --tag "flow abufs=2 --B 2000 --C 32" >> $O 2>> $E
python - <<'PY'
import json
for l in open('gpurun_out/s2_probe.jsonl'):
    r=json.loads(l); print(f"{r['tag']:45s} kernel={r['kernel_ms']:.3f} ms  {r['cand_per_s']/1e6:.1f} Mcand/s  {r['GBps']:.0f} GB/s hbm={r['hbm_frac']:.2f}")
PY
# role trace (cycle counters per warp role, mean/max over CTAs)
TS_S2_TRACE=1 $P --steps 2 --tag trace > gpurun_out/s2_trace.out 2> gpurun_out/s2_trace.err; grep "s2 trace" gpurun_out/s2_trace.err | tail -1
TS_S2_TRACE=1 $P --steps 2 --lo 180 --hi 180 --tag trace180 > /dev/null 2> gpurun_out/s2_trace180.err; grep "s2 trace" gpurun_out/s2_trace180.err | tail -1
run pipe tests/test_gpu_pipeline.py
run zshards tests/test_gpu_z_shards.py
# ncu: flow kernel and the first kernel on config #4
NCU="ncu --set full --clock-control none --import-source on"
CMD="python tools/s2_probe.py --steps 2"
timeout 300 $CMD > gpurun_out/plain_s2.log 2>&1 && \
timeout 600 $NCU -k regex:maxsim_flow -s 2 -c 1 -o gpurun_out/prof_s2_flow $CMD > gpurun_out/ncu_s2_flow.log 2>&1
echo "ncu flow rc=$?"
ls -la gpurun_out/*.ncu-rep
