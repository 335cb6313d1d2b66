#!/bin/bash
# Round 2, call 3 (1 GPU, ~20 min): the whole GPU suite with the new defaults (flow kernel, TS_FUSE, TS_PAIR, tf32),
# the rewritten bench.py (both arms), variant timings, ncu of the Stage-1 scan at B = 32 / 1024 and a launch list.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/gpu_suite.log 2>&1; echo "gpu suite rc=$? $(tail -1 gpurun_out/gpu_suite.log)"
grep -E "passed|failed|error" gpurun_out/gpu_suite.log | tail -3
( time timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err ) 2> gpurun_out/bench_n1.time; echo "bench rc=$? bytes=$(wc -c < gpurun_out/bench_n1.json) $(grep real gpurun_out/bench_n1.time)"
cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
( time timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2> gpurun_out/bench_ref.time; echo "ref rc=$? $(grep real gpurun_out/bench_ref.time)"
cat gpurun_out/bench_ref.json
# tensor-bound regime: CTA pairs (default) vs single-CTA tiles, 10M x 1024
PP="timeout 600 python tools/perf_probe.py --paths umma --rows 10000000 --dim 1024 --batches 128,256,512,1024 --steps 5"
$PP --tag pair > gpurun_out/pair_probe.jsonl 2> gpurun_out/pair_probe.err
TS_PAIR=0 $PP --tag single >> gpurun_out/pair_probe.jsonl 2>> gpurun_out/pair_probe.err
# BASELINE configs[1]: 1M x 768, bf16 and the reference's fp32 (tf32 tensor path by default for B > 4; exact CUDA-core scan)
timeout 300 python tools/perf_probe.py --rows 1000000 --dim 768 --dtype bf16 --paths auto --batches 1,32,1024 --steps 20 --tag c2_bf16 > gpurun_out/c2_probe.jsonl 2> gpurun_out/c2_probe.err
timeout 300 python tools/perf_probe.py --rows 1000000 --dim 768 --dtype fp32 --paths auto --batches 1,4,32,1024 --steps 20 --tag c2_fp32_auto --selfcheck >> gpurun_out/c2_probe.jsonl 2>> gpurun_out/c2_probe.err
TS_TF32=0 timeout 300 python tools/perf_probe.py --rows 1000000 --dim 768 --dtype fp32 --paths auto --batches 32 --steps 5 --tag c2_fp32_cuda_cores >> gpurun_out/c2_probe.jsonl 2>> gpurun_out/c2_probe.err
python - <<'PY'
import json
for f in ('gpurun_out/pair_probe.jsonl','gpurun_out/c2_probe.jsonl'):
    for l in open(f):
        r=json.loads(l)
        if 'B' in r: print(f"{r['tag']:20s} {r['dtype']} B={r['B']:5d} step={r['step_ms']:.3f} scan={r['scan_ms_per_launch']:.3f} x{r['scans_per_step']:.0f}  {r['qps']:.0f} q/s  {r['corpus_GBps_per_scan']:.0f} GB/s ({r['hbm_frac']:.2f})  {r['TFLOPs']:.0f} TF ({r['tensor_frac_sustained']:.2f})")
        else: print(r)
PY
# approximate mode
timeout 600 python tools/ivf_probe.py --rows 10000000 --dim 1024 --batches 1,8,32 --tag ivf_10M > gpurun_out/ivf_probe.jsonl 2> gpurun_out/ivf_probe.err
python - <<'PY'
import json
for l in open('gpurun_out/ivf_probe.jsonl'):
    r=json.loads(l)
    if r.get('what') in ('ivf','exact'): print(f"{r['tag']:8s} {r['what']:6s} B={r['B']:3d} step={r['step_ms']:.3f} ms scan={r['scan_ms']:.3f} ms {r['GBps']:.0f} GB/s ({r['hbm_frac']:.2f})")
    else: print(str(r)[:300])
PY
# ncu: launch list of a bench step, then the scan at B = 32 (headline) and B = 1024 (pair + single)
NCU="ncu --set full --clock-control none --import-source on"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-parity"
timeout 600 $CMD > gpurun_out/plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
timeout 600 $CMD > gpurun_out/plain_s1.log 2>&1 && \
timeout 900 $NCU -k regex:s1_umma -s 4 -c 1 -o gpurun_out/prof_s1_b32 $CMD > gpurun_out/ncu_s1_b32.log 2>&1
echo "s1 B=32 rc=$?"
CMD="python tools/perf_probe.py --paths umma --rows 4000000 --dim 1024 --batches 1024 --steps 2"
timeout 600 $CMD > gpurun_out/plain_s1_b1024.log 2>&1 && \
timeout 900 $NCU -k regex:s1_pair -s 3 -c 1 -o gpurun_out/prof_s1_b1024_pair $CMD > gpurun_out/ncu_s1_b1024_pair.log 2>&1
echo "s1 B=1024 pair rc=$?"
TS_PAIR=0 timeout 600 $CMD > gpurun_out/plain_s1_b1024s.log 2>&1 && \
TS_PAIR=0 timeout 900 $NCU -k regex:s1_umma -s 3 -c 1 -o gpurun_out/prof_s1_b1024_single $CMD > gpurun_out/ncu_s1_b1024_single.log 2>&1
echo "s1 B=1024 single rc=$?"
ls -la gpurun_out/*.ncu-rep
