#!/bin/bash
# Second GPU call of round 2 (1 GPU, ~10 min): parity + A/B timing of every opt-in kernel variant.
# Usage: gpurun --timeout 1500 -- tools/gpu/round2_variants.sh   (after round2_first.sh is green)
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
TS_TEST_EXPERIMENTAL=1 run zvariants tests/test_gpu_zzz_fullsize.py -k scan_variants
# select kernel: parallel count prefix (default since the emulator-validated rewrite) vs the first version
P8="timeout 300 python tools/perf_probe.py --paths umma --rows 1250000 --dim 1024 --batches 1,32 --steps 50"
$P8 --tag select_v2 > gpurun_out/select_probe.jsonl 2> gpurun_out/select_probe.err
TS_SELECT_V1=1 $P8 --tag select_v1 >> gpurun_out/select_probe.jsonl 2>> gpurun_out/select_probe.err
# opt-in single-launch scan: parity first, then its effect on small shards
TS_FUSE=1 run s1_fused tests/test_gpu_stage1.py -k "umma_path or planted or duplicates or cosine or merge_of_shards"
P="timeout 300 python tools/perf_probe.py --paths umma --rows 1250000 --dim 1024 --batches 1,32,128,1024"
$P --tag twolaunch > gpurun_out/fuse_probe.jsonl 2> gpurun_out/fuse_probe.err
TS_FUSE=1 $P --tag fused >> gpurun_out/fuse_probe.jsonl 2>> gpurun_out/fuse_probe.err
python - <<'PY'
import json
for l in open('gpurun_out/fuse_probe.jsonl'):
    r=json.loads(l); print(f"{r['tag']:10s} B={r['B']:5d} step={r['step_ms']:.3f} scan={r['scan_ms_per_launch']:.3f}")
PY
# opt-in Stage-2 epilogue (V2: LDS/STS, no -inf init, paired tcgen05.ld + max tree): parity, then A/B timing
TS_S2_V2=1 run s2_v2 tests/test_gpu_stage2.py
for v in "" 1; do
  for cfg in "--lo 16 --hi 180" "--lo 180 --hi 180" "--lo 16 --hi 40" "--Lq 128" "--dim 768 --ndocs 50000"; do
    TS_S2_V2=$v timeout 300 python tools/s2_probe.py $cfg --tag "v2=${v:-0} $cfg" >> gpurun_out/s2_v2_probe.jsonl 2>> gpurun_out/s2_v2_probe.err
  done
done
python - <<'PY'
import json
for l in open('gpurun_out/s2_v2_probe.jsonl'):
    r=json.loads(l); print(f"{r['tag']:45s} kernel={r['kernel_ms']:.3f} ms  {r['cand_per_s']/1e6:.1f} Mcand/s  hbm={r['hbm_frac']:.2f}")
PY
# two epilogue warpgroups (TS_S2_EPI2), alone and with V2: parity, then bit-equality + timing against the default
TS_S2_EPI2=1 run s2_epi2 tests/test_gpu_stage2.py
TS_S2_EPI2=1 TS_S2_V2=1 run s2_epi2_v2 tests/test_gpu_stage2.py
timeout 300 python tools/variant_ab.py --what s2 > gpurun_out/s2_ab.jsonl 2> gpurun_out/s2_ab.err; cat gpurun_out/s2_ab.jsonl
timeout 300 python tools/variant_ab.py --what s1 > gpurun_out/s1_ab.jsonl 2> gpurun_out/s1_ab.err; cat gpurun_out/s1_ab.jsonl
# BASELINE configs[1] with the reference's own storage dtype (fp32 corpus, CUDA-core stream scan): not timed in round 1
timeout 300 python tools/perf_probe.py --rows 1000000 --dim 768 --dtype fp32 --paths stream --batches 1,2,4 --tag c2_fp32 > gpurun_out/c2_fp32.jsonl 2> gpurun_out/c2_fp32.err; cat gpurun_out/c2_fp32.jsonl
# CTA pairs (cta_group::2) for B >= 129: parity, then A/B at the tensor-bound batch sizes
TS_PAIR=1 run s1_pair tests/test_gpu_stage1.py -k "umma_path"
PP="timeout 600 python tools/perf_probe.py --paths umma --rows 10000000 --dim 1024 --batches 256,512,1024 --steps 5"
$PP --tag single > gpurun_out/pair_probe.jsonl 2> gpurun_out/pair_probe.err
TS_PAIR=1 $PP --tag pair >> gpurun_out/pair_probe.jsonl 2>> gpurun_out/pair_probe.err
python - <<'PY'
import json
for l in open('gpurun_out/pair_probe.jsonl'):
    r=json.loads(l); print(f"{r['tag']:8s} B={r['B']:5d} scan={r['scan_ms_per_launch']:.3f} ms  {r['TFLOPs']:.0f} TFLOP/s  ({r['tensor_frac_sustained']:.2f} of sustained)")
PY
# fp32 storage on the tensor path (kind::tf32): parity, then BASELINE configs[1] (1M x 768 fp32) at B = 1/32/1024
TS_TEST_EXPERIMENTAL=1 run zz_tf32 tests/test_gpu_zz_tf32.py
timeout 300 python tools/perf_probe.py --rows 1000000 --dim 768 --dtype fp32 --paths umma --batches 1,32,128,1024 --tag c2_fp32_tf32 > gpurun_out/c2_tf32.jsonl 2> gpurun_out/c2_tf32.err; cat gpurun_out/c2_tf32.jsonl
# approximate mode (inverted lists over the resident rows): list scan vs the exact scan of the same shard
timeout 600 python tools/ivf_probe.py --rows 10000000 --dim 1024 --batches 1,8,32 --tag ivf_10M > gpurun_out/ivf_probe.jsonl 2> gpurun_out/ivf_probe.err
timeout 300 python tools/ivf_probe.py --rows 1000000 --dim 768 --batches 1,32 --tag ivf_1M >> gpurun_out/ivf_probe.jsonl 2>> gpurun_out/ivf_probe.err
python - <<'PY'
import json
for l in open('gpurun_out/ivf_probe.jsonl'):
    r=json.loads(l)
    if r['what'] in ('ivf','exact'): print(f"{r['tag']:8s} {r['what']:6s} B={r['B']:3d} step={r['step_ms']:.3f} ms scan={r['scan_ms']:.3f} ms {r['GBps']:.0f} GB/s ({r['hbm_frac']:.2f})")
    else: print(r)
PY
