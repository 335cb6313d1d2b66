#!/bin/bash
# First GPU call of round 2 (1 GPU, ~6 min): runs everything on the DEFAULT path that was written
# after the round-1 GPU budget ran out, then the bench.  Usage: gpurun --timeout 1200 -- tools/gpu/round2_first.sh
# The opt-in variants (TS_FUSE, TS_S2_V2, TS_PAIR, select A/B, fp32 timing) are in round2_variants.sh.
mkdir -p gpurun_out
run() { name=$1; shift; timeout 900 python -m pytest "$@" -q -m gpu -p no:cacheprovider > gpurun_out/$name.log 2>&1; echo "$name rc=$? $(tail -1 gpurun_out/$name.log)"; }
run pipe tests/test_gpu_pipeline.py
run s1 tests/test_gpu_stage1.py
run s2 tests/test_gpu_stage2.py
run zfull tests/test_gpu_zzz_fullsize.py
run zhybrid tests/test_gpu_z_hybrid.py
run zshards tests/test_gpu_z_shards.py
run zzivf tests/test_gpu_zz_ivf.py
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$? $(tail -1 gpurun_out/smoke.log)"
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 300 python tools/e2e_c5.py --docs 500000 > gpurun_out/e2e_c5_1gpu.json 2> gpurun_out/e2e_c5.err; echo "c5 rc=$?"; cat gpurun_out/e2e_c5_1gpu.json
