#!/bin/bash
# usage: run_gpu_multi.sh N   (N GPUs of one box); short timeouts: N x charging
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 240 $TR tools/dist_check.py > gpurun_out/dist_check_$N.log 2>&1; echo "dist_check rc=$? $(tail -1 gpurun_out/dist_check_$N.log)"
TS_P2P=1 timeout 240 $TR tools/dist_check.py > gpurun_out/dist_check_p2p_$N.log 2>&1; echo "dist_check p2p rc=$? $(tail -1 gpurun_out/dist_check_p2p_$N.log)"
TS_P2P=1 timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 5 --no-extra --no-cpu > gpurun_out/bench_p2p_n$N.json 2> gpurun_out/bench_p2p_n$N.err; echo "bench p2p N=$N rc=$?"; tail -c 700 gpurun_out/bench_p2p_n$N.json
timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench N=$N rc=$?"; tail -c 1500 gpurun_out/bench_n$N.json
timeout 300 $TR bench.py --gpus $N --steps 50 --warmup 5 --batch 1 --no-extra --no-cpu > gpurun_out/bench_n${N}_b1.json 2>> gpurun_out/bench_n$N.err; echo "bench B=1 rc=$?"; tail -c 900 gpurun_out/bench_n${N}_b1.json
tail -5 gpurun_out/bench_n$N.err
