#!/bin/bash
set -u
mkdir -p gpurun_out
N=${1:-2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node $N"
B200RAG_PIPELINE_TAIL=1 timeout 300 $TR --master-port 29541 tools/dist_check.py > gpurun_out/r02f_dist_check_n${N}_pipe1.log 2>&1; echo "dc p1 rc=$?"; tail -n 1 gpurun_out/r02f_dist_check_n${N}_pipe1.log
B200RAG_PIPELINE_TAIL=0 timeout 300 $TR --master-port 29542 tools/dist_check.py > gpurun_out/r02f_dist_check_n${N}_pipe0.log 2>&1; echo "dc p0 rc=$?"; tail -n 1 gpurun_out/r02f_dist_check_n${N}_pipe0.log
B200RAG_PIPELINE_TAIL=1 timeout 400 $TR --master-port 29543 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r02f_bench_n${N}_pipe1.json 2> gpurun_out/r02f_bench_n${N}_pipe1.err; echo "b p1 rc=$?"; cut -c1-160 gpurun_out/r02f_bench_n${N}_pipe1.json
B200RAG_PIPELINE_TAIL=0 timeout 400 $TR --master-port 29544 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/r02f_bench_n${N}_pipe0.json 2> gpurun_out/r02f_bench_n${N}_pipe0.err; echo "b p0 rc=$?"; cut -c1-160 gpurun_out/r02f_bench_n${N}_pipe0.json
