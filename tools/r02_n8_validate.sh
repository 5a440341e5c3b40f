#!/bin/bash
# 8-GPU validation of the pipelined tail: dist_check at 4 and 8 ranks, bench at 8 (tail off / on), 4 (on).
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B200RAG_PIPELINE_TAIL=1 timeout 300 $TR --nproc-per-node 8 --master-port 29541 tools/dist_check.py > gpurun_out/r02a_dist_check_n8_pipe1.log 2>&1; echo "dc8p1 rc=$?"
B200RAG_PIPELINE_TAIL=1 timeout 200 $TR --nproc-per-node 4 --master-port 29542 tools/dist_check.py > gpurun_out/r02a_dist_check_n4_pipe1.log 2>&1; echo "dc4p1 rc=$?"
timeout 200 $TR --nproc-per-node 8 --master-port 29543 tools/dist_check.py > gpurun_out/r02a_dist_check_n8_pipe0.log 2>&1; echo "dc8p0 rc=$?"
timeout 200 $TR --nproc-per-node 4 --master-port 29547 tools/dist_check.py > gpurun_out/r02a_dist_check_n4_pipe0.log 2>&1; echo "dc4p0 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r02a_bench_n8_pipe0.json 2> gpurun_out/r02a_bench_n8_pipe0.err; echo "b8p0 rc=$?"
B200RAG_PIPELINE_TAIL=1 timeout 300 $TR --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/r02a_bench_n8_pipe1.json 2> gpurun_out/r02a_bench_n8_pipe1.err; echo "b8p1 rc=$?"
B200RAG_PIPELINE_TAIL=1 timeout 300 $TR --nproc-per-node 4 --master-port 29546 bench.py --gpus 4 --steps 200 --warmup 20 > gpurun_out/r02a_bench_n4_pipe1.json 2> gpurun_out/r02a_bench_n4_pipe1.err; echo "b4p1 rc=$?"
tail -n 2 gpurun_out/r02a_dist_check_*.log
cat gpurun_out/r02a_bench_n8_pipe0.json gpurun_out/r02a_bench_n8_pipe1.json gpurun_out/r02a_bench_n4_pipe1.json | cut -c1-600
