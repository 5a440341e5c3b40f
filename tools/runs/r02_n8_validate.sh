#!/bin/bash
# N-GPU validation: dist_check (pipelined and classic) + bench at N (pipelined = default, classic) and at N/2
set -u
mkdir -p gpurun_out
N=${1:-8}
T=${2:-r02h}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node $N --master-port 29541 tools/dist_check.py > gpurun_out/${T}_dist_check_n${N}_pipe1.log 2>&1; echo "dc${N} p1 rc=$?"; tail -n 1 gpurun_out/${T}_dist_check_n${N}_pipe1.log
timeout 400 $TR --nproc-per-node $N --master-port 29544 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${T}_bench_n${N}_pipe1.json 2> gpurun_out/${T}_bench_n${N}_pipe1.err; echo "b${N} p1 rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n${N}_pipe1.json
B200RAG_PIPELINE_TAIL=0 timeout 400 $TR --nproc-per-node $N --master-port 29545 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/${T}_bench_n${N}_pipe0.json 2> gpurun_out/${T}_bench_n${N}_pipe0.err; echo "b${N} p0 rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n${N}_pipe0.json
B200RAG_PIPELINE_TAIL=0 timeout 300 $TR --nproc-per-node $N --master-port 29543 tools/dist_check.py > gpurun_out/${T}_dist_check_n${N}_pipe0.log 2>&1; echo "dc${N} p0 rc=$?"; tail -n 1 gpurun_out/${T}_dist_check_n${N}_pipe0.log
if [ $N -ge 4 ]; then
H=$((N/2))
timeout 300 $TR --nproc-per-node $H --master-port 29542 tools/dist_check.py > gpurun_out/${T}_dist_check_n${H}_pipe1.log 2>&1; echo "dc${H} p1 rc=$?"; tail -n 1 gpurun_out/${T}_dist_check_n${H}_pipe1.log
timeout 400 $TR --nproc-per-node $H --master-port 29546 bench.py --gpus $H --steps 200 --warmup 20 > gpurun_out/${T}_bench_n${H}_pipe1.json 2> gpurun_out/${T}_bench_n${H}_pipe1.err; echo "b${H} p1 rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n${H}_pipe1.json
fi
if [ $N -ge 8 ]; then
timeout 400 $TR --nproc-per-node 8 --master-port 29547 bench.py --gpus 8 --rows 100000000 --top-k 100 --steps 100 --warmup 10 > gpurun_out/${T}_bench_n8_100m_top100.json 2> gpurun_out/${T}_bench_n8_100m_top100.err; echo "b8 100M rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n8_100m_top100.json
timeout 300 python tools/plugin_group_bench.py --devices 0,1,2,3,4,5,6,7 --steps 200 > gpurun_out/${T}_plugin_group_n8.json 2> gpurun_out/${T}_plugin_group_n8.err; echo "plugin group 8 rc=$?"; cat gpurun_out/${T}_plugin_group_n8.json
fi
