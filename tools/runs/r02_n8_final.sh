#!/bin/bash
# final 8-GPU confirmation of the round's code: dist_check, the default bench at 8 and 4, the north-star config, the plugin group
set -u
mkdir -p gpurun_out
T=${1:-r02y}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29541 tools/dist_check.py > gpurun_out/${T}_dist_check_n8.log 2>&1; echo "dc8 rc=$?"; tail -n 1 gpurun_out/${T}_dist_check_n8.log
timeout 400 $TR --nproc-per-node 8 --master-port 29544 bench.py --gpus 8 --steps 200 --warmup 20 > gpurun_out/${T}_bench_n8.json 2> gpurun_out/${T}_bench_n8.err; echo "b8 rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n8.json
timeout 400 $TR --nproc-per-node 8 --master-port 29547 bench.py --gpus 8 --rows 100000000 --top-k 100 --steps 100 --warmup 10 > gpurun_out/${T}_bench_n8_100m_top100.json 2> gpurun_out/${T}_bench_n8_100m_top100.err; echo "b8 100M rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n8_100m_top100.json
timeout 400 $TR --nproc-per-node 4 --master-port 29546 bench.py --gpus 4 --steps 200 --warmup 20 > gpurun_out/${T}_bench_n4.json 2> gpurun_out/${T}_bench_n4.err; echo "b4 rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n4.json
timeout 400 $TR --nproc-per-node 2 --master-port 29548 bench.py --gpus 2 --steps 200 --warmup 20 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "b2 rc=$?"; cut -c1-150 gpurun_out/${T}_bench_n2.json
timeout 300 python tools/plugin_group_bench.py --devices 0,1,2,3,4,5,6,7 --steps 200 > gpurun_out/${T}_plugin_group_n8.json 2> gpurun_out/${T}_plugin_group_n8.err; echo "plugin group 8 rc=$?"; cat gpurun_out/${T}_plugin_group_n8.json
