#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02d}
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_retriever.py -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 15 gpurun_out/${T}_tests.log
B200RAG_PIPELINE_TAIL=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py > gpurun_out/${T}_dist_check_n1_pipe1.log 2>&1; echo "dc1p1 rc=$?"; tail -n 4 gpurun_out/${T}_dist_check_n1_pipe1.log
B200RAG_PIPELINE_TAIL=1 timeout 600 python bench.py --steps 200 --warmup 20 --rows 1250000 --no-cpu-baseline > gpurun_out/${T}_bench_1p25m_pipe1.json 2> gpurun_out/${T}_bench_1p25m_pipe1.err; echo "bench 1.25M p1 rc=$?"; tail -n 3 gpurun_out/${T}_bench_1p25m_pipe1.err; cut -c1-200 gpurun_out/${T}_bench_1p25m_pipe1.json
