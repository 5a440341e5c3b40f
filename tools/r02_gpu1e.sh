#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02e}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for NP in 2 4; do
B200RAG_SAME_DEVICE=1 B200RAG_PIPELINE_TAIL=1 timeout 400 $TR --nproc-per-node $NP --master-port 2956$NP tools/dist_check.py > gpurun_out/${T}_dist_check_same_n${NP}_pipe1.log 2>&1; echo "same-device n$NP p1 rc=$?"; tail -n 2 gpurun_out/${T}_dist_check_same_n${NP}_pipe1.log
done
B200RAG_SAME_DEVICE=1 B200RAG_PIPELINE_TAIL=0 timeout 400 $TR --nproc-per-node 2 --master-port 29571 tools/dist_check.py > gpurun_out/${T}_dist_check_same_n2_pipe0.log 2>&1; echo "same-device n2 p0 rc=$?"; tail -n 2 gpurun_out/${T}_dist_check_same_n2_pipe0.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 12 gpurun_out/${T}_tests.log
for R in 8192 4096; do
timeout 300 python tools/probe.py --rows 10000000 --modes sparse --batches 1,64 --R $R --iters 30 > gpurun_out/${T}_probe_sparse_R$R.log 2>&1; echo "probe R=$R rc=$?"; grep "mode=" gpurun_out/${T}_probe_sparse_R$R.log | cut -c1-220
done
B200RAG_SPARSE_THREADS=256 timeout 300 python tools/probe.py --rows 10000000 --modes sparse --batches 1,64 --iters 30 > gpurun_out/${T}_probe_sparse_T256.log 2>&1; grep "mode=" gpurun_out/${T}_probe_sparse_T256.log | cut -c1-220
timeout 300 python tools/probe.py --rows 10000000 --modes dense --batches 256 --iters 20 > gpurun_out/${T}_probe_dense_b256.log 2>&1; grep "mode=" gpurun_out/${T}_probe_dense_b256.log | cut -c1-260
