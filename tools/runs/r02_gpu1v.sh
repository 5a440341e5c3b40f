#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02v}
for PF in 1 0; do
B200RAG_GEMM_L2_PREFETCH=$PF timeout 300 python tools/probe.py --rows 10000000 --modes dense --batches 256,128 --iters 20 > gpurun_out/${T}_probe_dense_pf$PF.log 2>&1; echo "PF=$PF"; grep "mode=" gpurun_out/${T}_probe_dense_pf$PF.log | cut -c1-200
B200RAG_GEMM_L2_PREFETCH=$PF timeout 300 python tools/probe.py --rows 1000000 --modes dense --batches 256 --iters 30 > gpurun_out/${T}_probe_dense_1m_pf$PF.log 2>&1; grep "mode=" gpurun_out/${T}_probe_dense_1m_pf$PF.log | cut -c1-200
done
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/${T}_tests.log
