#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02x}
for rep in 1 2; do
for ST in 7 6 5; do
B200RAG_GEMM_STAGES=$ST timeout 300 python tools/probe.py --rows 10000000 --modes dense --batches 256 --iters 30 > gpurun_out/${T}_probe_st${ST}_$rep.log 2>&1; echo "stages=$ST rep=$rep"; grep "mode=" gpurun_out/${T}_probe_st${ST}_$rep.log | cut -c1-120
done
done
