#!/bin/bash
# 1-GPU round: smoke, the -m gpu suite, the default bench + its reference arm
set -u
mkdir -p gpurun_out
T=${1:-r02b}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 3 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 25 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 50 --warmup 10 > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"; tail -n 5 gpurun_out/${T}_bench_n1.err; cut -c1-3000 gpurun_out/${T}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
