#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02w}
timeout 300 python tools/probe.py --rows 10000000 --modes dense --batches 256,128 --iters 20 > gpurun_out/${T}_probe_dense.log 2>&1; grep "mode=" gpurun_out/${T}_probe_dense.log | cut -c1-200
timeout 300 python tools/probe.py --rows 1000000 --modes dense --batches 256 --iters 30 > gpurun_out/${T}_probe_dense_1m.log 2>&1; grep "mode=" gpurun_out/${T}_probe_dense_1m.log | cut -c1-200
timeout 600 python bench.py --steps 30 --warmup 5 --batch 256 --mode dense --no-cpu-baseline > gpurun_out/${T}_bench_n1_b256_dense.json 2> gpurun_out/${T}_bench_n1_b256_dense.err; echo "bench 10M B=256 dense rc=$?"
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/${T}_tests.log
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    d=json.load(open(f))
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"].get("mismatches"), "frac", round(d["roofline"]["frac"],3), d["roofline"].get("tensor_TFLOPs"), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
PY
