#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02t}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
timeout 300 python tools/probe.py --rows 10000000 --modes dense --batches 256,128,64 --iters 20 > gpurun_out/${T}_probe_dense.log 2>&1; grep "mode=" gpurun_out/${T}_probe_dense.log | cut -c1-230
timeout 300 python tools/probe.py --rows 1000000 --modes dense --batches 256 --iters 30 > gpurun_out/${T}_probe_dense_1m.log 2>&1; grep "mode=" gpurun_out/${T}_probe_dense_1m.log | cut -c1-230
timeout 600 python bench.py --steps 30 --warmup 5 --batch 256 --mode dense --no-cpu-baseline > gpurun_out/${T}_bench_n1_b256_dense.json 2> gpurun_out/${T}_bench_n1_b256_dense.err; echo "bench 10M B=256 dense rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 --batch 64 --no-cpu-baseline > gpurun_out/${T}_bench_n1_b64.json 2> gpurun_out/${T}_bench_n1_b64.err; echo "bench 10M B=64 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"]["rank0"].items()}, "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"].get("mismatches"), "frac", round(d["roofline"]["frac"],3), d["roofline"]["bound"], d["roofline"].get("tensor_TFLOPs"), d["clocks"]["sm_mhz"], d["clocks"]["power_w_max"])
PY
