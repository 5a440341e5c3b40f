#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02k}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/${T}_tests.log
for PT in 1 0; do
B200RAG_PIPELINE_TAIL=$PT timeout 600 python bench.py --steps 50 --warmup 10 --rows 12500000 --top-k 100 --no-cpu-baseline > gpurun_out/${T}_bench_12p5m_top100_pipe$PT.json 2> gpurun_out/${T}_bench_12p5m_top100_pipe$PT.err; echo "bench 12.5M top100 pipe$PT rc=$?"; tail -n 2 gpurun_out/${T}_bench_12p5m_top100_pipe$PT.err
done
timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench 10M rc=$?"
B200RAG_SCAN_DYNAMIC=0 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_n1_static.json 2> gpurun_out/${T}_bench_n1_static.err; echo "bench 10M static rc=$?"; timeout 600 python bench.py --steps 100 --warmup 10 --rows 1250000 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_1p25m.json 2> gpurun_out/${T}_bench_1p25m.err
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), "p50", round(d["p50_ms"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"]["rank0"].items()}, "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"].get("mismatches"), "sparse_ms", round(d["roofline"].get("sparse_scan_ms",0),4), "frac", round(d["roofline"]["frac"],3))
PY
