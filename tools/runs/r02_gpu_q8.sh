#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02q8}
timeout 900 python -m pytest tests/test_gpu_q8.py -x -q > gpurun_out/${T}_tests.log 2>&1; echo "q8 tests rc=$?"; tail -n 12 gpurun_out/${T}_tests.log | cut -c1-200
timeout 600 python bench.py --compressed --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_10m_q8.json 2> gpurun_out/${T}_bench_10m_q8.err; echo "bench q8 10M rc=$?"; tail -n 3 gpurun_out/${T}_bench_10m_q8.err
timeout 600 python bench.py --compressed --steps 50 --warmup 10 --rows 1250000 --no-cpu-baseline > gpurun_out/${T}_bench_1p25m_q8.json 2> gpurun_out/${T}_bench_1p25m_q8.err; echo "bench q8 1.25M rc=$?"
B200RAG_Q8_SLACK=236 timeout 600 python bench.py --compressed --steps 50 --warmup 10 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_10m_q8_slack236.json 2> gpurun_out/${T}_bench_10m_q8_slack236.err; echo "bench q8 slack 236 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"]["rank0"].items()}, "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"].get("mismatches"), "frac", round(d["roofline"]["frac"],3), round(d["roofline"]["achieved"],0), d["ambiguous_flags"], d["clocks"]["sm_mhz"])
PY
