#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02m}
run() { # name, env..., args
  name=$1; shift
  env "$@" > /dev/null 2>&1
}
for C in 0 144 140 132; do
B200RAG_SCAN_CTAS=$C timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_10m_ctas$C.json 2> gpurun_out/${T}_bench_10m_ctas$C.err; echo "10M ctas=$C rc=$?"
B200RAG_SCAN_CTAS=$C timeout 600 python bench.py --steps 50 --warmup 10 --rows 12500000 --top-k 100 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_12p5m_top100_ctas$C.json 2> gpurun_out/${T}_bench_12p5m_top100_ctas$C.err; echo "12.5M top100 ctas=$C rc=$?"
B200RAG_SCAN_CTAS=$C timeout 600 python bench.py --steps 100 --warmup 10 --rows 1250000 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_1p25m_ctas$C.json 2> gpurun_out/${T}_bench_1p25m_ctas$C.err; echo "1.25M ctas=$C rc=$?"
done
B200RAG_SCAN_DYNAMIC=1 timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_bench_10m_dyn8.json 2> gpurun_out/${T}_bench_10m_dyn8.err; echo "10M dyn chunk8 rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"]["rank0"].items()}, "e2e", round(d["e2e"]["value"],1), "frac", round(d["roofline"]["frac"],3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
