#!/bin/bash
# what the driver runs at round end on one GPU: smoke, -m gpu, the default bench (both arms)
set -u
mkdir -p gpurun_out
T=${1:-r02z}
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${T}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 2 gpurun_out/${T}_tests.log
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/${T}_bench_driver_args.json 2> gpurun_out/${T}_bench_driver_args.err; echo "bench (driver args) rc=$?"; tail -n 3 gpurun_out/${T}_bench_driver_args.err
timeout 600 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench (no flags) rc=$?"
python - <<PY
import json,glob
for f in ("gpurun_out/${T}_bench_driver_args.json","gpurun_out/${T}_bench_default.json"):
    d=json.load(open(f))
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "plugin", round(d["e2e_plugin"]["value"],1), d["e2e_plugin"].get("matches_c_abi_path"), "oracle", d["oracle_check"].get("mismatches"), "frac", round(d["roofline"]["frac"],3), "incl", round(d["value_incl_copies"]["value"],1), d["clocks"], d["gpu_launches"], "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["measured_10k"]["ms_per_query"])
    print("   compressed leg:", {k:v for k,v in (d.get("compressed_candidate_scan") or {}).items() if k!="note"})
PY
