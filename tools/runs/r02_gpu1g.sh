#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02g}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "tests rc=$?"; tail -n 6 gpurun_out/${T}_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29551 tools/dist_check.py > gpurun_out/${T}_dist_check_n1.log 2>&1; echo "dc1 rc=$?"; tail -n 1 gpurun_out/${T}_dist_check_n1.log
timeout 600 python bench.py --steps 50 --warmup 10 --rows 1250000 --no-cpu-baseline > gpurun_out/${T}_bench_1p25m.json 2> gpurun_out/${T}_bench_1p25m.err; echo "bench 1.25M rc=$?"; tail -n 3 gpurun_out/${T}_bench_1p25m.err
timeout 600 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --mode sparse > gpurun_out/${T}_bench_n1_sparse.json 2> gpurun_out/${T}_bench_n1_sparse.err; echo "bench sparse rc=$?"
timeout 300 python tools/plugin_group_bench.py --devices 0,0 --rows 2000000 --steps 50 --check > gpurun_out/${T}_plugin_group_same.json 2> gpurun_out/${T}_plugin_group_same.err; echo "plugin group rc=$?"; tail -n 2 gpurun_out/${T}_plugin_group_same.err; cat gpurun_out/${T}_plugin_group_same.json
python - <<'PY'
import json
for f in ("bench_1p25m","bench_n1","bench_n1_sparse"):
    try:
        d=json.load(open(f"gpurun_out/${T}_"+f+".json".replace("${T}","")))
    except Exception as e:
        import glob
        d=json.load(open(glob.glob("gpurun_out/*_"+f+".json")[-1]))
    print(f, round(d["value"],1), round(d["ms_per_step"],4), d["step_breakdown_ms"]["rank0"], "e2e", round(d["e2e"]["value"],1), "plugin", d.get("e2e_plugin",{}).get("value"), "oracle", d["oracle_check"].get("mismatches"), d["roofline"].get("sparse_scan_ms"))
PY
