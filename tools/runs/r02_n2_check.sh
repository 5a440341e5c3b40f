#!/bin/bash
# 2-GPU confirmation after the tail / fuse / compressed-leg changes: dist_check, the default bench (with the compressed leg),
# top-100 over 25M rows (12.5M per GPU: the north-star shard size, merged over 2 shards)
set -u
mkdir -p gpurun_out
T=${1:-r02n2}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 2 --master-port 29541 tools/dist_check.py > gpurun_out/${T}_dist_check_n2.log 2>&1; echo "dc2 rc=$?"; tail -n 1 gpurun_out/${T}_dist_check_n2.log | cut -c1-300
timeout 400 $TR --nproc-per-node 2 --master-port 29548 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/${T}_bench_n2.json 2> gpurun_out/${T}_bench_n2.err; echo "b2 rc=$?"; tail -n 2 gpurun_out/${T}_bench_n2.err | cut -c1-300
timeout 400 $TR --nproc-per-node 2 --master-port 29549 bench.py --gpus 2 --rows 25000000 --top-k 100 --steps 50 --warmup 10 > gpurun_out/${T}_bench_n2_25m_top100.json 2> gpurun_out/${T}_bench_n2_25m_top100.err; echo "b2 25M top-100 rc=$?"; tail -n 2 gpurun_out/${T}_bench_n2_25m_top100.err | cut -c1-300
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"].items() if isinstance(v,float)}, "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"].get("mismatches"), "frac", round(d["roofline"]["frac"],3), "amb", d["ambiguous_flags"], d["clocks"]["sm_mhz"])
    if "compressed_candidate_scan" in d: print("   compressed leg:", {k:v for k,v in d["compressed_candidate_scan"].items() if k!="note"})
PY
