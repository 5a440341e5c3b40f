#!/bin/bash
# the north-star configuration once more on the final code (rank-merge fuse, 32-warp dense tail): 100M rows over 8 GPUs, top-100
set -u
mkdir -p gpurun_out
T=${1:-r02ns}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 330 $TR --nproc-per-node 8 --master-port 29547 bench.py --gpus 8 --rows 100000000 --top-k 100 --steps 100 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_n8_100m_top100.json 2> gpurun_out/${T}_bench_n8_100m_top100.err; echo "b8 100M rc=$?"; tail -n 2 gpurun_out/${T}_bench_n8_100m_top100.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_n8_100m_top100.json"))
print(round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"].items() if isinstance(v,float)}, "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"], "frac", round(d["roofline"]["frac"],3), "step_frac", round(d["roofline"]["step_frac"],3), "amb", d["ambiguous_flags"], d["clocks"])
print("   compressed leg:", {k:v for k,v in (d.get("compressed_candidate_scan") or {}).items() if k!="note"})
PY
