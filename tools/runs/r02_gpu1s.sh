#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02s}
for BPC in 0 2 3; do
for NT in 128 256; do
B200RAG_SPARSE_BPC=$BPC B200RAG_SPARSE_THREADS=$NT timeout 300 python bench.py --steps 60 --warmup 10 --mode sparse --no-cpu-baseline --no-oracle-check > gpurun_out/${T}_sparse_bpc${BPC}_nt${NT}.json 2> gpurun_out/${T}_sparse_bpc${BPC}_nt${NT}.err; echo "bpc=$BPC nt=$NT rc=$?"
done
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_sparse_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), "e2e", round(d["e2e"]["value"],1), "sparse_ms", round(d["roofline"].get("sparse_scan_ms",0),4), round(d["roofline"].get("sparse_algorithmic_GBps",0),0))
PY
