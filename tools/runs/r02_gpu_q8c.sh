#!/bin/bash
# after the leg-tail threshold fix (rotated stride), the rank-merge fuse and the pipelined 8-bit scan: full GPU suite,
# headline bench, top-100 shard, per-kernel times
set -u
mkdir -p gpurun_out
T=${1:-r02q8d}
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 6 gpurun_out/${T}_tests.log | cut -c1-300
timeout 900 python bench.py --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench default rc=$?"; tail -n 3 gpurun_out/${T}_bench_default.err
timeout 900 python bench.py --rows 12500000 --top-k 100 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/${T}_bench_top100.json 2> gpurun_out/${T}_bench_top100.err; echo "bench top-100 12.5M rc=$?"; tail -n 3 gpurun_out/${T}_bench_top100.err
timeout 900 python bench.py --rows 1250000 --steps 50 --warmup 10 --no-cpu-baseline > gpurun_out/${T}_bench_1p25m.json 2> gpurun_out/${T}_bench_1p25m.err; echo "bench 1.25M rc=$?"
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_bench_*.json")):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, "ERR", e); continue
    print(f.split("/")[-1], round(d["value"],1), round(d["ms_per_step"],4), {k:round(v,4) for k,v in d["step_breakdown_ms"]["rank0"].items()}, "e2e", round(d["e2e"]["value"],1), "oracle", d["oracle_check"].get("mismatches"), "frac", round(d["roofline"]["frac"],3), "amb", d["ambiguous_flags"], d["clocks"]["sm_mhz"])
    if "compressed_candidate_scan" in d: print("   compressed leg:", {k:v for k,v in d["compressed_candidate_scan"].items() if k!="note"})
PY
K='regex:dense_scan_kernel|dense_scan_q8_kernel|sparse_scan_kernel|leg_tail_kernel|fuse_kernel|exchange_kernel|merge_lists_kernel|rescore_|finalize_leg_kernel'
A="python bench.py --compressed --steps 3 --warmup 3 --no-cpu-baseline --no-oracle-check"
$A > gpurun_out/${T}_plainA.json 2> gpurun_out/${T}_plainA.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file gpurun_out/${T}_launchesA.csv $A > gpurun_out/${T}_ncuA.log 2>&1
echo "A rc=$?"
B="python bench.py --rows 12500000 --top-k 100 --steps 3 --warmup 3 --no-cpu-baseline --no-oracle-check"
$B > gpurun_out/${T}_plainB.json 2> gpurun_out/${T}_plainB.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file gpurun_out/${T}_launchesB.csv $B > gpurun_out/${T}_ncuB.log 2>&1
echo "B rc=$?"
python - <<PY
import csv,re,collections
for tag in "AB":
    try:
        rows=[r for r in csv.reader(open("gpurun_out/${T}_launches%s.csv"%tag)) if len(r)>10 and r[0].isdigit()]
    except Exception as e:
        print(tag, "ERR", e); continue
    agg=collections.defaultdict(list)
    for r in rows:
        v=float(r[-1].replace(",","")); u=r[-2]
        v = v/1e3 if u in ("ns","nsecond") else (v*1e3 if u in ("ms","msecond") else v)
        agg[re.sub(r"\(.*","",r[4])].append(v)
    print(tag, {k:(len(v), round(sum(v[-6:])/len(v[-6:]),1)) for k,v in agg.items()})
PY
