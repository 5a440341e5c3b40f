#!/bin/bash
# per-kernel device times (ncu launch list) of two small bench commands; each first exits 0 without ncu
set -u
mkdir -p gpurun_out
T=${1:-r02q}
K='regex:dense_scan_kernel|dense_gemm_kernel|sparse_scan_kernel|leg_tail_kernel|fuse_kernel|exchange_kernel|merge_lists_kernel|rescore_|finalize_leg_kernel'
A="python bench.py --steps 3 --warmup 3 --rows 1250000 --no-cpu-baseline --no-oracle-check"
$A > gpurun_out/${T}_plainA.json 2> gpurun_out/${T}_plainA.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file gpurun_out/${T}_launchesA.csv $A > gpurun_out/${T}_ncuA.log 2>&1
echo "A rc=$?"
B="python bench.py --steps 3 --warmup 3 --mode sparse --no-cpu-baseline --no-oracle-check"
$B > gpurun_out/${T}_plainB.json 2> gpurun_out/${T}_plainB.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 300 --csv --log-file gpurun_out/${T}_launchesB.csv $B > gpurun_out/${T}_ncuB.log 2>&1
echo "B rc=$?"
python - <<PY
import csv,re,collections
for tag in "AB":
    rows=[r for r in csv.reader(open("gpurun_out/${T}_launches%s.csv"%tag)) if len(r)>10 and r[0].isdigit()]
    agg=collections.defaultdict(list)
    for r in rows:
        v=float(r[-1].replace(",","")); u=r[-2]
        v = v/1e3 if u in ("ns","nsecond") else (v*1e3 if u in ("ms","msecond") else v)
        agg[re.sub(r"\(.*","",r[4])].append(v)
    print(tag, {k:(len(v), round(sum(v[-6:])/len(v[-6:]),1)) for k,v in agg.items()})
PY
