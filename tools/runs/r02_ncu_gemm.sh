#!/bin/bash
set -u
mkdir -p gpurun_out
T=${1:-r02u}
G="python bench.py --batch 256 --mode dense --steps 2 --warmup 3 --no-cpu-baseline --no-oracle-check"
$G > gpurun_out/${T}_plain_b256.json 2> gpurun_out/${T}_plain_b256.err &&
ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 3 -c 1 -f -o gpurun_out/${T}_dense_gemm_pair_b256 $G > gpurun_out/${T}_ncu_full_gemm.log 2>&1
echo "gemm full rc=$?"
ncu -i gpurun_out/${T}_dense_gemm_pair_b256.ncu-rep --page raw --csv > gpurun_out/${T}_dense_gemm_pair_b256_raw.csv 2>/dev/null
ncu -i gpurun_out/${T}_dense_gemm_pair_b256.ncu-rep --page source --csv > gpurun_out/${T}_dense_gemm_pair_b256_source.csv 2>/dev/null
python - <<PY
import csv
rows=list(csv.reader(open("gpurun_out/${T}_dense_gemm_pair_b256_raw.csv")))
hdr,units,vals=rows[0],rows[1],rows[2]
for k in ("gpu__time_duration.sum","sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active","sm__cycles_elapsed.max","dram__bytes_read.sum","smsp__issue_active.avg.pct_of_peak_sustained_active","sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active","l1tex__m_xbar2l1tex_read_bytes.sum"):
    if k in hdr: print(k, vals[hdr.index(k)], units[hdr.index(k)])
PY
