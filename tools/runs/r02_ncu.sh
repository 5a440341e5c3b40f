#!/bin/bash
# ncu evidence of round 2 (ONE gpurun call, 1 GPU).  Every profiled command first exits 0 without ncu (&& directly before).
set -u
mkdir -p gpurun_out
T=${1:-r02n}
K='regex:dense_scan_kernel|dense_gemm_kernel|sparse_scan_kernel|leg_tail_kernel|fuse_kernel|exchange_kernel|merge_lists_kernel|rescore_|finalize_leg_kernel|pool_select_kernel|set_filter_thr_kernel'
BASE="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-oracle-check"
$BASE > gpurun_out/${T}_plain.json 2> gpurun_out/${T}_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${T}_launches.csv $BASE > gpurun_out/${T}_ncu_launch.log 2>&1
echo "launch list rc=$?"
$BASE > gpurun_out/${T}_plain2.json 2> gpurun_out/${T}_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:dense_scan_kernel -s 4 -c 1 -f -o gpurun_out/${T}_dense_scan $BASE > gpurun_out/${T}_ncu_full_dense.log 2>&1
echo "dense_scan full rc=$?"
$BASE > gpurun_out/${T}_plain3.json 2> gpurun_out/${T}_plain3.err &&
ncu --set full --clock-control none --import-source on -k regex:sparse_scan_kernel -s 4 -c 1 -f -o gpurun_out/${T}_sparse_scan_b1 $BASE > gpurun_out/${T}_ncu_full_sparse.log 2>&1
echo "sparse_scan b1 full rc=$?"
G="python bench.py --batch 256 --mode dense --steps 2 --warmup 3 --no-cpu-baseline --no-oracle-check"
$G > gpurun_out/${T}_plain_b256.json 2> gpurun_out/${T}_plain_b256.err &&
ncu --set full --clock-control none --import-source on -k regex:dense_gemm_kernel -s 3 -c 1 -f -o gpurun_out/${T}_dense_gemm_pair_b256 $G > gpurun_out/${T}_ncu_full_gemm.log 2>&1
echo "gemm full rc=$?"
for n in dense_scan sparse_scan_b1 dense_gemm_pair_b256; do
  ncu -i gpurun_out/${T}_$n.ncu-rep --page raw --csv > gpurun_out/${T}_${n}_raw.csv 2>/dev/null
done
ncu -i gpurun_out/${T}_dense_gemm_pair_b256.ncu-rep --page source --csv > gpurun_out/${T}_dense_gemm_pair_b256_source.csv 2>/dev/null
ls -la gpurun_out/${T}_*
