"""Derive the tracked profiles/ summaries from the ncu captures a gpurun call left in gpurun_out/.

    python tools/make_profiles.py r01f r02      # capture prefix in gpurun_out/, name prefix in profiles/

Inputs (gpurun_out/<cap>_*): launches.csv (--metrics gpu__time_duration.sum), *_raw.csv (ncu -i x.ncu-rep --page raw --csv).
Outputs (profiles/<name>_*): launch list (timed steps only), per-kernel step breakdown, raw pages, key metrics,
roofline_traffic.json (read by bench.py into roofline.traffic)."""
import csv
import json
import os
import re
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__bytes_read.sum.per_second',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'launch__occupancy_limit_shared_mem',
        'launch__occupancy_limit_registers', 'launch__grid_size', 'launch__block_size', 'launch__cluster_size',
        'launch__shared_mem_per_block_dynamic', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'sm__cycles_elapsed.max']


def key_metrics(raw_csv):
    rows = list(csv.reader(open(raw_csv)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")]}
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            out[k] = [vals[i], units[i]]
    return out


def launches(cap, name, steps):
    src = os.path.join(G, f"{cap}_launches.csv")
    rows = [r for r in csv.reader(open(src)) if len(r) > 10 and r[0].isdigit()]
    recs = [(r[4], float(r[-1].replace(",", "")), r[-2]) for r in rows if r[12] == "gpu__time_duration.sum"]
    # unit normalisation -> us
    norm = []
    for k, v, u in recs:
        v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
        norm.append((re.sub(r"\(.*", "", k), v))
    # the timed steps are the last `steps` occurrences of the fuse kernel
    fuse_idx = [i for i, (k, _) in enumerate(norm) if "fuse_kernel" in k]
    if not fuse_idx:
        print("launch list holds no search-path kernels (raise -c or filter with -k)")
        return {}, 0.0
    start = fuse_idx[-steps - 1] + 1 if len(fuse_idx) > steps else 0
    timed = norm[start:fuse_idx[-1] + 1]
    with open(os.path.join(P, f"{name}_launches_bench_n1.csv"), "w") as f:
        f.write("kernel,duration_us\n")
        for k, v in timed:
            f.write(f"\"{k}\",{v:.3f}\n")
    per = {}
    for k, v in timed:
        per[k] = per.get(k, 0.0) + v / steps
    tot = sum(per.values())
    json.dump({"cmd": "python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-oracle-check", "steps_summarised": steps,
               "us_per_step": {k: round(v, 2) for k, v in per.items()},
               "share": {k: round(v / tot, 4) for k, v in per.items()},
               "total_us_per_step_serialised": round(tot, 1)},
              open(os.path.join(P, f"{name}_step_breakdown_n1.json"), "w"), indent=1)
    return per, tot


def main():
    cap, name = sys.argv[1], sys.argv[2]
    os.makedirs(P, exist_ok=True)
    km = {}
    for f in sorted(os.listdir(G)):
        m = re.match(rf"{cap}_(.*)_raw\.csv$", f)
        if m:
            shutil.copy(os.path.join(G, f), os.path.join(P, f"{name}_{m.group(1)}_ncu_full_raw.csv"))
            km[m.group(1)] = key_metrics(os.path.join(G, f))
    json.dump(km, open(os.path.join(P, f"{name}_ncu_key_metrics.json"), "w"), indent=1)
    if os.path.exists(os.path.join(G, f"{cap}_launches.csv")):
        per, tot = launches(cap, name, 3)
        print("step breakdown (us):", {k: round(v, 1) for k, v in per.items()}, "total", round(tot, 1))
    if "dense_scan" in km:
        def gb(x):
            v, u = x
            v = float(v)
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        t = int(gb(km["dense_scan"]["dram__bytes_read.sum"]) + gb(km["dense_scan"]["dram__bytes_write.sum"]))
        json.dump({"dense_scan_rows_10000000": t,
                   "source": f"profiles/{name}_dense_scan_ncu_full_raw.csv: dram__bytes_read.sum + dram__bytes_write.sum, "
                             f"one launch of {km['dense_scan']['kernel']}, N=1, 10M rows (algorithmic bytes 20 480 000 000)"},
                  open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
    for k, v in km.items():
        print(k, v["kernel"], v.get("gpu__time_duration.sum"), v.get("dram__bytes_read.sum"),
              v.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"))


if __name__ == "__main__":
    main()
