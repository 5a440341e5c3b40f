"""Summarise an exported ncu report: key raw metrics + instruction/sample share per source line.
usage: python tools/ncu_summary.py raw.csv [src.csv] [top_n]"""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__registers_per_thread',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum', 'smsp__inst_executed.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_active.avg',
        'sm__cycles_elapsed.max', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'lts__t_bytes.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum']


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print("---", vals[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
        for i, h in enumerate(hdr):
            if h in WANT or ('issue_stalled' in h and h.endswith('per_issue_active.ratio')):
                try:
                    if 'issue_stalled' in h and float(vals[i]) < 0.3:
                        continue
                except ValueError:
                    pass
                print(f"  {h} [{units[i]}] = {vals[i]}")


def src(path, top):
    rows = list(csv.reader(open(path)))
    cur, agg = None, {}
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] in ('Function Name', 'Line No') or r[0] == '':
            continue
        try:
            agg[(cur, int(r[0]))] = (r[1][:100], int(r[7]), int(r[6]))
        except ValueError:
            pass
    tot = sum(v[1] for v in agg.values()) or 1
    ts = sum(v[2] for v in agg.values()) or 1
    print("total warp instructions", tot, "samples", ts)
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print(f"{k[0]}:{k[1]:<5d} {100 * v[1] / tot:5.1f}% inst {100 * v[2] / ts:5.1f}% samp  {v[0]}")


if __name__ == "__main__":
    raw(sys.argv[1])
    if len(sys.argv) > 2:
        src(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 30)
