#!/usr/bin/env python
"""Summarise an exported ncu report of ONE kernel launch: key raw metrics, stall-reason totals and the hottest SASS
instructions.   ncu -i X.ncu-rep --page raw --csv > raw.csv ; ncu -i X.ncu-rep --page source --csv --print-source sass > src.csv
    python tools/ncu_summary.py raw.csv src.csv [top_n]"""
import csv
import collections
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__m_xbar2l1tex_read_bytes.sum',
        'lts__t_sector_hit_rate.pct', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__cycles_active.avg', 'smsp__warps_eligible.avg.per_cycle_active',
        'smsp__warps_active.avg.per_cycle_active', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread',
        'launch__grid_size', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed']


def main():
    raw, src = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
    rows = list(csv.reader(open(raw)))
    hdr, vals = rows[0], rows[2]
    for h, u, v in zip(hdr, rows[1], vals):
        if h in WANT:
            print(f"{h:75s} {v} {u}")
    rows = list(csv.reader(open(src)))
    hdr = rows[1]
    iS, iE, iN = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('# Samples')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[2:] if len(r) > iE]
    tot = sum(int(r[iN] or 0) for r in data)
    print('samples', tot, 'warp instructions', sum(int(r[iE] or 0) for r in data), 'SASS lines', len(data))
    agg = collections.Counter()
    for r in data:
        for i in stall:
            agg[hdr[i]] += int(r[i] or 0)
    print('stalls:', ', '.join(f"{k[6:]} {v} ({100.0 * v / max(tot, 1):.0f}%)" for k, v in agg.most_common(10)))
    for idx, r in sorted(enumerate(data), key=lambda x: -int(x[1][iN] or 0))[:top]:
        st = {hdr[i][6:]: int(r[i] or 0) for i in stall if int(r[i] or 0) > 0}
        print(idx, r[iE], r[iN], r[iS][:64], st)


if __name__ == '__main__':
    main()
