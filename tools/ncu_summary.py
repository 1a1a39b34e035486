#!/usr/bin/env python
"""Summarise an .ncu-rep (run where ncu is installed, no GPU needed): headline counters, stall reasons and the
instructions with the most stall samples.  Usage: python tools/ncu_summary.py report.ncu-rep [top_n]"""
import csv
import io
import subprocess
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'dram__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum',
        'smsp__inst_executed_op_shared_atom.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.avg', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'local_load', 'smsp__inst_executed_op_local_ld.sum',
        'smsp__inst_executed_op_local_st.sum']


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    rows = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print('=== kernel:', r[hdr.index('Kernel Name')])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:82s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = []
        for i, k in enumerate(hdr):
            if 'pcsamp_warps_issue_stalled' in k and 'not_issued' not in k:
                try:
                    st.append((float(r[i]), k.replace('smsp__pcsamp_warps_issue_stalled_', '')))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1
        print('  stall samples:', ', '.join(f"{k}={v / tot * 100:.1f}%" for v, k in sorted(st, reverse=True)[:9]))
        break
    rows = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'source', '--csv']))))
    hdr = rows[1]
    isrc, iall, iex = hdr.index('Source'), hdr.index('Warp Stall Sampling (All Samples)'), hdr.index('Instructions Executed')
    data = []
    for r in rows[2:]:
        if r and r[0] == 'Kernel Name' and data:
            break
        if len(r) < len(hdr) or r[0] in ('Kernel Name', 'Address'):
            continue
        data.append(r)
    tot = sum(int(r[iall]) for r in data) or 1
    print(f"  source: {len(data)} SASS instructions, {tot} stall samples; top {top_n}:")
    top = sorted(range(len(data)), key=lambda i: -int(data[i][iall]))[:top_n]
    for i in sorted(top):
        r = data[i]
        print(f"  {i:5d} {int(r[iall]) / tot * 100:5.1f}% {int(r[iex]):10d}  {r[isrc].strip()[:95]}")


if __name__ == '__main__':
    main()
