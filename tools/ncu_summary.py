#!/usr/bin/env python3
"""Summarise an .ncu-rep (read here, no GPU): key counters + top stall sites.  Usage: ncu_summary.py rep [out.md]"""
import csv
import io
import subprocess
import sys
from collections import Counter

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'lts__t_bytes.sum', 'l1tex__t_bytes.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu.sum',
        'sm__warps_active.avg.per_cycle_active', 'sm__inst_executed_pipe_lsu.sum',
        'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_ltcfabric.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed']


def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    out = []
    rows = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'raw', '--csv']))))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        out.append('## ' + r[hdr.index('Kernel Name')][:100])
        for k in KEYS:
            if k in hdr:
                out.append(f'- {k}: {r[hdr.index(k)]} {units[hdr.index(k)]}')
    src = list(csv.reader(io.StringIO(run(['ncu', '-i', rep, '--page', 'source', '--csv']))))
    hi = next(i for i, r in enumerate(src) if 'Source' in r and 'Address' in r)
    h = src[hi]
    ix = {k: i for i, k in enumerate(h)}
    stalls = [k for k in h if k.startswith('stall_') and 'Not Issued' not in k]
    tot = Counter()
    byop = Counter()
    n = 0
    for r in src[hi + 1:]:
        if len(r) < len(h):
            continue
        try:
            s = float(r[ix['# Samples']])
        except ValueError:
            continue
        n += s
        op = r[ix['Source']].split()
        op = op[1] if op and op[0].startswith('@') and len(op) > 1 else (op[0] if op else '')
        byop[op.split('.')[0]] += s
        for k in stalls:
            try:
                tot[k] += float(r[ix[k]])
            except ValueError:
                pass
    out.append('### stall reasons (share of samples, first kernel instance in report order)')
    st = sum(tot.values()) or 1
    for k, v in tot.most_common(8):
        out.append(f'- {k}: {100 * v / st:.1f}%')
    out.append('### samples by opcode')
    for k, v in byop.most_common(10):
        out.append(f'- {k}: {100 * v / max(n, 1):.1f}%')
    text = '\n'.join(out)
    print(text)
    if len(sys.argv) > 2:
        open(sys.argv[2], 'w').write(text + '\n')


if __name__ == '__main__':
    main()
