#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per kernel (raw page) and the hottest SASS lines (source page).
usage: ncu_summary.py report.ncu-rep [--src KERNEL_ID] [--top N]"""
import csv, io, subprocess, sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__cycles_active.avg',
        'smsp__inst_executed.sum', 'sm__inst_executed_pipe_lsu.sum', 'lts__t_bytes.sum', 'lts__t_sectors_op_read.sum',
        'lts__t_sectors_op_write.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum']


def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print(f"[{r[0]}] {r[idx['Kernel Name']][:70]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}")
        for w in WANT:
            if w in idx:
                print(f"     {w:80s} {r[idx[w]]:>16s} {units[idx[w]]}")


def src(rep, kid, top):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-id', f':::{kid}'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    si, sc, ie = hdr.index('# Samples'), hdr.index('Source'), hdr.index('Instructions Executed')
    stall = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
    data = [r for r in rows[2:] if len(r) >= len(hdr) and r[si].isdigit()]
    data = data[:len(data) // 2] if len(data) % 2 == 0 and data[:len(data) // 2] == data[len(data) // 2:] else data
    tot = sum(int(r[si]) for r in data)
    print("total samples", tot, "instructions", len(data))
    order = sorted(range(len(data)), key=lambda i: -int(data[i][si]))[:top]
    for i in sorted(order):
        r = data[i]
        st = sorted([(int(r[c]), hdr[c][6:]) for c in stall if r[c].isdigit() and int(r[c]) > 0], reverse=True)[:2]
        print(f"{i:5d} {int(r[si]):6d} {100*int(r[si])/max(tot,1):5.1f}% x{r[ie]:>8s}  {r[sc].strip()[:80]:80s} {st}")


if __name__ == '__main__':
    rep = sys.argv[1]
    top = int(sys.argv[sys.argv.index('--top') + 1]) if '--top' in sys.argv else 40
    if '--src' in sys.argv:
        src(rep, sys.argv[sys.argv.index('--src') + 1], top)
    else:
        raw(rep)
