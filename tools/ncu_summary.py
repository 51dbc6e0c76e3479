#!/usr/bin/env python
"""Summarise an .ncu-rep: key metrics per captured launch and (with --source) the hottest source lines.
Usage: python tools/ncu_summary.py file.ncu-rep [--source N]"""
import csv, subprocess, sys, io, collections
KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'smsp__inst_executed.sum',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__cycles_elapsed.max', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__waves_per_multiprocessor']
def raw(rep):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h, u = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(h, r))
        print('==', d.get('Kernel Name', '')[:90])
        for name, unit, val in zip(h, u, r):
            if name in KEYS or 'issue_stalled' in name and name.endswith('per_issue_active.ratio') and float(val or 0) > 0.5:
                print(f'   {name:95s} {val:>16s} {unit}')
def source(rep, n):
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    fname = kname = ''; h = None; per = {}
    for r in rows:
        if not r: continue
        if r[0] == 'File Path': fname = r[1].split('/')[-1]; continue
        if r[0] == 'Function Name': kname = r[1][:40]; continue
        if r[0] == 'Line No': h = r; continue
        if h is None or r[0] == '' or len(r) < len(h): continue
        d = dict(zip(h[2:], r[2:]))
        try:
            key = (kname, fname, int(r[0]), r[1].strip()[:100])
            v = per.setdefault(key, [0.0, 0.0])
            v[0] += float(d.get('Instructions Executed', 0) or 0); v[1] += float(d.get('# Samples', 0) or 0)
        except ValueError:
            continue
    for kn in sorted({k[0] for k in per}):
        items = [(v, k) for k, v in per.items() if k[0] == kn]
        ti = sum(v[0] for v, _ in items); ts = sum(v[1] for v, _ in items)
        print(f'-- {kn}: {ti:.0f} warp instructions, {ts:.0f} samples')
        for v, k in sorted(items, key=lambda x: -x[0][0])[:n]:
            print(f'{100*v[0]/max(ti,1):5.1f}% inst {100*v[1]/max(ts,1):5.1f}% smp  {k[1]}:{k[2]:<5d} {k[3]}')

if __name__ == '__main__':
    raw(sys.argv[1])
    if '--source' in sys.argv:
        source(sys.argv[1], int(sys.argv[sys.argv.index('--source') + 1]))
