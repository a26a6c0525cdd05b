"""Compact table of the key per-kernel metrics from an `ncu --csv --page raw` log."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fmaheavy_cycles_active.avg.pct',
        'sm__inst_executed_pipe_alu.avg.pct', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct', 'smsp__average_warps_issue_stalled', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct']
cols = [j for j, h in enumerate(hdr) if any(w in h for w in want)]
ks = [r for r in rows[hi + 2:] if len(r) == len(hdr)]
names = [r[hdr.index('Kernel Name')].split('(')[0][-14:] for r in ks]
print(' ' * 60, *[f'{n:>14s}' for n in names])
for j in cols:
    vals = [r[j] for r in ks]
    try:
        if all(float(v.replace(',', '')) == 0 for v in vals):
            continue
    except ValueError:
        pass
    print(f'{hdr[j][-60:]:60s}', *[f'{v:>14s}' for v in vals])
