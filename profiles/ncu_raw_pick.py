"""Pick the headline metrics out of `ncu -i X.ncu-rep --page raw --csv` (stdin) into one line per profiled launch."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
rows = list(csv.reader(l for l in sys.stdin if not l.startswith('==')))
if len(rows) < 3:
    sys.exit('no rows')
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
name_i = idx.get('Kernel Name', 4)
print('# one line per launch: kernel | ' + ' | '.join(w.split('.')[0].replace('__', ':') for w in WANT if w in idx))
for r in rows[2:]:
    vals = []
    for w in WANT:
        if w in idx:
            vals.append(f'{r[idx[w]]} {units[idx[w]]}'.strip())
    print(r[name_i][:70].replace('void (anonymous namespace)::', ''), '|', ' | '.join(vals))
