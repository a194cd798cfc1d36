"""profiles/r2_traffic.json (read by bench.py for `roofline.traffic`): measured DRAM bytes per launch of the sparse-conv kernels
from the per-kernel ncu metric capture of one C2 training step (profiles/capture_r2.sh, part B)."""
import collections
import csv
import json
import sys

path = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/r2_step_metrics.csv'
rows = collections.defaultdict(dict)
with open(path) as f:
    for row in csv.DictReader(l for l in f if not l.startswith('==')):
        try:
            v = float(row['Metric Value'].replace(',', ''))
        except ValueError:
            continue
        unit = {'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}.get(row['Metric Unit'], 1.0)
        rows[(row['ID'], row['Kernel Name'])][row['Metric Name']] = v * unit
out = {}
for fam in ('spconv_tc_fwd_kernel', 'spconv_tc_wgrad_kernel', 'conv_tma_kernel', 'conv_tma_wgrad_kernel'):
    b = [m.get('dram__bytes_read.sum', 0.0) + m.get('dram__bytes_write.sum', 0.0) for (i, n), m in rows.items()
         if fam + '<' in n or fam + '(' in n]
    if b:
        out[fam] = {'dram_bytes_per_launch_avg': sum(b) / len(b), 'launches': len(b),
                    'source': f'ncu dram__bytes_read.sum + dram__bytes_write.sum over the {len(b)} launches of one C2 training step '
                              '(4 scans), profiles/r2_step_kernel_metrics.txt'}
json.dump(out, open('profiles/r2_traffic.json', 'w'), indent=1)
print(json.dumps(out, indent=1))
