"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (shares of the step)."""
import collections
import csv
import re
import sys


def main(path, top=40):
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        v = float(row['Metric Value'].replace(',', ''))
        v = v / 1e3 if row['Metric Unit'] == 'ns' else v * 1e3 if row['Metric Unit'] == 'ms' else v
        name = row['Kernel Name'].replace('void ', '').replace('<unnamed>::', '').replace('(anonymous namespace)::', '')
        name = re.sub(r'\(.*$', '', name)
        name = re.sub(r'\(int\)|\(bool\)', '', name)[:90]
        tot[name] += v
        cnt[name] += 1
    T = sum(tot.values())
    ours = sum(v for k, v in tot.items() if re.match(r'(spconv|norm_|seg_|paint|focal|topk|assign|count_|best_level|hash_|'
                                                      r'kernel_map|pair_|flag_|compact_|inverse_|voxelize|generative|tile_mask|'
                                                      r'maxpool|img_normalize|adamw|sumsq|clip_coef|cast_|segmented_nms|fill_kernel|'
                                                      r'unproject|depth_flag)', k))
    print(f'# {path}: {sum(cnt.values())} launches, {T / 1e3:.2f} ms total (serialised, cold cache: compare SHARES)')
    print(f'# esb200 hand-written kernels: {100 * ours / T:.1f}% of GPU time')
    print(f'{"us":>10} {"share":>6} {"n":>5}  kernel')
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
        print(f'{v:10.0f} {100 * v / T:5.1f}% {cnt[k]:5d}  {k}')


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
