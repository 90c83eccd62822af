"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.

    python tools/ncu_launches.py gpurun_out/launches.csv profiles/r01_launches.md "title" "command"
"""
import csv
import sys
from collections import OrderedDict


def main():
    src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    cmd = sys.argv[4] if len(sys.argv) > 4 else ''
    rows = [r for r in csv.reader(open(src)) if len(r) >= 15 and r[0].isdigit()]
    groups = OrderedDict()
    total = 0.0
    for r in rows:
        name, block, grid, ns = r[4], r[7], r[8], float(r[14])
        key = (name.split('(')[0][:60], grid, block)
        g = groups.setdefault(key, [0, 0.0])
        g[0] += 1
        g[1] += ns / 1e3
        total += ns / 1e3
    with open(out, 'w') as f:
        f.write('# %s\n\n' % title)
        if cmd:
            f.write('`%s`\n' % cmd)
        f.write('%d consecutive launches, %.1f us in total; cold-cache serialised times: compare SHARES.\n\n' % (len(rows), total))
        f.write('| kernel | grid | block | launches | total us | share | avg us |\n|---|---|---|---|---|---|---|\n')
        for (name, grid, block), (n, us) in sorted(groups.items(), key=lambda kv: -kv[1][1]):
            f.write('| `%s` | %s | %s | %d | %.1f | %.1f%% | %.1f |\n' % (name, grid, block, n, us, 100 * us / total, us / n))
    print('wrote', out)


if __name__ == '__main__':
    main()
