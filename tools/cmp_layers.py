"""Side-by-side per-layer times of several tools/layer_times.py outputs:  python tools/cmp_layers.py a.txt b.txt ..."""
import re
import sys

tabs = []
for f in sys.argv[1:]:
    rows = []
    for line in open(f):
        m = re.match(r'^(\S+)\s+(\d+->\s*\d+ k\d s\d\s+\d+x\d+)\s+([0-9.]+) ms', line)
        if m:
            rows.append((m.group(1), m.group(2), float(m.group(3))))
        elif line.startswith('<aux step>'):
            rows.append(('<aux>', '', float(line.split()[-2])))
    tabs.append(rows)
n = min(len(t) for t in tabs)
tot = [0.0] * len(tabs)
conv = [0.0] * len(tabs)
for i in range(n):
    name, shape = tabs[0][i][0], tabs[0][i][1]
    vals = [t[i][2] for t in tabs]
    for j, v in enumerate(vals):
        tot[j] += v
        if name != '<aux>':
            conv[j] += v
    print(f'{name:34s} {shape:28s} ' + ' '.join(f'{v:8.4f}' for v in vals) + ('   %+6.1f%%' % ((vals[-1] / vals[0] - 1) * 100) if len(vals) > 1 else ''))
print(f'{"total":63s} ' + ' '.join(f'{v:8.4f}' for v in tot))
print(f'{"conv only":63s} ' + ' '.join(f'{v:8.4f}' for v in conv))
