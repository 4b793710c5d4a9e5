"""Summarise an `ncu --csv` launch list (gpu__time_duration + dram bytes): per-kernel totals of ONE step of bench.py.
Usage: python tools/launch_summary.py <csv> [first_kernel_substring=stem_s2d] [out.json batch]"""
import csv, collections, glob, hashlib, json, os, sys


def _csrc_sha1():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = hashlib.sha1()
    for f in sorted(glob.glob(os.path.join(root, 'cvpytorch_b200', 'csrc', '*'))):
        if f.endswith(('.cu', '.cuh', '.h')) and not os.path.basename(f).startswith('train_'):  # (the training kernels are not on the measured inference path)
            h.update(open(f, 'rb').read())
    return h.hexdigest()[:12]


path = sys.argv[1]
anchor = sys.argv[2] if len(sys.argv) > 2 else 'stem_s2d'
with open(path) as f:
    lines = [l for l in f if not l.startswith('==')]
per = collections.OrderedDict()
for x in csv.DictReader(lines):
    k = (int(x['ID']), x['Kernel Name'])
    per.setdefault(k, {})[x['Metric Name']] = (float(x['Metric Value'].replace(',', '')), x['Metric Unit'])
ids = sorted(per)
SC = {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9, 'ns': 1e-3, 'us': 1, 'ms': 1e3, 'nsecond': 1e-3, 'usecond': 1, 'msecond': 1e3}
def val(d, m):
    if m not in d: return 0.0
    v, u = d[m]
    return v * SC[u]
starts = [i for i, k in enumerate(ids) if anchor in k[1]]
if len(starts) < 2:
    sys.exit(f'need two "{anchor}" launches to delimit a step, found {len(starts)}')
lo, hi = starts[-2], starts[-1]
agg = collections.OrderedDict(); tot = 0; conv_b = 0; conv_t = 0
for k in ids[lo:hi]:
    d = per[k]; n = k[1].split('(')[0][:58]
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += val(d, 'gpu__time_duration.sum'); a[2] += val(d, 'dram__bytes_read.sum'); a[3] += val(d, 'dram__bytes_write.sum')
print(f'one step = launches {lo}..{hi - 1} ({hi - lo} kernels)')
for n, a in agg.items():
    print(f'{n:58s} n={a[0]:3d} t={a[1]:8.1f} us  rd={a[2] / 1e6:8.1f} MB  wr={a[3] / 1e6:8.1f} MB')
    tot += a[1]
    if 'conv_tc' in n: conv_b += a[2] + a[3]; conv_t += a[1]
print(f'total {tot:.1f} us; conv_tc: {conv_t:.1f} us, DRAM traffic {conv_b / 1e9:.3f} GB per step')
if len(sys.argv) > 4:
    json.dump({'batch': int(sys.argv[4]), 'conv_dram_bytes_per_step': int(conv_b), 'conv_us_per_step_under_ncu': round(conv_t, 1),
               'step_us_under_ncu': round(tot, 1), 'kernels': {n: {'launches': a[0], 'us': round(a[1], 1), 'dram_read_MB': round(a[2] / 1e6, 1),
                                                                    'dram_write_MB': round(a[3] / 1e6, 1)} for n, a in agg.items()},
               'source': path.split('/')[-1], 'source_sha1': _csrc_sha1(), 'how': 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none; one step of bench.py'},
              open(sys.argv[3], 'w'), indent=1)
