"""Per-kernel totals of an ncu launch list: python tools/launch_list_summary.py gpurun_out/launches.csv  (csv of --metrics gpu__time_duration.sum)"""
import csv, collections, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
iK, iV, iU = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    v = float(r[iV].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[iU][:2], 1.0)
    a = agg.setdefault(r[iK][:64], [0, 0.0, 1e18, 0.0])
    a[0] += 1; a[1] += v; a[2] = min(a[2], v); a[3] = max(a[3], v)
tot = sum(a[1] for a in agg.values())
print(f'| kernel | launches | total ms | share | avg us | min us | max us |\n|---|---:|---:|---:|---:|---:|---:|')
for k, a in agg.items():
    print(f'| `{k}` | {a[0]} | {a[1] / 1e3:.3f} | {100 * a[1] / tot:.1f} % | {a[1] / a[0]:.1f} | {a[2]:.1f} | {a[3]:.1f} |')
print(f'| total | {sum(a[0] for a in agg.values())} | {tot / 1e3:.3f} | | | | |')
