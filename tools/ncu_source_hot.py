"""Where the warp-stall samples of one kernel fall: ncu -i X.ncu-rep --page source --csv --kernel-name K | python tools/ncu_source_hot.py [window]"""
import csv, sys
win = int(sys.argv[1]) if len(sys.argv) > 1 else 250
rows = list(csv.reader(sys.stdin))
hdr = next(r for r in rows if r and r[0] == 'Address')
data = [r for r in rows if r and r[0].startswith('0x') and len(r) == len(hdr)]
iS, iSrc = hdr.index('# Samples'), hdr.index('Source')
stall_cols = [i for i, h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot = sum(int(r[iS]) for r in data)
print('total samples', tot, 'instructions', len(data))
for w in range(0, len(data), win):
    chunk = data[w:w + win]
    s = sum(int(r[iS]) for r in chunk)
    agg, ops = {}, {}
    for r in chunk:
        for c in stall_cols:
            agg[hdr[c]] = agg.get(hdr[c], 0) + int(r[c])
        t = r[iSrc].split()
        op = t[1] if t[0].startswith('@') else t[0]
        ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + 1
    print(w, f'{100 * s / max(tot, 1):5.1f}%', sorted(agg.items(), key=lambda kv: -kv[1])[:3], sorted(ops.items(), key=lambda kv: -kv[1])[:5])
