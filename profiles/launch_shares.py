"""Kernel shares from an `ncu --metrics gpu__time_duration.sum --csv` launch list.  usage: launch_shares.py LIST.csv"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
start = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[start]; ix = {h: i for i, h in enumerate(hdr)}
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) < len(hdr): continue
    try: v = float(r[ix['Metric Value']].replace(',', ''))
    except ValueError: continue
    k = r[ix['Kernel Name']].split('(')[0][-60:]
    agg[k][0] += 1; agg[k][1] += v
tot = sum(v[1] for v in agg.values())
print("kernel,launches,total_ms,share_pct")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"\"{k}\",{n},{t / 1e6:.3f},{100 * t / tot:.1f}")
