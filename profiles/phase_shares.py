"""Per-phase shares of a kernel from an `ncu --page source --csv --print-source sass` dump: the SASS is cut at the
barriers that were executed; for every phase the share of executed warp-instructions, of the PC samples, and the
top stall reasons.  usage: python profiles/phase_shares.py SRC.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
num = lambda r, k: int(r[ix[k]] or 0)
tot = sum(num(r, 'Instructions Executed') for r in data)
ts = sum(num(r, '# Samples') for r in data)
print(f"\n# phases between executed barriers: {tot} warp-instructions, {ts} PC samples, {len(data)} SASS instructions")
bars = [k for k, r in enumerate(data) if 'BAR' in r[ix['Source']] and num(r, 'Instructions Executed') > 0]
stall = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
bounds = [0] + [b + 1 for b in bars] + [len(data)]
for a, b in zip(bounds[:-1], bounds[1:]):
    seg = data[a:b]
    n = sum(num(r, 'Instructions Executed') for r in seg); smp = sum(num(r, '# Samples') for r in seg)
    if smp < ts * 0.01 and n < tot * 0.01:
        continue
    st = sorted(((h[6:], sum(num(r, h) for r in seg)) for h in stall), key=lambda t: -t[1])[:4]
    print(f"sass [{a:5d},{b:5d})  inst {100 * n / tot:5.1f} %  samples {100 * smp / ts:5.1f} %   top stalls: "
          + ", ".join(f"{k} {100 * v // max(smp, 1)} %" for k, v in st))
ops = {}
for r in data:
    s = r[ix['Source']].split()
    if not s: continue
    op = s[1] if s[0].startswith('@') and len(s) > 1 else s[0]
    ops[op.split('.')[0]] = ops.get(op.split('.')[0], 0) + num(r, 'Instructions Executed')
print("# executed warp-instructions by opcode: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in sorted(ops.items(), key=lambda t: -t[1])[:16]))
