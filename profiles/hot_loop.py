"""Hot-loop listing of one kernel from an `ncu --page source --csv --print-source sass` dump (several kernels may be
concatenated): instructions executed more than FRAC x the most executed one, with samples and top stalls.
usage: python profiles/hot_loop.py SRC.csv KERNEL_SUBSTRING [FRAC]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
want = sys.argv[2]; frac = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
blocks = []; cur = None
for r in rows:
    if r and r[0] == "Kernel Name": cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None: cur["rows"].append(r)
for b in blocks:
    if want not in b["name"]: continue
    hdr = b["rows"][0]; ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in b["rows"][1:] if len(r) == len(hdr)]
    num = lambda r, k: int(r[ix[k]] or 0)
    tot = sum(num(r, "Instructions Executed") for r in data); ts = sum(num(r, "# Samples") for r in data)
    stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg = sorted(((sum(num(r, h) for r in data), h[6:]) for h in stall), reverse=True)[:8]
    print(b["name"][:80]); print("warp-instructions", tot, "samples", ts, "stalls", [(k, round(100 * v / ts, 1)) for v, k in agg])
    mx = max(num(r, "Instructions Executed") for r in data)
    hot = [(i, r) for i, r in enumerate(data) if num(r, "Instructions Executed") > frac * mx]
    print("hot instructions", len(hot), "share of executed", round(sum(num(r, "Instructions Executed") for i, r in hot) / tot, 3),
          "share of samples", round(sum(num(r, "# Samples") for i, r in hot) / ts, 3))
    for i, r in hot:
        st = sorted(((num(r, h), h[6:]) for h in stall), reverse=True)[:2]
        print(f"{i:5d} {r[ix['Source']].strip()[:58]:58s} x{num(r, 'Instructions Executed'):9d} s{num(r, '# Samples'):6d} {st}")
