"""Top stall locations of the first kernel in an `ncu --page source --csv --print-source sass` dump."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
seg = rows[starts[0]: starts[1] if len(starts) > 1 else len(rows)]
hdr = seg[1]
ix = {h: i for i, h in enumerate(hdr)}
data = [r for r in seg[2:] if len(r) == len(hdr)]
S = lambda r, h: int(float(r[ix[h]] or 0))
tot = sum(S(r, "# Samples") for r in data)
print("total samples", tot, "instructions", len(data))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {h: sum(S(r, h) for r in data) for h in stalls}
print("by reason:", sorted(((v, k) for k, v in agg.items() if v), reverse=True)[:8])
for r in sorted(data, key=lambda r: -S(r, "# Samples"))[:n]:
    s = S(r, "# Samples")
    st = sorted(((S(r, h), h[6:]) for h in stalls), reverse=True)[:2]
    print("%6d %5.1f%% %s %-80s %s" % (s, 100.0 * s / max(tot, 1), r[ix["Address"]][-5:], r[ix["Source"]][:80], st))
