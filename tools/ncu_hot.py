"""Print the hot SASS lines (stall samples) of an `ncu --page source --csv` export."""
import csv
import sys

f = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.012
rows = list(csv.reader(open(f)))
# the export may hold several kernel instances: take the first
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
lo = starts[which]
hi = starts[which + 1] if which + 1 < len(starts) else len(rows)
print(rows[lo][1][:100])
h = rows[lo + 1]
si, ss, ie, te = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
sc = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
I = lambda x: int(x) if x.isdigit() else 0
body = [r for r in rows[lo + 2:hi] if len(r) > max(ss, ie)]
tot = sum(I(r[ss]) for r in body)
print("total samples", tot, "instr lines", len(body))
agg = {}
for r in body:
    for i in sc:
        agg[h[i][6:]] = agg.get(h[i][6:], 0) + I(r[i])
print("stall mix:", ", ".join("%s %.1f%%" % (k, 100 * v / max(1, sum(agg.values()))) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for n, r in enumerate(body):
    s = I(r[ss])
    if s > tot * thr:
        top = sorted([(I(r[i]), h[i][6:]) for i in sc if I(r[i]) > 0], reverse=True)[:2]
        print("%5d %7d %5.1f%% %10s %11s  %-56s %s" % (n, s, 100 * s / tot, r[ie], r[te], r[si].strip()[:56], top))
