"""Summarise an `ncu -i <rep> --page source --csv` export of k_dec_main: stall mix of the kernel, and for every hot loop
(instructions grouped by their execution count) the instructions per iteration, sampled cycles per iteration and stall mix.

usage: ncu -i prof.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv [symbols-per-role]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    I = lambda r, k: int(r[ix[k]]) if r[ix[k]].isdigit() else 0
    print(rows[0][1])
    tot = sum(I(r, "# Samples") for r in data)
    mix = collections.Counter()
    for r in data:
        for s in stalls:
            mix[s[6:]] += I(r, s)
    print("samples %d; stall mix: %s" % (tot, ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(mix.values()))) for k, v in mix.most_common(8))))
    # loops: bucket instructions by execution count (the per-symbol loops execute millions of times)
    by = collections.defaultdict(list)
    for i, r in enumerate(data):
        ex = I(r, "Instructions Executed")
        if ex > 200000:
            by[ex].append(i)
    groups = sorted(by.items(), key=lambda kv: -len(kv[1]))[:6]
    sel_per_instr = None
    for ex, idx in sorted(groups, key=lambda kv: -kv[0]):
        s = sum(I(data[i], "# Samples") for i in idx)
        m = collections.Counter()
        for i in idx:
            for st in stalls:
                m[st[6:]] += I(data[i], st)
        issue = m["selected"]
        if not issue:
            continue
        # `selected` samples = issue cycles: one per instruction per iteration
        cyc = s / issue * len(idx)
        print("\nloop executed %d times: %d instructions on the common path, %.0f sampled cycles per iteration (lines %d..%d)"
              % (ex, len(idx), cyc, idx[0], idx[-1]))
        print("  stall mix: " + ", ".join("%s %.1f%%" % (k, 100.0 * v / max(1, sum(m.values()))) for k, v in m.most_common(7)))
        hot = sorted(idx, key=lambda i: -I(data[i], "# Samples"))[:10]
        per = issue / len(idx)
        for i in sorted(hot):
            r = data[i]
            top = sorted(((I(r, st), st[6:]) for st in stalls if I(r, st) > 0), reverse=True)[:2]
            print("  %6d %5.1f cyc  %-58s %s" % (i, I(r, "# Samples") / per, r[ix["Source"]].strip()[:58], top))


if __name__ == "__main__":
    main()
