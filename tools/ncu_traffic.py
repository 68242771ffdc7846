"""Turn an `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,... --csv` log of the bench command into
profiles/traffic.json (the per-launch DRAM bytes bench.py quotes in `roofline.traffic`) and a short text summary.

usage: python tools/ncu_traffic.py <ncu.csv> <kernel regex> <stage> <config> <job_blocks> <world> <usemodel> <source label>
"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    path, kre, stage, config, blocks, world, usemodel, label = sys.argv[1:9]
    rows = [r for r in csv.reader(open(path)) if len(r) > 12 and r[0].isdigit()]
    per = collections.OrderedDict()
    for r in rows:
        if not re.search(kre, r[4]):
            continue
        per.setdefault(r[0], {})[r[-3]] = float(r[-1].replace(",", ""))
    if not per:
        raise SystemExit("no launch of %s in %s" % (kre, path))
    last = list(per.values())[-1]                    # the timed step's launch (the first is the warm-up step)
    traffic = int(last.get("dram__bytes_read.sum", 0) + last.get("dram__bytes_write.sum", 0))
    out = os.path.join(ROOT, "profiles", "traffic.json")
    t = json.load(open(out)) if os.path.exists(out) else {}
    t["%s:%s:%s:%s:%s" % (stage, config, blocks, world, usemodel)] = {"bytes": traffic, "source": label, "metrics": last}
    json.dump(t, open(out, "w"), indent=1, sort_keys=True)
    print(json.dumps({"launches": len(per), "traffic_bytes": traffic, "metrics": last}))


if __name__ == "__main__":
    main()
