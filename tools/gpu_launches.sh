#!/bin/bash
# launch list (device time of every kernel launch) of a short bench run
mkdir -p gpurun_out
B=${1:-64}
python bench.py --blocks $B --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/plain_launch.json 2> gpurun_out/plain_launch.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 2000 --csv --log-file gpurun_out/launches.csv \
    python bench.py --blocks $B --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_launch.log 2>&1
echo "rc=$?"
python - <<'PY'
import csv, collections
rows=[r for r in csv.reader(open("gpurun_out/launches.csv")) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict()
for r in rows:
    name=r[4].split("(")[0]; v=float(r[-1].replace(",","")); unit=r[-2]
    ms = v/1e6 if unit in ("ns","nsecond") else (v/1e3 if unit in ("us","usecond") else v)
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms
tot=sum(a[1] for a in agg.values())
for k,(n,ms) in sorted(agg.items(), key=lambda kv:-kv[1][1]):
    print("%-40s %5d %10.2f ms %5.1f%%" % (k[:40], n, ms, 100*ms/tot))
PY
