#!/bin/bash
# decode / encode time against the number of co-resident blocks
mkdir -p gpurun_out
for b in ${BLOCKS:-74 148 200 274}; do
  timeout 600 python bench.py --blocks $b --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/b_$b.json 2> gpurun_out/b_$b.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$b.json").read().strip().splitlines()[-1])
    print("blocks $b: value", d["value"], "comp", d["compress_MBps"], "decomp", d["decompress_MBps"], {k:v for k,v in d["stage_ms"].items() if v>100})
except Exception as e:
    print("blocks $b failed", e); print(open("gpurun_out/b_$b.err").read()[-1500:])
PY
done
