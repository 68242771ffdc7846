#!/bin/bash
# one GPU session: parity suite, then decode / encode timings (device-resident, no CPU arm)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest.log
tail -5 gpurun_out/pytest.log
for b in ${BLOCKS:-1 64 274}; do
  timeout 600 python bench.py --blocks $b --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/b_$b.json 2> gpurun_out/b_$b.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_$b.json").read().strip().splitlines()[-1])
    print("blocks $b: value", d["value"], "comp", d["compress_MBps"], "decomp", d["decompress_MBps"], {k:v for k,v in d["stage_ms"].items() if v>50})
except Exception as e:
    print("blocks $b failed", e); print(open("gpurun_out/b_$b.err").read()[-1500:])
PY
done
if [ -f _variants/timing.so ]; then
  cp fastqueeze_b200/libseqarc_b200.so /tmp/keep.so; cp _variants/timing.so fastqueeze_b200/libseqarc_b200.so
  timeout 300 python tools/prof_run.py --blocks 1 --pairs 8000 --decode 2>&1 | grep -E "dec warp|decode" | tee gpurun_out/timing.log
  cp /tmp/keep.so fastqueeze_b200/libseqarc_b200.so
fi
