#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest4.log
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/b_274.json 2> gpurun_out/b_274.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_274.json").read().strip().splitlines()[-1])
    print("274: value", d["value"], "comp", d["compress_MBps"], "decomp", d["decompress_MBps"], {k:v for k,v in d["stage_ms"].items() if v>20})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/b_274.err").read()[-1500:])
PY
timeout 1500 python tools/scale_check.py --config pe150 --mode records > gpurun_out/scale_pe150_records3.json 2> gpurun_out/scale_pe150_records3.err; echo "scale rc=$?"
cat gpurun_out/scale_pe150_records3.json; tail -3 gpurun_out/scale_pe150_records3.err
