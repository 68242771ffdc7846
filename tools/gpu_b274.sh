timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/b_274.json 2> gpurun_out/b_274.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/b_274.json").read().strip().splitlines()[-1])
    print("274: value", d["value"], "comp", d["compress_MBps"], "decomp", d["decompress_MBps"], {k:v for k,v in d["stage_ms"].items() if v>10})
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/b_274.err").read()[-1500:])
PY
