python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 2>&1 | tail -1
for b in 1 148; do timeout 300 python bench.py --blocks $b --steps 1 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d[\"value\"], d[\"compress_MBps\"], d[\"decompress_MBps\"], d[\"stage_ms\"][\"dec.decode\"])"; done
