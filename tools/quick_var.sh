# A/B helper: time the decode of N blocks (default 1) with each library variant under _variants/ (gpurun_out is not shipped)
# time the decode of 1 block with each library variant under _variants/
for v in _variants/*.so; do
  cp $v fastqueeze_b200/libseqarc_b200.so
  echo "$v: $(timeout 300 python bench.py --blocks ${1:-1} --steps 1 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d[\"decompress_MBps\"], d[\"stage_ms\"][\"dec.decode\"])")"
done
