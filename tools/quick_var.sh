# A/B helper: one bench step with each library variant under _variants/ (blocks = $1, default the full job)
keep=/tmp/keep_lib.so; cp fastqueeze_b200/libseqarc_b200.so $keep
for v in _variants/*.so; do
  cp $v fastqueeze_b200/libseqarc_b200.so
  echo "$v: $(timeout 300 python bench.py --blocks ${1:-274} --steps 1 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('comp', d['compress_MBps'], 'decomp', d['decompress_MBps'], {k: round(v) for k, v in d['stage_ms'].items() if k.startswith('enc.') and v > 90})")"
done
cp $keep fastqueeze_b200/libseqarc_b200.so
