# per-chain cycle counts of block 0 with each library variant under _variants/ (built with -DSQ_DEBUG_TIMING)
mkdir -p gpurun_out/diag
keep=/tmp/keep_lib.so; cp fastqueeze_b200/libseqarc_b200.so $keep
for v in _variants/*.so; do
  cp $v fastqueeze_b200/libseqarc_b200.so
  for nb in ${@:-274}; do
    echo "== $(basename $v) blocks $nb"
    timeout 400 python bench.py --config ${TICK_CONFIG:-pe150} --blocks $nb --steps 1 --warmup 1 --no-cpu --no-e2e 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('dec warp') and ('qual' in l or 'seq' in l): print('  ', l.strip())
    if l.startswith('{'): d = json.loads(l); print('   decompress', d['decompress_MBps'], 'MB/s, dec.decode', round(d['stage_ms']['dec.decode']), 'ms')
" | sort -u
  done
done
cp $keep fastqueeze_b200/libseqarc_b200.so
