# A/B helper: decode timing for several values of SQ_DEC_CARVEOUT ("auto" = computed from the co-resident CTAs)
# usage: quick_env.sh <blocks> <carve values...>
b=$1; shift
for c in "$@"; do
  if [ "$c" = auto ]; then unset SQ_DEC_CARVEOUT; else export SQ_DEC_CARVEOUT=$c; fi
  echo "blocks $b carve $c: $(timeout 300 python bench.py --blocks $b --steps 1 --warmup 1 --no-cpu --no-e2e 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d[\"decompress_MBps\"], d[\"stage_ms\"][\"dec.decode\"])")"
done
