#!/bin/bash
# Round-2 evidence session on one B200: plain runs first, then the same commands under ncu (B200_PROFILING.md recipe)
O=gpurun_out/${EV_OUT:-r2}; mkdir -p $O
python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest_gpu.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 500 > $O/clocks.csv &
SMI=$!
python bench.py > $O/bench_n1.json 2> $O/bench_n1.err; echo "bench rc=$?"; tail -c 600 $O/bench_n1.json; echo
kill $SMI
python bench.py --impl reference --steps 1 --warmup 1 > $O/bench_reference_arm.json 2> $O/bench_reference_arm.err; echo "ref rc=$?"; tail -c 400 $O/bench_reference_arm.json; echo
for c in ${EV_CONFIGS-se100_q8 pe150_lossy pe150_ref}; do
  timeout 900 python bench.py --config $c --steps 1 --warmup 1 > $O/bench_$c.json 2> $O/bench_$c.err; echo "$c rc=$?"
  python - <<PY
import json
try:
    d=json.loads(open("$O/bench_$c.json").read().strip().splitlines()[-1])
    print("$c: value", d["value"], "comp", d["compress_MBps"], "decomp", d["decompress_MBps"], "e2e", (d["e2e"] or {}).get("value"), "cpu", (d["cpu_baseline"] or {}).get("value"))
except Exception as e:
    print("$c failed", e); print(open("$O/bench_$c.err").read()[-800:])
PY
done
# launch list + DRAM traffic of the dominant kernel, on the bench command (device-only legs)
CMD="python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e"
$CMD > $O/plain_for_ncu.json 2> $O/plain_for_ncu.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 900 --csv --log-file $O/launches_full_job.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct --clock-control none \
    -k regex:k_dec_main -c 2 --csv --log-file $O/traffic_k_dec_main.csv $CMD > $O/ncu_traffic.log 2>&1
echo "traffic rc=$?"
python tools/ncu_traffic.py $O/traffic_k_dec_main.csv k_dec_main dec.decode pe150 274 1 1 "profiles/r2_traffic_k_dec_main.csv (ncu, k_dec_main, the timed step's launch)" | tee $O/traffic_summary.json
cp profiles/traffic.json $O/traffic.json
# stall profile of the decoder on one small block
python tools/prof_run.py --blocks 1 --pairs 8000 --decode > $O/plain_dec.log 2>&1 &&
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on \
    -k regex:k_dec_main -c 1 -o $O/prof_dec -f python tools/prof_run.py --blocks 1 --pairs 8000 --decode > $O/ncu_dec.log 2>&1
echo "stall profile rc=$?"
ls -la $O
