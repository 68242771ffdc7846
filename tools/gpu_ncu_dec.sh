#!/bin/bash
# source-level stall profile of the decoder on one small block (plain run first, then ncu)
mkdir -p gpurun_out
python tools/prof_run.py --blocks 1 --pairs 8000 --decode > gpurun_out/plain_dec.log 2>&1 &&
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on \
    -k regex:k_dec_main -c 1 -o gpurun_out/prof_dec -f python tools/prof_run.py --blocks 1 --pairs 8000 --decode > gpurun_out/ncu_dec.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_dec.log; ls -la gpurun_out/prof_dec.ncu-rep
