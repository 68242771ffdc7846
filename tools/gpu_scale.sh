#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest3.log
free -g | head -2; df -h /dev/shm | tail -1; nproc
timeout 1500 python tools/scale_check.py --config pe150 --mode records > gpurun_out/scale_pe150_records.json 2> gpurun_out/scale_pe150_records.err; echo "scale rc=$?"
cat gpurun_out/scale_pe150_records.json; tail -3 gpurun_out/scale_pe150_records.err
