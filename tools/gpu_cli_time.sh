#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python tools/scale_check.py --config pe150 --mode records > gpurun_out/scale_pe150_records2.json 2> gpurun_out/scale_pe150_records2.err; echo "scale rc=$?"
cat gpurun_out/scale_pe150_records2.json; tail -3 gpurun_out/scale_pe150_records2.err
