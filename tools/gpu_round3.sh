#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_edge.py -x -q -m gpu > gpurun_out/pytest5.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest5.log
bash tools/gpu_b274.sh
