#!/bin/bash
# 2-GPU session: GPU test-suite (sharded CLI cases run at 2 GPUs), every bench config on a small job, a 2-rank bench
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest2.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest2.log
for c in pe150 pe150_lossy se100_q8 pe150_ref; do
  timeout 900 python bench.py --config $c --blocks 8 --steps 1 --warmup 1 --cpu-sample-blocks 4 > gpurun_out/cfg_$c.json 2> gpurun_out/cfg_$c.err
  echo "$c rc=$?"; tail -c 1500 gpurun_out/cfg_$c.json; echo; tail -2 gpurun_out/cfg_$c.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --blocks 16 --steps 1 --warmup 1 > gpurun_out/n2_small.json 2> gpurun_out/n2_small.err
echo "n2 rc=$?"; tail -c 2500 gpurun_out/n2_small.json; echo; tail -3 gpurun_out/n2_small.err
