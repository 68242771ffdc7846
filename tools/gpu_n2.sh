#!/bin/bash
# two GPUs: the sharded CLI tests, then the bench line at N = 2 (strong scaling + replica figure)
O=gpurun_out/r2b; mkdir -p $O
nvidia-smi -L
python -m pytest tests/test_gpu_cli.py -q -m gpu -k "sharded or small_batches" > $O/pytest_n2_cli.log 2>&1; echo "cli rc=$?"; tail -2 $O/pytest_n2_cli.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 2 --warmup 1 > $O/bench_n2.json 2> $O/bench_n2.err
echo "bench rc=$?"; tail -c 1500 $O/bench_n2.json; tail -5 $O/bench_n2.err
