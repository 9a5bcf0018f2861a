#!/bin/bash
# usage: multi.sh N  -- multi-GPU checks on an N-GPU box: bit-identity of sharded runs, bench.py under torchrun
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
if [ "$N" = "2" ]; then timeout 600 python -m pytest tests/test_gpu_multi.py -q 2>&1 | tail -3 | tee gpurun_out/multi_tests_n$N.log; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -c 2500 gpurun_out/r02_bench_n$N.json; tail -4 gpurun_out/r02_bench_n$N.err
