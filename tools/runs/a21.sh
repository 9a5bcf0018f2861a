#!/bin/bash
# 16-warp geometry of the compressor kernels and the EQ forward for batches of at most one item per SM
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_dynamics.py tests/test_gpu_eq.py tests/test_gpu_processors.py -q 2>&1 | tail -4 | tee gpurun_out/a21_tests.log
for b in 128 1024; do
timeout 600 python bench.py --batch $b --steps 20 --warmup 3 --no-extras 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('batch $b ms_per_step=%.4f value=%.4g stages=%s' % (d['ms_per_step'], d['value'], json.dumps({k: v['ms'] for k, v in d['stages'].items()})))"
done 2>&1 | tee gpurun_out/a21_bench.log
