#!/bin/bash
mkdir -p gpurun_out
P=$PWD/dasp_pytorch_b200
for v in default eqscalar default eqscalar; do
  if [ $v = default ]; then E=""; else E="DASP_LIB_PATH=$P/libdasp_b200_eqscalar.so"; fi
  env $E timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep -E "^parametric" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('$v', k, 'fwd_ms=%.4f bwd_ms=%.4f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms']))"
done
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/a10_tests.log; tail -2 gpurun_out/a10_tests.log
timeout 900 python bench.py --steps 30 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 600 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
