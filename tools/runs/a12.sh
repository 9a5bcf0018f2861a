#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dynamics.py tests/test_gpu_processors.py -q 2>&1 | tail -3 > gpurun_out/a12_tests.log; tail -2 gpurun_out/a12_tests.log
for g in 0 1 0 1; do DASP_DYN_GENERIC=$g timeout 300 python tools/quick_bench.py --ops comp --bs 1024 2>&1 | grep -E "^compressor" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('generic=$g', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))"; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dynamics_bwd" -s 1 -c 1 -o gpurun_out/r02_dynbwd python tools/quick_bench.py --ops comp --bs 1024 > gpurun_out/a12_ncu.log 2>&1
ls -la gpurun_out/r02_dynbwd.ncu-rep
