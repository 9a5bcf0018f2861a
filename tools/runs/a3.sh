#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eq.py tests/test_gpu_dynamics.py -q 2>&1 | tail -8 > gpurun_out/a3_tests.log; tail -3 gpurun_out/a3_tests.log
./tools/probe/ffma2_probe2 > gpurun_out/a3_ffma2_probe.txt 2>&1; cat gpurun_out/a3_ffma2_probe.txt
timeout 300 python tools/quick_bench.py --ops eq,comp 2>&1 | grep -E "parametric|compressor" | cut -c1-260
timeout 900 python bench.py --steps 20 > gpurun_out/a3_bench.json 2> gpurun_out/a3_bench.err; tail -c 5000 gpurun_out/a3_bench.json; tail -5 gpurun_out/a3_bench.err
