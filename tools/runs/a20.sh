#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none -k regex:eq_bwd -s 1 -c 1 -o gpurun_out/r02_eq_bwd_final python tools/quick_bench.py --ops eq --bs 1024 > gpurun_out/a20_ncu.log 2>&1
tail -3 gpurun_out/a20_ncu.log
