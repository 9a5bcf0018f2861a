#!/bin/bash
# GPU run A1: new EQ kernels -- correctness first (bounded), then variant timings, then the rest of the suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a1_gpu.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_eq.py -q -x 2>&1 | tail -30 > gpurun_out/a1_eq_tests.log
echo "eq tests rc=$?" >> gpurun_out/a1_eq_tests.log
tail -5 gpurun_out/a1_eq_tests.log
for fw in 1 2 4; do
  DASP_EQ_FWD_W=$fw DASP_EQ_BWD_W=$fw DASP_EQ_BWD_S=1 timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep parametric | sed "s/^/W=$fw S=1 /" >> gpurun_out/a1_eq_variants.log
  DASP_EQ_FWD_W=$fw DASP_EQ_BWD_W=$fw DASP_EQ_BWD_S=2 timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep parametric | sed "s/^/W=$fw S=2 /" >> gpurun_out/a1_eq_variants.log
done
DASP_EQ_FWD_W=8 DASP_EQ_BWD_W=8 timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep parametric | sed "s/^/W=8 S=1 /" >> gpurun_out/a1_eq_variants.log
for fw in 2 4 8; do
  DASP_EQ_FWD_W=$fw DASP_EQ_BWD_W=$fw DASP_EQ_BWD_S=1 timeout 300 python tools/quick_bench.py --ops eq --bs 256 2>&1 | grep parametric | sed "s/^/W=$fw S=1 /" >> gpurun_out/a1_eq_variants.log
done
cat gpurun_out/a1_eq_variants.log
timeout 1500 python -m pytest tests -m gpu -q --deselect tests/test_gpu_eq.py 2>&1 | tail -40 > gpurun_out/a1_all_tests.log
tail -15 gpurun_out/a1_all_tests.log
timeout 600 python tools/quick_bench.py --ops reverb,comp --bs 1024 2>&1 | grep -E "reverb|compressor" > gpurun_out/a1_reverb.log; cat gpurun_out/a1_reverb.log
# ncu full-set capture of the EQ kernels (default variant choice) at 1024 x 2 x 48000
timeout 600 ncu --set full --clock-control none --import-source on -k regex:eq_ -s 2 -c 2 -o gpurun_out/a1_eq python tools/quick_bench.py --ops eq --bs 1024 > gpurun_out/a1_ncu.log 2>&1
ls -la gpurun_out | tail -5
