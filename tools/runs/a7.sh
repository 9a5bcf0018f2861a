#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spectral_gen|ifft_shape|x_fft|partition_mac|ifft_mix|g_fft|ifft_dx|ifft_irgrad|vector_fft|regular_fft|reverb_param" --launch-skip 12 -c 12 -o gpurun_out/r02_reverb_b148 python tools/debug/reverb_step.py 148 > gpurun_out/a7_ncu1.log 2>&1
ls -la gpurun_out/r02_reverb_b148.ncu-rep
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_chain_launches.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/a7_ncu2.log 2>&1
wc -l gpurun_out/r02_chain_launches.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"dynamics_" -s 2 -c 2 -o gpurun_out/r02_dyn python tools/quick_bench.py --ops comp --bs 1024 > gpurun_out/a7_ncu3.log 2>&1
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a7_ref.json 2> gpurun_out/a7_ref.err; tail -c 1500 gpurun_out/a7_ref.json
