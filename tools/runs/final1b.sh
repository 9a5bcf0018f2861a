#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 30 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 300 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 900 ncu --set full --clock-control none -k regex:"spectral_gen|ifft_shape|x_fft|partition_mac|ifft_mix|g_fft|ifft_dx|ifft_irgrad|vector_fft|regular_fft|reverb_param" --launch-skip 11 -c 11 -o gpurun_out/r02_reverb_b148_final python tools/debug/reverb_step.py 148 > gpurun_out/final_ncu1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_chain_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/final_ncu2.log 2>&1
du -sh gpurun_out; ls -la gpurun_out
