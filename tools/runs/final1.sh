#!/bin/bash
# final 1-GPU evidence run: full GPU test suite, smoke(), contract bench (both arms), ncu captures of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/final_tests.log; tail -2 gpurun_out/final_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py --steps 30 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_n1.json 2> gpurun_out/r02_ref.err; tail -c 300 gpurun_out/r02_bench_reference_n1.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"spectral_gen|ifft_shape|x_fft|partition_mac|ifft_mix|g_fft|ifft_dx|ifft_irgrad|vector_fft|regular_fft|reverb_param" --launch-skip 11 -c 11 -o gpurun_out/r02_reverb_b148_final python tools/debug/reverb_step.py 148 > gpurun_out/final_ncu1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_chain_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/final_ncu2.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eq_fwd|eq_bwd|dynamics_" -s 0 -c 8 -o gpurun_out/r02_scan_final python tools/debug/scan_step.py > gpurun_out/final_ncu3.log 2>&1
ls -la gpurun_out/*final*
