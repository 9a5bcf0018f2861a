#!/bin/bash
# final validation of the committed state: full GPU test-suite, smoke(), default bench line, EQ ncu capture + launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 | tee gpurun_out/final2_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2 | tee gpurun_out/final2_smoke.log
timeout 900 python bench.py --steps 30 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; tail -c 400 gpurun_out/r02_bench_n1.json; tail -2 gpurun_out/r02_bench_n1.err
timeout 600 ncu --set full --clock-control none -k regex:"eq_fwd|eq_bwd" -s 4 -c 3 -o gpurun_out/r02_eq_final python tools/quick_bench.py --ops eq --bs 1024 > gpurun_out/final2_ncu_eq.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_chain_launches_final.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/final2_ncu2.log 2>&1
ls -la gpurun_out | head -40
