#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eq.py tests/test_gpu_processors.py -q 2>&1 | tail -8 > gpurun_out/a5_tests.log; tail -3 gpurun_out/a5_tests.log
L=gpurun_out/a5_variants.log; : > $L
run() { name=$1; ops=$2; bs=$3; shift; shift; shift
  env "$@" timeout 300 python tools/quick_bench.py --ops $ops --bs $bs 2>&1 | grep -E "^(parametric|reverb)" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('$name', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))" >> $L
}
run float_auto eq 1024
run pair_tables eq 1024 DASP_EQ_PAIR_TABLES=1
run float_fW8 eq 1024 DASP_EQ_FWD_W=8
run float_fW2 eq 1024 DASP_EQ_FWD_W=2
run float_fW4S1 eq 1024 DASP_EQ_FWD_S=1
run float_bW4S1 eq 1024 DASP_EQ_BWD_W=4
run float_bW4S2 eq 1024 DASP_EQ_BWD_W=4 DASP_EQ_BWD_S=2
run float_bW2S2 eq 1024 DASP_EQ_BWD_W=2 DASP_EQ_BWD_S=2
run float_auto eq 256
cat $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eq_fwd|eq_bwd" -s 4 -c 2 -o gpurun_out/a5_eq python tools/quick_bench.py --ops eq --bs 1024 > gpurun_out/a5_ncu.log 2>&1
ls -la gpurun_out/a5_eq.ncu-rep
