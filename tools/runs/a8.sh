#!/bin/bash
mkdir -p gpurun_out
P=$PWD/dasp_pytorch_b200
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5 > gpurun_out/a8_tests.log; tail -3 gpurun_out/a8_tests.log
L=gpurun_out/a8_variants.log; : > $L
run() { name=$1; ops=$2; bs=$3; shift; shift; shift
  env "$@" timeout 300 python tools/quick_bench.py --ops $ops --bs $bs 2>&1 | grep -E "^(parametric|reverb|compressor)" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('$name', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))" >> $L
}
run twiddle_recurrence reverb 1024
run twiddle_tables reverb 1024 DASP_LIB_PATH=$P/libdasp_b200_notw.so
run twiddle_recurrence reverb 1024
run twiddle_tables reverb 1024 DASP_LIB_PATH=$P/libdasp_b200_notw.so
run dyn_e7 comp 1024
run dyn_e11 comp 1024 DASP_LIB_PATH=$P/libdasp_b200_dyn11.so
run dyn_e15 comp 1024 DASP_LIB_PATH=$P/libdasp_b200_dyn15.so
cat $L
