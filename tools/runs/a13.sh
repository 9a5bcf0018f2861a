#!/bin/bash
mkdir -p gpurun_out
P=$PWD/dasp_pytorch_b200
for v in "" _dyn_m5s3 _dyn_m5s2 _dyn_m6s2 _dyn_m7s2 "" _dyn_m5s3 _dyn_m5s2 _dyn_m6s2 _dyn_m7s2; do
  if [ -z "$v" ]; then E=""; else E="DASP_LIB_PATH=$P/libdasp_b200$v.so"; fi
  env $E timeout 300 python tools/quick_bench.py --ops comp --bs 1024 2>&1 | grep -E "^compressor" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('variant=[$v]', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))"; done
