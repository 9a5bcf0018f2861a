#!/bin/bash
# decoupled-warp compressor backward: register-budget variants against the legacy kernels
mkdir -p gpurun_out
P=$PWD/dasp_pytorch_b200
for rep in 1 2; do
for cfg in "|DASP_DYN_LEGACY=1" "_d2w16|DASP_DYN_W=4" "_d2w16|DASP_DYN_W=8" "_d2w20|DASP_DYN_W=4" "_d2w12|DASP_DYN_W=4" "_d2w16|DASP_DYN_W=2"; do
  v=${cfg%%|*}; e=${cfg##*|}
  env DASP_LIB_PATH=$P/libdasp_b200$v.so $e timeout 300 python tools/quick_bench.py --ops comp --bs 1024 2>&1 | grep -E "^compressor" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('[$cfg]', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))"; done; done 2>&1 | tee gpurun_out/a15_bench.log
