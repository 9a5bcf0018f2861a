#!/bin/bash
# EQ: packed (FFMA2) vs scalar arithmetic after the collective brackets are gone
mkdir -p gpurun_out
P=$PWD/dasp_pytorch_b200
run() { env "$@" timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep -E "^parametric_eq" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('[$*]', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))"; }
{
run X=packed
run DASP_LIB_PATH=$P/libdasp_b200_eqscalar.so
run DASP_LIB_PATH=$P/libdasp_b200_eqscalar.so DASP_EQ_FWD_W=8
run DASP_LIB_PATH=$P/libdasp_b200_eqscalar.so DASP_EQ_BWD_W=4
run X=packed
run DASP_LIB_PATH=$P/libdasp_b200_eqscalar.so
} 2>&1 | tee gpurun_out/a18_bench.log
DASP_LIB_PATH=$P/libdasp_b200_eqscalar.so timeout 300 python -m pytest tests/test_gpu_eq.py -x -q 2>&1 | tail -2 | tee -a gpurun_out/a18_bench.log
