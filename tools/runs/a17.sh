#!/bin/bash
# EQ kernels with the warp index read from lane 0 (no collective brackets around the scan shuffles): parity, then timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eq.py tests/test_gpu_processors.py -x -q 2>&1 | tail -4 | tee gpurun_out/a17_tests.log
run() { env "$@" timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep -E "^parametric_eq" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('[$*]', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))"; }
{
run DASP_LIB_PATH=$PWD/dasp_pytorch_b200/libdasp_b200_before.so
run X=1
run DASP_EQ_FWD_W=8
run DASP_EQ_FWD_W=4 DASP_EQ_FWD_S=1
run DASP_EQ_FWD_W=2
run DASP_EQ_BWD_W=4
run DASP_EQ_BWD_W=4 DASP_EQ_BWD_S=2
run DASP_EQ_BWD_W=3
run DASP_EQ_BWD_W=2
run DASP_EQ_PAIR_TABLES=1
run X=1
} 2>&1 | tee gpurun_out/a17_bench.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('bench ms_per_step=%.3f value=%.4g stages=%s' % (d['ms_per_step'], d['value'], json.dumps({k: (v['ms'], v['frac']) for k, v in d['stages'].items()})))" 2>&1 | tee -a gpurun_out/a17_bench.log
