#!/bin/bash
mkdir -p gpurun_out
P=$PWD/dasp_pytorch_b200
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/a4_tests.log; tail -3 gpurun_out/a4_tests.log
./tools/probe/ffma2_probe2 > gpurun_out/a4_ffma2_probe.txt 2>&1; tail -8 gpurun_out/a4_ffma2_probe.txt
L=gpurun_out/a4_variants.log; : > $L
run() { name=$1; ops=$2; shift; shift
  env "$@" timeout 300 python tools/quick_bench.py --ops $ops --bs 1024 2>&1 | grep -E "^(parametric|reverb)" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('$name', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))" >> $L
}
run scalar_e15 eq,reverb
run packed_e15 eq,reverb DASP_LIB_PATH=$P/libdasp_b200_packed.so
run scalar_e15_fW8 eq DASP_EQ_FWD_W=8
run scalar_e15_fW2 eq DASP_EQ_FWD_W=2
run scalar_e15_bW4S1 eq DASP_EQ_BWD_W=4 DASP_EQ_BWD_S=1
run scalar_e23 eq DASP_LIB_PATH=$P/libdasp_b200_e23.so
run scalar_e23_fW3_bW4 eq DASP_LIB_PATH=$P/libdasp_b200_e23.so DASP_EQ_FWD_W=3 DASP_EQ_BWD_W=4
run scalar_e31 eq DASP_LIB_PATH=$P/libdasp_b200_e31.so
run scalar_e31_fW2_bW3 eq DASP_LIB_PATH=$P/libdasp_b200_e31.so DASP_EQ_FWD_W=2 DASP_EQ_BWD_W=3
run scalar_e31_fW3S1 eq DASP_LIB_PATH=$P/libdasp_b200_e31.so DASP_EQ_FWD_W=3 DASP_EQ_FWD_S=1
cat $L
timeout 600 python bench.py --steps 20 --no-extras > gpurun_out/a4_bench.json 2> gpurun_out/a4_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/a4_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'eager', d['eager_ms_per_step'], 'e2e', d['e2e']['value'] / 1e9)
print({k: (v['ms'], v['frac']) for k, v in d['stages'].items()})
PY
tail -3 gpurun_out/a4_bench.err
