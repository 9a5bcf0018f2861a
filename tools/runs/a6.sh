#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_eq.py -q 2>&1 | tail -4 > gpurun_out/a6_tests.log; tail -2 gpurun_out/a6_tests.log
L=gpurun_out/a6_variants.log; : > $L
run() { name=$1; ops=$2; bs=$3; shift; shift; shift
  env "$@" timeout 300 python tools/quick_bench.py --ops $ops --bs $bs 2>&1 | grep -E "^(parametric|reverb)" | python -c "
import sys, json
for l in sys.stdin:
    k, d = l.split(' ', 1); d = json.loads(d); print('$name', k, 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))" >> $L
}
run fW3S2 eq 1024 DASP_EQ_FWD_W=3
run fW3S1 eq 1024 DASP_EQ_FWD_W=3 DASP_EQ_FWD_S=1
run fW4S2_bW3S1 eq 1024 DASP_EQ_BWD_W=3
cat $L
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"eq_bwd" -s 1 -c 1 -o gpurun_out/a6_eqbwd python tools/quick_bench.py --ops eq --bs 1024 > gpurun_out/a6_ncu.log 2>&1
timeout 900 python bench.py --steps 20 > gpurun_out/a6_bench.json 2> gpurun_out/a6_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/a6_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'eager', d['eager_ms_per_step'], 'e2e', d['e2e']['value'] / 1e9)
print({k: (v['ms'], v['frac']) for k, v in d['stages'].items()})
print(d.get('cpu_baseline')); print(d.get('reference_gpu')); print(d.get('configs'))
PY
tail -3 gpurun_out/a6_bench.err
