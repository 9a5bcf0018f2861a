#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_reverb.py tests/test_gpu_processors.py -q 2>&1 | tail -3 > gpurun_out/a11_tests.log; tail -2 gpurun_out/a11_tests.log
timeout 300 python tools/quick_bench.py --ops reverb --bs 1024 2>&1 | grep -E "^reverb" | cut -c1-200
timeout 600 python bench.py --steps 20 --no-extras > gpurun_out/a11_bench.json 2> gpurun_out/a11_bench.err; python - <<'PY'
import json
d = json.loads(open('gpurun_out/a11_bench.json').read().strip().splitlines()[-1])
print('ms_per_step', d['ms_per_step'], 'eager', d['eager_ms_per_step'], 'e2e', d['e2e']['value'] / 1e9)
print({k: (v['ms'], v['frac']) for k, v in d['stages'].items()})
PY
