#!/bin/bash
# GPU run A2: EQ variants (E, W, S), reverb launch list, new bench.py at N=1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_eq.py -q 2>&1 | tail -15 > gpurun_out/a2_eq_tests.log; tail -3 gpurun_out/a2_eq_tests.log
L=gpurun_out/a2_eq_variants.log; : > $L
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 300 python tools/quick_bench.py --ops eq --bs 1024 2>&1 | grep parametric | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l.split(' ', 1)[1]); print('$name', 'fwd_ms=%.4f bwd_ms=%.4f fwd_frac=%.3f bwd_frac=%.3f' % (d['fwd_ms'], d['fwdbwd_ms'] - d['fwd_ms'], d['fwd_frac'], d['bwd_frac']))" >> $L
}
P=$PWD/dasp_pytorch_b200
run e15_auto
run e15_fS1 DASP_EQ_FWD_S=1
run e15_fW3S1_bW1 DASP_EQ_FWD_S=1 DASP_EQ_FWD_W=3 DASP_EQ_BWD_W=1 DASP_EQ_BWD_S=1
run e15_bW8 DASP_EQ_BWD_W=8
run e15_bW6 DASP_EQ_BWD_W=6
run e15_bW4S1 DASP_EQ_BWD_W=4 DASP_EQ_BWD_S=1
run e13_auto DASP_LIB_PATH=$P/libdasp_b200_e13.so
run e13_fS1_bW8 DASP_LIB_PATH=$P/libdasp_b200_e13.so DASP_EQ_FWD_S=1 DASP_EQ_BWD_W=8
run e13_fW4S1 DASP_LIB_PATH=$P/libdasp_b200_e13.so DASP_EQ_FWD_S=1 DASP_EQ_FWD_W=4
run e11_auto DASP_LIB_PATH=$P/libdasp_b200_e11.so
run e11_fW4S2_bW8 DASP_LIB_PATH=$P/libdasp_b200_e11.so DASP_EQ_FWD_W=4 DASP_EQ_BWD_W=8
run e11_fW4S1 DASP_LIB_PATH=$P/libdasp_b200_e11.so DASP_EQ_FWD_W=4 DASP_EQ_FWD_S=1
cat $L
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/a2_reverb_launches.csv python tools/debug/reverb_step.py 148 > gpurun_out/a2_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/a2_reverb_launches.csv')) if len(r) > 5]
hdr = rows[0]; ik = hdr.index('Kernel Name'); iv = hdr.index('Metric Value'); iid = hdr.index('ID')
n = len(rows) - 1
half = rows[1 + n // 2:]
agg = collections.OrderedDict()
for r in half:
    k = r[ik].split('(')[0][-60:]
    agg.setdefault(k, [0, 0.0]); agg[k][0] += 1; agg[k][1] += float(r[iv].replace(',', ''))
tot = sum(v[1] for v in agg.values())
for k, v in agg.items(): print('%-62s x%d %9.1f us %5.1f%%' % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
print('total us', tot / 1e3)
PY
timeout 900 python bench.py --steps 10 > gpurun_out/a2_bench.json 2> gpurun_out/a2_bench.err; tail -c 3000 gpurun_out/a2_bench.json; tail -5 gpurun_out/a2_bench.err
