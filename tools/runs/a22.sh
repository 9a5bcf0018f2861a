#!/bin/bash
# compute-sanitizer memcheck over the smoke chain and the small-shape variant tests (16-warp paths included)
mkdir -p gpurun_out
timeout 150 compute-sanitizer --tool memcheck --error-exitcode 7 python -c "import __graft_entry__ as g; g.smoke(); print('smoke under memcheck ok')" > gpurun_out/a22_memcheck_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/a22_memcheck_smoke.log
tail -5 gpurun_out/a22_memcheck_smoke.log
timeout 160 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_dynamics.py::test_compressor_every_warps_per_item_variant "tests/test_gpu_eq.py::test_eq_every_warps_per_pair_variant" -x -q > gpurun_out/a22_memcheck_variants.log 2>&1; echo "rc=$?" >> gpurun_out/a22_memcheck_variants.log
tail -5 gpurun_out/a22_memcheck_variants.log
