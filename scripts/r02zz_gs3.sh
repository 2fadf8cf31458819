#!/bin/bash
# GPU check after the _layer_output refactor: descriptor-extraction tests of the stacked engine + smoke().
OUT=gpurun_out/r02zz_gs3
mkdir -p $OUT
s0=$(date +%s)
timeout 60 python -m pytest -q tests/test_step_gpu.py tests/test_ops_gpu.py tests/test_decisions_gpu.py -k "predict or expand or descriptor or video or open_world" > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 6 $OUT/pytest_gpu.log | cut -c1-300
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 2 $OUT/smoke.log
