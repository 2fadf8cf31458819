#!/bin/bash
OUT=gpurun_out/r02zz_gs4
mkdir -p $OUT
s0=$(date +%s)
timeout 55 python -m pytest -q "tests/test_compat_gpu.py::test_standalone_branch_builders_and_small_entry_points" > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 25 $OUT/pytest_gpu.log | cut -c1-300
