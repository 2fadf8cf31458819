#!/bin/bash
# GPU check of postriplet == 2 with GaitSet branches + the _layer_output refactor (new tests first).
OUT=gpurun_out/r02zz_gs2
mkdir -p $OUT
s0=$(date +%s)
timeout 150 python -m pytest -q "tests/test_gaitset_gpu.py::test_gaitset_postriplet2_graph_fp32" \
    "tests/test_compat_gpu.py::test_postriplet2_gaitset_builder" "tests/test_compat_gpu.py::test_postriplet2_builder" \
    tests/test_gaitset_gpu.py tests/test_compat_gpu.py tests/test_edge_gpu.py > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 30 $OUT/pytest_gpu.log | cut -c1-400
