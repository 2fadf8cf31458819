#!/bin/bash
# GPU check of the 1-modality GaitSet graph (new tests first) + the whole GaitSet suite and the GaitSet builder protocol.
OUT=gpurun_out/r02zz_gs1
mkdir -p $OUT
s0=$(date +%s)
timeout 185 python -m pytest -q "tests/test_gaitset_gpu.py::test_gaitset_single_modality_graph_fp32" \
    "tests/test_compat_gpu.py::test_gaitset_single_modality_builder" "tests/test_compat_gpu.py::test_gaitset_builder_protocol" \
    tests/test_gaitset_gpu.py > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 25 $OUT/pytest_gpu.log
