#!/bin/bash
mkdir -p gpurun_out/r02zz_gs5
timeout 25 python -m pytest -q "tests/test_compat_gpu.py::test_shim_gpu_knn_is_what_the_test_mains_import" > gpurun_out/r02zz_gs5/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -n 15 gpurun_out/r02zz_gs5/pytest_gpu.log | cut -c1-250
