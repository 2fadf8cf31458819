#!/bin/bash
# GPU-box driver for the fp16 storage / mixed-precision backward checks.
mkdir -p gpurun_out/f16
O=gpurun_out/f16
python scripts/tc_check.py > $O/check_bf16.log 2>&1
UGN_CHECK_DT=f16 python scripts/tc_check.py > $O/check_f16.log 2>&1
UGN_CHECK_DT=f16 UGN_CHECK_MAG=0.01 python scripts/tc_check.py gemm > $O/check_f16_small.log 2>&1
UGN_CHECK_DT=f16 UGN_CHECK_MIXED=1 UGN_CHECK_GS=1024 python scripts/tc_check.py conv > $O/check_f16_mixed.log 2>&1
tail -n 40 $O/check_f16.log $O/check_f16_small.log $O/check_f16_mixed.log
timeout 900 python -m pytest tests/test_step_gpu.py -x -q -s -k "tensor_core" > $O/pytest_tc.log 2>&1
grep -v "^   \[" $O/pytest_tc.log | tail -n 15
grep "^   \[f16mix\]\|^   \[bf16x3\]\|^\[" $O/pytest_tc.log | tail -n 80
for m in bf16x3 f16x3 f16mix; do
  timeout 600 python bench.py --mode $m --no-knn > $O/bench_$m.json 2> $O/bench_$m.err || tail -n 5 $O/bench_$m.err
  python - <<PY
import json
for l in open("$O/bench_$m.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("$m", d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("op_ms_per_step"), d.get("roofline"))
PY
done
