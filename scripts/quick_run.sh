#!/bin/bash
# GPU-box: full GPU test-suite, default bench, launch list of one eager step.  usage: quick_run.sh <tag> [pytest-args]
TAG=${1:-q}; shift
O=gpurun_out/$TAG; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q -s "$@" > $O/pytest_gpu.log 2>&1
grep -v "^   " $O/pytest_gpu.log | tail -n 25
grep "update cosine\|losses" $O/pytest_gpu.log | tail -n 50
timeout 600 python bench.py --no-knn > $O/bench_n1.json 2> $O/bench_n1.err || tail -n 20 $O/bench_n1.err
python - <<PY
import json
for l in open("$O/bench_n1.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["value"], d["ms_per_step"], d["e2e"], d.get("op_ms_per_step"), d.get("roofline"))
PY
CMD="python bench.py --steps 2 --warmup 3 --no-graph --lite"
$CMD > $O/plain.log 2>&1 || { tail -5 $O/plain.log; exit 1; }
NL=$(python -c "import json,sys; print(json.loads(open('$O/plain.log').read().strip().splitlines()[-1])['gpu_launches']//2)")
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((NL*3)) -c $NL --csv \
    --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
python scripts/launch_table.py $O/launches.csv
