#!/bin/bash
# Last GPU call of round 2 (runs on the box via gpurun): full GPU test suite, smoke(), default bench line and the ncu launch
# list of one eager step of the FINAL code.  Everything is bounded by `timeout`; outputs land in gpurun_out/r02zz/.
OUT=gpurun_out/r02zz
mkdir -p $OUT
s0=$(date +%s)
timeout 200 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 3 $OUT/pytest_gpu.log
timeout 60 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 2 $OUT/smoke.log
timeout 150 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench rc=$? t=$(( $(date +%s) - s0 ))s"; python scripts/bline.py $OUT/bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-graph --lite"
timeout 60 $CMD > $OUT/plain.log 2>&1 || { tail -5 $OUT/plain.log; exit 1; }
NL=$(python -c "import json,sys; print(json.loads(open('$OUT/plain.log').read().strip().splitlines()[-1])['gpu_launches']//2)")
timeout 90 ncu --metrics gpu__time_duration.sum --clock-control none -s $((NL*3)) -c $NL --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
echo "ncu rc=$? NL=$NL t=$(( $(date +%s) - s0 ))s"
