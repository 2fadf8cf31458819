#!/bin/bash
mkdir -p gpurun_out
CMD="python scripts/conv_bench.py fwd"
$CMD > gpurun_out/cb_plain.log 2>&1 || { tail -5 gpurun_out/cb_plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:tc_convp -s 3 -c 1 -o gpurun_out/prof_convp $CMD > gpurun_out/ncu_cb.log 2>&1
ncu -i gpurun_out/prof_convp.ncu-rep --page raw --csv > gpurun_out/prof_convp_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_convp.ncu-rep --page source --csv > gpurun_out/prof_convp_src.csv 2>/dev/null
ls -la gpurun_out | tail -8
tail -3 gpurun_out/ncu_cb.log
