#!/bin/bash
# Runs on the GPU box (via gpurun): plain run, launch list, then two small --set full captures.
set -x
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-graph --lite"
$CMD > gpurun_out/plain.log 2>&1 || { tail -5 gpurun_out/plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 480 -c 320 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 153 -c 10 -o gpurun_out/prof_fwd $CMD > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_kernel -s 175 -c 7 -o gpurun_out/prof_bwd $CMD > gpurun_out/ncu3.log 2>&1
for f in prof_fwd prof_bwd; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/${f}_raw.csv 2>/dev/null
done
ls -la gpurun_out
du -sm gpurun_out
# keep the payload under the 64 MiB copy-back limit
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_bwd.ncu-rep; fi
if [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; then rm -f gpurun_out/prof_fwd.ncu-rep; fi
tail -n 3 gpurun_out/plain.log
