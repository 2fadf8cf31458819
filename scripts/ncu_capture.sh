#!/bin/bash
# Runs on the GPU box (via gpurun): GPU tests, bench line, ncu launch list of one eager step, then
# `--set full` captures of the dominant kernels.  Everything lands in gpurun_out/ (scratch); the
# summaries that are to be judged are condensed into profiles/ by scripts/summarise_profiles.py.
#   usage: scripts/ncu_capture.sh [tag] [notests]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
set -x
if [ "$2" != "gsonly" ]; then
if [ "$2" != "notests" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  tail -n 5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err || tail -n 20 $OUT/bench_n1.err
tail -c 3000 $OUT/bench_n1.json
CMD="python bench.py --steps 2 --warmup 3 --no-graph --lite"
$CMD > $OUT/plain.log 2>&1 || { tail -5 $OUT/plain.log; exit 1; }
# launch list: skip the 3 warm-up steps (per-step launch count is printed by the plain run)
NL=$(python -c "import json,sys; print(json.loads(open('$OUT/plain.log').read().strip().splitlines()[-1])['gpu_launches']//2)")
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s $((NL*3)) -c $NL --csv \
    --log-file $OUT/launches.csv $CMD > $OUT/ncu_list.log 2>&1
for KS in tc_convp_kernel:45:15 tc_wgradv_kernel:27:9 tc_kernel:80:20 optim_kernel:3:1 knn_tc_scan_kernel:2:2; do
  K=${KS%%:*}; R=${KS#*:}; S=${R%%:*}; C=${R#*:}
  RUN="$CMD"; if [ $K = knn_tc_scan_kernel ]; then RUN="python bench.py --knn-only"; fi
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:^$K\$ -s $S -c $C -o $OUT/prof_$K $RUN > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_${K}_raw.csv 2>/dev/null
done
fi
# ---- GaitSet branch type: launch list of one eager step at 24 rows + captures of its own kernels
GCMD="python scripts/gs_bench.py 24 f16mix 1"
GS_LITE=1 $GCMD > $OUT/gs_plain.log 2>&1 || tail -5 $OUT/gs_plain.log
GL=$(python -c "import json; print(json.loads(open('$OUT/gs_plain.log').read().strip().splitlines()[-1])['launches_per_step'])")
# the measured step is bracketed by cudaProfilerStart/Stop inside the script (torch's own fill / copy kernels of the
# plan construction would otherwise shift a launch-count window)
GS_LITE=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
    --log-file $OUT/gs_launches.csv $GCMD > $OUT/gs_ncu_list.log 2>&1
for KS in gs_conv1_fwd_kernel:3:3 gs_conv1_wgrad_kernel:3:3 gs_setmax_bwd_kernel:9:3 gs_pad_kernel:24:4; do
  K=${KS%%:*}; R=${KS#*:}; S=${R%%:*}; C=${R#*:}
  GS_LITE=1 timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:$K -c $C -o $OUT/prof_$K $GCMD > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_${K}_raw.csv 2>/dev/null
  rm -f $OUT/prof_$K.ncu-rep
done
# the 3x3 'same' layers of the GaitSet branch on the shared conv kernels (first modality: a2,b1,b2,a3,a4,b3,b4,a5,a6)
GS_LITE=1 timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:tc_convp_kernel -c 9 -o $OUT/prof_gs_convp $GCMD > $OUT/ncu_gs_convp.log 2>&1
ncu -i $OUT/prof_gs_convp.ncu-rep --page raw --csv > $OUT/prof_gs_convp_raw.csv 2>/dev/null
rm -f $OUT/prof_gs_convp.ncu-rep
du -sm $OUT
# keep the payload under the 64 MiB copy-back limit
for K in optim_kernel knn_tc_scan_kernel tc_kernel tc_wgradv_kernel tc_convp_kernel; do
  if [ $(du -sm gpurun_out | cut -f1) -gt 50 ]; then rm -f $OUT/prof_$K.ncu-rep; fi
done
ls -la $OUT
