#!/bin/bash
# GPU-box: conv kernel correctness sweep (fp16, mixed backward) + per-layer timings with A/B switches
O=gpurun_out/conv; mkdir -p $O
UGN_CHECK_DT=f16 UGN_CHECK_MIXED=1 UGN_CHECK_GS=1024 python scripts/tc_check.py conv > $O/check.log 2>&1; grep "fwd rel\|dgrad rel\|ERROR" $O/check.log
python scripts/tc_check.py conv > $O/check_bf16.log 2>&1; grep "fwd rel\|ERROR" $O/check_bf16.log
echo "== default"; python scripts/conv_bench.py "$@" 2>&1 | tail -7
echo "== UGN_NO_CONCAT"; UGN_NO_CONCAT=1 python scripts/conv_bench.py fwd 2>&1 | tail -7
echo "== UGN_NO_CONCAT UGN_NO_N192"; UGN_NO_CONCAT=1 UGN_NO_N192=1 python scripts/conv_bench.py fwd 2>&1 | tail -7
bash scripts/conv_prof.sh 2>&1 | grep "convp\|==" | awk '!seen[$0]++' | cut -c1-330
