mkdir -p gpurun_out/r02i
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $T scripts/dp_check.py fused > gpurun_out/r02i/dpcheck.log 2>&1; tail -2 gpurun_out/r02i/dpcheck.log | cut -c1-300; grep -c identical gpurun_out/r02i/dpcheck.log
UGN_DP_TIMING=1 timeout 600 $T bench.py --gpus 2 --steps 30 --warmup 5 --no-knn > gpurun_out/r02i/b2.json 2> gpurun_out/r02i/b2.err; python scripts/bline.py gpurun_out/r02i/b2.json || tail -5 gpurun_out/r02i/b2.err
python -c "
import json;d=json.loads(open('gpurun_out/r02i/b2.json').read().strip().splitlines()[-1]);print(d['config'].get('dp_timing'), d['config'].get('rank_check'))"
UGN_DP_PUSH=0 UGN_DP_TIMING=1 timeout 600 $T bench.py --gpus 2 --steps 30 --warmup 5 --no-knn > gpurun_out/r02i/b2np.json 2> gpurun_out/r02i/b2np.err; python scripts/bline.py gpurun_out/r02i/b2np.json
