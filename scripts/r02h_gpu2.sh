mkdir -p gpurun_out/r02h
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
UGN_DP_TIMING=1 timeout 600 $T bench.py --gpus 2 --steps 30 --warmup 5 --no-knn > gpurun_out/r02h/b2.json 2> gpurun_out/r02h/b2.err; python scripts/bline.py gpurun_out/r02h/b2.json || tail -5 gpurun_out/r02h/b2.err
python -c "
import json;d=json.loads(open('gpurun_out/r02h/b2.json').read().strip().splitlines()[-1]);print(d['config'].get('dp_timing'), d['config'].get('rank_check'))"
