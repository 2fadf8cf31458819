N=${1:-8}
mkdir -p gpurun_out/r02j
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514"
timeout 300 $T scripts/dp_check.py fused > gpurun_out/r02j/dpcheck_n$N.log 2>&1; tail -1 gpurun_out/r02j/dpcheck_n$N.log | cut -c1-200; grep -c "identical" gpurun_out/r02j/dpcheck_n$N.log
UGN_DP_TIMING=1 timeout 600 $T bench.py --gpus $N --steps 30 --warmup 5 > gpurun_out/r02j/b$N.json 2> gpurun_out/r02j/b$N.err; python scripts/bline.py gpurun_out/r02j/b$N.json || grep -n "Error\|error" gpurun_out/r02j/b$N.err | head
python - <<PY
import json
d=json.loads(open('gpurun_out/r02j/b$N.json').read().strip().splitlines()[-1])
print(d['config'].get('dp_timing'), d['config'].get('rank_check'), d['config'].get('dp_exchange'))
print('e2e', d['e2e']['value'], d['e2e']['full_batch']['value'])
for k in ('knn','knn_d2048'):
    kn=d.get(k,{})
    print(k, {q:(round(kn[q]['queries_per_s']), round(kn[q]['ms'],3)) for q in ('Q4096','Q64') if q in kn})
PY
