N=${1:-2}
mkdir -p gpurun_out/r02p
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29516"
[ "$N" != 8 ] && timeout 300 $T scripts/dp_check.py fused > gpurun_out/r02p/dpcheck_n$N.log 2>&1; tail -1 gpurun_out/r02p/dpcheck_n$N.log | cut -c1-300; grep -c "identical" gpurun_out/r02p/dpcheck_n$N.log
for D in 1 0; do
UGN_DP_DEFER=$D UGN_DP_TIMING=1 timeout 600 $T bench.py --gpus $N --steps 30 --warmup 5 --no-configs > gpurun_out/r02p/b${N}_d$D.json 2> gpurun_out/r02p/b${N}_d$D.err; python scripts/bline.py gpurun_out/r02p/b${N}_d$D.json || grep -n "Error\|error" gpurun_out/r02p/b${N}_d$D.err | head
python - <<PY
import json
d=json.loads(open('gpurun_out/r02p/b${N}_d$D.json').read().strip().splitlines()[-1])
print('defer=$D', d['ms_per_step'], d['value'], d['config'].get('dp_timing'), d['config'].get('rank_check'))
print('e2e', d['e2e']['value'], d['e2e']['full_batch']['value'])
PY
done
