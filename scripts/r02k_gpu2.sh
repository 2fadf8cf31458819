mkdir -p gpurun_out/r02k
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515"
timeout 600 $T bench.py --gpus 2 --knn-only --knn-d 2048 > gpurun_out/r02k/k2.json 2> gpurun_out/r02k/k2.err; python -c "
import json;d=json.loads(open('gpurun_out/r02k/k2.json').read().strip().splitlines()[-1])['knn'];print({q:(round(d[q]['queries_per_s']), round(d[q]['ms'],3)) for q in ('Q4096','Q64')})"
UGN_KNN_GRAPH=0 timeout 600 $T bench.py --gpus 2 --knn-only --knn-d 2048 > gpurun_out/r02k/k2ng.json 2> gpurun_out/r02k/k2ng.err; python -c "
import json;d=json.loads(open('gpurun_out/r02k/k2ng.json').read().strip().splitlines()[-1])['knn'];print('nograph',{q:(round(d[q]['queries_per_s']), round(d[q]['ms'],3)) for q in ('Q4096','Q64')})"
