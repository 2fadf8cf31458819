mkdir -p gpurun_out/r02c
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $T scripts/dp_check.py fused > gpurun_out/r02c/dpcheck.log 2>&1; tail -3 gpurun_out/r02c/dpcheck.log | cut -c1-600
UGN_DP_ONEGRAPH=1 timeout 600 $T scripts/dp_check.py fused > gpurun_out/r02c/dpcheck1g.log 2>&1; tail -3 gpurun_out/r02c/dpcheck1g.log | cut -c1-600
timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 3 --no-knn > gpurun_out/r02c/b2.json 2> gpurun_out/r02c/b2.err; python scripts/bline.py gpurun_out/r02c/b2.json || tail -5 gpurun_out/r02c/b2.err
UGN_DP_ONEGRAPH=1 timeout 600 $T bench.py --gpus 2 --steps 20 --warmup 3 --no-knn > gpurun_out/r02c/b2g.json 2> gpurun_out/r02c/b2g.err; python scripts/bline.py gpurun_out/r02c/b2g.json || tail -5 gpurun_out/r02c/b2g.err
