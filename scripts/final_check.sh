# final single-GPU validation: GPU tests, smoke(), default bench, reference arm
mkdir -p gpurun_out/final
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final/pytest_gpu.log 2>&1; tail -3 gpurun_out/final/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final/smoke.log 2>&1; tail -2 gpurun_out/final/smoke.log
s=$(date +%s); timeout 900 python bench.py > gpurun_out/final/bench.json 2> gpurun_out/final/bench.err; echo "bench rc=$? wall=$(( $(date +%s) - s ))s"; python scripts/bline.py gpurun_out/final/bench.json
s=$(date +%s); timeout 900 python bench.py --impl reference > gpurun_out/final/ref.json 2> gpurun_out/final/ref.err; echo "ref rc=$? wall=$(( $(date +%s) - s ))s"; tail -c 600 gpurun_out/final/ref.json
