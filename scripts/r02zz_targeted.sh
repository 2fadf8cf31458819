#!/bin/bash
# The targeted GPU checks that followed the full validation of round 2 (scripts/r02zz_final.sh), one per gpurun call
# (`scripts/grun.sh <timeout> 'bash scripts/r02zz_targeted.sh <step>'`); logs under gpurun_out/r02zz_<step>/ and, condensed,
# profiles/r02zz_<step>_pytest_gpu.log.
#   gs1  the 1-modality GaitSet graph: new tests + the whole GaitSet suite + the GaitSet builder protocol
#   gs2  postriplet == 2 with GaitSet branches + the _layer_output refactor: GaitSet, compat and edge suites
#   gs3  descriptor-extraction tests of the stacked engine + smoke()
#   gs4  the small entry points of nets/mj_uwyhNets_ba.py (stand-alone branch builders, fc_loadBranch, MatMul ...)
#   gs5  tf_shim.install(gpu_knn=True) against scikit-learn
STEP=${1:?step}
OUT=gpurun_out/r02zz_$STEP
mkdir -p $OUT
s0=$(date +%s)
T=tests/test_compat_gpu.py
case $STEP in
  gs1) SEL="tests/test_gaitset_gpu.py::test_gaitset_single_modality_graph_fp32 $T::test_gaitset_single_modality_builder $T::test_gaitset_builder_protocol tests/test_gaitset_gpu.py" ;;
  gs2) SEL="tests/test_gaitset_gpu.py::test_gaitset_postriplet2_graph_fp32 $T::test_postriplet2_gaitset_builder $T::test_postriplet2_builder tests/test_gaitset_gpu.py $T tests/test_edge_gpu.py" ;;
  gs3) SEL="tests/test_step_gpu.py tests/test_ops_gpu.py tests/test_decisions_gpu.py -k predict+or+expand+or+descriptor+or+video+or+open_world" ;;
  gs4) SEL="$T::test_standalone_branch_builders_and_small_entry_points" ;;
  gs5) SEL="$T::test_shim_gpu_knn_is_what_the_test_mains_import" ;;
  *) echo "unknown step $STEP"; exit 2 ;;
esac
timeout 170 python -m pytest -q ${SEL//+/ } > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$? t=$(( $(date +%s) - s0 ))s"; tail -n 25 $OUT/pytest_gpu.log | cut -c1-300
if [ $STEP = gs3 ]; then
  timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 $OUT/smoke.log
fi
