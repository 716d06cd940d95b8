#!/bin/bash
# compute-sanitizer evidence (SURVEY 5 / VERDICT r1 item 4): memcheck, racecheck, synccheck and initcheck over smoke()
# and a trimmed subset of the GPU tests (the warp-synchronous shared-memory queues of k_fast_cells, the quadtree
# kernel's cell tables handed over through L2 atomics, the matcher, the RGB-D stage).  Run on the GPU box:
#   bash tools/sanitize.sh <tag>        -> gpurun_out/<tag>_sanitizer_<tool>.txt
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
SUBSET='tests/test_gpu_contract.py::test_stage_calls_are_rerunnable tests/test_gpu_contract.py::test_separate_angle_and_orb_entry_points tests/test_gpu_contract.py::test_match_knn_batch_fixed_stride tests/test_gpu_parity.py::test_degenerate_inputs tests/test_gpu_rgbd.py::test_rgbd_frame_stage_sequence tests/test_gpu_rgbd.py::test_compute_stereo_matches'
python -c "import __graft_entry__ as g; g.build()" || exit 1
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ $tool = memcheck ] && extra="--leak-check no --report-api-errors no"
  [ $tool = initcheck ] && extra="--track-unused-memory no"
  f=$out/${tag}_sanitizer_${tool}.txt
  echo "== compute-sanitizer --tool $tool : smoke()" > $f
  timeout 900 compute-sanitizer --tool $tool $extra --print-limit 30 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep -vE "^=+ *$" | tail -40 >> $f
  echo "== compute-sanitizer --tool $tool : pytest subset" >> $f
  timeout 1500 compute-sanitizer --tool $tool $extra --print-limit 30 python -m pytest -x -q $SUBSET 2>&1 | grep -vE "^=+ *$" | tail -60 >> $f
  grep -E "ERROR SUMMARY|passed|failed|smoke ok" $f
done
